echo "== th20"; timeout 300 python tools/perf_persist.py 2>&1 | tail -1
for th in 8 28; do echo "== th$th"; IRIS_B200_LIB=$GRAFT_REPO_ROOT/iris_b200/_lib/ab/libiris_th$th.so timeout 300 python tools/perf_persist.py 2>&1 | tail -1; done
