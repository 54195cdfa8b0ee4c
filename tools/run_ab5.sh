echo "== default"; timeout 400 python tools/perf_builders.py 2>&1 | tail -4
for so in iris_b200/_lib/ab/*.so; do echo "== $so"; IRIS_B200_LIB=$GRAFT_REPO_ROOT/$so timeout 400 python tools/perf_builders.py 2>&1 | tail -4; done
