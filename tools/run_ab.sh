# usage: bash tools/run_ab.sh <script> ; runs it with the default library and every A/B build under iris_b200/_lib/ab
echo "== default"; IRIS_PERF_QUICK=1 timeout 300 python $1 2>&1 | tail -1
for so in iris_b200/_lib/ab/*.so; do echo "== $so"; IRIS_PERF_QUICK=1 IRIS_B200_LIB=$GRAFT_REPO_ROOT/$so timeout 300 python $1 2>&1 | tail -1; done
