run() { timeout 400 python tools/quick_perf.py 1000000 2>&1 | grep -E "field_fwd|single_fwd_M|single_fwd_rec|Error|error" | tr -d '\n'; echo; }
echo "== default"; run
for so in iris_b200/_lib/ab/*.so; do echo "== $so"; IRIS_B200_LIB=$GRAFT_REPO_ROOT/$so run; done
