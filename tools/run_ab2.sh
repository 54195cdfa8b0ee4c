# quick_perf with the default library and every A/B build: forward / adjoint rates of path_tracing_single
run() { timeout 400 python tools/quick_perf.py 1000000 2>&1 | grep -E "single_bwd_brdf|field_bwd|Error|error" | tr -d '\n'; echo; }
if [ -z "$SKIP_DEFAULT" ]; then echo "== default"; run; fi
for so in iris_b200/_lib/ab/*.so; do echo "== $so"; IRIS_B200_LIB=$GRAFT_REPO_ROOT/$so run; done
