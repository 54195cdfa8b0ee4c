# r2a: baseline state on the GPU + first C5 numbers + ncu of the fused field adjoint (never profiled) and of k_trace_queue at 5M triangles
TAG=r2a
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --workload c5 --views 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_c5v4.json 2> gpurun_out/${TAG}_c5v4.err; tail -c 3000 gpurun_out/${TAG}_c5v4.json; tail -5 gpurun_out/${TAG}_c5v4.err
timeout 600 python bench.py --workload c3 --views 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_c3v2.json 2> gpurun_out/${TAG}_c3v2.err; tail -c 3000 gpurun_out/${TAG}_c3v2.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_field_backward_tc5|k_field_backward_scatter|k_field_forward_tc5|k_trace_queue" -s 8 -c 8 -o gpurun_out/prof_${TAG} -f python tools/prof_step.py > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_full_${TAG}.log
IRIS_PROF_TRIS=5000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_queue" -s 2 -c 2 -o gpurun_out/prof_${TAG}_5m -f python tools/prof_step.py > gpurun_out/ncu_full_${TAG}_5m.log 2>&1; tail -2 gpurun_out/ncu_full_${TAG}_5m.log
