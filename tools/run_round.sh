# end-of-round measurement set -> gpurun_out/*_$TAG.*   (TAG e.g. r1f)
TAG=${1:-r1f}
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_${TAG}_c2.err
timeout 600 python bench.py --workload brdf --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_brdf.json 2> gpurun_out/bench_${TAG}_brdf.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_queue|k_field_backward_scatter|k_field_backward_dgrad|k_field_forward_tc5|k_field_backward_wgrad|k_single_gen|k_single_shade|k_primary" -s 17 -c 12 -o gpurun_out/prof_${TAG} -f python tools/prof_step.py > gpurun_out/ncu_full_${TAG}.log 2>&1
cat gpurun_out/bench_${TAG}*.json | python tools/show_bench.py
tail -2 gpurun_out/ncu_full_${TAG}.log
