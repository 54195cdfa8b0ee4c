TAG=${1:-r1h}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err
timeout 600 python bench.py --workload brdf --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_brdf.json 2> gpurun_out/bench_${TAG}_brdf.err
cat gpurun_out/bench_${TAG}.json gpurun_out/bench_${TAG}_c4.json gpurun_out/bench_${TAG}_brdf.json gpurun_out/bench_${TAG}_ref.json | python tools/show_bench.py
