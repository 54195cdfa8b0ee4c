# end-of-round measurement set -> gpurun_out/r2z_*   (ncu reports are digested on the box: only text comes back, gpurun_out is capped at 64 MiB)
TAG=r2z
if [ "$1" != "ncu" ]; then
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/${TAG}_c2.json 2> gpurun_out/${TAG}_c2.err
timeout 600 python bench.py --workload brdf --steps 3 --warmup 3 > gpurun_out/${TAG}_brdf.json 2> gpurun_out/${TAG}_brdf.err
cat gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_c4.json gpurun_out/${TAG}_c2.json gpurun_out/${TAG}_brdf.json gpurun_out/${TAG}_ref.json | python tools/show_bench.py
else
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --views 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt; rm -f gpurun_out/${TAG}_launches.csv
for k in k_trace_queue k_field_backward_scatter k_field_forward_tc5 k_field_backward_tc5v2 k_primary; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/${TAG}_prof_$k -f python tools/prof_step.py > gpurun_out/${TAG}_ncu_$k.log 2>&1
  python tools/ncu_summary.py raw /tmp/${TAG}_prof_$k.ncu-rep > gpurun_out/${TAG}_${k}_raw.txt 2>&1
  ncu -i /tmp/${TAG}_prof_$k.ncu-rep --page source --csv --print-source sass > /tmp/${TAG}_$k.csv 2>/dev/null
  python tools/ncu_sass_hot.py /tmp/${TAG}_$k.csv 30 > gpurun_out/${TAG}_${k}_sass_hot.txt 2>&1
done
head -12 gpurun_out/${TAG}_launches.txt
fi
