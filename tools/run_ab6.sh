# c3 step (2 views) with each BVH builder variant: headline + per-kernel rates
run() { timeout 300 python bench.py --views 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_samples_per_s']
print('%.1f M  '%(d['value']/1e6)+'  '.join('%s %.2f'%(n.replace('k_',''),v/1e9) for n,v in k.items() if n in ('k_primary','k_trace_queue')), d['scene'])"; }
echo "== device: SAH top + treelets (default)"; run
echo "== device: LBVH top + SAH treelets"; IRIS_BENCH_OPTIONS=lbvh_sah_top=0 run
echo "== device: Morton LBVH"; IRIS_BENCH_OPTIONS=lbvh_sah_top=0,lbvh_sah_treelets=0 run
echo "== host binned SAH"; IRIS_BENCH_BUILDER=sah run
