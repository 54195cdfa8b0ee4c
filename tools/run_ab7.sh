# c3 step (2 views) with different amounts of L2 set aside for the persisting hash-grid window
run() { timeout 300 python bench.py --views 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_samples_per_s']
print('%.1f M  '%(d['value']/1e6)+'  '.join('%s %.2f'%(n.replace('k_',''),v/1e9) for n,v in k.items()))"; }
for mb in 0 32 64 96 128; do echo "== l2_persist_mb=$mb"; IRIS_BENCH_OPTIONS=l2_persist_mb=$mb run; done
