# c3 step (2 views, 2 steps) with the default library and with every A/B build under iris_b200/_lib/ab/: headline + per-kernel rates
run() { timeout 300 python bench.py --views 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_samples_per_s']
print('%.1f M  '%(d['value']/1e6)+'  '.join('%s %.2f'%(n.replace('k_',''),v/1e9) for n,v in k.items()))"; }
echo "== default"; run
for so in iris_b200/_lib/ab/*.so; do echo "== $so"; IRIS_B200_LIB=$GRAFT_REPO_ROOT/$so run; done
