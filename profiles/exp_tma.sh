./profiles/micro/graph_pdl
for cfg in "SIFTCUDA_BLUR_TMA=1" "SIFTCUDA_BLUR_TMA=0"; do
  echo "== $cfg"; env SIFTCUDA_GRAPH=0 $cfg python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f roof %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']), d['roofline']['isolated_launch_ms'], d['stage_ms_per_step'])"
done
