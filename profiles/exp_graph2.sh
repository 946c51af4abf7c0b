for cfg in "SIFTCUDA_GRAPH=1" "SIFTCUDA_GRAPH=0"; do
  echo "== $cfg"; env $cfg python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f sync %.1f roof %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['sync_call']['value'],d['roofline']['frac']))"
done
