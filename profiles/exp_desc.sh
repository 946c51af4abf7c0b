for parts in 0 3 4 8; do
  echo "== SIFTCUDA_DESC_PARTS=$parts"; env SIFTCUDA_DESC_PARTS=$parts python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done
