for cfg in "SIFTCUDA_BANDS=1" "SIFTCUDA_BANDS=2" "SIFTCUDA_BANDS=3" "SIFTCUDA_BANDS=4" "SIFTCUDA_PDL_TILES=600" "SIFTCUDA_PDL_TILES=2100"; do
  echo "== $cfg"; env $cfg python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f roof %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done
