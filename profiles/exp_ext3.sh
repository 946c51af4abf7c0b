for wl in "vga256 20" "4k64 4" "8k 4"; do set -- $wl
for cfg in "SIFTCUDA_EXTREMA_TMA=0" "SIFTCUDA_EXTREMA_TMA=1"; do
  echo "== $1 $cfg"; env $cfg python bench.py --workload $1 --steps $2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done; done
