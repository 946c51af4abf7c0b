for cfg in "SIFTCUDA_TAIL=0" "SIFTCUDA_TAIL=1" "SIFTCUDA_TAIL_PIXELS=9000"; do
  echo "== $cfg"; env $cfg python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']//d['steps']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done
for cfg in "SIFTCUDA_TAIL=0" "SIFTCUDA_TAIL=1"; do
  echo "== vga256 $cfg"; env $cfg python bench.py --workload vga256 --steps 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tailOctaves --csv -c 2 python bench.py --steps 1 --quick 2>&1 | grep tailOctaves | awk -F'","' '{print $5, $NF}' | head -3
SIFTCUDA_TAIL_PIXELS=9000 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tailOctaves --csv -c 2 python bench.py --steps 1 --quick 2>&1 | grep tailOctaves | awk -F'","' '{print $5, $NF}' | head -3
