for cfg in "SIFTCUDA_EXTREMA_TMA=0" "SIFTCUDA_EXTREMA_TMA=1" "SIFTCUDA_EXTREMA_ROWS=30" "SIFTCUDA_EXTREMA_ROWS=90"; do
  echo "== $cfg"; env $cfg python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms %.4f e2e %.1f'%(d['value'],d['ms_per_step'],d['e2e']['value']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()})"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:extremaMask --csv -c 12 python bench.py --steps 1 --quick 2>&1 | grep extremaMask | awk -F'","' '{print $5, $12, $NF}' | head -12
