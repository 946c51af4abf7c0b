# e2e with host-resident outputs (zero-copy descriptor stores + early keypoint copy) vs explicit D2H
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for h in 1 0; do
  SIFTCUDA_HOST_OUT=$h python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('host_out', $h, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v,4) for k,v in s.items()})"
done
