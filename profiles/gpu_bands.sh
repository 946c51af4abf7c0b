for nb in 1 2 3 4; do
  SIFTCUDA_BANDS=$nb python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bands', $nb, 'fps', round(d['value'],1), 'pyramid', round(d['stage_ms_per_step']['pyramid'],4), 'blur5', round(sum(d['roofline']['per_tap_launch_ms']),4), 'frac', round(d['roofline']['frac'],3))"
done
SIFTCUDA_BANDS=3 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
