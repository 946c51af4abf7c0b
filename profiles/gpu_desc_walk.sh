# A/B of the descriptor walk modes (SIFTCUDA_DESC_WALK = 0 flattened spans, 4/8/16 tile columns)
for w in 0 16 8 4; do
  SIFTCUDA_DESC_WALK=$w python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('walk', $w, 'fps', round(d['value'],1), 'desc', round(s['descriptor'],4), 'ori', round(s['orientation'],4), 'pyr', round(s['pyramid'],4))"
  SIFTCUDA_DESC_WALK=$w timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "butterfly or synthetic or batch" 2>&1 | tail -1
done
