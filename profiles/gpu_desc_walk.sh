# A/B of the descriptor walk modes (SIFTCUDA_DESC_WALK = 0 flattened spans, 1 flattened aligned pairs,
# 4/8/16 tile columns) and warps per CTA
run() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('$1', 'fps', round(d['value'],1), 'desc', round(s['descriptor'],4), 'ori', round(s['orientation'],4), 'pyr', round(s['pyramid'],4))"
}
for w in 0 1; do for wp in 2 7 1; do
  export SIFTCUDA_DESC_WALK=$w SIFTCUDA_DESC_WARPS=$wp
  run "walk $w warps $wp"
done; done
SIFTCUDA_DESC_WALK=1 SIFTCUDA_DESC_WARPS=7 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
