run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'pyramid', round(d['stage_ms_per_step']['pyramid'],4), 'blur5', round(sum(d['roofline']['per_tap_launch_ms']),4))"; }
run all
SIFTCUDA_DEBUG_MAX_OCTAVE=0 run oct0_only
SIFTCUDA_DEBUG_MAX_OCTAVE=1 run oct01
SIFTCUDA_DEBUG_MAX_OCTAVE=0 SIFTCUDA_DEBUG_SKIP=1 run oct0_nograd
SIFTCUDA_DEBUG_MAX_OCTAVE=0 SIFTCUDA_DEBUG_SKIP=2 run oct0_noext
SIFTCUDA_DEBUG_MAX_OCTAVE=0 SIFTCUDA_DEBUG_SKIP=3 run oct0_bluronly
SIFTCUDA_DEBUG_SKIP=3 run all_bluronly
