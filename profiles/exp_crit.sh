for cfg in "SIFTCUDA_DEBUG_MAX_OCTAVE=0" "SIFTCUDA_DEBUG_MAX_OCTAVE=1" "SIFTCUDA_DEBUG_MAX_OCTAVE=2" "SIFTCUDA_DEBUG_MAX_OCTAVE=6" "SIFTCUDA_DEBUG_SKIP=1" "SIFTCUDA_DEBUG_SKIP=2" "SIFTCUDA_DEBUG_SKIP=3" "SIFTCUDA_DEBUG_SKIP=3 SIFTCUDA_DEBUG_MAX_OCTAVE=0"; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --no-cpu-baseline --quick 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms %.4f'%(d['ms_per_step']), {k: round(v,4) for k,v in d['stage_ms_per_step'].items()}, round(d['roofline']['avg_launch_ms']*5,4))"
done
