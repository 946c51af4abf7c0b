timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streaming" 2>&1 | tail -5
timeout 200 python profiles/blur_stream_ab.py 3 2 1
for v in 0 1; do
SIFTCUDA_BLUR_STREAM=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('stream $v', 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), {k: round(v,4) for k,v in s.items()})"
done
