# correctness + default bench + other workloads (single GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err; tail -3 gpurun_out/bench_1080p.err
python bench.py --workload vga256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vga256.json 2> gpurun_out/bench_vga256.err; tail -3 gpurun_out/bench_vga256.err
python bench.py --workload 4k64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_4k64.json 2> gpurun_out/bench_4k64.err; tail -3 gpurun_out/bench_4k64.err
python bench.py --workload 8k --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8k.json 2> gpurun_out/bench_8k.err; tail -3 gpurun_out/bench_8k.err
python - <<'PY'
import json
for n in ('1080p','vga256','4k64','8k'):
    try:
        d=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), d['roofline']['per_tap_launch_ms'], 'stages', {k:round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'kp', d['config']['keypoints_per_step_per_gpu'], 'chunk', d['config']['resident_chunk'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
