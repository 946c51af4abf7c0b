mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python profiles/blur_modes.py
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_1080p.json'))
print({k:d[k] for k in ('value','ms_per_step','stage_ms_per_step')})
print(d['roofline']['frac'], d['roofline']['per_tap_launch_ms'], d['e2e']['value'])
PY
tail -3 gpurun_out/bench_1080p.err
