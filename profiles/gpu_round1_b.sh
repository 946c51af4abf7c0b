# tests + bench + ncu launch list + full capture of the octave-0 blur launches
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_1080p.json'))
print({k:d[k] for k in ('value','ms_per_step','wall_ms_per_step','stage_ms_per_step','gpu_launches')})
print(d['roofline']); print(d['e2e']); print(d.get('cpu_baseline')); print(d['clocks'])
PY
tail -5 gpurun_out/bench_1080p.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --quick > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blurKernel -s 1 -c 5 -o gpurun_out/prof_blur python bench.py --steps 1 --quick > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
