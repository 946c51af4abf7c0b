# Round-1 record run: GPU tests, default bench line, ncu launch list, ncu --set full captures of the hot
# kernels, the other BASELINE workloads. Everything lands in gpurun_out/ (summarised by profiles/summarize.py).
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --quick > gpurun_out/ncu_launch.log 2>&1
# full-plane blur launches (row bands off so that one launch = one scale of the whole plane)
SIFTCUDA_BANDS=1 ncu --set full --clock-control none --import-source on -k regex:blurKernel -s 1 -c 5 -o gpurun_out/prof_blur python bench.py --steps 1 --quick > gpurun_out/ncu_full.log 2>&1
for k in descriptorKernel orientationKernel extremaMaskKernel gradientKernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_$k python bench.py --steps 1 --quick > gpurun_out/ncu_$k.log 2>&1
done
python bench.py --workload vga256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vga256.json 2> gpurun_out/bench_vga256.err
python bench.py --workload 4k64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_4k64.json 2> gpurun_out/bench_4k64.err
python bench.py --workload 8k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8k.json 2> gpurun_out/bench_8k.err
python - <<'PY'
import json
for n in ('1080p','vga256','4k64','8k'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/bench_{n}.json') if l.startswith('{')][-1])
        print(n, 'fps', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(n, 'FAILED', e)
PY
