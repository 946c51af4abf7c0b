# Round-2 record run: GPU tests, default bench line, ncu launch list of the same command, ncu --set full
# captures of the hot kernels, the other BASELINE workloads, sanitizer passes. Everything lands in
# gpurun_out/r2/ (summarised by profiles/summarize.py profiles/r2).
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_1080p.json 2> $O/bench_1080p.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --quick > $O/ncu_launch.log 2>&1
# full-plane blur launches (row bands off so that one launch = one scale of the whole plane)
SIFTCUDA_BANDS=1 ncu --set full --clock-control none --import-source on -k regex:blurKernel -s 1 -c 5 -o $O/prof_blur python bench.py --steps 2 --quick > $O/ncu_blur.log 2>&1
for k in descriptorKernel orientationKernel extremaMaskTmaKernel tailOctavesKernel gradientKernel grayUpsample2xKernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o $O/prof_$k python bench.py --steps 2 --quick > $O/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:matchKernel -c 1 -o $O/prof_matchKernel python profiles/match_bench.py 20000 > $O/ncu_match.log 2>&1
python profiles/match_bench.py 35000 > $O/match_bench.json 2> $O/match_bench.err; cat $O/match_bench.json
python bench.py --workload vga256 --steps 30 --no-cpu-baseline > $O/bench_vga256.json 2> $O/bench_vga256.err
python bench.py --workload 4k64 --steps 6 --no-cpu-baseline > $O/bench_4k64.json 2> $O/bench_4k64.err
python bench.py --workload 8k --steps 6 --no-cpu-baseline > $O/bench_8k.json 2> $O/bench_8k.err
python bench.py --input-format gray8 --steps 100 --no-cpu-baseline > $O/bench_1080p_gray8.json 2> $O/bench_gray8.err
SIFTCUDA_GRAPH=1 python bench.py --steps 100 --no-cpu-baseline > $O/bench_1080p_graph.json 2> $O/bench_graph.err
python bench.py --impl reference --steps 5 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'PY'
import json
for n in ('1080p','vga256','4k64','8k','1080p_gray8','1080p_graph','reference'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r2/bench_{n}.json') if l.startswith('{')][-1])
        print(n, 'fps', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d.get('roofline',{}).get('frac',0),3), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(n, 'FAILED', e)
PY
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python __graft_entry__.py --smoke > $O/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' $O/sanitize_$tool.log | tr '\n' ' ')"
done
