mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p_n2.json 2> gpurun_out/bench_n2.err
tail -2 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload vga256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vga256_n2.json 2>> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>> gpurun_out/bench_n2.err
python - <<'PY'
import json
for n in ('1080p_n2','vga256_n2','ref_n2'):
    try:
        txt=[l for l in open(f'gpurun_out/bench_{n}.json') if l.startswith('{')]
        d=json.loads(txt[-1]); print(n, 'lines', len(txt), 'fps', round(d['value'],1), 'n_gpus', d['n_gpus'], 'scaling', d['scaling'], 'e2e', round(d['e2e']['value'],1), d['config']['workload'][:60])
    except Exception as e: print(n, 'FAILED', e)
PY
