# N-GPU runs (N = $1): default workload (replicas), vga256 (sharded batch), reference arm
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload vga256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vga256_n$N.json 2>> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err
python - <<PY
import json
for n in ('1080p_n$N','vga256_n$N','ref_n$N'):
    try:
        txt=[l for l in open(f'gpurun_out/bench_{n}.json') if l.startswith('{')]
        d=json.loads(txt[-1]); print(n, 'lines', len(txt), 'fps', round(d['value'],1), 'n_gpus', d['n_gpus'], 'scaling', d['scaling'], 'e2e', round(d['e2e']['value'],1), d['config']['workload'][:60])
    except Exception as e: print(n, 'FAILED', e)
PY
