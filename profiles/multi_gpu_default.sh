mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1080p_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
txt=[l for l in open('gpurun_out/bench_1080p_n$N.json') if l.startswith('{')]
d=json.loads(txt[-1]); print('lines', len(txt), 'fps', round(d['value'],1), 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],1), 'ms', d['ms_per_step'])
PY
nproc
