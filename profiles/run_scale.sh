#!/bin/bash
# usage: run_scale.sh N  — bench lines of the three shardable / replicated workloads on N GPUs of one box
N=$1
run() {  # workload steps extra-env
  if [ "$N" = "1" ]; then
    env $3 python bench.py --gpus 1 --workload $1 --steps $2 --no-cpu-baseline
  else
    env $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --workload $1 --steps $2 --no-cpu-baseline
  fi
}
mkdir -p gpurun_out/scale
for wl in "1080p 100" "vga256 30" "4k64 6"; do
  set -- $wl
  run $1 $2 "SIFTCUDA_GRAPH=0" > gpurun_out/scale/${1}_n${N}.json 2> gpurun_out/scale/${1}_n${N}.err
  tail -2 gpurun_out/scale/${1}_n${N}.err
done
run 1080p 100 "SIFTCUDA_GRAPH=1" > gpurun_out/scale/1080p_graph_n${N}.json 2> gpurun_out/scale/1080p_graph_n${N}.err
run vga256 30 "SIFTCUDA_GRAPH=1" > gpurun_out/scale/vga256_graph_n${N}.json 2> gpurun_out/scale/vga256_graph_n${N}.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale/*_n*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1f"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], "sync %.1f"%d["e2e"]["sync_call"]["value"], d["e2e"].get("gather"))
    except Exception as e:
        print(f, "ERR", e)
PY
