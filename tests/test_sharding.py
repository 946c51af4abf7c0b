"""Multi-GPU path on CPU: shard arithmetic and a world_size-2 gloo run of the bench's
barrier / max-over-ranks / host-gather pattern (no collective on the data path)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT
from siftmetal_b200.sharding import all_shards, chunks, shard_range


@pytest.mark.parametrize("n,world", [(256, 1), (256, 2), (256, 8), (64, 4), (7, 8), (1, 2), (0, 3), (13, 5)])
def test_shards_partition_the_batch(n, world):
    sh = all_shards(n, world)
    assert sh[0][0] == 0 and sh[-1][1] == n
    for (a0, a1), (b0, b1) in zip(sh, sh[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in sh]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def test_shard_argument_checks():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        shard_range(10, 0, 0)
    assert chunks(10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert chunks(0, 4) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(
        """
        import os, sys, json
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        from siftmetal_b200.sharding import shard_range
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 37
        a, b = shard_range(n, rank, world)
        # stand-in for per-frame results: (frame id, keypoint count); each rank only touches its frames
        local = [(f, 100 + f) for f in range(a, b)]
        dist.barrier()
        t = torch.tensor([0.5 + rank], dtype=torch.float64)      # per-rank device time
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, local)                   # host-side result gathering only
        if rank == 0:
            flat = [x for part in gathered for x in part]
            print(json.dumps({"frames": [f for f, _ in flat], "max_time": float(t), "world": world}))
        dist.destroy_process_group()
        """ % ROOT))
    port = _free_port()
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=240, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["frames"] == list(range(37)) and out["max_time"] == 1.5


def test_two_rank_shared_memory_gather(tmp_path):
    """ShmGather (the bench's host gather) in copy mode: two gloo ranks publish fake result columns
    of their frame shards call by call; rank 0 reads every frame of the job in frame order from the
    mapped segments. (The zero-copy mode, slots bound to the segment, needs a GPU.)"""
    script = tmp_path / "gworker.py"
    script.write_text(textwrap.dedent(
        """
        import os, sys, json
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from siftmetal_b200.sharding import ShmGather, shard_range, block_layout
        from siftmetal_b200.api import BatchResult, KeypointColumns, DescriptorColumns
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        total = 9
        a, b = shard_range(total, rank, world)
        sizes = np.ones((7, 2), np.float32)

        def fake(frames, step):      # frame f owns f + 1 keypoints in octave 0, each with one descriptor
            n = sum(f + 1 for f in frames)
            kc = np.zeros((len(frames), 7), np.int32); kc[:, 0] = [f + 1 for f in frames]
            ids = np.concatenate([np.full(f + 1, 100 * step + f, np.float32) for f in frames])
            kv = KeypointColumns(ids, ids, ids, ids, ids, np.zeros((n, 2), np.int16), np.zeros((n, 2), np.uint8), sizes)
            dv = DescriptorColumns(np.repeat(ids.astype(np.uint8)[:, None], 128, 1), ids, np.arange(n, dtype=np.int32))
            return BatchResult(kv, dv, kc, kc.copy(), kc.copy())

        g = ShmGather(rank, world, max_frames_per_call=3, calls_per_step=2, cap_kp=64, cap_desc=64)
        starts = [shard_range(total, r, world)[0] for r in range(world)]
        out = []
        for step in range(3):
            mine = list(range(a, b))
            half = len(mine) // 2               # two calls per step
            for part, off in ((mine[:half], 0), (mine[half:], half)):
                g.before_submit()
                g.publish(fake(part, step), off)
                g.call_done()
            if rank == 0:
                for f in range(total):
                    r = max(i for i in range(world) if starts[i] <= f)
                    kp, desc = g.frame(r, f - starts[r])
                    assert len(kp["sigma"]) == f + 1 and np.all(kp["sigma"] == 100 * step + f), (step, f)
                    assert desc["features"].shape == (f + 1, 128) and np.all(desc["theta"] == 100 * step + f)
                out.append(g.stats["keypoints"])
        if rank == 0:
            print(json.dumps({"sums": out, "stats": g.summary(), "layout": block_layout(1000, 2000)[1]}))
        g.close()
        dist.destroy_process_group()
        """ % ROOT))
    port = _free_port()
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=240, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-3000:]
    import json

    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["sums"] == [45, 90, 135] and out["stats"]["frames"] == 27 and out["stats"]["calls"] == 6
    assert out["layout"] == 5 * 4096 + 4096 + 2048 + 256000 + 2 * 8192
