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
