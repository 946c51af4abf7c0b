"""Frame sharding across ranks: frames are independent units, so a batch is split
into contiguous blocks, one per GPU, with no collective on the data path. Results are gathered on
the host in frame order."""
import os
from typing import List, Tuple


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the frames rank `rank` owns: frame f goes to rank f * world // n_frames,
    i.e. contiguous blocks whose sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shards(n_frames: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n_frames, r, world) for r in range(world)]


def chunks(n_local: int, chunk: int) -> List[Tuple[int, int]]:
    """(offset, count) pieces of a rank's frames that fit the context's max_batch."""
    if chunk < 1:
        raise ValueError("chunk must be >= 1")
    return [(s, min(chunk, n_local - s)) for s in range(0, n_local, chunk)]


# ------------------------------------------------------------------------------------------------
# Host-side result gathering ("results are D2H-copied per GPU and concatenated on
# the host in frame order"; no NCCL, no NVLink traffic).
import numpy as np

_COLS = (  # (name, dtype, elements per item) of the keypoint / descriptor columns of the wire format
    ("kp", (("absolute_x", "<f4", 1), ("absolute_y", "<f4", 1), ("sigma", "<f4", 1), ("value", "<f4", 1),
            ("sub_scale", "<f4", 1), ("scaled_xy", "<i2", 2), ("octave_scale", "u1", 2))),
    ("desc", (("features", "u1", 128), ("theta", "<f4", 1), ("keypoint", "<i4", 1))),
)


class ShmGather:
    """One process per GPU, every rank on the same box: each rank publishes the result columns of
    its frame shard in a POSIX shared-memory segment; rank 0 maps all segments and sees the whole
    job's results in frame order (rank order = frame order, shards are contiguous). Data moves once
    (pinned result columns → the rank's own segment, all ranks in parallel); the only collective is
    one host barrier per step on a gloo side group. Segments are double-buffered by step parity so
    that one barrier per step is enough."""

    def __init__(self, rank, world, engine=None, n_local=1, cap_kp=None, cap_desc=None, tag=None, group=None):
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.rank, self.world, self.n_local = rank, world, n_local
        if cap_kp is None:      # half of the per-frame list capacities (themselves >= 4x the 1/f-noise density)
            cap_kp = max(1024, engine.info.max_keypoints_per_frame * n_local // 2)
            cap_desc = max(1024, engine.info.max_descriptors_per_frame * n_local // 2)
        self.cap_kp, self.cap_desc = int(cap_kp), int(cap_desc)
        self.group = group if group is not None else dist.new_group(backend="gloo")
        tag = tag or os.environ.get("MASTER_PORT", "0")
        # every rank must know every other rank's shard size and capacities to map its segment
        mine = (n_local, self.cap_kp, self.cap_desc)
        self.shapes = [None] * world
        dist.all_gather_object(self.shapes, mine, group=self.group)
        self.name = lambda r: f"siftgather_{tag}_{r}"
        self.own = shared_memory.SharedMemory(name=self.name(rank), create=True, size=2 * self._half_bytes(*mine))
        dist.barrier(group=self.group)
        self.peers = []
        if rank == 0:
            self.peers = [self.own] + [shared_memory.SharedMemory(name=self.name(r)) for r in range(1, world)]
            # attaching registers the segment with this process's resource tracker (Python < 3.13),
            # which would unlink it a second time at exit: the owner rank unlinks its own segment
            try:
                from multiprocessing import resource_tracker
                for p in self.peers[1:]:
                    resource_tracker.unregister(p._name, "shared_memory")
            except Exception:
                pass
        self.step = 0
        self._reset_cursor()
        self.stats = {"steps": 0, "frames": 0, "keypoints": 0, "descriptors": 0}
        self.last = None

    @staticmethod
    def _layout(n_frames, cap_kp, cap_desc):
        """name -> (byte offset, dtype, shape) inside one half of a segment."""
        off, lay = 0, {}

        def put(name, dtype, shape):
            nonlocal off
            lay[name] = (off, np.dtype(dtype), shape)
            off += int(np.dtype(dtype).itemsize * int(np.prod(shape)))
            off = (off + 63) // 64 * 64

        put("totals", "<i8", (2,))
        for c in ("keypoint_counts", "descriptor_counts", "candidate_counts"):
            put(c, "<i4", (n_frames, 7))
        for group, cols in _COLS:
            cap = cap_kp if group == "kp" else cap_desc
            for name, dtype, per in cols:
                put(f"{group}.{name}", dtype, (cap, per) if per > 1 else (cap,))
        return lay, off

    def _half_bytes(self, n_frames, cap_kp, cap_desc):
        return self._layout(n_frames, cap_kp, cap_desc)[1]

    def _views(self, shm, shape, half):
        lay, size = self._layout(*shape)
        base = half * size
        return {k: np.ndarray(s, dtype=d, buffer=shm.buf, offset=base + o) for k, (o, d, s) in lay.items()}

    def _reset_cursor(self):
        self.kp_at = self.desc_at = 0

    def publish(self, result, frame_offset):
        """Copies one call's result (frames [frame_offset, +n) of this rank's shard) into the segment."""
        v = self._views(self.own, self.shapes[self.rank], self.step & 1)
        n = result.keypoint_counts.shape[0]
        for c in ("keypoint_counts", "descriptor_counts", "candidate_counts"):
            v[c][frame_offset:frame_offset + n] = getattr(result, c)
        kc, dc = result.keypoint_columns, result.descriptor_columns
        nk, nd = len(kc), len(dc)
        if self.kp_at + nk > self.cap_kp or self.desc_at + nd > self.cap_desc:
            raise RuntimeError("ShmGather: segment too small for this step's results")
        for name, _, _ in _COLS[0][1]:
            v[f"kp.{name}"][self.kp_at:self.kp_at + nk] = getattr(kc, name)
        for name, _, _ in _COLS[1][1]:
            v[f"desc.{name}"][self.desc_at:self.desc_at + nd] = getattr(dc, name)
        self.kp_at += nk
        self.desc_at += nd
        v["totals"][:] = (self.kp_at, self.desc_at)

    def step_done(self):
        """All shards of the step are published: barrier, then rank 0 holds the job's results in
        frame order (views into the segments, nothing is copied again)."""
        import torch.distributed as dist

        dist.barrier(group=self.group)
        if self.rank == 0:
            parts = [self._views(self.peers[r], self.shapes[r], self.step & 1) for r in range(self.world)]
            counts = np.concatenate([p["keypoint_counts"] for p in parts])      # [total frames, 7], frame order
            dcounts = np.concatenate([p["descriptor_counts"] for p in parts])
            self.last = {"keypoint_counts": counts, "descriptor_counts": dcounts, "parts": parts}
            self.stats["steps"] += 1
            self.stats["frames"] += int(counts.shape[0])
            self.stats["keypoints"] += int(sum(int(p["totals"][0]) for p in parts))
            self.stats["descriptors"] += int(sum(int(p["totals"][1]) for p in parts))
            assert int(counts.sum()) == sum(int(p["totals"][0]) for p in parts)
        self.step += 1
        self._reset_cursor()

    def frame(self, f):
        """Rank 0: (keypoint column views, descriptor column views) of job frame f of the last step."""
        starts = np.cumsum([0] + [s[0] for s in self.shapes])
        r = int(np.searchsorted(starts, f, side="right")) - 1
        p, lf = self.last["parts"][r], f - int(starts[r])
        k0 = int(p["keypoint_counts"][:lf].sum()); k1 = k0 + int(p["keypoint_counts"][lf].sum())
        d0 = int(p["descriptor_counts"][:lf].sum()); d1 = d0 + int(p["descriptor_counts"][lf].sum())
        kp = {name: p[f"kp.{name}"][k0:k1] for name, _, _ in _COLS[0][1]}
        desc = {name: p[f"desc.{name}"][d0:d1] for name, _, _ in _COLS[1][1]}
        return kp, desc

    def summary(self):
        s = dict(self.stats)
        s["transport"] = "POSIX shared memory, one segment per rank, one gloo barrier per step"
        return s

    def close(self):
        import torch.distributed as dist

        self.last = None
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        for p in self.peers[1:]:
            p.close()
        self.own.close()
        try:
            self.own.unlink()
        except FileNotFoundError:
            pass
