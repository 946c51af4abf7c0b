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
import ctypes as _C

import numpy as np

# columns of the wire format in block order (SiftKeypointColumns, then SiftDescriptorColumns):
# (group, name, dtype, elements per row)
_COLS = (("kp", "absolute_x", "<f4", 1), ("kp", "absolute_y", "<f4", 1), ("kp", "sigma", "<f4", 1),
         ("kp", "value", "<f4", 1), ("kp", "sub_scale", "<f4", 1), ("kp", "scaled_xy", "<i2", 2),
         ("kp", "octave_scale", "u1", 2), ("desc", "features", "u1", 128), ("desc", "theta", "<f4", 1),
         ("desc", "keypoint", "<i4", 1))


def block_layout(cap_kp, cap_desc):
    """Byte offsets of the ten columns in a result block and its size — the layout of
    sift_result_layout (every column 256-byte aligned, in wire-format order)."""
    off, offsets = 0, []
    for group, _, dtype, per in _COLS:
        offsets.append(off)
        rows = cap_kp if group == "kp" else cap_desc
        off += (rows * np.dtype(dtype).itemsize * per + 255) // 256 * 256
    return offsets, max(off, 256)


class ShmGather:
    """One process per GPU, every rank on the same box. Each rank owns a POSIX shared-memory
    segment of a few result regions; with an engine its in-flight slots are BOUND to those regions
    (sift_register_host_memory / sift_bind_result_memory), so the kernels' own stores put the
    result columns where rank 0 can read them: the gather moves no data on the host. Rank 0 maps
    every segment and sees each call's results of all ranks in frame order (rank order = frame
    order, shards are contiguous). The only collective is one host barrier per call on a gloo side
    group. Regions rotate per call, so a call's columns stay intact until `regions` further calls
    have been submitted. Without an engine (CPU tests) publish() copies the columns instead."""

    HEADER_INTS = 4   # int64: frames in the call, frame offset in the shard, keypoints, descriptors

    def __init__(self, rank, world, engine=None, max_frames_per_call=1, calls_per_step=1, regions=4, cap_kp=None,
                 cap_desc=None, tag=None, group=None):
        import torch.distributed as dist
        from multiprocessing import shared_memory

        self.rank, self.world, self.engine, self.regions = rank, world, engine, regions
        if engine is not None:
            lay = engine.result_layout()
            cap_kp, cap_desc = int(lay.capacity_keypoints), int(lay.capacity_descriptors)
            offsets, nbytes = block_layout(cap_kp, cap_desc)
            assert offsets == list(lay.offset) and nbytes == int(lay.bytes), "block layout differs from the library's"
        self.cap_kp, self.cap_desc = int(cap_kp), int(cap_desc)
        self.group = group if group is not None else dist.new_group(backend="gloo")
        tag = tag or os.environ.get("MASTER_PORT", "0")
        mine = (int(max_frames_per_call), self.cap_kp, self.cap_desc, int(calls_per_step))
        self.shapes = [None] * world          # every rank's (frames per call, capacities, calls per step)
        dist.all_gather_object(self.shapes, mine, group=self.group)
        self.calls_per_step = max(s[3] for s in self.shapes)
        self.name = lambda r: f"siftgather_{tag}_{r}"
        self.own = shared_memory.SharedMemory(name=self.name(rank), create=True,
                                              size=regions * self._region_bytes(mine) + 4096)
        self._base = _C.addressof(_C.c_char.from_buffer(self.own.buf))
        self._pad = (-self._base) % 4096       # regions start page-aligned
        if engine is not None:
            engine.register_host_memory(self._base + self._pad, regions * self._region_bytes(mine))
        dist.barrier(group=self.group)
        self.peers, self._peer_pad = [], []
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
            for p in self.peers:
                a = _C.addressof(_C.c_char.from_buffer(p.buf))
                self._peer_pad.append((-a) % 4096)
        # every rank pads by its own mapping address; rank 0 needs the owners' paddings
        pads = [None] * world
        dist.all_gather_object(pads, self._pad, group=self.group)
        self._owner_pad = pads
        self.call = 0                 # calls submitted
        self.done = 0                 # calls gathered
        self._inflight = []           # region of each submitted, not yet published call
        self.stats = {"calls": 0, "frames": 0, "keypoints": 0, "descriptors": 0}
        self.step_calls = []          # rank 0: gathered calls of the current step

    # -- layout of one region: header | counts | result block ---------------------------------------
    def _header_bytes(self, shape):
        frames = shape[0]
        return (8 * self.HEADER_INTS + 3 * frames * 7 * 4 + 255) // 256 * 256

    def _region_bytes(self, shape):
        return (self._header_bytes(shape) + block_layout(shape[1], shape[2])[1] + 4095) // 4096 * 4096

    def _region_views(self, shm, pad, shape, region):
        base = pad + region * self._region_bytes(shape)
        frames, cap_kp, cap_desc = shape[0], shape[1], shape[2]
        v = {"meta": np.ndarray((self.HEADER_INTS,), "<i8", shm.buf, base)}
        o = base + 8 * self.HEADER_INTS
        for c in ("keypoint_counts", "descriptor_counts", "candidate_counts"):
            v[c] = np.ndarray((frames, 7), "<i4", shm.buf, o)
            o += frames * 7 * 4
        block = base + self._header_bytes(shape)
        offsets, _ = block_layout(cap_kp, cap_desc)
        for (group, name, dtype, per), off in zip(_COLS, offsets):
            rows = cap_kp if group == "kp" else cap_desc
            v[f"{group}.{name}"] = np.ndarray((rows, per) if per > 1 else (rows,), dtype, shm.buf, block + off)
        v["block_offset"] = block
        return v

    # -- per call --------------------------------------------------------------------------------------
    def before_submit(self):
        """Call right before engine.submit*: points the slot that submit will use at the next region."""
        region = self.call % self.regions
        if self.engine is not None:
            shape = self.shapes[self.rank]
            addr = self._base + self._pad + region * self._region_bytes(shape) + self._header_bytes(shape)
            self.engine.bind_result_memory(self.engine.next_slot(), addr, block_layout(shape[1], shape[2])[1])
        self._inflight.append(region)
        self.call += 1

    def publish(self, result, frame_offset):
        """The oldest submitted call has completed: record what it produced (and, without a bound
        engine, copy its columns into the region)."""
        region = self._inflight.pop(0)
        v = self._region_views(self.own, self._pad, self.shapes[self.rank], region)
        n = result.keypoint_counts.shape[0]
        for c in ("keypoint_counts", "descriptor_counts", "candidate_counts"):
            v[c][:n] = getattr(result, c)
        kc, dc = result.keypoint_columns, result.descriptor_columns
        nk, nd = len(kc), len(dc)
        if self.engine is None:
            if nk > self.cap_kp or nd > self.cap_desc:
                raise RuntimeError("ShmGather: region too small for this call's results")
            for group, name, _, _ in _COLS:
                src = getattr(kc if group == "kp" else dc, name)
                v[f"{group}.{name}"][:len(src)] = src
        v["meta"][:] = (n, frame_offset, nk, nd)

    def publish_empty(self):
        """A rank with fewer calls in the step than its peers still takes part in every barrier."""
        region = self.call % self.regions
        self.call += 1
        v = self._region_views(self.own, self._pad, self.shapes[self.rank], region)
        v["meta"][:] = (0, 0, 0, 0)

    def call_done(self):
        """Every rank has published call number `done`: barrier; rank 0 then holds that call's
        results of all ranks (views into the segments, nothing is copied)."""
        import torch.distributed as dist

        dist.barrier(group=self.group)
        region = self.done % self.regions
        if self.rank == 0:
            parts = [self._region_views(self.peers[r], self._owner_pad[r] if r else self._pad, self.shapes[r], region)
                     for r in range(self.world)]
            if self.done % self.calls_per_step == 0:
                self.step_calls = []
            self.step_calls.append(parts)
            self.stats["calls"] += 1
            self.stats["frames"] += int(sum(int(p["meta"][0]) for p in parts))
            self.stats["keypoints"] += int(sum(int(p["meta"][2]) for p in parts))
            self.stats["descriptors"] += int(sum(int(p["meta"][3]) for p in parts))
            for p in parts:   # the columns the kernels stored and the counts agree
                n = int(p["meta"][0])
                assert int(p["keypoint_counts"][:n].sum()) == int(p["meta"][2])
        self.done += 1

    def frame(self, rank, local_frame):
        """Rank 0: (keypoint column views, descriptor column views) of frame `local_frame` of rank
        `rank`'s shard, from the calls of the current step."""
        for parts in self.step_calls:
            p = parts[rank]
            n, off = int(p["meta"][0]), int(p["meta"][1])
            if off <= local_frame < off + n:
                lf = local_frame - off
                kc, dc = p["keypoint_counts"], p["descriptor_counts"]
                k0 = int(kc[:lf].sum()); k1 = k0 + int(kc[lf].sum())
                d0 = int(dc[:lf].sum()); d1 = d0 + int(dc[lf].sum())
                kp = {name: p[f"kp.{name}"][k0:k1] for g, name, _, _ in _COLS if g == "kp"}
                desc = {name: p[f"desc.{name}"][d0:d1] for g, name, _, _ in _COLS if g == "desc"}
                return kp, desc
        raise IndexError("frame not in the gathered calls of this step")

    def summary(self):
        s = dict(self.stats)
        s["transport"] = ("POSIX shared memory, one segment per rank; " +
                          ("the kernels store the result columns into it (zero host copies); "
                           if self.engine is not None else "columns copied once into it; ") +
                          "one gloo barrier per call")
        return s

    def close(self):
        import torch.distributed as dist

        self.step_calls = []
        try:
            dist.barrier(group=self.group)
        except Exception:
            pass
        for p in self.peers[1:]:
            p.close()
        try:
            self.own.close()
        except BufferError:
            pass              # ctypes / numpy views of the buffer are still alive: the unlink below frees the name
        try:
            self.own.unlink()
        except FileNotFoundError:
            pass
