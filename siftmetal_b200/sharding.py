"""Frame sharding across ranks: frames are independent units, so a batch is split
into contiguous blocks, one per GPU, with no collective on the data path. Results are gathered on
the host in frame order."""
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
