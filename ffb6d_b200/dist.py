"""Multi-GPU plumbing of the hot path: frames are independent (every KNN / gather stays inside
one batch item: NN/knn_.cxx:109-113, models/ffb6d.py:172-174), so ranks take disjoint frame
ranges and the only cross-rank traffic is the timing reduction of bench.py.  No data-path
collective exists ("replicas / weak scaling")."""
import torch
import torch.distributed as dist


def frame_shard(frames_per_rank, rank, world):
    """Global frame ids of one rank under weak scaling: rank r owns [r*F, (r+1)*F)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return range(rank * frames_per_rank, (rank + 1) * frames_per_rank)


def split_frames(n_frames, rank, world):
    """Strong-scaling split of a fixed set of frames: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
