"""Clip-wise sharding across the GPUs of one box.

Clips are independent -- all state (background, weights, frame ring, track ids) is per ``Clip`` -- so rank r of
``world`` takes a contiguous range of clips and runs the whole path on its own GPU.  There is NO collective on the
data path; ``torch.distributed`` (NCCL on GPUs, gloo on CPU for the tests) only provides the barrier, the
max-over-ranks of a timing and the sum of counters.  The reference's analogue is ``multiprocessing.Pool`` over files
(src/track/trackextractor.py:80-85).
"""
import os


def clip_range(n_clips, rank, world):
    """[lo, hi) of the clips rank ``rank`` owns: contiguous, sizes differ by at most one, in rank order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank {} / world {}".format(rank, world))
    base, extra = divmod(int(n_clips), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items, rank, world):
    lo, hi = clip_range(len(items), rank, world)
    return items[lo:hi]


def env_rank():
    """(rank, local_rank, world) from the torchrun environment (1 process: 0, 0, 1)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_distributed(backend=None):
    """Initialise the default process group when WORLD_SIZE > 1; returns torch.distributed or None."""
    rank, local_rank, world = env_rank()
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return dist


def _reduce(value, op_name, dist):
    import torch

    if dist is None:
        return value
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(value, dist):
    """Timing rule: a multi-GPU duration is the max over ranks."""
    return _reduce(float(value), "MAX", dist)


def sum_over_ranks(value, dist):
    return _reduce(float(value), "SUM", dist)


def extract_sharded(paths_or_clips, parse_clips, dist=None):
    """Run ``parse_clips(list_of_clips)`` on this rank's shard; returns (lo, hi, result)."""
    rank, _, world = env_rank()
    lo, hi = clip_range(len(paths_or_clips), rank, world)
    return lo, hi, parse_clips(paths_or_clips[lo:hi])
