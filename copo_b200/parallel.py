"""Host-side data-parallel arithmetic (one process per GPU, `torch.distributed`): scene sharding, agreeing on the
number of minibatches, and the small reductions around the kernels.  No CUDA in here, so the world-size-2 `gloo`
tests exercise exactly the code the trainers run under NCCL.

Replaces the reference's actor parallelism (`num_rollout_workers`, `synchronous_parallel_sample`, `sync_weights`,
LCF broadcast; torch_copo/algo_ippo.py:32-36, algo_copo.py:518-525, 572-613) - see SURVEY.md 8e.
"""
import math

import torch


def active(dist):
    return dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def scene_offset(rank, scenes_per_rank):
    """Global index of a rank's first scene: scenes shard contiguously, RNG streams are keyed by the global index."""
    return int(rank) * int(scenes_per_rank)


def num_minibatches(n_rows, minibatch_size, dist=None, device="cpu"):
    """Every rank must run the same number of gradient all-reduces per epoch: the max over ranks of
    ceil(rows / minibatch_size)."""
    k = max(1, math.ceil(n_rows / max(1, minibatch_size)))
    if active(dist):
        t = torch.tensor([k], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k = int(t.item())
    return k


def minibatch_bounds(n_rows, k):
    """k contiguous [begin, end) slices covering n_rows (the last ones may be shorter, never empty if n_rows > 0)."""
    size = max(1, math.ceil(n_rows / k))
    return [(min(j * size, max(n_rows - 1, 0)), min((j + 1) * size, n_rows)) if j * size < n_rows else (0, min(1, n_rows))
            for j in range(k)]


def allreduce_sum_(t, dist=None):
    if active(dist):
        dist.all_reduce(t)
    return t


def allreduce_mean_(t, dist=None):
    """Gradient averaging over ranks (the flat gradient buffer, or g_new / g_old before their dot product)."""
    if active(dist):
        dist.all_reduce(t)
        t /= dist.get_world_size()
    return t


def mean_std_from_sums(s, s2, n):
    """Population mean / std from (sum x, sum x^2, n) - the statistics behind rllib's `standardized`."""
    mean = s / max(n, 1.0)
    var = max(s2 / max(n, 1.0) - mean * mean, 0.0)
    return mean, math.sqrt(var)
