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
    """Minibatches per epoch when every rank cuts its own `n_rows` into pieces of `minibatch_size`: the max over ranks of
    ceil(rows / minibatch_size) (every rank must run the same number of gradient all-reduces)."""
    k = max(1, math.ceil(n_rows / max(1, minibatch_size)))
    if active(dist):
        t = torch.tensor([k], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k = int(t.item())
    return k


def minibatch_bounds(n_rows, k):
    """k contiguous [begin, end) slices covering n_rows, sizes differing by at most one row (numpy.array_split);
    slices are empty only when n_rows < k."""
    base, extra = divmod(int(n_rows), int(k))
    out, lo = [], 0
    for j in range(k):
        hi = lo + base + (1 if j < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_row_counts(n_rows, dist=None, device="cpu"):
    """[world] row counts of every rank (one small all-gather per training iteration)."""
    if not active(dist):
        return [int(n_rows)]
    t = torch.tensor([int(n_rows)], dtype=torch.int64, device=device)
    parts = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return [int(p.item()) for p in parts]


def minibatch_plan_all(row_counts, minibatch_size):
    """The reference cuts the WHOLE train batch (all workers' rows) into consecutive minibatches of
    `sgd_minibatch_size` rows, the last one ragged (rllib.utils.sgd.minibatches, called from train_one_step;
    algo_copo.py:555).  Data parallel: the global batch is the union of the ranks' rows and every rank contributes a
    FIXED quota of q = minibatch_size // world of its own (shuffled) rows to each minibatch while it has rows left -
    with one rank this is exactly rllib's slicing, and the per-rank minibatch shape repeats from step to step and from
    iteration to iteration (so the captured gradient step, policy._graphed_step, is reused; only the ragged tail runs
    eagerly).  k = max over ranks of ceil(n_r / q); a rank that has run out contributes an empty slice.  Returns
    (per-rank lists of k [begin, end) slices, the k GLOBAL minibatch sizes): losses and gradients are normalised by the
    global size, so the all-reduced sum of the ranks' gradients is the gradient of the whole-minibatch mean whatever the
    split."""
    world = max(1, len(row_counts))
    q = max(1, int(minibatch_size) // world)
    k = max(1, max(math.ceil(int(n) / q) for n in row_counts))
    per_rank = [[(min(j * q, int(n)), min((j + 1) * q, int(n))) for j in range(k)] for n in row_counts]
    sizes = [sum(b[j][1] - b[j][0] for b in per_rank) for j in range(k)]
    return per_rank, sizes


def minibatch_plan(row_counts, rank, minibatch_size):
    """This rank's slices of minibatch_plan_all and the global minibatch sizes."""
    per_rank, sizes = minibatch_plan_all(row_counts, minibatch_size)
    return per_rank[rank], sizes


class AllReduceTimer:
    """CUDA-event time spent inside the all-reduces of one training iteration (on the launching stream)."""

    def __init__(self):
        self.pairs, self.count = [], 0

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, e0):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.pairs.append((e0, e1))
        self.count += 1

    def total_ms(self):
        """Synchronises; returns (milliseconds, number of all-reduces) since the last call."""
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.pairs)
        n = self.count
        self.pairs, self.count = [], 0
        return ms, n


def allreduce_sum_(t, dist=None, timer=None):
    if active(dist):
        e = timer.begin() if timer is not None and t.is_cuda else None
        dist.all_reduce(t)
        if e is not None:
            timer.end(e)
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
