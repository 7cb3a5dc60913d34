"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic the trainers run under NCCL: minibatch-count
agreement, advantage statistics from partial sums, "all-reduce the gradient vectors before the dot product" for the
meta-gradient (checked with the torch-CPU oracle), scene sharding offsets, and the reference arm's rank gating."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from copo_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import models as om
    out = {}
    # 1. every rank runs the same number of minibatches
    rows = [1000, 3000][rank]
    out["k"] = parallel.num_minibatches(rows, 512, dist)
    out["bounds_ok"] = all(hi > lo for lo, hi in parallel.minibatch_bounds(rows, out["k"]))
    # 1b. the global minibatch plan: ranks with unequal row counts; the gradient of the whole-minibatch mean loss is the
    # all-reduced SUM of per-rank gradients when every rank normalises by the GLOBAL minibatch size (and the learner
    # statistic / KL that drives update_kl is the same on every rank)
    counts = parallel.gather_row_counts([30, 70][rank], dist)
    bounds, sizes = parallel.minibatch_plan(counts, rank, 32)
    out["counts"], out["sizes"], out["bounds"] = counts, sizes, bounds
    torch.manual_seed(0)
    net = om.CoPOModel(10, hiddens=(8, 8)).double()
    gen = torch.Generator().manual_seed(5)
    R = 100
    full = dict(obs=torch.rand(R, 10, generator=gen, dtype=torch.float64),
                actions=torch.randn(R, 2, generator=gen, dtype=torch.float64),
                action_logp=-2 + 0.1 * torch.randn(R, generator=gen, dtype=torch.float64),
                action_dist_inputs=0.1 * torch.randn(R, 4, generator=gen, dtype=torch.float64))
    for kcol in ("advantages", "normalized_advantages", "vf_preds", "value_targets", "nei_values", "nei_target",
                 "global_values", "global_target", "nei_advantage", "global_advantages"):
        full[kcol] = torch.randn(R, generator=gen, dtype=torch.float64)
    full["centralized_critic_obs"] = full["obs"]
    mine = {k: (v[:30] if rank == 0 else v[30:]) for k, v in full.items()}
    grads, kls = [], []
    for j, ((lo, hi), rows) in enumerate(zip(bounds, sizes)):
        net.zero_grad()
        kl = torch.zeros(1, dtype=torch.float64)
        if hi > lo:
            part = {k: v[lo:hi] for k, v in mine.items()}
            loss, st = om.ppo_loss(net, part, dict(om.DEFAULT_CFG), "copo")
            (loss * (hi - lo) / rows).backward()                  # this rank's share of the global-minibatch mean
            kl = (st["mean_kl_loss"].detach() * (hi - lo) / rows).reshape(1)
        g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])
        parallel.allreduce_sum_(g, dist)
        parallel.allreduce_sum_(kl, dist)
        grads.append(g)
        kls.append(float(kl))
    # single-process reference over the same global minibatches (rank 0 rows then rank 1 rows of each slice)
    all_bounds = parallel.minibatch_plan_all(counts, 32)[0]
    err = 0.0
    for j in range(len(sizes)):
        idx = list(range(all_bounds[0][j][0], all_bounds[0][j][1])) + [30 + r for r in range(all_bounds[1][j][0],
                                                                                               all_bounds[1][j][1])]
        net.zero_grad()
        part = {k: v[idx] for k, v in full.items()}
        loss, st = om.ppo_loss(net, part, dict(om.DEFAULT_CFG), "copo")
        loss.backward()
        g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])
        err = max(err, float((g - grads[j]).abs().max()) / (float(g.abs().max()) + 1e-30),
                  abs(float(st["mean_kl_loss"].detach()) - kls[j]))
    out["dp_grad_err"] = err
    # 2. whole-batch statistics from per-rank partial sums (algo_copo.py:547-551 needs the global mean / std)
    rng = np.random.default_rng(0)
    x = rng.normal(1.5, 2.0, 4000).astype(np.float32)
    mine = x[:1000] if rank == 0 else x[1000:]
    st = torch.tensor([mine.astype(np.float64).sum(), (mine.astype(np.float64) ** 2).sum(), float(mine.size)],
                      dtype=torch.float64)
    parallel.allreduce_sum_(st, dist)
    out["mean_std"] = parallel.mean_std_from_sums(*st.tolist())
    out["want_mean_std"] = (float(x.mean()), float(x.std()))
    # 3. the meta-gradient is bilinear in (g_new, g_old): reduce the vectors first, then take the dot
    torch.manual_seed(0)
    m, tgt = om.CoPOModel(12, hiddens=(16, 16)).double(), om.CoPOModel(12, hiddens=(16, 16)).double()
    g = torch.Generator().manual_seed(1)
    B = 64
    batch = dict(obs=torch.rand(B, 12, generator=g, dtype=torch.float64),
                 actions=torch.randn(B, 2, generator=g, dtype=torch.float64),
                 action_logp=-2 + 0.1 * torch.randn(B, generator=g, dtype=torch.float64),
                 global_advantages=torch.randn(B, generator=g, dtype=torch.float64),
                 advantages=torch.randn(B, generator=g, dtype=torch.float64),
                 nei_advantage=torch.randn(B, generator=g, dtype=torch.float64))
    eps = torch.randn(B, generator=g, dtype=torch.float64)
    full = om.meta_gradient(m, tgt, batch, om.DEFAULT_CFG, 0.0, 1.0, eps)
    half = {k: v[rank * 32:(rank + 1) * 32] for k, v in batch.items()}
    _, _, _, (g_new, g_old) = om.meta_gradient(m, tgt, half, om.DEFAULT_CFG, 0.0, 1.0, eps[rank * 32:(rank + 1) * 32])
    fn = torch.cat([t.reshape(-1) for t in g_new])
    fo = torch.cat([t.reshape(-1) for t in g_old])
    wrong = float((fn * fo).sum())                            # per-rank dot, to be averaged: NOT the reference value
    parallel.allreduce_mean_(fn, dist)
    parallel.allreduce_mean_(fo, dist)
    out["grad_value"] = float((fn * fo).sum())
    out["want_grad_value"] = full[2]["grad_value"]
    w = torch.tensor([wrong], dtype=torch.float64)
    parallel.allreduce_mean_(w, dist)
    out["mean_of_dots"] = float(w)
    out["offset"] = parallel.scene_offset(rank, 4096)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["k"] == res[1]["k"] == 6 and res[0]["bounds_ok"] and res[1]["bounds_ok"]
    assert res[0]["counts"] == res[1]["counts"] == [30, 70] and res[0]["sizes"] == res[1]["sizes"] == [32, 30, 16, 16, 6]
    # a fixed quota of 32 // 2 rows per rank and minibatch; the rank with fewer rows runs out and contributes empty slices
    assert res[0]["bounds"] == [(0, 16), (16, 30), (30, 30), (30, 30), (30, 30)]
    assert res[1]["bounds"] == [(0, 16), (16, 32), (32, 48), (48, 64), (64, 70)]
    assert res[0]["dp_grad_err"] < 1e-12 and res[1]["dp_grad_err"] < 1e-12
    for r in (0, 1):
        assert np.allclose(res[r]["mean_std"], res[r]["want_mean_std"], rtol=1e-6)
        assert abs(res[r]["grad_value"] - res[r]["want_grad_value"]) < 1e-12 + 1e-9 * abs(res[r]["want_grad_value"])
    assert abs(res[0]["mean_of_dots"] - res[0]["want_grad_value"]) > 1e-6 * abs(res[0]["want_grad_value"])
    assert (res[0]["offset"], res[1]["offset"]) == (0, 4096)


def test_single_process_helpers():
    assert not parallel.active(None)
    assert parallel.num_minibatches(0, 512) == 1 and parallel.num_minibatches(1025, 512) == 3
    b = parallel.minibatch_bounds(10, 3)
    assert b == [(0, 4), (4, 7), (7, 10)]
    assert parallel.minibatch_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]      # empty slices only when rows < k
    # one rank: rllib's slicing (full minibatches, ragged tail); two ranks: a fixed quota of minibatch_size // 2 rows each
    assert parallel.minibatch_plan([10], 0, 4) == ([(0, 4), (4, 8), (8, 10)], [4, 4, 2])
    assert parallel.minibatch_plan([5, 9], 1, 8) == ([(0, 4), (4, 8), (8, 9)], [8, 5, 1])
    assert parallel.minibatch_plan([5, 9], 0, 8) == ([(0, 4), (4, 5), (5, 5)], [8, 5, 1])
    assert parallel.minibatch_plan([0, 5], 0, 512) == ([(0, 0)], [5])
    assert parallel.gather_row_counts(7) == [7]
    t = torch.ones(3)
    assert parallel.allreduce_mean_(t) is t


def test_reference_arm_rank_gating():
    """Under torchrun only rank 0 runs and prints the reference arm; other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
