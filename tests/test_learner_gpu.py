"""Parity of the learner-side CUDA kernels (through the C ABI) with the CPU oracle: MLP forward against the reference's
golden vectors, linear fwd/bwd, Gaussian sampling, GAE x3, LCF mix + standardize, critic-obs fusion, PPO / CoPO loss
and gradients, meta-gradient + LCF Adam step.  Tolerances: 1e-4 relative on returns / advantages (north star), 1e-5 on
forward values, 2e-4 on gradients (fp32 summation order differs from torch-CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import gae as og
from oracle import models as om

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mlp_golden.npz")


def close(a, b, rtol, atol):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), "max violation %g at %s (got %r want %r)" % (
        err.max(), np.unravel_index(err.argmax(), err.shape), a.flat[err.argmax()], b.flat[err.argmax()])


@pytest.mark.parametrize("precision", ["fp32", "bf16_split"])
@pytest.mark.parametrize("name", ["copo_inter", "ccppo_round", "ippo_tollgate", "cl_bottle", "copo_tollgate",
                                  "ccppo_parking"])
def test_policy_forward_golden(name, precision):
    from copo_b200.models import CCModel
    z = np.load(GOLD)
    sub = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    obs, want = sub.pop("obs"), sub.pop("mean")
    m = CCModel(obs.shape[1], precision=precision)
    m.load_policy_npz(sub)
    logits = m.forward(torch.from_numpy(obs).cuda())
    if precision == "fp32":
        close(logits[:, :2], want, 1e-5, 2e-6)
    else:
        close(logits[:, :2], want, 2e-4, 5e-5)          # tensor-core path: operands carry 16 mantissa bits (hi + lo)
    # checkpoint round trip through RLlib names
    m2 = CCModel(obs.shape[1], seed=5, precision=precision)
    m2.load_state_dict({k: v.cpu().numpy() for k, v in m.state_dict().items()})
    assert torch.equal(m2.forward(torch.from_numpy(obs).cuda()), logits)
    with pytest.raises(ValueError):
        m.value_function()


@pytest.mark.parametrize("M,K,N", [(777, 92, 256), (300, 256, 256), (1000, 157, 256), (129, 184, 256), (513, 256, 4),
                                   (64, 256, 1), (5000, 463, 256), (1, 92, 256)])
def test_linear_forward_backward(M, K, N):
    from copo_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.tanh(torch.randn(M, K, generator=g))
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    act = 1 if N > 8 else 0
    xr, Wr, br = x.double().requires_grad_(), W.double().requires_grad_(), b.double().requires_grad_()
    y = xr @ Wr.T + br
    if act:
        y = torch.tanh(y)
    y.backward(dy.double())
    got = ops.linear_forward(x.cuda(), W.cuda(), b.cuda(), act)
    close(got, y, 1e-5, 1e-5)
    if act:   # backward is defined on the pre-activation gradient: feed dz = dy * (1 - y^2)
        dz = (dy.double() * (1 - y.detach() ** 2)).float()
    else:
        dz = dy
    dW, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    dx = ops.linear_backward(dz.cuda(), x.cuda(), W.cuda(), dW, db, h_prev_is_tanh=True)
    close(dW, Wr.grad, 2e-4, 2e-4 * M ** 0.5 / 10)
    close(db, br.grad, 2e-4, 2e-4 * M ** 0.5 / 10)
    close(dx, xr.grad * (1 - x.double() ** 2), 1e-4, 1e-5)
    dx2 = ops.linear_backward(dz.cuda(), x.cuda(), W.cuda(), torch.zeros_like(dW), None, h_prev_is_tanh=False)
    close(dx2, xr.grad, 1e-4, 1e-5)


def test_gaussian_sample_and_logp():
    from copo_b200 import ops
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1000, 4, generator=g) * 0.5
    eps = torch.randn(1000, 2, generator=g)
    a, lp = ops.gaussian_sample(logits.cuda(), eps.cuda())
    d = om.DiagGaussian(logits)
    want_a = d.sample(eps)
    close(a, want_a, 1e-6, 1e-6)
    close(lp, d.logp(want_a), 1e-5, 1e-5)
    a_det, _ = ops.gaussian_sample(logits.cuda(), deterministic=True)
    close(a_det, logits[:, :2], 0, 0)
    # in-kernel generator: standard normal moments, reproducible per (seed, step)
    big = torch.zeros(200000, 4, device="cuda")
    a1, _, e1 = ops.gaussian_sample(big, seed=3, step=7, want_eps=True)
    a2, _ = ops.gaussian_sample(big, seed=3, step=7)
    a3, _ = ops.gaussian_sample(big, seed=3, step=8)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3) and torch.equal(a1, e1)
    assert abs(float(a1.mean())) < 0.01 and abs(float(a1.std()) - 1) < 0.01
    assert abs(float((a1[:, 0] * a1[:, 1]).mean())) < 0.01


def _random_flags(T, N, rng):
    flags = np.zeros((T, N), np.uint8)
    for n in range(N):
        t = rng.integers(0, 5)
        while t < T:
            length = rng.integers(1, 40)
            end = min(T, t + length)
            flags[t:end, n] = 1
            if end < T or rng.random() < 0.3:
                flags[end - 1, n] |= 2                      # done closes the trajectory
            else:
                pass                                         # cut by the fragment end
            t = end + rng.integers(0, 30)
    flags[:, 0] = 1                                         # one column alive throughout, never done
    flags[:, 1] = 0                                         # one empty column
    return flags


def test_gae3_matches_reference_restatement():
    from copo_b200 import ops
    rng = np.random.default_rng(0)
    T, S, A = 120, 6, 10
    N = S * A
    flags = _random_flags(T, N, rng)
    f = lambda *s: rng.normal(size=s).astype(np.float32)
    r, v, nr, nv, gv = f(T, N), f(T, N) * 3, f(T, N), f(T, N) * 3, f(T, N) * 3
    gr_scene = f(T, S)
    gr = np.repeat(gr_scene, A, axis=1)
    want = og.rollout_gae3(flags, r, v, nr, nv, gr, gv, 0.99, 0.95)
    c = lambda a: torch.from_numpy(a).cuda()
    adv, tgt = ops.gae3(c(flags), [c(r), c(nr), c(gr_scene)], [c(v), c(nv), c(gv)], 0.99, 0.95,
                        global_reward_per_scene=A)
    for k, (a, t) in enumerate((("advantages", "value_targets"), ("nei_advantage", "nei_target"),
                                ("global_advantages", "global_target"))):
        close(adv[k], want[a], 1e-4, 1e-5)
        close(tgt[k], want[t], 1e-4, 1e-5)
    adv1, _ = ops.gae3(c(flags), [c(r)], [c(v)], 0.99, 0.95)
    close(adv1[0], want["advantages"], 1e-4, 1e-5)
    # IPPO: trajectories the fragment end cuts bootstrap with the value of the next observation (stock rllib)
    nxt = f(N) * 3
    want_ippo = og.rollout_gae3(flags, r, v, nr, nv, gr, gv, 0.99, 0.95, heads=1, next_values=nxt)
    adv_i, tgt_i = ops.gae3(c(flags), [c(r)], [c(v)], 0.99, 0.95, bootstrap=[c(nxt)])
    close(adv_i[0], want_ippo["advantages"], 1e-4, 1e-5)
    close(tgt_i[0], want_ippo["value_targets"], 1e-4, 1e-5)
    assert not torch.equal(adv_i[0], adv1[0])
    # full-size property: linearity in the rewards (GAE is linear for fixed values / flags)
    T2, N2 = 50, 40 * 512
    fl = c(_random_flags(T2, 64, rng)).repeat(1, N2 // 64).contiguous()
    r1, r2, vv = (torch.randn(T2, N2, device="cuda") for _ in range(3))
    a1, _ = ops.gae3(fl, [r1], [vv], 0.99, 0.95)
    a2, _ = ops.gae3(fl, [r2], [vv], 0.99, 0.95)
    a12, _ = ops.gae3(fl, [r1 + r2], [vv * 2], 0.99, 0.95)
    close(a12[0], a1[0] + a2[0], 1e-4, 1e-4)


def test_lcf_mix_standardize():
    from copo_b200 import ops
    rng = np.random.default_rng(1)
    R = 50000
    flags = (rng.random(R) < 0.8).astype(np.uint8)
    adv, nei, gadv = (rng.normal(size=R).astype(np.float32) * s for s in (2.0, 1.0, 3.0))
    lcf = np.clip(rng.normal(0.2, 0.3, R), -1, 1).astype(np.float32)
    v = flags > 0
    norm, mean, std, g = og.lcf_mix_standardize(adv[v], nei[v], lcf[v], gadv[v])
    c = lambda a: torch.from_numpy(a).cuda()
    st = ops.lcf_mix_stats(c(flags), c(adv), c(nei), c(lcf), c(gadv)).tolist()
    m, s = ops.stats_to_mean_std(st[0], st[1], st[2])
    gm, gs = ops.stats_to_mean_std(st[3], st[4], st[2])
    assert st[2] == v.sum() and abs(m - mean) < 1e-5 and abs(s - std) < 1e-5
    gt = c(gadv)
    got = ops.lcf_mix_apply(c(flags), c(adv), c(nei), c(lcf), gt, m, s, gm, gs)
    close(got[c(v)], norm, 1e-4, 1e-5)
    close(gt[c(v)], g, 1e-4, 1e-5)
    assert float(got[~c(v)].abs().max()) == 0.0


def _cc_reference(obs, act, flags, nei_sorted, nei_dist, A, mode, mf_dist=10.0, num_neighbours=4):
    """Row-wise restatement of concat_ccppo_process / mean_field_ccppo_process (algo_ccppo.py:225-311) on arrays:
    a neighbour contributes when it has a row at the same t (valid flag)."""
    R, D = obs.shape
    W = D + 2
    C = D + (1 if mode == "mf" else 4) * W
    out = np.zeros((R, C), np.float32)
    for r in range(R):
        if not flags[r] & 1:
            continue
        out[r, :D] = obs[r]
        base = r - r % A
        if mode == "mf":
            ol, al = [], []
            for j, d in zip(nei_sorted[r], nei_dist[r]):
                if d > mf_dist:
                    continue
                if flags[base + j] & 1:
                    ol.append(obs[base + j]); al.append(act[base + j])
            if ol:
                out[r, D:2 * D] = np.mean(ol, axis=0)
                out[r, 2 * D:2 * D + 2] = np.mean(al, axis=0)
        else:
            for cnt, j in enumerate(nei_sorted[r]):
                if cnt >= num_neighbours:
                    break
                if flags[base + j] & 1:
                    s = D + cnt * W
                    out[r, s:s + D] = obs[base + j]
                    out[r, s + D:s + W] = act[base + j]
    return out


@pytest.mark.parametrize("mode", ["mf", "concat"])
def test_cc_obs_fuse(mode, monkeypatch):
    from copo_b200 import ops
    rng = np.random.default_rng(2)
    T, S, A, D = 5, 7, 12, 23
    R = T * S * A
    obs, act = rng.random((R, D)).astype(np.float32), rng.normal(size=(R, 2)).astype(np.float32)
    flags = (rng.random(R) < 0.8).astype(np.uint8)
    pos = rng.random((R, 2)) * 30
    nei_sorted, nei_dist = [], []
    mf_mask = np.zeros(R, np.uint64)
    nei_list = np.full((R, 4), -1, np.int8)
    for r in range(R):
        base = r - r % A
        d = np.linalg.norm(pos[base:base + A] - pos[r], axis=1)
        order = [j for j in np.argsort(d, kind="stable") if j != r % A and d[j] < 20]
        nei_sorted.append(order); nei_dist.append([d[j] for j in order])
        for j in order:
            if d[j] <= 10.0:
                mf_mask[r] |= np.uint64(1) << np.uint64(j)
        nei_list[r, :min(4, len(order))] = order[:4]
    want = _cc_reference(obs, act, flags, nei_sorted, nei_dist, A, mode)
    c = lambda a: torch.from_numpy(a).cuda()
    got = ops.cc_obs_fuse(c(obs), c(act), c(flags), c(mf_mask.view(np.int64)), c(nei_list), A, mode, True)
    assert got.shape[1] == om.centralized_critic_obs_dim(D, 2, True, 4, mode)
    close(got, want, 1e-5, 1e-6)
    if mode == "mf":
        # the mean-field form has two kernels: one CTA per (t, scene) with the scene staged in shared memory (whenever the
        # rows are whole scenes - the call above), and one warp per row (B2C_FUSE_ROWWISE, or a ragged row count)
        monkeypatch.setenv("B2C_FUSE_ROWWISE", "1")
        rowwise = ops.cc_obs_fuse(c(obs), c(act), c(flags), c(mf_mask.view(np.int64)), c(nei_list), A, mode, True)
        monkeypatch.delenv("B2C_FUSE_ROWWISE")
        assert torch.equal(rowwise, got)                             # same sums in the same order
        ragged = ops.cc_obs_fuse(c(obs[:-A // 2]), c(act[:-A // 2]), c(flags[:-A // 2]), c(mf_mask.view(np.int64)[:-A // 2]),
                                 None, A, mode, True)
        assert torch.equal(ragged[:R - A], got[:R - A])


def test_mean_field_fusion_at_the_c3_shape():
    """40 slots x 91 observation columns + 2 action columns (three column slots per lane), every scene of a step: the
    scene-wise kernel against the row-wise one, and the means against numpy on sampled rows."""
    from copo_b200 import ops
    S, A, D = 512, 40, 91
    R = S * A
    g = torch.Generator(device="cuda").manual_seed(4)
    obs, act = torch.rand(R, D, device="cuda", generator=g), torch.randn(R, 2, device="cuda", generator=g)
    flags = (torch.rand(R, device="cuda", generator=g) < 0.85).to(torch.uint8)
    mf = torch.randint(0, 2 ** 40, (R,), device="cuda", dtype=torch.int64, generator=g) & \
        torch.randint(0, 2 ** 40, (R,), device="cuda", dtype=torch.int64, generator=g) & \
        torch.randint(0, 2 ** 40, (R,), device="cuda", dtype=torch.int64, generator=g)
    mf &= ~(torch.ones(R, dtype=torch.int64, device="cuda") << (torch.arange(R, device="cuda") % A))      # not oneself
    got = ops.cc_obs_fuse(obs, act, flags, mf, None, A, "mf", True)
    os.environ["B2C_FUSE_ROWWISE"] = "1"
    try:
        rowwise = ops.cc_obs_fuse(obs, act, flags, mf, None, A, "mf", True)
    finally:
        del os.environ["B2C_FUSE_ROWWISE"]
    assert got.shape == (R, 2 * D + 2) and torch.equal(got, rowwise)
    # the operand output: the bits tc_split_rows makes of the fused rows (zero padding included)
    got2, split = ops.cc_obs_fuse(obs, act, flags, mf, None, A, "mf", True, want_split=True)
    assert torch.equal(got2, got) and split.shape == (R, 2 * ops.tc_padded_k(2 * D + 2))
    assert torch.equal(split.view(torch.int16), ops.tc_split_rows(got).view(torch.int16))
    o, a, f, m, out = obs.cpu().numpy(), act.cpu().numpy(), flags.cpu().numpy(), mf.cpu().numpy(), got.cpu().numpy()
    for r in (0, 1, 39, 40, 777, R - 41, R - 1):
        base = r - r % A
        nb = [base + j for j in range(A) if (int(m[r]) >> j) & 1 and f[base + j] & 1]
        want = np.zeros(2 * D + 2, np.float32)
        if f[r] & 1:
            want[:D] = o[r]
            if nb:
                want[D:2 * D] = o[nb].astype(np.float64).mean(0)
                want[2 * D:] = a[nb].astype(np.float64).mean(0)
        assert np.allclose(out[r], want, rtol=1e-5, atol=1e-6), r


def _copo_batch(B, D, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(obs=torch.rand(B, D, generator=g), centralized_critic_obs=torch.rand(B, D, generator=g),
                actions=0.5 * r(B, 2), action_logp=-1.5 + 0.3 * r(B), action_dist_inputs=0.3 * r(B, 4), advantages=r(B),
                normalized_advantages=r(B), vf_preds=r(B), value_targets=r(B), nei_values=r(B), nei_target=r(B),
                global_values=r(B), global_target=r(B), nei_advantage=r(B), global_advantages=r(B))


def _oracle_model_from(model, cls):
    ref = cls(model.obs_dim, cdim=model.cobs_dim)
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref.load_state_dict(sd)
    return ref


def _flat_grad(model, ref):
    """Oracle gradients laid out like the product's flat buffer."""
    named = dict(ref.named_parameters())
    parts = []
    for net in model.nets.values():
        for name in net.names:
            for suffix in (".weight", ".bias"):
                gr = named[name + suffix].grad
                parts.append((gr if gr is not None else torch.zeros_like(named[name + suffix])).reshape(-1))
    return torch.cat(parts)


@pytest.mark.parametrize("algo,B,precision", [("copo", 2048, "fp32"), ("ccppo", 700, "fp32"), ("ippo", 512, "fp32"),
                                              ("copo", 3000, "bf16_split"), ("ippo", 129, "bf16_split")])
def test_loss_and_gradients_match_oracle(algo, B, precision):
    from copo_b200 import policy as P
    D = 92
    cls = {"copo": P.CoPOPolicy, "ccppo": P.CCPPOPolicy, "ippo": P.IPPOPolicy}[algo]
    cfg = cls.default_config()
    cfg["fuse_mode"] = "none"
    cfg["precision"] = precision
    cfg["vf_clip_param"] = 0.8          # small enough that the clipped branch of the value loss is exercised
    pol = cls(D, 2, cfg)
    pol.kl_coeff = 0.3
    batch = _copo_batch(B, D, 1)
    if algo == "ippo":
        batch["centralized_critic_obs"] = batch["obs"]
    ref = _oracle_model_from(pol.model, om.CoPOModel if algo == "copo" else om.CCModel)
    ocfg = dict(om.DEFAULT_CFG, vf_clip_param=0.8, kl_coeff=0.3)
    total, st = om.ppo_loss(ref, batch, ocfg, algo)
    total.backward()
    gb = {k: v.cuda() for k, v in batch.items()}
    pol.model.zero_grad()
    got = pol.loss(pol.model, None, gb)
    k = 1.0 if precision == "fp32" else 10.0
    close(got, total, 2e-5 * k, 1e-6 * k)
    close(pol.model.tower_stats["mean_kl_loss"], st["mean_kl_loss"], 1e-4 * k, 1e-6 * k)
    close(pol.model.tower_stats["mean_entropy"], st["mean_entropy"], 1e-5 * k, 1e-6 * k)
    close(pol.model.tower_stats["mean_vf_loss"], st["mean_vf_loss"], 1e-5 * k, 1e-6 * k)
    want_g = _flat_grad(pol.model, ref)
    scale = float(want_g.abs().max())
    if precision == "fp32":
        close(pol.model.grad, want_g, 2e-4, 2e-5 * scale)
    else:
        close(pol.model.grad, want_g, 1e-3, 2e-4 * scale)
    # one optimiser step = torch.optim.Adam(lr)
    opt = torch.optim.Adam(ref.parameters(), lr=cfg["lr"])
    opt.step()
    pol._optimizer.apply(pol.model.grad)
    if precision == "fp32":      # Adam's first step is lr * sign(g): tiny gradient differences flip nothing in fp32
        close(pol.model.state_dict()["_hidden_layers.0._model.0.weight"],
              ref.state_dict()["_hidden_layers.0._model.0.weight"], 1e-5, 2e-6)


@pytest.mark.parametrize("algo", ["copo", "ippo"])
def test_plain_value_loss_branch(algo):
    """old_value_loss=False: vf_loss = clamp((v - target)^2, 0, vf_clip_param) (algo_ippo.py:146-148,
    algo_copo.py:358-363) - loss and gradients against the torch-CPU oracle."""
    from copo_b200 import policy as P
    D, B = 92, 900
    cls = {"copo": P.CoPOPolicy, "ippo": P.IPPOPolicy}[algo]
    cfg = cls.default_config()
    cfg.update(fuse_mode="none", precision="fp32", vf_clip_param=1.5, old_value_loss=False)
    pol = cls(D, 2, cfg)
    batch = _copo_batch(B, D, 4)
    batch["centralized_critic_obs"] = batch["obs"]
    ref = _oracle_model_from(pol.model, om.CoPOModel if algo == "copo" else om.CCModel)
    ocfg = dict(om.DEFAULT_CFG, vf_clip_param=1.5, old_value_loss=False)
    total, st = om.ppo_loss(ref, batch, ocfg, algo)
    total.backward()
    pol.model.zero_grad()
    got = pol.loss(pol.model, None, {k: v.cuda() for k, v in batch.items()})
    close(got, total, 2e-5, 1e-6)
    close(pol.model.tower_stats["mean_vf_loss"], st["mean_vf_loss"], 1e-5, 1e-6)
    assert float(st["mean_vf_loss"]) < 1.5 - 1e-3          # some rows are clamped, some are not
    want_g = _flat_grad(pol.model, ref)
    close(pol.model.grad, want_g, 2e-4, 2e-5 * float(want_g.abs().max()))


@pytest.mark.parametrize("precision", ["fp32", "bf16_split"])
def test_global_rows_normalisation_sums_to_the_whole_batch_gradient(precision):
    """Data-parallel rule: each rank normalises by the GLOBAL minibatch row count; the sum of the ranks' gradients,
    losses and KL is then the whole-minibatch mean even when the ranks hold different numbers of rows (here: one
    process plays both ranks, 700 + 1348 rows, plus an empty share)."""
    from copo_b200 import policy as P
    D, B = 92, 2048
    cfg = P.copo_config(precision=precision)
    pol = P.CoPOPolicy(D, 2, cfg)
    pol.kl_coeff = 0.3
    batch = {k: v.cuda() for k, v in _copo_batch(B, D, 6).items()}
    pol.model.zero_grad()
    whole = pol.loss(pol.model, None, batch).clone()
    g_whole, v_whole = pol.model.grad.clone(), pol.stats_vector.clone()
    g_sum, l_sum, v_sum = torch.zeros_like(g_whole), 0.0, torch.zeros_like(v_whole)
    for lo, hi in ((0, 700), (700, 2048), (2048, 2048)):
        pol.model.zero_grad()
        part = {k: v[lo:hi].contiguous() for k, v in batch.items()}
        l_sum = l_sum + pol.loss(pol.model, None, part, global_rows=B)
        g_sum += pol.model.grad
        v_sum += pol.stats_vector
    k = 1.0 if precision == "fp32" else 20.0
    close(l_sum, whole, 1e-6 * k, 1e-7 * k)
    close(v_sum, v_whole, 1e-6 * k, 1e-7 * k)
    close(g_sum, g_whole, 1e-4 * k, 1e-6 * k * float(g_whole.abs().max()))


@pytest.mark.parametrize("precision", ["fp32", "bf16_split"])
def test_meta_update_matches_oracle(precision):
    from copo_b200 import policy as P
    D, B = 92, 1500
    cfg = P.copo_config(precision=precision)
    pol = P.CoPOPolicy(D, 2, cfg)
    # make theta_new differ from theta_old (as after the SGD epochs)
    pol.model.flat.add_(0.01 * torch.randn_like(pol.model.flat))
    pol.model.lcf_parameters.copy_(torch.tensor([0.3, -1.7]))
    pol._raw_lcf_adv_mean, pol._raw_lcf_adv_std = 0.05, 1.3
    batch = _copo_batch(B, D, 2)
    eps = torch.randn(B, generator=torch.Generator().manual_seed(9))
    ref, tgt = _oracle_model_from(pol.model, om.CoPOModel), _oracle_model_from(pol.target_model, om.CoPOModel)
    final, g_lcf, st, _ = om.meta_gradient(ref, tgt, batch, om.DEFAULT_CFG, 0.05, 1.3, eps)
    opt = torch.optim.Adam([ref.lcf_parameters], lr=cfg["lcf_lr"])
    ref.lcf_parameters.grad = g_lcf
    opt.step()
    out = pol.meta_update({k: v.cuda() for k, v in batch.items()}, eps=eps.cuda())
    assert abs(out["grad_value"] - st["grad_value"]) <= 2e-4 * abs(st["grad_value"]) + 1e-9
    assert abs(out["lcf_final_loss"] - st["lcf_final_loss"]) <= 3e-4 * abs(st["lcf_final_loss"]) + 1e-9
    assert abs(out["new_policy_ego_loss"] - st["new_policy_ego_loss"]) < 1e-5
    assert abs(out["old_policy_logp_loss"] - st["old_policy_logp_loss"]) < 1e-4
    close(pol.model.lcf_grad, g_lcf, 5e-4, 1e-9)
    close(pol.model.lcf_parameters, ref.lcf_parameters, 1e-6, 1e-7)
    # theta_old <- theta_new, LCF hand-over (algo_copo.py:442-471)
    pol.update_old_policy()
    assert torch.equal(pol.target_model.flat, pol.model.flat)
    pol.assign_lcf(pol.model.lcf_parameters.clone(), float(pol.model.lcf_mean), float(pol.model.lcf_std))


def test_meta_update_outside_the_lcf_clamps():
    """lcf_mean = clamp(tanh(p0)) and lcf_std = exp(clamp(p1, -20, 2)) (algo_copo.py:171-177) have zero derivative outside
    their clamps: with p = (10, 3) the meta step's LCF gradient is exactly zero, Adam leaves the parameters where they
    are, and the logged LCF is the clamped one."""
    import math
    from copo_b200 import policy as P
    D, B = 92, 600
    pol = P.CoPOPolicy(D, 2, P.copo_config())
    pol.model.flat.add_(0.01 * torch.randn_like(pol.model.flat))
    pol.model.mark_weights_changed()
    pol.model.lcf_parameters.copy_(torch.tensor([10.0, 3.0]))
    pol._raw_lcf_adv_mean, pol._raw_lcf_adv_std = 0.0, 1.0
    batch = {k: v.cuda() for k, v in _copo_batch(B, D, 6).items()}
    out = pol.meta_update(batch, eps=torch.randn(B, generator=torch.Generator().manual_seed(1)).cuda())
    assert torch.equal(pol.model.lcf_grad.cpu(), torch.zeros(2))
    assert torch.equal(pol.model.lcf_parameters.cpu(), torch.tensor([10.0, 3.0]))
    assert abs(out["lcf"] - (1 - 1e-6)) < 1e-7 and abs(out["lcf_std"] - math.exp(2.0)) < 1e-5
    assert out["lcf_param"] == 10.0 and out["lcf_std_param"] == 3.0 and abs(out["lcf_deg"] - 90.0) < 1e-3
    assert math.isfinite(out["grad_value"]) and out["grad_value"] != 0.0
    assert abs(out["global_adv"] - float(batch["global_advantages"].double().mean())) < 1e-6


def test_gather_cols():
    from copo_b200 import ops
    g = torch.Generator().manual_seed(0)
    src = torch.randn(5000, 11, generator=g).cuda()
    idx = torch.randperm(5000, generator=g)[:1777].cuda()
    out = ops.gather_cols(src, idx)
    assert out.shape == (11, 1777) and out.is_contiguous() and torch.equal(out, src[idx].t())
    wide = torch.randn(5000, 16, generator=g).cuda()
    assert torch.equal(ops.gather_cols(wide[:, :11], idx), wide[idx][:, :11].t())       # strided source rows


def test_adam_and_dot():
    from copo_b200 import ops
    p = torch.nn.Parameter(torch.randn(10001))
    opt = torch.optim.Adam([p], lr=3e-4)
    q, m, v = p.detach().clone().cuda(), torch.zeros(10001, device="cuda"), torch.zeros(10001, device="cuda")
    for step in range(1, 5):
        g = torch.randn(10001)
        p.grad = g.clone()
        opt.step()
        ops.adam_step(q, g.cuda(), m, v, 3e-4, step)
    close(q, p, 1e-6, 1e-7)
    a, b = torch.randn(100000, device="cuda"), torch.randn(100000, device="cuda")
    assert abs(float(ops.dot(a, b)) - float((a.double() * b.double()).sum())) < 1e-6


def test_empty_inputs_are_accepted():
    """Zero-row calls are legal everywhere (a rank can hold no valid rows in a minibatch slice)."""
    from copo_b200 import ops
    e = lambda *shape: torch.empty(shape, device="cuda")
    W, b = torch.randn(256, 92, device="cuda"), torch.zeros(256, device="cuda")
    assert ops.linear_forward(e(0, 92), W, b, 1).shape == (0, 256)
    assert ops.tc_split_rows(e(0, 92)).shape == (0, 256)
    y, _ = ops.tc_linear(ops.tc_split_rows(e(0, 92)), ops.tc_prep_weight(W), b, act=1)
    assert y.shape == (0, 256)
    a, lp = ops.gaussian_sample(e(0, 4))
    assert a.shape == (0, 2) and lp.shape == (0,)
    dW = torch.zeros(256, 92, device="cuda")
    ops.linear_backward(e(0, 256), e(0, 92), W, dW, torch.zeros(256, device="cuda"), False, need_dx=False)
    assert float(dW.abs().max()) == 0.0
    assert ops.gather_rows(e(10, 7), torch.empty(0, dtype=torch.int64, device="cuda")).shape == (0, 7)
    st = ops.lcf_mix_stats(torch.empty(0, dtype=torch.uint8, device="cuda"), e(0), e(0), e(0), e(0))
    assert st.tolist() == [0.0] * 5
