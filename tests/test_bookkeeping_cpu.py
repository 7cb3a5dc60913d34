"""Pins the CPU restatements in oracle/ (GAE x3, LCF mix + standardize, networks, action distribution, losses):
against scipy's lfilter / closed forms, against the reference's own numpy forward on the shipped weights
(tests/golden/mlp_golden.npz), and against torch.distributions for the Gaussian algebra."""
import os

import numpy as np
import pytest
import torch

from oracle import gae as og
from oracle import models as om

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mlp_golden.npz")


def test_discount_cumsum_closed_form():
    rng = np.random.default_rng(0)
    x = rng.normal(size=37)
    g = 0.99 * 0.95
    want = np.array([sum(x[t + k] * g ** k for k in range(len(x) - t)) for t in range(len(x))])
    assert np.allclose(og.discount_cumsum(x, g), want, rtol=1e-12)


def test_gae_bootstrap_quirk_and_gamma_one():
    r = np.array([1.0, 2.0, 3.0], np.float32)
    v = np.array([0.5, 0.25, 4.0], np.float32)
    # not done: bootstrap with the value of the LAST row itself (algo_ccppo.py:362-365)
    adv, tgt = og.compute_advantages(r, v, v[-1], 0.99, 0.95)
    d2 = 3.0 + 0.99 * 4.0 - 4.0
    d1 = 2.0 + 0.99 * 4.0 - 0.25
    d0 = 1.0 + 0.99 * 0.25 - 0.5
    want = np.array([d0 + 0.9405 * (d1 + 0.9405 * d2), d1 + 0.9405 * d2, d2])
    assert np.allclose(adv, want, rtol=1e-6) and adv.dtype == np.float32
    assert np.allclose(tgt, want + v, rtol=1e-6)
    # done: last_r = 0
    adv, _ = og.compute_advantages(r, v, 0.0, 0.99, 0.95)
    assert np.isclose(adv[-1], 3.0 - 4.0)
    # global head: gamma = 1 (algo_copo.py:498-500)
    adv, _ = og.compute_global_advantage(r, v, 0.0, 1.0, 0.95)
    assert np.isclose(adv[-1], -1.0) and np.isclose(adv[1], (2.0 + 4.0 - 0.25) + 0.95 * -1.0)


def test_rollout_columns_are_cut_into_trajectories():
    V, D = og.FLAG_VALID, og.FLAG_VALID | og.FLAG_DONE
    col = np.array([V, V, D, 0, 0, V, V], np.uint8)
    assert og.trajectories(col) == [[0, 1, 2], [5, 6]]
    assert og.trajectories(np.array([0, 0], np.uint8)) == []
    assert og.trajectories(np.array([D, D], np.uint8)) == [[0], [1]]
    T, N = 7, 3
    rng = np.random.default_rng(1)
    flags = np.stack([col, np.full(T, V, np.uint8), np.zeros(T, np.uint8)], 1)
    arr = lambda: rng.normal(size=(T, N)).astype(np.float32)
    r, v, nr, nv, gr, gv = arr(), arr(), arr(), arr(), arr(), arr()
    out = og.rollout_gae3(flags, r, v, nr, nv, gr, gv)
    # column 0, first trajectory ends done; second is cut by the fragment end and bootstraps with its own value
    a, t = og.compute_advantages(r[[0, 1, 2], 0], v[[0, 1, 2], 0], 0.0)
    assert np.array_equal(out["advantages"][[0, 1, 2], 0], a) and np.array_equal(out["value_targets"][[0, 1, 2], 0], t)
    a, _ = og.compute_advantages(r[[5, 6], 0], v[[5, 6], 0], v[6, 0])
    assert np.array_equal(out["advantages"][[5, 6], 0], a)
    a, _ = og.compute_global_advantage(gr[[5, 6], 0], gv[[5, 6], 0], gv[6, 0], 1.0, 0.95)
    assert np.array_equal(out["global_advantages"][[5, 6], 0], a)
    assert (out["advantages"][3:5, 0] == 0).all() and (out["nei_advantage"][:, 2] == 0).all()


def test_ippo_bootstraps_with_the_next_observation():
    """Stock rllib PPO (IPPOPolicy inherits its postprocessing): a trajectory that is not done bootstraps with
    V(NEXT_OBS of its last row); done trajectories and trajectories that end before the fragment does are untouched."""
    V, D = og.FLAG_VALID, og.FLAG_VALID | og.FLAG_DONE
    T, N = 6, 4
    flags = np.stack([np.full(T, V, np.uint8),                       # alive throughout: cut by the fragment end
                      np.array([V, V, D, 0, V, V], np.uint8),        # second trajectory is cut
                      np.array([0, V, V, V, V, D], np.uint8),        # ends done on the last row: last_r = 0
                      np.zeros(T, np.uint8)], 1)
    rng = np.random.default_rng(5)
    r, v = rng.normal(size=(T, N)).astype(np.float32), rng.normal(size=(T, N)).astype(np.float32)
    nxt = rng.normal(size=N).astype(np.float32)
    z = np.zeros((T, N), np.float32)
    out = og.rollout_gae3(flags, r, v, z, z, z, z, heads=1, next_values=nxt)
    a, t = og.compute_advantages(r[:, 0], v[:, 0], nxt[0])
    assert np.array_equal(out["advantages"][:, 0], a) and np.array_equal(out["value_targets"][:, 0], t)
    a, _ = og.compute_advantages(r[[4, 5], 1], v[[4, 5], 1], nxt[1])
    assert np.array_equal(out["advantages"][[4, 5], 1], a)
    a, _ = og.compute_advantages(r[[0, 1, 2], 1], v[[0, 1, 2], 1], 0.0)
    assert np.array_equal(out["advantages"][[0, 1, 2], 1], a)
    a, _ = og.compute_advantages(r[1:, 2], v[1:, 2], 0.0)
    assert np.array_equal(out["advantages"][1:, 2], a)
    # the last delta is r + gamma * V(next) - V(last)
    assert np.isclose(out["advantages"][5, 0], r[5, 0] + 0.99 * nxt[0] - v[5, 0], rtol=1e-6)
    own = og.rollout_gae3(flags, r, v, z, z, z, z, heads=1)
    assert not np.array_equal(own["advantages"][:, 0], out["advantages"][:, 0])      # the CCPPO / CoPO rule differs
    assert np.array_equal(own["advantages"][:, 2], out["advantages"][:, 2])


def test_lcf_mix_standardize():
    rng = np.random.default_rng(2)
    adv, nei, gadv = (rng.normal(size=1000).astype(np.float32) for _ in range(3))
    lcf = np.clip(rng.normal(0.2, 0.1, 1000), -1, 1).astype(np.float32)
    norm, mean, std, g = og.lcf_mix_standardize(adv, nei, lcf, gadv)
    raw = np.cos(lcf * np.pi / 2) * adv + np.sin(lcf * np.pi / 2) * nei
    assert np.isclose(mean, raw.mean()) and np.isclose(std, raw.std())
    assert abs(norm.mean()) < 1e-5 and abs(norm.std() - 1) < 1e-4 and abs(g.std() - 1) < 1e-4
    # the 1e-4 floor on the std (rllib standardized)
    assert np.array_equal(og.standardized(np.full(5, 3.0, np.float32)), np.zeros(5, np.float32))
    _, _, std0, _ = og.lcf_mix_standardize(np.ones(4, np.float32), np.ones(4, np.float32), np.zeros(4, np.float32),
                                           np.ones(4, np.float32))
    assert std0 == 1e-4


GOLDEN_MODELS = ["copo_inter", "ccppo_round", "ippo_tollgate", "cl_bottle", "copo_tollgate", "ccppo_parking"]


@pytest.mark.parametrize("name", GOLDEN_MODELS)
def test_policy_forward_matches_reference_golden(name):
    z = np.load(GOLD)
    sub = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    obs, want = sub.pop("obs"), sub.pop("mean")
    path = os.path.join(HERE, "golden", "_tmp_%s.npz" % name)
    np.savez(path, **sub)
    try:
        layers = om.load_policy_npz(path)
    finally:
        os.remove(path)
    assert layers[0][0].shape == (obs.shape[1], 256) and layers[2][0].shape == (256, 4)
    got = om.mlp_forward_np(layers, obs)[:, :2]
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    if name == "copo_inter":
        assert np.allclose(got[0], [-0.5872524, -0.51996565], atol=1e-6)
    # the torch restatement of the model, loaded with the same weights, agrees too
    m = om.CCModel(obs.shape[1])
    with torch.no_grad():
        for lin, (W, b) in zip([m._hidden_layers[0]._model[0], m._hidden_layers[1]._model[0], m._logits._model[0]],
                               layers):
            lin.weight.copy_(torch.from_numpy(W.T))
            lin.bias.copy_(torch.from_numpy(b))
        out = m(torch.from_numpy(obs)).numpy()
    assert np.allclose(out[:, :2], want, rtol=1e-4, atol=1e-5)


def test_model_shapes_and_names():
    m = om.CoPOModel(92)
    n = sum(p.numel() for p in m.parameters())
    assert n == 360199 + 2                      # 360 199 network parameters + lcf_parameters
    names = set(m.state_dict().keys())
    for k in ("_hidden_layers.0._model.0.weight", "_logits._model.0.bias", "_value_branch_separate.1._model.0.weight",
              "_value_branch._model.0.weight", "nei_value_network.2._model.0.weight",
              "global_value_network.0._model.0.bias", "lcf_parameters"):
        assert k in names, k
    assert om.centralized_critic_obs_dim(91, 2, True, 4, "mf") == 184
    assert om.centralized_critic_obs_dim(91, 2, True, 4, "concat") == 463
    assert om.centralized_critic_obs_dim(92, 2, True, 4, "none") == 92
    assert abs(m.lcf_std.item() - 0.1) < 1e-7 and m.lcf_mean.item() == 0.0
    with pytest.raises(ValueError):
        m.value_function()
    w = m._hidden_layers[0]._model[0].weight
    assert torch.allclose(w.pow(2).sum(1), torch.ones(256), atol=1e-5)       # normc rows


def test_diag_gaussian_algebra():
    torch.manual_seed(0)
    a, b = torch.randn(64, 4), torch.randn(64, 4)
    x = torch.randn(64, 2)
    da, db = om.DiagGaussian(a), om.DiagGaussian(b)
    na = torch.distributions.Normal(a[:, :2], a[:, 2:].exp())
    nb = torch.distributions.Normal(b[:, :2], b[:, 2:].exp())
    assert torch.allclose(da.logp(x), na.log_prob(x).sum(1), atol=1e-5)
    assert torch.allclose(da.entropy(), na.entropy().sum(1), atol=1e-5)
    assert torch.allclose(da.kl(db), torch.distributions.kl_divergence(na, nb).sum(1), atol=1e-4, rtol=1e-4)


def _batch(B, odim, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(obs=torch.rand(B, odim, generator=g), centralized_critic_obs=torch.rand(B, odim, generator=g),
                actions=r(B, 2), action_logp=-2 + 0.1 * r(B), action_dist_inputs=0.3 * r(B, 4), advantages=r(B),
                normalized_advantages=r(B), vf_preds=r(B), value_targets=r(B), nei_values=r(B), nei_target=r(B),
                global_values=r(B), global_target=r(B), nei_advantage=r(B), global_advantages=r(B))


def test_losses_reduce_to_known_values():
    torch.manual_seed(1)
    m = om.CoPOModel(92)
    b = _batch(128, 92)
    # make the stored behaviour distribution equal to the current one: ratio = 1, KL = 0
    with torch.no_grad():
        logits = m(b["obs"])
        b["action_dist_inputs"] = logits.clone()
        b["action_logp"] = om.DiagGaussian(logits).logp(b["actions"])
    total, st = om.ppo_loss(m, b, om.DEFAULT_CFG, "copo")
    assert abs(st["mean_kl_loss"].item()) < 1e-6
    assert abs(st["mean_policy_loss"].item() + b["normalized_advantages"].mean().item()) < 1e-5
    parts = st["mean_policy_loss"] + st["mean_vf_loss"] + st["mean_nei_vf_loss"] + st["mean_global_vf_loss"]
    assert abs(total.item() - parts.item()) < 1e-4
    t2, s2 = om.ppo_loss(m, b, om.DEFAULT_CFG, "ccppo")
    assert "mean_nei_vf_loss" not in s2
    total.backward()
    assert m.lcf_parameters.grad is None and m._logits._model[0].weight.grad is not None


def test_meta_gradient_matches_finite_difference():
    torch.manual_seed(2)
    m, tgt = om.CoPOModel(92).double(), om.CoPOModel(92).double()
    b = {k: v.double() for k, v in _batch(64, 92, seed=3).items()}
    eps = torch.randn(64, dtype=torch.float64)
    final, g, st, (g_new, g_old) = om.meta_gradient(m, tgt, b, om.DEFAULT_CFG, 0.1, 1.3, eps)
    assert len(g_new) == len(g_old) == 6 and g.shape == (2,)
    h = 1e-6
    for k in range(2):
        with torch.no_grad():
            m.lcf_parameters[k] += h
        fp = om.meta_gradient(m, tgt, b, om.DEFAULT_CFG, 0.1, 1.3, eps)[0].item()
        with torch.no_grad():
            m.lcf_parameters[k] -= 2 * h
        fm = om.meta_gradient(m, tgt, b, om.DEFAULT_CFG, 0.1, 1.3, eps)[0].item()
        with torch.no_grad():
            m.lcf_parameters[k] += h
        assert abs((fp - fm) / (2 * h) - g[k].item()) < 1e-5 * max(1.0, abs(g[k].item()))


def test_adam_and_kl_controller():
    p = torch.nn.Parameter(torch.tensor([0.3, -1.2]))
    opt = torch.optim.Adam([p], lr=1e-4)
    q, m, v = p.detach().clone(), torch.zeros(2), torch.zeros(2)
    for step in range(1, 4):
        g = torch.tensor([0.5 * step, -0.1])
        p.grad = g.clone()
        opt.step()
        q, m, v = om.adam_step(q, g, m, v, step, 1e-4)
        assert torch.allclose(p.detach(), q, atol=1e-7)
    assert om.update_kl(0.2, 0.03) == pytest.approx(0.3) and om.update_kl(0.2, 0.004) == pytest.approx(0.1)
    assert om.update_kl(0.2, 0.01) == 0.2


def test_product_kl_controller():
    """IPPOPolicy.update_kl (the product's, not the oracle's) is rllib 2.2.0's KLCoeffMixin rule (called at
    algo_copo.py:631-632): x1.5 above 2 x kl_target, x0.5 below 0.5 x kl_target; inclusive edges leave it alone.
    The shipped training log shows the rule at work: cur_kl_coeff = 0.675 = 0.2 x 1.5^3."""
    from types import SimpleNamespace
    from copo_b200 import policy as P
    pol = SimpleNamespace(config=P.ippo_config(), kl_coeff=0.2)
    up = lambda kl: P.IPPOPolicy.update_kl(pol, kl)
    assert up(0.03) == pytest.approx(0.3) and pol.kl_coeff == pytest.approx(0.3)
    assert up(0.004) == pytest.approx(0.15)
    assert up(0.01) == pytest.approx(0.15) and up(0.02) == pytest.approx(0.15) and up(0.005) == pytest.approx(0.15)
    pol.kl_coeff = 0.2
    for _ in range(3):
        up(1.0)
    assert pol.kl_coeff == pytest.approx(0.675)
    for kl in (0.0, 0.0049, 0.0051, 0.0199, 0.0201, 5.0):
        pol.kl_coeff = 0.2
        assert up(kl) == pytest.approx(om.update_kl(0.2, kl))


REF_CKPT = "/root/reference/copo_code/copo/best_checkpoints"


@pytest.mark.skipif(not os.path.isdir(REF_CKPT), reason="the reference tree is only mounted in the build container")
def test_every_shipped_checkpoint_loads_and_matches_the_reference_forward():
    """All 20 `best_checkpoints/*.npz` (three naming schemes, five observation widths): the wire-format reader of the
    product (`copo_b200.checkpoint`, host-side) + the oracle forward reproduce the reference's own numpy forward
    (`eval/get_policy_function.py:54-98`, imported from the reference tree)."""
    import sys
    sys.path.insert(0, "/root/reference/copo_code")
    try:
        from copo.eval import get_policy_function as ref
    finally:
        sys.path.pop(0)
    from copo_b200 import checkpoint as ck
    files = sorted(f for f in os.listdir(REF_CKPT) if f.endswith(".npz"))
    assert len(files) == 20
    rng = np.random.default_rng(7)
    widths = set()
    for f in files:
        w = dict(np.load(os.path.join(REF_CKPT, f)))
        sd = ck.state_dict_from_policy_npz(w)
        odim = sd[ck.TORCH_POLICY[0] + ".weight"].shape[1]
        widths.add(odim)
        obs = rng.uniform(0, 1, (8, odim)).astype(np.float32)
        if f.startswith("ccppo"):
            want = ref._compute_actions_for_torch_policy(w, obs, deterministic=True)
        else:
            want = ref._compute_actions_for_tf_policy(w, obs, deterministic=True, policy_name="default",
                                                      layer_name_suffix="_1" if f.startswith("copo") else "")
        layers = [(sd[n + ".weight"].T, sd[n + ".bias"]) for n in ck.TORCH_POLICY]
        got = om.mlp_forward_np(layers, obs)[:, :2]
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6), f
        # and back out in both namings
        for naming, sfx in (("torch", ""), ("tf", "_1")):
            again = ck.state_dict_from_policy_npz(ck.policy_npz_from_state_dict(sd, naming, suffix=sfx))
            assert all(np.array_equal(again[k], sd[k]) for k in sd), (f, naming)
    assert widths == {91, 92, 96, 97, 156, 157}
