"""The CUDA path against tests/golden/ref_golden.npz - outputs of the REFERENCE's own loss / meta-update / advantage /
fusion code executed in the build container (tests/golden/make_ref_golden.py).  Same fixtures as
tests/test_ref_golden_cpu.py, which pins the oracle to them; here the kernels are held to the reference's numbers
directly: losses, statistics and the full flat gradient of IPPO / CCPPO / CoPO (fp32 kernels, 32-32 hidden layers as
in the fixture), the CoPO meta update incl. the Adam step on the LCF parameters, the three GAE streams, the batched
critic-obs fusion."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
META = json.loads(bytes(G["meta_json"]).decode())


def _policy(key, algo, prefix="w", **over):
    from copo_b200 import policy as P
    m = META[key]
    cls = {"copo": P.CoPOPolicy, "ccppo": P.CCPPOPolicy, "ippo": P.IPPOPolicy}[algo]
    cfg = cls.default_config()
    cfg.update(precision="fp32", fcnet_hiddens=list(m["hiddens"]), fuse_mode="mf" if m.get("cdim", m["odim"]) != m["odim"] else "none")
    cfg.update(over)
    pol = cls(m["odim"], 2, cfg)
    sd = {k[len(key) + len(prefix) + 2:]: G[k] for k in G.files if k.startswith("%s/%s/" % (key, prefix))}
    pol.model.load_state_dict(sd)
    return pol


def _flat(model, key, prefix="grad"):
    """The fixture's flat gradient (torch parameter order of the oracle model) re-laid like the product's flat buffer."""
    from oracle import models as om
    m = META[key]
    ref = (om.CoPOModel if "lcf_parameters" in [k.split("/")[-1] for k in G.files if k.startswith(key + "/w/")] else om.CCModel)(
        m["odim"], hiddens=tuple(m["hiddens"]), cdim=m.get("cdim", m["odim"]))
    g, off, named = torch.from_numpy(G[key + "/" + prefix]), 0, {}
    for name, p in ref.named_parameters():
        named[name] = g[off:off + p.numel()]
        off += p.numel()
    parts = []
    for net in model.nets.values():
        for name in net.names:
            parts += [named[name + ".weight"], named[name + ".bias"]]
    return torch.cat(parts)


@pytest.mark.parametrize("algo", ["copo", "ccppo", "ippo"])
@pytest.mark.parametrize("cname", ["default", "plain_vf"])
def test_losses_and_gradients_against_the_reference_code(algo, cname):
    key = "loss/%s/%s" % (algo, cname)
    m = META[key]
    c = m["cfg"]
    pol = _policy(key, algo, clip_param=c["clip_param"], vf_clip_param=c["vf_clip_param"], vf_loss_coeff=c["vf_loss_coeff"],
                  old_value_loss=c["old_value_loss"])
    pol.kl_coeff, pol.entropy_coeff = c["kl_coeff"], c["entropy_coeff"]
    batch = {k[len(key) + 7:]: torch.from_numpy(G[k]).cuda() for k in G.files if k.startswith(key + "/batch/")}
    if algo == "ippo":
        batch["centralized_critic_obs"] = batch["obs"]
    pol.model.zero_grad()
    total = float(pol.loss(pol.model, None, batch))
    st = m["stats"]
    assert abs(total - st["total_loss"]) <= 2e-5 * max(1.0, abs(st["total_loss"]))
    ts = pol.model.tower_stats
    for k in ("mean_policy_loss", "mean_vf_loss", "mean_entropy", "mean_kl_loss", "mean_nei_vf_loss", "mean_global_vf_loss"):
        if k in st and k in ts:
            assert abs(float(ts[k]) - st[k]) <= 2e-5 * max(1.0, abs(st[k])), (k, float(ts[k]), st[k])
    want = _flat(pol.model, key).cuda()
    err = float((pol.model.grad - want).abs().max())
    assert err <= 2e-4 * float(want.abs().max()), err


def test_meta_update_against_the_reference_code():
    from copo_b200 import policy as P
    m = META["meta"]
    cfg = P.copo_config(precision="fp32", fcnet_hiddens=list(m["hiddens"]), clip_param=m["clip_param"], lcf_lr=m["lcf_lr"])
    pol = P.CoPOPolicy(m["odim"], 2, cfg)
    pol.model.load_state_dict({k[len("meta/w/"):]: G[k] for k in G.files if k.startswith("meta/w/")})
    pol.target_model.load_state_dict({k[len("meta/w_old/"):]: G[k] for k in G.files if k.startswith("meta/w_old/")})
    pol._raw_lcf_adv_mean, pol._raw_lcf_adv_std = m["raw_mean"], m["raw_std"]
    batch = {k[len("meta/batch/"):]: torch.from_numpy(G[k]).cuda() for k in G.files if k.startswith("meta/batch/")}
    out = pol.meta_update(batch, eps=torch.from_numpy(G["meta/eps"]).cuda())
    st = m["stats"]
    for k, tol in (("new_policy_ego_loss", 1e-5), ("old_policy_logp_loss", 1e-5), ("lcf_lcf_adv_loss", 1e-5),
                   ("grad_value", 3e-4), ("lcf_final_loss", 3e-4), ("coordinated_adv", 1e-5), ("global_adv", 1e-5),
                   ("lcf", 1e-5), ("lcf_std", 1e-5), ("lcf_param", 1e-5), ("lcf_std_param", 1e-5)):
        assert abs(out[k] - st[k]) <= tol * max(1.0, abs(st[k])), (k, out[k], st[k])
    got = pol.model.lcf_parameters.cpu().numpy()
    assert np.allclose(got, G["meta/lcf_parameters_after"], rtol=0, atol=2e-6)        # torch.optim.Adam step of the reference


def test_advantage_streams_against_the_reference_code():
    """compute_nei_advantage / compute_global_advantage (algo_copo.py:189-204) on one trajectory, cut (bootstrapped with
    its own last value, algo_ccppo.py:362-365) and terminated, through the gae3 kernel."""
    from copo_b200 import ops
    T = len(G["adv/nei_values"])
    for tag in ("cut", "done"):
        flags = torch.ones((T, 1), dtype=torch.uint8, device="cuda")
        if tag == "done":
            flags[-1, 0] = 3
        col = lambda k: torch.from_numpy(G["adv/" + k]).reshape(T, 1).cuda().contiguous()
        z = torch.zeros((T, 1), device="cuda")
        adv, tgt = ops.gae3(flags, [z, col("nei_rewards"), col("global_rewards")], [z, col("nei_values"), col("global_values")],
                            0.99, 0.95)
        for h, (ka, kt) in ((1, ("nei_advantage", "nei_target")), (2, ("global_advantages", "global_target"))):
            wa, wt = G["adv/%s/%s" % (tag, ka)], G["adv/%s/%s" % (tag, kt)]
            assert np.allclose(adv[h].reshape(T).cpu().numpy(), wa, rtol=1e-5, atol=1e-6), (tag, ka)
            assert np.allclose(tgt[h].reshape(T).cpu().numpy(), wt, rtol=1e-5, atol=1e-6), (tag, kt)


@pytest.mark.parametrize("mode", ["concat", "mf"])
@pytest.mark.parametrize("cf", [True, False])
def test_batched_critic_obs_fusion_against_the_reference_code(mode, cf):
    """The fixture's six trajectories laid out as one scene of six slots x twelve steps: the cc_obs_fuse kernel (masks /
    nearest-four lists derived from the fixture's infos) gives the rows concat_ccppo_process / mean_field_ccppo_process
    gave."""
    from copo_b200 import ops
    f = META["fuse"]
    names = sorted(f["infos"])
    A, T, odim, adim = len(names), 12, f["odim"], f["adim"]
    slot = {n: i for i, n in enumerate(names)}
    obs, act = np.zeros((T, A, odim), np.float32), np.zeros((T, A, adim), np.float32)
    flags = np.zeros((T, A), np.uint8)
    mf = np.zeros((T, A), np.uint64)
    nl = -np.ones((T, A, 4), np.int8)
    for n in names:
        for i, t in enumerate(G["fuse/in/%s/t" % n]):
            obs[t, slot[n]], act[t, slot[n]], flags[t, slot[n]] = G["fuse/in/%s/obs" % n][i], G["fuse/in/%s/actions" % n][i], 1
            info = f["infos"][n][i]
            for rank, (m_, d) in enumerate(zip(info["neighbours"], info["neighbours_distance"])):
                if d <= f["mf_nei_distance"]:
                    mf[t, slot[n]] |= np.uint64(1) << np.uint64(slot[m_])
                if rank < 4:
                    nl[t, slot[n], rank] = slot[m_]
    dev = lambda a, dt=None: torch.from_numpy(a if dt is None else a.view(dt)).cuda()
    cobs = ops.cc_obs_fuse(dev(obs.reshape(T * A, odim)), dev(act.reshape(T * A, adim)), dev(flags.reshape(-1)),
                           dev(mf.reshape(-1), np.int64) if mode == "mf" else None,
                           dev(nl.reshape(T * A, 4)) if mode == "concat" else None, A, mode, cf).cpu().numpy().reshape(T, A, -1)
    for n in names:
        want = G["fuse/%s/cf%d/%s" % (mode, int(cf), n)]
        for i, t in enumerate(G["fuse/in/%s/t" % n]):
            assert np.allclose(cobs[t, slot[n]], want[i], rtol=1e-6, atol=1e-7), (mode, cf, n, t)
