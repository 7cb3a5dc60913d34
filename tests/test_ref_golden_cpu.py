"""Pins the oracle - and the product's host-side logic - against tests/golden/ref_golden.npz: outputs of the REFERENCE's
own functions executed in the build container (tests/golden/make_ref_golden.py extracts them from /root/reference with
`ast` and runs them on stand-ins for the RLlib names).  Covered: CCEnv distance map / neighbour lists, neighbourhood and
global advantages, both centralized-critic fusions, the IPPO / CCPPO / CoPO losses with their gradients, the CoPO meta
update.  The GPU twin (tests/test_ref_golden_gpu.py) holds the CUDA kernels to the same numbers."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import gae as og
from oracle import models as om
from oracle import wrappers as ow

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
META = json.loads(bytes(G["meta_json"]).decode())


def _model(key, algo, prefix="w"):
    m = META[key]
    cls = om.CoPOModel if algo == "copo" else om.CCModel
    net = cls(m["odim"], hiddens=tuple(m["hiddens"]), cdim=m.get("cdim", m["odim"]))
    sd = {k[len(key) + len(prefix) + 2:]: torch.from_numpy(G[k]) for k in G.files if k.startswith("%s/%s/" % (key, prefix))}
    net.load_state_dict(sd)
    return net


def _batch(key):
    return {k[len(key) + 7:]: torch.from_numpy(G[k]) for k in G.files if k.startswith(key + "/batch/")}


def test_distance_map_and_neighbour_lists():
    w = META["wrappers"]
    pos = G["wrappers/pos"]
    veh = {n: (None if n in w["none"] else pos[i]) for i, n in enumerate(w["names"])}
    dm = ow.update_distance_map(veh)
    for key, (ids, ds) in w["neighbours"].items():
        name, dist = key.split("@")
        got_ids, got_ds = ow.find_in_range(dm, name, int(dist))
        assert got_ids == ids and got_ds == ds, key             # same order (stable on ties), same float64 distances


def test_neighbourhood_and_global_advantages():
    nv, nr = G["adv/nei_values"], G["adv/nei_rewards"]
    gv, gr = G["adv/global_values"], G["adv/global_rewards"]
    for tag in ("cut", "done"):
        adv, tgt = og.compute_nei_advantage(nr, nv, float(nv[-1]) if tag == "cut" else 0.0, 0.99, 0.95)
        assert np.array_equal(adv, G["adv/%s/nei_advantage" % tag]) and np.array_equal(tgt, G["adv/%s/nei_target" % tag])
        adv, tgt = og.compute_global_advantage(gr, gv, float(gv[-1]) if tag == "cut" else 0.0, 1.0, 0.95)
        assert np.array_equal(adv, G["adv/%s/global_advantages" % tag])
        assert np.array_equal(tgt, G["adv/%s/global_target" % tag])


@pytest.mark.parametrize("mode", ["concat", "mf"])
@pytest.mark.parametrize("cf", [True, False])
def test_critic_obs_fusion_of_one_trajectory(mode, cf):
    """The product's per-trajectory hook (host logic of CCPPOPolicy.postprocess_trajectory) against the reference's
    concat_ccppo_process / mean_field_ccppo_process."""
    from copo_b200 import policy as P
    f = META["fuse"]
    names = sorted(f["infos"])
    batches = {n: {"t": G["fuse/in/%s/t" % n], "obs": G["fuse/in/%s/obs" % n], "actions": G["fuse/in/%s/actions" % n],
                   "infos": f["infos"][n]} for n in names}
    odim, adim = f["odim"], f["adim"]
    cdim = odim + (f["num_neighbours"] if mode == "concat" else 1) * (odim + (adim if cf else 0))
    fake = types.SimpleNamespace(config=dict(fuse_mode=mode, counterfactual=cf, num_neighbours=f["num_neighbours"],
                                             mf_nei_distance=f["mf_nei_distance"]),
                                 model=types.SimpleNamespace(cobs_dim=cdim))
    for n in names:
        others = {m: (None, batches[m]) for m in names if m != n}
        got = P.CCPPOPolicy._trajectory_critic_obs(fake, dict(batches[n]), batches[n]["obs"], others, episode=object())
        want = G["fuse/%s/cf%d/%s" % (mode, int(cf), n)]
        assert got.shape == want.shape and np.array_equal(got, want), (mode, cf, n)


@pytest.mark.parametrize("algo", ["copo", "ccppo", "ippo"])
@pytest.mark.parametrize("cname", ["default", "plain_vf"])
def test_losses_and_gradients(algo, cname):
    key = "loss/%s/%s" % (algo, cname)
    net, batch, m = _model(key, algo), _batch(key), META[key]
    total, st = om.ppo_loss(net, batch, dict(om.DEFAULT_CFG, **m["cfg"]), algo)
    net.zero_grad()
    total.backward()
    g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])
    want = torch.from_numpy(G[key + "/grad"])
    assert float((g - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))
    for k, v in m["stats"].items():
        if k in st:
            assert abs(float(st[k]) - v) <= 1e-6 * max(1.0, abs(v)), (k, float(st[k]), v)
    assert {"total_loss", "mean_policy_loss", "mean_vf_loss", "mean_entropy", "mean_kl_loss"} <= set(m["stats"])


def test_meta_update():
    m = META["meta"]
    net = om.CoPOModel(m["odim"], hiddens=tuple(m["hiddens"]))
    old = om.CoPOModel(m["odim"], hiddens=tuple(m["hiddens"]))
    net.load_state_dict({k[len("meta/w/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("meta/w/")})
    old.load_state_dict({k[len("meta/w_old/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("meta/w_old/")})
    batch = {k[len("meta/batch/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("meta/batch/")}
    final, g_lcf, st, _ = om.meta_gradient(net, old, batch, dict(clip_param=m["clip_param"]), m["raw_mean"], m["raw_std"],
                                           torch.from_numpy(G["meta/eps"]))
    for k in ("new_policy_ego_loss", "old_policy_logp_loss", "lcf_lcf_adv_loss", "lcf_final_loss", "grad_value",
              "coordinated_adv", "global_adv"):
        assert abs(st[k] - m["stats"][k]) <= 2e-6 * max(1.0, abs(m["stats"][k])), (k, st[k], m["stats"][k])
    # one Adam step on the LCF parameters (torch.optim.Adam, lr = lcf_lr) lands where the reference's optimizer did
    p = net.lcf_parameters.detach().double()
    p2, _, _ = om.adam_step(p, g_lcf.double(), torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64), 1,
                            m["lcf_lr"])
    p2 = np.asarray(p2)
    assert np.allclose(p2, G["meta/lcf_parameters_after"], rtol=0, atol=1e-6)
    assert abs(float(np.tanh(p2[0])) - m["stats"]["lcf"]) < 1e-6 and abs(float(np.exp(p2[1])) - m["stats"]["lcf_std"]) < 1e-6


@pytest.mark.parametrize("scenario", range(4))
def test_wrapper_steps_against_the_reference_methods(scenario):
    """CCEnv.step + LCFEnv.step + _add_lcf of the reference, executed over a scripted base env (agents leaving and
    joining; angle / linear mixing, native / coordinated return, forced + normal redraw every step, uniform draws),
    against the oracle's restatement fed with the same random stream."""
    sc = META["lcf_env"][scenario]
    cfg = sc["config"]
    rng = np.random.RandomState(sc["rng_seed"])
    lcf_map = {}
    for step in sc["steps"]:
        names = list(step["reward"].keys())
        pos = {k: np.asarray(step["pos"][k], np.float64) for k in names}
        infos = ow.cc_step(pos, step["reward"], cfg["neighbours_distance"])
        now = {}
        for k in names:                                          # the reference draws inside its per-agent loop, in order
            lcf, obs = ow.add_lcf(np.asarray(step["obs"][k], np.float32), lcf_map.get(k), rng, True, cfg["force_lcf"],
                                  cfg["lcf_dist"], sc["lcf_mean"], sc["lcf_std"])
            lcf_map.setdefault(k, lcf)                           # the episode value is set once (:341-342)
            now[k] = lcf
            assert abs(float(obs[-1]) - step["out_obs_last"][k]) < 1e-7
        got_r = ow.lcf_step(dict(step["reward"]), infos, now, cfg["lcf_mode"], cfg["return_native_reward"])
        for k in names:
            want = step["info"][k]
            assert infos[k]["neighbours"] == want["neighbours"]
            assert infos[k]["neighbours_distance"] == want["neighbours_distance"]
            for q in ("nei_rewards", "global_rewards", "lcf", "coordinated_rewards", "native_rewards"):
                assert abs(infos[k][q] - want[q]) <= 1e-12 * max(1.0, abs(want[q])), (k, q)
            assert abs(got_r[k] - step["out_reward"][k]) <= 1e-12 * max(1.0, abs(step["out_reward"][k]))
    if cfg["force_lcf"] != -100:                                 # forced + normal: a fresh value every step
        a = [s["info"]["agent0"]["lcf"] for s in sc["steps"]]
        assert len(set(a)) == len(a)


def test_metadrive_crosscheck_fixture_is_the_shipped_data():
    """tests/golden/metadrive_crosscheck.npz (tools/metadrive_crosscheck.py): the policy in it is the shipped
    `copo_inter.npz` (zero observation -> the mean SURVEY.md 8c quotes), and the MetaDrive statistics are the reference's
    evaluate_results CSVs (success rates of the CoPO / IPPO Intersection populations as DrawEvalResult.ipynb shows them)."""
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metadrive_crosscheck.npz"))
    x = np.zeros((1, 92), np.float32)
    for layer in ("fc_1_1", "fc_2_1", "fc_out_1"):
        x = x @ fx["copo_inter/default/%s/kernel" % layer] + fx["copo_inter/default/%s/bias" % layer]
        if layer != "fc_out_1":
            x = np.tanh(x)
    assert np.allclose(x[0, :2], [-0.5872524, -0.51996565], atol=1e-6)
    assert fx["ippo_inter/default/fc_1/kernel"].shape == (91, 256)
    assert abs(fx["copo_inter/lcf"][0] - 0.36824979071031544) < 1e-12
    assert tuple(fx["reference/copo/episodes"]) == (100, 5) and tuple(fx["reference/ippo/episodes"]) == (120, 6)
    assert abs(fx["reference/copo/success_rate"][0] - 0.7825) < 5e-4 and abs(fx["reference/ippo/success_rate"][0] - 0.4805) < 5e-4


def test_shipped_metadrive_policies_follow_the_road_under_the_specified_conventions():
    """The two conventions of the simulator specification that reference-held data pins (crosscheck_probe.py,
    profiles/r02_e_metadrive_crosscheck.md): the reference's shipped MetaDrive-trained Intersection policies, alone on the
    road, reach their destinations in the specified simulator - and never do when the steering sign and the order of the
    two lateral-distance observations are put back the way they were."""
    import crosscheck_probe as probe
    for name, floor in (("ippo_inter", 0.7), ("copo_inter", 0.5)):
        spec = probe.run(name, (), S=8, A=1, T=300)
        undone = probe.run(name, ("undo_steer", "undo_lat"), S=8, A=1, T=300)
        assert spec["finished"] >= 10 and spec["success"] >= floor and spec["crash"] == 0.0, (name, spec)
        assert undone["success"] == 0.0 and undone["out"] >= 0.9, (name, undone)
