"""Checkpoint wire formats: both npz namings and the RLlib pickle layout round-trip, and - when
the reference tree is present (build container) - the REFERENCE's own loader + numpy forward reads what we write."""
import os
import sys

import numpy as np
import pytest
import torch

from copo_b200 import checkpoint as ck
from oracle import models as om

REF = "/root/reference/copo_code"


def _state(odim=91, seed=0):
    torch.manual_seed(seed)
    m = om.CCModel(odim)
    return m, {k: v.detach().numpy() for k, v in m.state_dict().items()}


def test_npz_namings_round_trip(tmp_path):
    m, sd = _state()
    for naming, suffix in (("torch", ""), ("tf", ""), ("tf", "_1")):
        p = str(tmp_path / ("p_%s%s.npz" % (naming, suffix)))
        ck.save_policy_npz(p, sd, naming=naming, suffix=suffix)
        back = ck.state_dict_from_policy_npz(np.load(p))
        assert set(back) == {n + s for n in ck.TORCH_POLICY for s in (".weight", ".bias")}
        for k, v in back.items():
            assert np.array_equal(v, sd[k]), k
        # the oracle's loader (used by the golden tests) reads the same file
        layers = om.load_policy_npz(p)
        obs = np.random.default_rng(0).random((5, 91)).astype(np.float32)
        want = m(torch.from_numpy(obs)).detach().numpy()
        assert np.allclose(om.mlp_forward_np(layers, obs), want, atol=1e-5)
    with pytest.raises(ValueError):
        ck.policy_npz_from_state_dict(sd, naming="onnx")


def test_rllib_pickle_round_trip(tmp_path):
    _, sd = _state(92, seed=1)
    p = str(tmp_path / "checkpoint-1")
    ck.save_rllib_checkpoint(p, dict(sd, _optimizer_variables=[1, 2]))
    back = ck.load_rllib_checkpoint(p)
    assert "_optimizer_variables" not in back and set(back) == set(sd)
    assert all(np.array_equal(back[k], sd[k]) for k in sd)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_loader_reads_our_files(tmp_path):
    sys.path.insert(0, REF)
    try:
        from copo.eval import get_policy_function as gpf
        from copo.eval.get_policy_function_from_checkpoint import get_policy_function_from_checkpoint
    finally:
        sys.path.remove(REF)
    m, sd = _state(91, seed=2)
    obs = np.random.default_rng(1).random((7, 91)).astype(np.float32)
    want = m(torch.from_numpy(obs)).detach().numpy()[:, :2]
    # torch-era npz through the reference's numpy forward
    w = ck.policy_npz_from_state_dict(sd, "torch")
    assert np.allclose(gpf._compute_actions_for_torch_policy(w, obs, deterministic=True), want, atol=1e-5)
    # TF-era npz with the CoPO suffix
    w = ck.policy_npz_from_state_dict(sd, "tf", suffix="_1")
    got = gpf._compute_actions_for_tf_policy(w, obs, deterministic=True, policy_name="default", layer_name_suffix="_1")
    assert np.allclose(got, want, atol=1e-5)
    # RLlib trial checkpoint through the reference's checkpoint evaluator ("ccppo" = torch naming branch)
    p = str(tmp_path / "checkpoint-7")
    ck.save_rllib_checkpoint(p, sd)
    fn = get_policy_function_from_checkpoint("ccppo", p, deterministic=True)
    act = fn({"agent%d" % i: obs[i] for i in range(7)}, {})
    got = np.stack([act["agent%d" % i] for i in range(7)])
    assert np.allclose(got, want, atol=1e-5)


def test_tf_copo_state_round_trip():
    torch.manual_seed(3)
    m = om.CoPOModel(92)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    tf = ck.tf_copo_state_from_state_dict(sd)
    assert len(tf) == 24 and tf["default/fc_value_nei_1_1/kernel"].shape == (92, 256)
    back = ck.state_dict_from_tf_copo_state(dict(tf, _optimizer_variables=None))
    assert set(back) == set(sd) - {"lcf_parameters"}
    assert all(np.array_equal(back[k], sd[k]) for k in back)
    with pytest.raises(KeyError):
        ck.state_dict_from_tf_copo_state({k: v for k, v in tf.items() if "value_out_nei" not in k})
    with pytest.raises(KeyError):
        ck.state_dict_from_tf_copo_state(dict(tf, **{"default/extra/kernel": np.zeros(1)}))


DEMO = os.path.join(REF, "copo", "eval", "demo_raw_checkpoints", "copo")


@pytest.mark.skipif(not os.path.isdir(DEMO), reason="reference tree not present")
@pytest.mark.parametrize("step", [490, 625])
def test_shipped_full_copo_checkpoint_loads_into_the_four_networks(step):
    """the shipped trial checkpoints hold the full TF-era CoPO state (24 arrays, 360 199 parameters);
    they load - without ray - into the four networks of CoPOModel, and the policy part reproduces the reference's own
    numpy forward on the very same arrays."""
    import glob
    path = glob.glob(os.path.join(DEMO, "*", "checkpoint_%d" % step, "checkpoint-%d" % step))[0]
    state = ck.load_rllib_checkpoint(path, stub_ray=True)
    assert len(state) == 24 and sum(v.size for v in state.values()) == 360199
    sd = ck.state_dict_from_tf_copo_state(state)
    m = om.CoPOModel(92)
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert list(missing.missing_keys) == ["lcf_parameters"] and not missing.unexpected_keys
    sys.path.insert(0, REF)
    try:
        from copo.eval import get_policy_function as gpf
    finally:
        sys.path.remove(REF)
    obs = np.random.default_rng(step).random((9, 92)).astype(np.float32)
    obs[0] = 0.0
    want = gpf._compute_actions_for_tf_policy({k: v for k, v in state.items() if "value" not in k}, obs,
                                              deterministic=True, policy_name="default", layer_name_suffix="_1")
    with torch.no_grad():
        got = m(torch.from_numpy(obs)).numpy()[:, :2]
        values = [f(torch.from_numpy(obs)).numpy() for f in (m.central_value_function, m.get_nei_value, m.get_global_value)]
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5)
    # the three value heads: restated forward on the shipped arrays ([in, out] kernels, tanh hidden layers)
    for v, (h1, h2, out) in zip(values, (("fc_value_1", "fc_value_2", "value_out"),
                                         ("fc_value_nei_1", "fc_value_nei_2", "value_out_nei"),
                                         ("fc_value_global_1", "fc_value_global_2", "value_out_global"))):
        x = obs.astype(np.float64)
        for layer, act in ((h1, True), (h2, True), (out, False)):
            x = x @ state["default/%s_1/kernel" % layer].astype(np.float64) + state["default/%s_1/bias" % layer]
            x = np.tanh(x) if act else x
        assert v.shape == (9,) and np.allclose(v, x[:, 0], rtol=1e-4, atol=1e-5)
