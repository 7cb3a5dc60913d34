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
