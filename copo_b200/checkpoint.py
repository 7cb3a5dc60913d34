"""Checkpoint wire formats of the reference, host-side and dependency-free.

  * policy-only `.npz` as shipped in `best_checkpoints/` - two namings (copo/eval/get_policy_function.py:54-98):
      torch era  `_hidden_layers.{0,1}._model.0.{weight,bias}`, `_logits._model.0.{weight,bias}`  (weights [out, in])
      TF era     `{policy}/fc_1{sfx}/kernel|bias`, `fc_2{sfx}`, `fc_out{sfx}`                       (kernels [in, out])
  * the RLlib trial checkpoint the reference's evaluator unpickles
    (copo/eval/get_policy_function_from_checkpoint.py:12-50):
      pickle({"worker": pickle({"state": {policy_name: {param name: ndarray}}, "filters": {}}), ...})

  * the full TF-era CoPO policy state inside such a trial checkpoint (the shipped
    `eval/demo_raw_checkpoints/copo/.../checkpoint-{490,625}`): 24 arrays `{policy}/fc_{1,2}_1`, `fc_out_1`,
    `fc_value_{1,2}_1`, `value_out_1`, `fc_value_nei_{1,2}_1`, `value_out_nei_1`, `fc_value_global_{1,2}_1`,
    `value_out_global_1` (kernels [in, out]) = the four networks of `CoPOModel` (360 199 parameters at 92 inputs).
    Those files were pickled with ray classes inside; `load_rllib_checkpoint(..., stub_ray=True)` reads them without ray.

`state_dict` below is a mapping name -> array under RLlib's torch names (what `CCModel.state_dict()` returns).
"""
import io
import pickle

import numpy as np

TORCH_POLICY = ("_hidden_layers.0._model.0", "_hidden_layers.1._model.0", "_logits._model.0")
TF_LAYERS = ("fc_1", "fc_2", "fc_out")
# the other three networks of CoPOModel: RLlib torch names <- TF-era layer names (before the `_1` suffix)
TORCH_VALUE = ("_value_branch_separate.0._model.0", "_value_branch_separate.1._model.0", "_value_branch._model.0")
TORCH_NEI = ("nei_value_network.0._model.0", "nei_value_network.1._model.0", "nei_value_network.2._model.0")
TORCH_GLOBAL = ("global_value_network.0._model.0", "global_value_network.1._model.0",
                "global_value_network.2._model.0")
TF_COPO_NETS = ((TORCH_POLICY, ("fc_1", "fc_2", "fc_out")),
                (TORCH_VALUE, ("fc_value_1", "fc_value_2", "value_out")),
                (TORCH_NEI, ("fc_value_nei_1", "fc_value_nei_2", "value_out_nei")),
                (TORCH_GLOBAL, ("fc_value_global_1", "fc_value_global_2", "value_out_global")))


def _np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=np.float32)


def policy_npz_from_state_dict(state_dict, naming="torch", policy_name="default", suffix=""):
    """Policy-only arrays in the requested naming (`suffix="_1"` is what shipped CoPO files use)."""
    out = {}
    for layer, name in zip(TF_LAYERS, TORCH_POLICY):
        W, b = _np(state_dict[name + ".weight"]), _np(state_dict[name + ".bias"])
        if naming == "torch":
            out[name + ".weight"], out[name + ".bias"] = W, b
        elif naming == "tf":
            out["%s/%s%s/kernel" % (policy_name, layer, suffix)] = np.ascontiguousarray(W.T)
            out["%s/%s%s/bias" % (policy_name, layer, suffix)] = b
        else:
            raise ValueError("naming must be 'torch' or 'tf'")
    return out


def state_dict_from_policy_npz(arrays):
    """Either naming -> RLlib torch names (policy only)."""
    arrays = dict(arrays)
    keys = list(arrays.keys())
    if TORCH_POLICY[0] + ".weight" in keys:
        return {k: _np(arrays[k]) for k in keys if any(k.startswith(n) for n in TORCH_POLICY)}
    suffix = "_1" if any(k.endswith("fc_1_1/kernel") for k in keys) else ""
    out = {}
    for layer, name in zip(TF_LAYERS, TORCH_POLICY):
        k = [x for x in keys if x.endswith("/%s%s/kernel" % (layer, suffix))][0]
        out[name + ".weight"] = np.ascontiguousarray(_np(arrays[k]).T)
        out[name + ".bias"] = _np(arrays[k.replace("kernel", "bias")])
    return out


def state_dict_from_tf_copo_state(arrays, suffix="_1"):
    """The 24 arrays of a TF-era CoPO policy state (see the module docstring) -> RLlib torch names of `CoPOModel`
    (policy, value, neighbourhood value, global value; `lcf_parameters` is not part of that state - the reference reads
    the LCF from `progress.csv`, get_policy_function_from_checkpoint.py:53-63)."""
    arrays = {k: v for k, v in dict(arrays).items() if k != "_optimizer_variables"}
    out, used = {}, set()
    for torch_names, tf_names in TF_COPO_NETS:
        for name, layer in zip(torch_names, tf_names):
            ks = [k for k in arrays if k.endswith("/%s%s/kernel" % (layer, suffix))]
            if len(ks) != 1:
                raise KeyError("expected exactly one '%s%s/kernel' in the state, found %d" % (layer, suffix, len(ks)))
            kb = ks[0].replace("kernel", "bias")
            out[name + ".weight"] = np.ascontiguousarray(_np(arrays[ks[0]]).T)      # TF kernels are [in, out]
            out[name + ".bias"] = _np(arrays[kb])
            used.update((ks[0], kb))
    extra = set(arrays) - used
    if extra:
        raise KeyError("arrays that are not part of a CoPO policy state: %s" % sorted(extra)[:4])
    return out


def tf_copo_state_from_state_dict(state_dict, policy_name="default", suffix="_1"):
    """Inverse of `state_dict_from_tf_copo_state` (what a TF-era consumer of a full CoPO state expects)."""
    out = {}
    for torch_names, tf_names in TF_COPO_NETS:
        for name, layer in zip(torch_names, tf_names):
            out["%s/%s%s/kernel" % (policy_name, layer, suffix)] = np.ascontiguousarray(_np(state_dict[name + ".weight"]).T)
            out["%s/%s%s/bias" % (policy_name, layer, suffix)] = _np(state_dict[name + ".bias"])
    return out


def save_policy_npz(path, state_dict, naming="torch", policy_name="default", suffix=""):
    np.savez(path, **policy_npz_from_state_dict(state_dict, naming, policy_name, suffix))


def save_rllib_checkpoint(path, state_dict, policy_name="default", extra=None):
    """The nested pickle the reference's `get_policy_function_from_checkpoint` reads."""
    state = {k: _np(v) for k, v in state_dict.items()}
    worker = {"state": {policy_name: state}, "filters": {}}
    blob = {"worker": pickle.dumps(worker), "train_exec_impl": None}
    if extra:
        blob.update(extra)
    with open(path, "wb") as f:
        pickle.dump(blob, f)


class _RayStub:
    """Stands in for any ray class referenced by a pickled trial checkpoint (only `NoFilter` objects in practice)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "ray" or module.startswith("ray."):
            return _RayStub
        return super().find_class(module, name)


def load_rllib_checkpoint(path, policy_name="default", stub_ray=False):
    """The policy state {param name: ndarray} of a trial checkpoint.  stub_ray=True reads files whose pickles reference
    ray classes (the shipped demo checkpoints do) without ray being installed."""
    load = (lambda b: _StubUnpickler(io.BytesIO(b)).load()) if stub_ray else pickle.loads
    with open(path, "rb") as f:
        blob = load(f.read())
    worker = load(blob["worker"])
    state = dict(worker["state"][policy_name])
    state.pop("_optimizer_variables", None)
    return state
