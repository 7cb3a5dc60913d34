"""Checkpoint wire formats of the reference, host-side and dependency-free.

  * policy-only `.npz` as shipped in `best_checkpoints/` - two namings (copo/eval/get_policy_function.py:54-98):
      torch era  `_hidden_layers.{0,1}._model.0.{weight,bias}`, `_logits._model.0.{weight,bias}`  (weights [out, in])
      TF era     `{policy}/fc_1{sfx}/kernel|bias`, `fc_2{sfx}`, `fc_out{sfx}`                       (kernels [in, out])
  * the RLlib trial checkpoint the reference's evaluator unpickles
    (copo/eval/get_policy_function_from_checkpoint.py:12-50):
      pickle({"worker": pickle({"state": {policy_name: {param name: ndarray}}, "filters": {}}), ...})

`state_dict` below is a mapping name -> array under RLlib's torch names (what `CCModel.state_dict()` returns).
"""
import pickle

import numpy as np

TORCH_POLICY = ("_hidden_layers.0._model.0", "_hidden_layers.1._model.0", "_logits._model.0")
TF_LAYERS = ("fc_1", "fc_2", "fc_out")


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


def load_rllib_checkpoint(path, policy_name="default"):
    with open(path, "rb") as f:
        blob = pickle.load(f)
    worker = pickle.loads(blob["worker"])
    state = dict(worker["state"][policy_name])
    state.pop("_optimizer_variables", None)
    return state
