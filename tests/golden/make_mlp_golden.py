"""Generates tests/golden/mlp_golden.npz by running the REFERENCE's own numpy policy forward
(copo/eval/get_policy_function.py:54-98, importable without ray/metadrive) on its shipped weights
(copo/best_checkpoints/*.npz).  Run in the build container only (needs /root/reference):

    python tests/golden/make_mlp_golden.py

Stored per model: the three layer matrices as shipped (so the test does not need /root/reference), a seeded
observation batch in [0, 1], and the reference's deterministic output (the Gaussian mean).  The all-zero
observation row is row 0 (SURVEY.md 8c quotes copo_inter: [-0.5872524, -0.51996565]).
"""
import os
import sys

import numpy as np

REF = "/root/reference/copo_code"
sys.path.insert(0, REF)
from copo.eval import get_policy_function as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# every naming scheme and every observation width the reference ships (best_checkpoints/*.npz):
#   copo_*  TF-era naming with the _1 suffix (92 / 97 / 157 inputs: LCF appended)
#   ippo_*, cl_*  TF-era naming without suffix (91 / 96 / 156)
#   ccppo_*  torch-era naming
MODELS = {"copo_inter": 92, "ccppo_round": 91, "ippo_tollgate": 156, "cl_bottle": 96, "copo_tollgate": 157,
          "ccppo_parking": 91}
ROWS = {"copo_inter": 48, "ccppo_round": 48}          # the others: 16 rows


def main():
    out = {}
    rng = np.random.default_rng(2022)
    for name, odim in MODELS.items():
        path = os.path.join(REF, "copo", "best_checkpoints", name + ".npz")
        w = np.load(path)
        w = {k: w[k] for k in w.files}
        obs = rng.uniform(0, 1, (ROWS.get(name, 16), odim)).astype(np.float32)
        obs[0] = 0.0
        if name.startswith("ccppo"):
            mean = ref._compute_actions_for_torch_policy(w, obs, deterministic=True)
            names = ["_hidden_layers.0._model.0", "_hidden_layers.1._model.0", "_logits._model.0"]
            for n in names:
                out["%s/%s.weight" % (name, n)] = w[n + ".weight"]
                out["%s/%s.bias" % (name, n)] = w[n + ".bias"]
        else:
            sfx = "_1" if name.startswith("copo") else ""
            mean = ref._compute_actions_for_tf_policy(w, obs, deterministic=True, policy_name="default",
                                                      layer_name_suffix=sfx)
            for layer in ("fc_1" + sfx, "fc_2" + sfx, "fc_out" + sfx):
                out["%s/default/%s/kernel" % (name, layer)] = w["default/%s/kernel" % layer]
                out["%s/default/%s/bias" % (name, layer)] = w["default/%s/bias" % layer]
        out[name + "/obs"] = obs
        out[name + "/mean"] = np.asarray(mean)
        print(name, "zero-obs mean", np.asarray(mean)[0])
    np.savez(os.path.join(HERE, "mlp_golden.npz"), **out)


if __name__ == "__main__":
    main()
