"""Generates tests/golden/metadrive_crosscheck.npz in the build container (needs /root/reference): what the reference
holds about REAL MetaDrive behaviour on the Intersection map, for tools/metadrive_crosscheck.py to put next to this
repository's simulator -

  * the shipped policies the reference evaluated (copo/best_checkpoints/{copo_inter, ippo_inter}.npz, weights as
    shipped) and the LCF the reference's evaluator feeds the CoPO policy (eval/get_policy_function.py:31,
    meta_svo_lookup_table["copo_inter"]),
  * the per-episode evaluation results the reference ships for its CoPO / IPPO Intersection populations
    (eval/demo_results/evaluate_results/{copo,ippo}_inter_*.csv: RecorderEnv reports of 20 MetaDrive episodes per
    population member), reduced to mean and standard deviation per column.

    python tests/golden/make_crosscheck_fixture.py
"""
import glob
import os
import re

import numpy as np
import pandas as pd

REF = "/root/reference/copo_code/copo"
HERE = os.path.dirname(os.path.abspath(__file__))
COLUMNS = ("success_rate", "crash_rate", "out_rate", "velocity_step_mean_episode_mean", "num_neighbours_mean_episode_mean",
           "episode_length_mean", "num_agents_total", "episode_reward_mean", "episode_cost_mean")


def main():
    out = {}
    for name in ("copo_inter", "ippo_inter"):
        w = np.load(os.path.join(REF, "best_checkpoints", name + ".npz"))
        for k in w.files:
            out["%s/%s" % (name, k)] = w[k]
    src = open(os.path.join(REF, "eval", "get_policy_function.py")).read()
    m = re.search(r'"copo_inter":\s*\(([0-9.eE+-]+),\s*([0-9.eE+-]+)\)', src)
    out["copo_inter/lcf"] = np.array([float(m.group(1)), float(m.group(2))])
    for algo in ("copo", "ippo"):
        files = sorted(glob.glob(os.path.join(REF, "eval", "demo_results", "evaluate_results", "%s_inter_*.csv" % algo)))
        d = pd.concat([pd.read_csv(f) for f in files])
        out["reference/%s/episodes" % algo] = np.array([len(d), len(files)])
        for c in COLUMNS:
            out["reference/%s/%s" % (algo, c)] = np.array([d[c].mean(), d[c].std()])
    np.savez_compressed(os.path.join(HERE, "metadrive_crosscheck.npz"), **out)
    print({k: (v.shape if v.size > 4 else v.tolist()) for k, v in out.items()})


if __name__ == "__main__":
    main()
