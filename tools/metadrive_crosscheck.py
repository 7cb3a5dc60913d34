"""Behavioural cross-check of this repository's simulator against REAL MetaDrive, through what the reference holds
(SURVEY.md 7 step 0e): the reference's shipped Intersection policies (trained in MetaDrive) drive THIS simulator for
whole episodes, and the evaluation report is put next to the per-episode results the reference ships for the same
populations in MetaDrive (eval/demo_results/evaluate_results/*.csv, reduced by tests/golden/make_crosscheck_fixture.py).
This is a distributional comparison, not parity: the simulator here is a kinematic restatement ("parity
unpinned"); the table says how far its behaviour under a MetaDrive-trained policy is from MetaDrive's.

usage (GPU): python tools/metadrive_crosscheck.py [scenes] > profiles/rNN_metadrive_crosscheck.md"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from copo_b200.evaluate import evaluate  # noqa: E402
from copo_b200.models import CCModel  # noqa: E402

ROWS = (("success_rate", "success_rate"), ("crash_rate", "crash_rate"), ("out_rate", "out_rate"),
        ("velocity_step_mean_episode_mean", "velocity_step_mean_episode_mean"),
        ("num_neighbours_mean_episode_mean", "num_neighbours_step_mean"), ("num_agents_total", "agents_per_episode"),
        ("episode_cost_mean", None), ("episode_length_mean", None), ("episode_reward_mean", None))


def main(scenes=64):
    fx = np.load(os.path.join(ROOT, "tests", "golden", "metadrive_crosscheck.npz"))
    print("# Shipped MetaDrive-trained Intersection policies in this simulator vs their MetaDrive evaluation\n")
    print("`python tools/metadrive_crosscheck.py %d`: %d scenes x one 1000-step episode each, 30 agents (MetaDrive's "
          "Intersection default), stochastic actions as in the reference's evaluator (`eval/evaluate_population.py`), "
          "RecorderEnv's 20 m neighbourhood.  Reference columns: mean +- standard deviation over the episodes the "
          "reference ships (`tests/golden/make_crosscheck_fixture.py`).\n" % (scenes, scenes))
    for algo, name in (("copo", "copo_inter"), ("ippo", "ippo_inter")):
        w = {k.split("/", 1)[1]: fx[k] for k in fx.files if k.startswith(name + "/") and not k.endswith("/lcf")}
        odim = [v for k, v in w.items() if k.endswith("/kernel") and "fc_1" in k][0].shape[0]
        model = CCModel(odim, 2)
        model.load_policy_npz(w)
        kw = {}
        if algo == "copo":                               # the evaluator appends the population's mean LCF
            kw = dict(lcf_mean=float(fx["copo_inter/lcf"][0]), lcf_std=1e-3)
        rep = evaluate(model, "MultiAgentIntersectionEnv", num_scenes=scenes, num_agents=30, horizon=1000, seed=7, **kw)
        rep["agents_per_episode"] = rep["num_agents_total"] / scenes
        n_ep, n_pop = fx["reference/%s/episodes" % algo]
        print("## %s (`best_checkpoints/%s.npz`; reference: %d MetaDrive episodes of %d population members)\n" %
              (algo.upper(), name, n_ep, n_pop))
        print("| metric | MetaDrive (reference CSVs) | this simulator |\n|---|---|---|")
        for ref_key, my_key in ROWS:
            m, s = fx["reference/%s/%s" % (algo, ref_key)]
            mine = "%.3f" % rep[my_key] if my_key else "-"
            print("| %s | %.3f +- %.3f | %s |" % (ref_key, m, s, mine))
        print("| max_step_rate | - | %.3f |" % rep["max_step_rate"])
        print()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
