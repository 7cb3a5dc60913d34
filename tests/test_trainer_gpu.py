"""End-to-end checks of the host API on the GPU: the dict-API environments (reference's env interface) against the
oracle, and full training iterations of the three trainers on a small scene batch."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_dict_env_follows_reference_interface():
    from copo_b200 import envs
    from copo_b200.maps import build_map
    from oracle import sim as osim
    cls = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)
    assert cls.__name__ == "LCFMultiAgentIntersectionEnv" and cls.default_config()["neighbours_distance"] == 40
    name = envs.get_rllib_compatible_env(cls)
    env = envs.make_env(name, {"num_agents": 12, "start_seed": 5})
    A = 12
    cfg = osim.SimConfig(seed=5, auto_reset=False)
    cfg.num_agents = A
    ref = osim.OracleSim(build_map("intersection"), 1, A, cfg)
    ref.episode[:] = 0
    o = env.reset()
    r0 = ref.reset(new_episode=True)
    assert len(o) == A and all(v.shape == (92,) and v.dtype == np.float32 for v in o.values())
    assert set(o.keys()) == set(env.observation_space.keys()) == set("agent%d" % i for i in range(A))
    rng = np.random.default_rng(0)
    seen_done = 0
    for t in range(150):
        acts = {k: np.array([rng.uniform(-0.3, 0.3), rng.uniform(0, 1)], np.float32) for k in env.vehicles}
        arr = np.zeros((1, A, 2), np.float32)
        for k, a in acts.items():
            arr[0, env._slot_of[k]] = a
        o, r, d, i = env.step(acts)
        w = ref.step(arr)
        assert set(o.keys()) == set(i.keys()) == set(r.keys()) and "__all__" in d
        for k in o:
            s = env._slot_now[k]
            assert k == "agent%d" % w["agent_id"][0, s]
            assert np.array_equal(o[k], w["obs"][0, s])
            inf = i[k]
            if w["flags"][0, s] & osim.F_VALID:
                assert r[k] == float(w["reward"][0, s]) and d[k] == bool(w["flags"][0, s] & osim.F_DONE)
            assert inf["nei_rewards"] == float(w["nei_reward"][0, s])
            assert inf["global_rewards"] == float(w["global_reward"][0])
            assert inf["lcf"] == float(w["lcf"][0, s]) and 0.0 <= o[k][-1] <= 1.0
            mask = 0
            for n in inf["neighbours"]:
                mask |= 1 << env._slot_now[n]
            assert mask == int(w["nei_mask"][0, s])
            assert inf["neighbours_distance"] == sorted(inf["neighbours_distance"])
            want = math.cos(inf["lcf"] * math.pi / 2) * r[k] + math.sin(inf["lcf"] * math.pi / 2) * inf["nei_rewards"]
            assert abs(inf["coordinated_rewards"] - want) < 1e-9
            if w["flags"][0, s] & osim.F_VALID:
                for key in ("velocity", "steering", "acceleration", "step_reward", "cost", "episode_length",
                            "episode_reward", "arrive_dest", "crash", "out_of_road", "route_completion", "all_agents"):
                    assert key in inf
            else:                              # freshly spawned: only what the wrappers add (eval/recoder.py:128 keys on it)
                assert "step_reward" not in inf and "all_agents" in inf
            seen_done += int(d[k])
    assert seen_done > 0
    env.set_lcf_dist(0.3, 0.05)
    with pytest.raises(AssertionError):
        env.set_lcf_dist(0.3, 0.0)
    env.close()


@pytest.mark.parametrize("algo", ["copo", "ccppo", "ippo"])
def test_training_iterations(algo):
    from copo_b200 import trainer as T
    cls = {"copo": T.CoPOTrainer, "ccppo": T.CCPPOTrainer, "ippo": T.IPPOTrainer}[algo]
    tr = cls(dict(env="MultiAgentRoundaboutEnv" if algo == "ccppo" else "MultiAgentIntersectionEnv", num_scenes=16,
                  rollout_fragment_length=24, sgd_minibatch_size=2048, num_sgd_iter=2, lcf_num_iters=2,
                  env_config={"num_agents": 20}, seed=1))
    assert tr.policy.model.obs_dim == (92 if algo == "copo" else 91)
    if algo == "ccppo":
        assert tr.policy.model.cobs_dim == 184
    p0 = tr.policy.model.flat.clone()
    for it in range(2):
        res = tr.train()
        st = res["info"]["learner"]["default"]["learner_stats"]
        for k in ("total_loss", "policy_loss", "vf_loss", "kl", "entropy", "cur_kl_coeff"):
            assert math.isfinite(st[k]), (k, st[k])
        cm = res["custom_metrics"]
        assert 0 <= cm["success_rate"] <= 1 and cm["agent_steps"] > 0
    assert res["training_iteration"] == 2 and res["timesteps_total"] == 2 * 24 * 16
    assert not torch.equal(p0, tr.policy.model.flat) and torch.isfinite(tr.policy.model.flat).all()
    if algo == "copo":
        mu = res["custom_metrics"]["meta_update"]
        for k in ("grad_value", "lcf", "lcf_std", "lcf_final_loss", "raw_lcf_adv_mean_value"):
            assert math.isfinite(mu[k])
        assert torch.equal(tr.policy.target_model.flat, tr.policy.model.flat)          # theta_old handed over
        assert abs(tr.env.lcf_mean - float(tr.policy.model.lcf_mean)) < 1e-6           # envs draw from the new LCF
        assert float(tr.policy.model.lcf_parameters[0]) != 0.0
    tr.stop()


def test_tensor_core_values_keep_gae_within_tolerance():
    """North-star tolerance: advantages / value targets computed from tensor-core (split-bf16) value predictions stay
    within 1e-4 (relative to the advantage scale) of those from the exact fp32 kernels, on a real rollout."""
    from copo_b200 import trainer as T
    tr = T.CoPOTrainer(dict(env="MultiAgentIntersectionEnv", num_scenes=32, rollout_fragment_length=40,
                            env_config={"num_agents": 40}, seed=3))
    # non-trivial value heads: scale the output layers up so values are O(1..10) like trained critics
    for name in ("value", "nei", "global"):
        tr.policy.model.nets[name].W[2].mul_(300.0)
    tr.policy.model.mark_weights_changed()
    ro = tr.sample()
    res = {}
    for prec in ("fp32", "bf16_split"):
        tr.policy.model.precision = prec
        out = tr.policy.postprocess_rollout(dict(ro))
        res[prec] = {k: out[k].clone() for k in ("vf_preds", "advantages", "value_targets", "nei_advantage",
                                                  "global_advantages", "global_target")}
    valid = (ro["flags"] & 1) > 0
    assert float(res["fp32"]["vf_preds"][valid].abs().mean()) > 0.5
    for k, a in res["fp32"].items():
        b = res["bf16_split"][k]
        scale = float(a[valid].abs().mean()) + 1e-12
        err = float((a - b)[valid].abs().max())
        assert err <= 1e-4 * scale + 1e-6, (k, err, scale)
    tr.stop()


def test_batched_evaluation_of_a_shipped_policy():
    """a shipped CoPO policy (golden fixture weights, TF-era naming) loads and is evaluated on
    the batched simulator; the report carries the reference's evaluation columns."""
    import os
    from copo_b200.evaluate import evaluate
    from copo_b200.models import CCModel
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mlp_golden.npz"))
    sub = {k[len("copo_inter") + 1:]: z[k] for k in z.files if k.startswith("copo_inter/")}
    m = CCModel(92)
    m.load_policy_npz(sub)
    res = evaluate(m, "MultiAgentIntersectionEnv", num_scenes=16, num_agents=30, horizon=150, seed=1, lcf_mean=0.225,
                   lcf_std=0.1)
    for k in ("success_rate", "crash_rate", "out_rate", "max_step_rate", "velocity_step_mean_episode_mean",
              "num_neighbours_step_mean", "step_reward_mean", "num_agents_success_per_300_steps"):
        assert math.isfinite(res[k]), k
    total = res["success_rate"] + res["crash_rate"] + res["out_rate"] + res["max_step_rate"]
    assert 0.99 <= total <= 2.0              # an agent can crash and leave the road in the same step
    assert res["agent_steps"] > 16 * 20 * 100 and res["velocity_step_mean_episode_max"] > 1.0


def test_evaluation_report_equals_the_oracle_replay():
    """the report `evaluate()` reduces on the device against an independent computation - the same
    actions replayed through the numpy oracle simulator (bit-identical step outputs), the report's quantities recomputed
    from the oracle's outputs with plain numpy following eval/recoder.py:177-349 (rates over finished agents, step means
    over acting agents, neighbour counts within the evaluation radius)."""
    import os
    from copo_b200.evaluate import evaluate
    from copo_b200.maps import build_map
    from copo_b200.models import CCModel
    from oracle import sim as osim
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mlp_golden.npz"))
    sub = {k[len("copo_inter") + 1:]: z[k] for k in z.files if k.startswith("copo_inter/")}
    m = CCModel(92)
    m.load_policy_npz(sub)
    S, A, H = 3, 30, 120
    trace = []
    res = evaluate(m, "MultiAgentIntersectionEnv", num_scenes=S, num_agents=A, horizon=H, seed=4, lcf_mean=0.225,
                   lcf_std=0.1, trace=trace)
    assert len(trace) == H
    cfg = osim.SimConfig(num_agents=A, horizon=H, auto_reset=False, append_lcf=True, seed=4, lcf_mean=0.225, lcf_std=0.1,
                         neighbours_distance=20.0)
    ref = osim.OracleSim(build_map("intersection"), S, A, cfg)
    ref.reset()
    acc = dict(done=0, success=0, crash=0, out=0, max_step=0, steps=0, reward=0.0, cost=0, nei=0)
    vels = []
    for t in range(H):
        o = ref.step(trace[t])
        f = o["flags"].astype(np.int64)
        valid, done = (f & 1) > 0, (f & 2) > 0
        acc["steps"] += int(valid.sum()); acc["done"] += int(done.sum())
        acc["success"] += int(((f & 4) > 0).sum()); acc["crash"] += int(((f & 8) > 0).sum())
        acc["out"] += int(((f & 16) > 0).sum()); acc["max_step"] += int((((f & 32) > 0) & done).sum())
        acc["reward"] += float((o["reward"].astype(np.float64) * valid).sum())
        acc["cost"] += int((((f & 8) > 0) & valid).sum())
        bits = o["nei_mask"].astype(np.uint64)
        cnt = np.zeros(bits.shape, np.int64)
        for j in range(A):
            cnt += ((bits >> np.uint64(j)) & np.uint64(1)).astype(np.int64)
        acc["nei"] += int((cnt * valid).sum())
        if valid.any():
            vels.append(float((o["obs"][..., 3].astype(np.float64) * valid).sum() / valid.sum() * 22.22222137451172 * 3.6))
    d = max(acc["done"], 1)
    want = {"success_rate": acc["success"] / d, "crash_rate": acc["crash"] / d, "out_rate": acc["out"] / d,
            "max_step_rate": acc["max_step"] / d, "num_agents_total": acc["done"], "agent_steps": acc["steps"],
            "step_reward_mean": acc["reward"] / acc["steps"], "cost_step_mean": acc["cost"] / acc["steps"],
            "num_neighbours_step_mean": acc["nei"] / acc["steps"], "velocity_step_mean_episode_mean": np.mean(vels),
            "velocity_step_mean_episode_min": np.min(vels), "velocity_step_mean_episode_max": np.max(vels),
            "num_agents_success_per_300_steps": acc["success"] / S / H * 300.0}
    assert acc["done"] > 10 and acc["steps"] > S * 20 * H // 2
    for k, w in want.items():
        assert abs(res[k] - w) <= 1e-5 * max(1.0, abs(w)), (k, res[k], w)


def test_curriculum_changes_the_population_by_slot_masking():
    from copo_b200 import trainer as T
    from copo_b200.curriculum import ChangeNCallback
    tr = T.IPPOTrainer(dict(env="MultiAgentIntersectionEnv", num_scenes=8, rollout_fragment_length=10,
                            sgd_minibatch_size=1024, num_sgd_iter=1, env_config={"num_agents": 20}, seed=0))
    cb = ChangeNCallback(total_time_step=8 * 10 * 8, target_num_agents=20)
    pops = []
    for it in range(8):
        res = tr.train()
        cb.on_train_result(tr, res)
        valid = (tr.ro["flags"] & 1) > 0
        pops.append(int(valid.view(10, 8, 20).sum(2).max()))
    assert [n for _, n in cb.history] == [5, 10, 15, 20]
    assert pops[1] <= 5 and pops[3] <= 10 and pops[-1] > 10      # the next fragment runs with the new population
    tr.stop()
