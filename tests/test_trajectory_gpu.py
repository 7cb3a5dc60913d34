"""The hook RLlib calls - `postprocess_trajectory(sample_batch, other_agent_batches, episode)` on ONE agent's trajectory
(torch_copo/algo_ccppo.py:322-374, algo_copo.py:473-502) - against the batched `postprocess_rollout` the trainers use,
on the same rollout: a dict-API episode (one scene, agents keyed by name, respawns under fresh names) is recorded twice,
as per-agent SampleBatch-style dicts and as the [T, N] rollout columns of the scene's slots."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rollout(env, pol, T, seed=0):
    """Steps the dict env T times with `pol` acting for every agent.  Returns (per-agent batches, rollout columns)."""
    from copo_b200 import policy as P
    A, D = env.A, env.D
    dev = pol.device
    obs = env.reset()
    rows = {}                                               # agent name -> list of row dicts
    cols = {k: [] for k in ("obs", "actions", "rewards", "flags", "nei_rewards", "global_rewards", "step_lcf", "mf_mask",
                            "nei_list", "action_logp", "action_dist_inputs")}
    slot_rows = []                                          # per step: {agent name: slot}
    for t in range(T):
        # agents that act this step (a terminated agent's last observation is returned once more, without a slot)
        names = sorted((k for k in obs if k in env._slot_of), key=lambda k: env._slot_of[k])
        slots = [env._slot_of[k] for k in names]
        x = torch.as_tensor(np.stack([obs[k] for k in names]), device=dev)
        act, logp, logits = pol.compute_actions(x, step=t)
        act_np = act.cpu().numpy()
        full_obs = np.zeros((A, D), np.float32)
        full_act = np.zeros((A, 2), np.float32)
        full_logp, full_logits = np.zeros(A, np.float32), np.zeros((A, 4), np.float32)
        full_obs[slots], full_act[slots] = x.cpu().numpy(), act_np
        full_logp[slots], full_logits[slots] = logp.cpu().numpy(), logits.cpu().numpy()
        new_obs, rew, done, info = env.step({k: act_np[n] for n, k in enumerate(names)})
        out = env._last_out
        cols["obs"].append(full_obs); cols["actions"].append(full_act)
        cols["action_logp"].append(full_logp); cols["action_dist_inputs"].append(full_logits)
        cols["rewards"].append(out["reward"][0].cpu().numpy().copy())
        cols["flags"].append(out["flags"][0].cpu().numpy().copy())
        cols["nei_rewards"].append(out["nei_reward"][0].cpu().numpy().copy())
        cols["global_rewards"].append(out["global_reward"].cpu().numpy().copy())
        cols["step_lcf"].append(out["lcf"][0].cpu().numpy().copy())
        cols["mf_mask"].append(out["mf_mask"][0].cpu().numpy().copy())
        cols["nei_list"].append(out["nei_list"][0].cpu().numpy().copy())
        slot_rows.append(dict(zip(names, slots)))
        for n, k in enumerate(names):
            if k not in rew or not (int(cols["flags"][-1][slots[n]]) & 1):
                continue
            rows.setdefault(k, []).append(dict(t=t, obs=obs[k], actions=act_np[n], rewards=rew[k], dones=done[k],
                                               infos=info[k], new_obs=new_obs.get(k, obs[k])))
        obs = new_obs
        if done["__all__"]:
            break
    batches = {}
    for k, rs in rows.items():
        batches[k] = {c: (np.stack([r[c] for r in rs]) if c in ("obs", "actions", "new_obs") else
                          [r[c] for r in rs] if c == "infos" else np.asarray([r[c] for r in rs]))
                      for c in ("t", "obs", "actions", "rewards", "dones", "infos", "new_obs")}
    c = lambda k, dt=None: torch.as_tensor(np.stack(cols[k]), device=dev)
    ro = {P.OBS: c("obs"), P.ACTIONS: c("actions"), P.REWARDS: c("rewards"), "flags": c("flags"),
          P.NEI_REWARDS: c("nei_rewards"), P.GLOBAL_REWARDS: c("global_rewards"), "step_lcf": c("step_lcf"),
          "mf_mask": c("mf_mask"), "nei_list": c("nei_list"), "slots": A}
    # observation after the last row, slot-major (IPPO bootstraps cut trajectories with its value)
    nxt = np.zeros((A, D), np.float32)
    for k, v in obs.items():
        nxt[env._slot_now[k]] = v
    ro["next_obs"] = torch.as_tensor(nxt, device=dev)
    return batches, ro, slot_rows


@pytest.mark.parametrize("algo,env_name,fuse", [("copo", "MultiAgentIntersectionEnv", "none"),
                                                ("ccppo", "MultiAgentRoundaboutEnv", "mf"),
                                                ("ccppo", "MultiAgentIntersectionEnv", "concat"),
                                                ("ippo", "MultiAgentIntersectionEnv", "none")])
def test_per_trajectory_postprocess_equals_the_batched_rollout_form(algo, env_name, fuse):
    from copo_b200 import envs as E
    from copo_b200 import policy as P
    base = getattr(E, env_name)
    cls = {"copo": E.get_lcf_env(base), "ccppo": E.get_ccenv(base), "ippo": base}[algo]
    env = cls({"num_agents": 16, "horizon": 90, "delay_done": 3, "start_seed": 3})
    pcls = {"copo": P.CoPOPolicy, "ccppo": P.CCPPOPolicy, "ippo": P.IPPOPolicy}[algo]
    cfg = pcls.default_config()
    cfg.update(fuse_mode=fuse, precision="fp32", seed=2)
    pol = pcls(env.D, 2, cfg)
    for name in pol.model.nets:                             # non-trivial critics: O(1) values
        if name != "policy":
            pol.model.nets[name].W[2].mul_(200.0)
    # a policy that floors the throttle and steers to one side (std 0.3): agents leave the road or crash within a few
    # dozen steps, so slots are respawned under fresh names inside the recorded window
    pol.model.nets["policy"].b[2].copy_(torch.tensor([0.5, 1.0, -1.2, -1.2], device=pol.device))
    batches, ro, slot_rows = _rollout(env, pol, T=70)
    assert len(batches) > 16                                # respawned agents: more names than slots
    ro = pol.postprocess_rollout(ro)
    want = {k: v.cpu().numpy() for k, v in ro.items() if torch.is_tensor(v) and v.dim() == 2}
    keys = ["vf_preds", "advantages", "value_targets"]
    if algo == "copo":
        keys += ["nei_values", "nei_advantage", "nei_target", "global_values", "global_advantages", "global_target"]
    checked, fused_rows = 0, 0
    for name, b in batches.items():
        others = {k: (pol, v) for k, v in batches.items() if k != name}
        out = pol.postprocess_trajectory(dict(b), others, episode=object())
        if algo != "ippo":
            own = out["centralized_critic_obs"]
            assert own.shape[1] == pol.model.cobs_dim
            fused_rows += int((np.abs(own[:, env.D:]).sum(1) > 0).sum())
        for i, t in enumerate(b["t"]):
            s = slot_rows[t][name]
            for k in keys:
                a, w = float(out[k][i]), float(want[k][t, s])
                assert abs(a - w) <= 2e-5 * max(1.0, abs(w)), (name, t, k, a, w)
            if algo != "ippo":
                np.testing.assert_allclose(out["centralized_critic_obs"][i],
                                           ro["centralized_critic_obs"][t, s].cpu().numpy(), rtol=1e-5, atol=1e-6)
            checked += 1
    assert checked > 500
    if fuse != "none":
        assert fused_rows > 50                              # the neighbour fusion did something
    env.close()
