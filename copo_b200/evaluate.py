"""Batched evaluation of a policy on the device: the episode statistics the reference's evaluation path reports.

The reference evaluates one MetaDrive env at a time through `RecorderEnv` (eval/recoder.py:73-355) and
`MultiAgentDrivingCallbacks` (torch_copo/utils/callbacks.py:14-148): per-episode success / crash / out-of-road /
max-step rates over finished agents, step-mean velocity / reward / cost / neighbour count, episode reward.  Here S
scenes run one episode each side by side and the same quantities are reduced from the step outputs (SURVEY.md 8f
rank 1).  Keys follow the reference's CSV columns (eval.py:205-232, recoder.py:177-349)."""
import torch

from . import ops
from .batched_env import (BatchedDrivingEnv, FLAG_ARRIVE, FLAG_CRASH, FLAG_DONE, FLAG_MAXSTEP, FLAG_OUT, FLAG_VALID,
                          MAP_OF_ENV)


def evaluate(model, env="MultiAgentIntersectionEnv", num_scenes=64, num_agents=None, horizon=1000, seed=0,
             deterministic=False, lcf_mean=0.0, lcf_std=0.1, neighbours_distance=20.0, device=None, append_lcf=None,
             trace=None):
    """Runs one episode of `horizon` steps in every scene with `model` (CCModel / CoPOModel) acting for all agents.
    `neighbours_distance` is RecorderEnv's evaluation radius (default 20, recoder.py:75).  `append_lcf`: whether the
    policy takes the LCF as its last observation entry (CoPO policies); default: a CoPOModel does, any other model
    does when its input is one wider than the map's base observation.  `trace`: a list that receives every step's
    actions (numpy [S, A, 2]) - tests replay them through the oracle simulator and recompute the report."""
    from .maps import build_map
    from .models import CoPOModel
    map_name = MAP_OF_ENV.get(env, env)
    if append_lcf is None:
        append_lcf = isinstance(model, CoPOModel) or model.obs_dim == build_map(map_name).base_obs_dim + 1
    sim = BatchedDrivingEnv(map_name, num_scenes=num_scenes, num_agents=num_agents, num_slots=num_agents,
                            horizon=horizon, auto_reset=False, append_lcf=append_lcf, seed=seed, lcf_mean=lcf_mean,
                            lcf_std=lcf_std, neighbours_distance=neighbours_distance, device=device or model.device)
    assert sim.D == model.obs_dim, "policy expects %d observations, env gives %d" % (model.obs_dim, sim.D)
    S, A = sim.S, sim.A
    out = sim.reset()
    dev = sim.device
    z = lambda: torch.zeros((), dtype=torch.float64, device=dev)
    acc = dict(done=z(), success=z(), crash=z(), out=z(), max_step=z(), steps=z(), reward=z(), cost=z(), nei=z(),
               ep_reward=z(), ep_len=z(), vel_min=torch.full((), 1e9, dtype=torch.float64, device=dev),
               vel_max=z(), vel_sum=z(), vel_n=z())
    for t in range(horizon):
        obs = out["obs"].reshape(S * A, -1)
        if deterministic:
            actions = model.forward(obs)[:, :2].contiguous()
        else:
            _, actions, _ = model.forward_sample(obs, seed, t)
        if trace is not None:
            trace.append(actions.view(S, A, 2).cpu().numpy().copy())
        out = sim.step(actions.view(S, A, 2))
        f = out["flags"]
        valid = (f & FLAG_VALID) > 0
        done = (f & FLAG_DONE) > 0
        n = valid.sum()
        acc["steps"] += n
        acc["done"] += done.sum()
        acc["success"] += ((f & FLAG_ARRIVE) > 0).sum()
        acc["crash"] += ((f & FLAG_CRASH) > 0).sum()
        acc["out"] += ((f & FLAG_OUT) > 0).sum()
        acc["max_step"] += (((f & FLAG_MAXSTEP) > 0) & done).sum()
        acc["reward"] += (out["reward"] * valid).sum()
        acc["cost"] += (((f & FLAG_CRASH) > 0) & valid).sum()
        bits = out["nei_mask"]
        cnt = torch.zeros_like(bits)
        for j in range(A):
            cnt += (bits >> j) & 1
        acc["nei"] += (cnt * valid).sum()
        # velocity is observation entry 3 (speed / max speed); step mean over active agents (recoder.py:186-199)
        v = (out["obs"][..., 3] * valid).sum() / n.clamp(min=1) * 22.22222137451172 * 3.6
        acc["vel_sum"] += v
        acc["vel_n"] += (n > 0)
        acc["vel_min"] = torch.minimum(acc["vel_min"], torch.where(n > 0, v.double(), acc["vel_min"]))
        acc["vel_max"] = torch.maximum(acc["vel_max"], v.double())
    st = sim.get_state()
    sim.close()
    a = {k: float(v) for k, v in acc.items()}
    d = max(a["done"], 1.0)
    steps = max(a["steps"], 1.0)
    return {
        "success_rate": a["success"] / d, "crash_rate": a["crash"] / d, "out_rate": a["out"] / d,
        "max_step_rate": a["max_step"] / d, "num_agents_total": a["done"], "episode_length": float(horizon),
        "step_reward_mean": a["reward"] / steps, "cost_step_mean": a["cost"] / steps,
        "num_neighbours_step_mean": a["nei"] / steps,
        "velocity_step_mean_episode_mean": a["vel_sum"] / max(a["vel_n"], 1.0),
        "velocity_step_mean_episode_min": a["vel_min"], "velocity_step_mean_episode_max": a["vel_max"],
        "num_agents_success_per_300_steps": a["success"] / S / horizon * 300.0,
        "agent_steps": a["steps"], "num_scenes": S,
    }
