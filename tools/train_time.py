"""Per-iteration timers of CoPOTrainer at the bench's training shape (1024 scenes x 40 agents x T rows per GPU): wall
time of an iteration and of its phases (a device synchronisation on both sides of each)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from copo_b200.trainer import CoPOTrainer
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tr = CoPOTrainer(dict(env="MultiAgentIntersectionEnv", num_scenes=1024, rollout_fragment_length=T, sgd_minibatch_size=65536,
                      num_sgd_iter=5, lcf_num_iters=5, env_config={"num_agents": 40}, seed=0))
phase = {}


def timed(obj, name, label):
    fn = getattr(obj, name)

    def wrapper(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); phase[label] = phase.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    setattr(obj, name, wrapper)


timed(tr, "sample", "sample")
timed(tr.policy, "postprocess_rollout", "postprocess")
timed(tr.policy, "standardize_advantages", "standardize")
timed(tr, "_learn", "learn")
timed(tr, "_after_sgd", "meta")
timed(tr, "_episode_metrics", "metrics")
for it in range(iters):
    phase.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = tr.train()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    row = dict(it=it, wall_ms=round(dt * 1e3, 1), agent_steps=res["custom_metrics"]["agent_steps"],
               Msteps_per_s=round(res["custom_metrics"]["agent_steps"] / dt / 1e6, 2))
    row.update({k + "_ms": round(v, 1) for k, v in phase.items()})
    print(json.dumps(row), flush=True)
tr.stop()
