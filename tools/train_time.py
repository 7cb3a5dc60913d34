"""Per-iteration timers of CoPOTrainer at the bench's training shape (1024 scenes x 40 agents x T rows per GPU)."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from copo_b200.trainer import CoPOTrainer
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
tr = CoPOTrainer(dict(env="MultiAgentIntersectionEnv", num_scenes=1024, rollout_fragment_length=T, sgd_minibatch_size=65536,
                      num_sgd_iter=5, lcf_num_iters=5, env_config={"num_agents": 40}, seed=0))
out = []
for it in range(iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = tr.train()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out.append(dict(it=it, wall_ms=round(dt * 1e3, 1), sample_ms=round(tr._timers["sample_time_ms"], 1),
                    learn_ms=round(tr._timers["learn_time_ms"], 1), agent_steps=res["custom_metrics"]["agent_steps"],
                    Msteps_per_s=round(res["custom_metrics"]["agent_steps"] / dt / 1e6, 2)))
    print(json.dumps(out[-1]), flush=True)
tr.stop()
