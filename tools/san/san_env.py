import sys, torch
sys.path.insert(0, ".")
from copo_b200.batched_env import BatchedDrivingEnv
for name, S, A in (("intersection", 7, 40), ("parking_lot", 9, 10), ("tollgate", 4, 40)):
    env = BatchedDrivingEnv(name, num_scenes=S, num_slots=A, num_agents=A, seed=1)
    env.reset()
    out = dict(env.out); out["obs_split"] = env.alloc_obs_split()
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(6):
        a = torch.rand((S, A, 2), device="cuda", generator=g) * 2 - 1
        env.step(a, out=out)
    torch.cuda.synchronize()
    print(name, "ok", float(out["obs"].sum()))
