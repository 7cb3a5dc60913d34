"""Times the scene step alone (CUDA events) at a given batch: without and with the policy-operand output, and the
lidar kernel alone (two-kernel mode); prints agent-steps/s and algorithmic HBM GB/s."""
import sys, json
import torch
sys.path.insert(0, ".")
from copo_b200.batched_env import BatchedDrivingEnv


def timed(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(steps):
        fn(t)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main(S=4096, A=40, name="intersection", steps=200):
    env = BatchedDrivingEnv(name, num_scenes=S, num_slots=A, num_agents=A, seed=0)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.rand((S, A, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
    for a in acts: a[..., 0] *= 0.2
    for t in range(50): env.step(acts[t % 8])
    ms = timed(lambda t: env.step(acts[t % 8]), steps)
    out2 = dict(env.out)
    out2["obs_split"] = env.alloc_obs_split()
    for t in range(10): env.step(acts[t % 8], out=out2)
    ms_split = timed(lambda t: env.step(acts[t % 8], out=out2), steps)
    ms_lidar = ms_lidar_split = None
    if env.kernels_per_step > 1:
        ms_lidar = timed(lambda t: env.relaunch_lidar(), steps)
        ms_lidar_split = timed(lambda t: env.relaunch_lidar(out2), steps)
    nbytes = S * A * (4 * env.D + 153)
    print(json.dumps(dict(map=name, S=S, A=A, D=env.D, ms_per_step=ms, agent_steps_per_s=S * A / ms * 1e3,
                          algo_GBps=nbytes / ms / 1e6, ms_with_operand=ms_split, ms_lidar=ms_lidar,
                          ms_lidar_with_operand=ms_lidar_split)))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 40,
         sys.argv[3] if len(sys.argv) > 3 else "intersection")
