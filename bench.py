#!/usr/bin/env python
"""Headline benchmark: agent-env-steps/s of the batched multi-agent driving step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--config c2|c3|c4|c5]   this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N ... [--config ...]           the CPU restatement (oracle/) on all host cores

Workloads (`--config`, per-GPU shapes of BASELINE.json `configs[1..4]`; scenes shard across GPUs without a data-path
collective, weak scaling):
  c2 (default)  CoPO Intersection, 40 agents x 4096 scenes              obs 92
  c3            CCPPO mean-field Roundabout, 40 agents x 4096 scenes     obs 91, critic obs 184 (fuse + value head in the step)
  c4            CoPO Tollgate, 40 agents x 1024 scenes (8192 over 8)     obs 157
  c5            CoPO Parking Lot, 10 agents x 4096 scenes (32768 over 8) obs 92 (small-observation path, fused kernel)

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream around every step; L2 is flushed
(256 MiB write) between steps, outside the timed events; max over ranks.  `e2e` goes through the host-buffer API
(pinned host actions -> device, step, every output -> pinned host) with the copies inside the timed region.
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent_env_steps_per_s"
UNIT = "agent-env-steps/s"

CONFIGS = {
    "c2": dict(algo="copo", map="intersection", slots=40, scenes=4096,
               label="CoPO Intersection 40 agents x %d scenes per GPU"),
    "c3": dict(algo="ccppo", map="roundabout", slots=40, scenes=4096,
               label="CCPPO (mean-field) Roundabout 40 agents x %d scenes per GPU"),
    "c4": dict(algo="copo", map="tollgate", slots=40, scenes=1024,
               label="CoPO Tollgate 40 agents x %d scenes per GPU (8192 scenes over 8 GPUs)"),
    "c5": dict(algo="copo", map="parking_lot", slots=10, scenes=4096,
               label="CoPO Parking Lot 10 agents x %d scenes per GPU (32768 scenes over 8 GPUs)"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json, burst)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def _kernel_source_hash():
    """Identifies the scene-step kernel build a committed ncu capture belongs to: a hash of the two sources without their
    comments and blank space (a comment edit compiles to the same SASS and keeps the capture valid)."""
    h = hashlib.sha1()
    for f in ("env_step.cu", "sim_core.cuh"):
        txt = open(os.path.join(ROOT, "copo_b200", "csrc", f)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"//[^\n]*", "", txt)
        h.update(" ".join(txt.split()).encode())
    return h.hexdigest()[:12]


def _ncu_traffic(config):
    """DRAM bytes per scene step from a committed `ncu --set full` capture (profiles/env_step_traffic.json), only when
    that capture was taken on this very kernel source and workload; else None."""
    p = os.path.join(ROOT, "profiles", "env_step_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    e = d.get(config)
    if not isinstance(e, dict) or e.get("kernel_source_hash") != _kernel_source_hash():
        return None, None
    return e.get("dram_bytes_per_launch"), e.get("source")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference; the simulator part is this repo's own spec - MetaDrive is not here)
# ---------------------------------------------------------------------------------------------------------
def _normc_layers(rng, dims, last_std=0.01):
    import numpy as np
    layers = []
    for k in range(len(dims) - 1):
        i, o = dims[k], dims[k + 1]
        std = last_std if k == len(dims) - 2 else 1.0
        w = rng.normal(size=(o, i)).astype(np.float32)
        w *= std / np.sqrt((w ** 2).sum(1, keepdims=True))
        layers.append((np.ascontiguousarray(w.T), np.zeros(o, np.float32)))
    return layers


def _mlp(layers, x):
    import numpy as np
    for n, (W, b) in enumerate(layers):
        x = x @ W + b
        if n < len(layers) - 1:
            x = np.tanh(x)
    return x


def _oracle_worker(args):
    config, scenes, steps, warmup, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"          # the reference sets this for its workers (utils/utils.py:183)
    import numpy as np
    from copo_b200.maps import build_map
    from oracle import sim as osim
    c = CONFIGS[config]
    A = c["slots"]
    cfg = osim.SimConfig(seed=seed, append_lcf=(c["algo"] == "copo"))
    cfg.num_agents = A
    sim = osim.OracleSim(build_map(c["map"]), scenes, A, cfg, scene_offset=seed * 1000)
    out = sim.reset()
    rng = np.random.default_rng(seed)
    # the same rollout step as the GPU arm: policy MLP forward (D-256-256-4, normc init) + Gaussian sample + env step
    # (+ for CCPPO mean-field: critic-obs fusion, algo_ccppo.py:266-311, and the central value head)
    D = out["obs"].shape[-1]
    policy = _normc_layers(rng, (D, 256, 256, 4))
    value = _normc_layers(rng, (2 * D + 2, 256, 256, 1)) if c["algo"] == "ccppo" else None
    bits = (np.uint64(1) << np.arange(A, dtype=np.uint64))

    def act(obs):
        x = _mlp(policy, obs.reshape(-1, D))
        mean, log_std = x[:, :2], x[:, 2:]
        a = mean + np.exp(log_std) * rng.standard_normal(mean.shape).astype(np.float32)
        return a.reshape(scenes, A, 2).astype(np.float32)

    def critic(obs, a, o):
        m = ((o["mf_mask"][..., None] & bits) != 0) & ((o["flags"] & 1) != 0)[:, None, :]       # [S, A, A]
        n = np.maximum(m.sum(-1, keepdims=True), 1).astype(np.float32)
        mf = m.astype(np.float32)
        cobs = np.concatenate([obs, mf @ obs / n, mf @ a / n], -1)
        return _mlp(value, cobs.reshape(-1, 2 * D + 2))

    def step(obs):
        a = act(obs)
        o = sim.step(a)
        if value is not None:
            critic(obs, a, o)
        return o

    for t in range(warmup):
        out = step(out["obs"])
    n0 = int(sim.agent_steps.sum())
    t0 = time.perf_counter()
    for t in range(steps):
        out = step(out["obs"])
    dt = time.perf_counter() - t0
    return int(sim.agent_steps.sum()) - n0, dt


def cpu_oracle_throughput(config, procs, scenes_per_proc, steps, warmup):
    """Runs `procs` oracle processes side by side; returns (agent-steps/s aggregate, seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if procs == 1:
        res = [_oracle_worker((config, scenes_per_proc, steps, warmup, 0))]
    else:
        with ctx.Pool(procs) as pool:
            res = pool.map(_oracle_worker, [(config, scenes_per_proc, steps, warmup, k) for k in range(procs)])
    wall = time.perf_counter() - t0
    total = sum(n for n, _ in res)
    slowest = max(dt for _, dt in res)
    return total / slowest, slowest, wall, total


def _host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _real_reference_probe():
    """is a real MetaDrive + the reference's wrappers importable (from a driver-provided
    baseline/_ref or the environment)?  Returns (env factory, why-not)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import metadrive                                                     # noqa: F401
        from metadrive.envs.marl_envs import MultiAgentIntersectionEnv       # noqa: F401
        from copo.torch_copo.utils.env_wrappers import get_lcf_env           # noqa: F401
    except Exception as e:                                                   # ImportError and anything its import raises
        return None, "%s: %s" % (type(e).__name__, e)
    return (lambda cls_name, n: get_lcf_env(getattr(__import__("metadrive.envs.marl_envs", fromlist=[cls_name]),
                                                    cls_name))({"num_agents": n})), None


def _real_reference_worker(args):
    """The reference's own FPS loop (eval/recoder.py:379-404): U(-1,1)^2 actions for every live vehicle."""
    cls_name, n_agents, steps, warmup, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    make, why = _real_reference_probe()
    env = make(cls_name, n_agents)
    env.reset()
    rng = np.random.default_rng(seed)
    total, t0 = 0, None
    for t in range(warmup + steps):
        if t == warmup:
            t0, total = time.perf_counter(), 0
        o, r, d, i = env.step({k: rng.uniform(-1, 1, 2) for k in env.vehicles.keys()})
        total += len(r)
        if d["__all__"]:
            env.reset()
    dt = time.perf_counter() - t0
    env.close()
    return total, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    cores = _host_cores()
    procs = max(1, min(cores, 64))
    scenes = 16 if c["slots"] >= 20 else 64
    # keep the whole run to a couple of minutes: ~80 ms per step of 16 scenes x 40 agents per process
    steps = max(1, min(args.steps, 200))
    warmup = max(1, min(args.warmup, 20))
    make, why_not = _real_reference_probe()
    if make is not None:
        import multiprocessing as mp
        cls = {"intersection": "MultiAgentIntersectionEnv", "roundabout": "MultiAgentRoundaboutEnv",
               "tollgate": "MultiAgentTollgateEnv", "parking_lot": "MultiAgentParkingLotEnv"}[c["map"]]
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_real_reference_worker, [(cls, c["slots"], steps, warmup, k) for k in range(procs)])
        total, slowest = sum(n for n, _ in res), max(dt for _, dt in res)
        value, kind = total / slowest, "reference"
        sample = "%d processes x 1 MetaDrive scene x %d agents x %d env steps (the reference's own FPS loop)" % (
            procs, c["slots"], steps)
        scenes = 1
    else:
        value, slowest, wall, total = cpu_oracle_throughput(args.config, procs, scenes, steps, warmup)
        kind = "port"
        sample = "%d processes x %d scenes x %d agents x %d rollout steps of the numpy oracle (policy forward + env step%s)" % (
            procs, scenes, c["slots"], steps, " + mean-field critic-obs fusion + value head" if c["algo"] == "ccppo" else "")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": slowest / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (c["label"] % scenes).replace("per GPU", "per process") +
                               ", one rollout step: policy MLP forward + Gaussian sample + env step with neighbour/LCF "
                               "bookkeeping, CPU restatement (oracle/; MetaDrive + RLlib are not installable here)",
                   "name": args.config, "map": c["map"], "agents_per_scene": c["slots"], "scenes": procs * scenes,
                   "real_reference": "not importable (%s)" % why_not if why_not else "MetaDrive + reference wrappers"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from copo_b200 import _lib, ops
    from copo_b200 import policy as P
    from copo_b200.batched_env import BatchedDrivingEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG stays as the caller set it; NCCL logs to file descriptor 1, which main() has pointed at stderr
        dist.init_process_group("nccl", device_id=dev)
    c = CONFIGS[args.config]
    S, A = (args.scenes or c["scenes"]), c["slots"]
    N = S * A
    copo = c["algo"] == "copo"
    env = BatchedDrivingEnv(c["map"], num_scenes=S, num_slots=A, num_agents=A, seed=args.seed, scene_offset=rank * S,
                            append_lcf=copo, device=dev)
    D = env.D
    if copo:
        pol = P.CoPOPolicy(D, 2, P.copo_config(seed=args.seed), device=dev, dist=dist if world > 1 else None)
    else:
        pol = P.CCPPOPolicy(D, 2, P.ccppo_config(seed=args.seed), device=dev, dist=dist if world > 1 else None)
    # rollout ring: the env writes observations / rewards / flags of step t straight into slot t % RING
    RING = 4
    obs = torch.zeros((RING + 1, N, D), device=dev)
    acts = torch.zeros((RING, N, 2), device=dev)
    outs = []
    for r in range(RING):
        o = env.alloc_outputs()
        o["obs"] = obs[r + 1].view(S, A, D)
        o["nei_list"] = None                 # optional outputs, computed only on request: the nearest-neighbour list
        if copo:                             # feeds CCPPO's concat fusion, the mean-field mask its mean-field fusion
            o["mf_mask"] = None
        outs.append(o)
    split = [env.alloc_obs_split() for _ in range(2)]      # the policy's [hi | lo] operand, written by the env
    first = dict(env.out)
    first["obs_split"] = split[0]
    env.reset(out=first)
    obs[0].copy_(env.out["obs"].reshape(N, D))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    state = {"t": 0}

    def critic_step(src, a, o):
        """CCPPO mean-field (C3): critic-obs fusion of this step's rows + the central value head on them"""
        cobs, cobs_split = ops.cc_obs_fuse(src, a, o["flags"].view(-1), o["mf_mask"].view(-1), None, A, "mf", True,
                                           want_split=True)
        return pol.model.central_value_function(cobs, cobs_split)

    def rollout_step():
        """policy forward (one tcgen05 kernel; logits + Gaussian sample in its epilogue) -> fused scene
        step (which also emits the next observation as the policy's bf16 operand); everything stays in HBM"""
        t = state["t"]
        r = t % RING
        src = obs[r] if r or t == 0 else obs[RING]
        lg, actions, logp = pol.model.forward_sample(src, args.seed + rank * 7919, t, obs_split=split[t % 2].view(N, -1),
                                                     actions=acts[r])
        outs[r]["obs_split"] = split[(t + 1) % 2]
        env.step(actions.view(S, A, 2), out=outs[r])
        if not copo:
            critic_step(src, actions, outs[r])
        state["t"] = t + 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, RING)):
        rollout_step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = env.agent_steps()
    l0 = _lib.LAUNCHES
    barrier()
    for t in range(args.steps):
        flush.fill_(t & 0xff)                        # L2 flush, outside the timed events
        ev[t][0].record(stream)
        rollout_step()
        ev[t][1].record(stream)
    barrier()
    launches = _lib.LAUNCHES - l0
    clk = clocks.stop() if rank == 0 else None
    n1 = env.agent_steps()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    agent_steps = n1 - n0
    # the same steps back to back without the flush (secondary figure)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    b0.record(stream)
    for t in range(args.steps):
        rollout_step()
    b1.record(stream)
    barrier()
    noflush_ms = b0.elapsed_time(b1) / args.steps

    # ---- per-kernel timing (same inputs, L2 flushed before each) ------------------------------------------------
    def time_kernel(fn, reps=20):
        ms = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    act_buf = torch.rand((S, A, 2), device=dev) * 2 - 1
    outs[0]['obs_split'] = split[0]
    env_ms = time_kernel(lambda: env.step(act_buf, out=outs[0]))
    lidar_ms = time_kernel(lambda: env.relaunch_lidar(outs[0])) if env.kernels_per_step == 2 else None
    net = pol.model.nets["policy"]
    w1, w2, _ = net.tc_weights(pol.model.weights_version)
    a1 = ops.tc_split_rows(obs[0])
    _, s1 = ops.tc_linear(a1, w1, net.b[0], act=1, want_f32=False, want_split=True)
    l2_ms = time_kernel(lambda: ops.tc_linear_head(s1, w2, net.b[1], net.W[2], net.b[2], act=1, sample=(1, 1)))
    l1_ms = time_kernel(lambda: ops.tc_linear(a1, w1, net.b[0], act=1, want_f32=False, out_split=s1, want_split=True))
    # the rollout's policy forward: ONE kernel (mlp_fused.cu), hidden layers on chip
    mlp_ms = time_kernel(lambda: ops.tc_mlp2_head(a1, w1, net.b[0], w2, net.b[1], net.W[2], net.b[2], sample=(1, 1)))
    fuse_ms = None
    if not copo:
        fuse_ms = time_kernel(lambda: ops.cc_obs_fuse(obs[0], acts[0], outs[0]["flags"].view(-1),
                                                      outs[0]["mf_mask"].view(-1), None, A, "mf", True))

    # ---- end to end, host in the loop: pinned host actions -> device, scene step, every env output -> pinned host,
    # policy forward + sample on the new observations, sampled actions -> pinned host (they are the next step's input)
    host_act = torch.zeros((S, A, 2)).pin_memory()

    def e2e_step(t):
        # the scene step's outputs start crossing PCIe (one copy of the output arena, on the env's copy stream) as soon
        # as its kernels end; the policy forward on the new observations is queued behind the scene step meanwhile
        env.step_host(host_act, obs_split=split[0], wait=False)
        o = env.host_step_out
        src = o["obs"].view(N, D)
        lg, a, lp = pol.model.forward_sample(src, args.seed + rank * 7919, 100000 + t, obs_split=split[0].view(N, -1))
        if not copo:
            critic_step(src, a, o)
        host_act.copy_(a.view(S, A, 2), non_blocking=True)
        env.wait_host()                                  # every env output is in pinned host memory
        torch.cuda.current_stream().synchronize()        # ... and so are the next step's actions

    for t in range(3):
        e2e_step(t)
    e2e_steps = max(3, min(args.steps, 50))
    m0 = env.agent_steps()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(e2e_steps):
        e2e_step(3 + t)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_agent_steps = env.agent_steps() - m0
    # what the PCIe link alone takes for one step's outputs (the arena copy, nothing else in flight)
    pc = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        env._arena_host.copy_(env._arena_dev, non_blocking=True)
        b.record(stream)
        torch.cuda.synchronize()
        pc.append(a.elapsed_time(b))
    pcie_ms = float(np.median(pc))

    stats = torch.tensor([total_ms, e2e_ms, float(agent_steps), float(e2e_agent_steps)], dtype=torch.float64,
                         device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = float(mx[0]), float(mx[1])
        agent_steps, e2e_agent_steps = float(sm[2]), float(sm[3])

    train = None
    if args.train_iters > 0:
        train = time_training(args, dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = agent_steps / (total_ms * 1e-3)
    e2e_value = e2e_agent_steps / (e2e_ms * 1e-3)
    hbm_peak, tc_peak, peak_src = _peaks()
    # 4*D + 153 bytes per agent slot and step (action 8 + state 64 in + 64 out + obs 4*D + reward 4 +
    # nei reward 4 + flags 1 + mask 8).  The kernel ALSO writes the observation a second time as the policy's bf16
    # [hi | lo] operand (2 * Kp * 2 B per slot): reported separately, not part of the roofline fraction.
    algo_bytes = N * (4 * D + 153)
    operand_bytes = N * 2 * env.split_width
    env_gbs = algo_bytes / (env_ms * 1e-3) / 1e9
    l2_flops = 2.0 * N * 256 * 256                   # algorithmic fp32-equivalent flops of the 256x256 layer
    l2_tf = l2_flops / (l2_ms * 1e-3) / 1e12
    mlp_flops = 2.0 * N * (D * 256 + 256 * 256 + 4 * 256)           # the whole policy network, fp32-equivalent
    mlp_issued = 4 * 2.0 * N * 256 * (env.split_width // 2 + 256)    # bf16 flops the tensor cores execute (4 products)
    mlp_tf = mlp_flops / (mlp_ms * 1e-3) / 1e12
    traffic, traffic_src = _ncu_traffic(args.config)
    env_roof = {"bound": "hbm", "achieved": env_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": env_gbs / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "scene step: env_step_kernel<state> + env_lidar_kernel (two launches, timed together)"
                          if env.kernels_per_step == 2 else "env_step_kernel (fused)",
                "algorithmic_bytes_per_launch": algo_bytes, "bytes_per_slot_step": 4 * D + 153, "kernel_ms": env_ms,
                "with_policy_operand": {"bytes_per_launch": algo_bytes + operand_bytes,
                                        "achieved": (algo_bytes + operand_bytes) / (env_ms * 1e-3) / 1e9,
                                        "frac": (algo_bytes + operand_bytes) / (env_ms * 1e-3) / 1e9 / hbm_peak,
                                        "what": "counts the bf16 [hi | lo] copy of the observation the kernel also writes "
                                                "for the policy's first layer (not in SURVEY 8d's figure)"},
                "note": "issue-bound, not HBM-bound (72-laser lidar + neighbour search are ALU work): see profiles/ for the "
                        "instruction mix and issue-slot utilisation"}
    mlp_roof = {"bound": "tensor", "achieved": mlp_tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": mlp_tf / tc_peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "tc_mlp2_kernel (policy network %d-256-256-4 + sample in one launch, bf16_split)" % D,
                "algorithmic_flops_per_launch": mlp_flops, "kernel_ms": mlp_ms, "tensor_flops_issued": mlp_issued,
                "frac_of_issued_bf16_flops": mlp_issued / (mlp_ms * 1e-3) / 1e12 / tc_peak,
                "note": "achieved counts fp32-equivalent flops; the kernel executes 4x as many bf16 flops on zero-padded "
                        "reduction lengths (hi/lo split operands, four products): frac_of_issued_bf16_flops"}
    layer_roof = {"bound": "tensor", "achieved": l2_tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": l2_tf / tc_peak,
                  "traffic": None, "peak_source": peak_src,
                  "kernel": "tc_linear_kernel (one 256x256 layer + logits + sample; the learner's forward kernel)",
                  "algorithmic_flops_per_launch": l2_flops, "kernel_ms": l2_ms, "tensor_flops_issued": 4 * l2_flops}
    others = [mlp_roof, layer_roof]
    if fuse_ms is not None:
        fb = N * (4 * D + 4 * (2 * D + 2))           # read own obs + write the fused critic obs (1100 B at D = 91)
        others.append({"bound": "hbm", "kernel": "cc_obs_fuse_kernel (mean-field)", "kernel_ms": fuse_ms,
                       "algorithmic_bytes_per_launch": fb, "achieved": fb / (fuse_ms * 1e-3) / 1e9, "peak": hbm_peak,
                       "unit": "GB/s", "frac": fb / (fuse_ms * 1e-3) / 1e9 / hbm_peak})
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = max(1, min(_host_cores(), 64))
        cs = 16 if A >= 20 else 64
        v, slowest, wall, total = cpu_oracle_throughput(args.config, cores, cs, 150, 3)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "numpy oracle, %d processes x %d scenes x %d agents x 150 rollout steps (policy forward + env "
                         "step) on %d cores (%.1f s)" % (cores, cs, A, cores, slowest)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "ms_per_step_back_to_back_no_flush": noflush_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (c["label"] % S) + ", one rollout step: policy MLP forward (%d-256-256-4, one "
                               "tcgen05 split-bf16 kernel, logits + Gaussian sample in its epilogue) + scene step (dynamics, "
                               "crash/out/arrive, respawn, neighbours + nei/global reward, 72-laser lidar obs%s)%s"
                               % (D, " + LCF" if copo else "",
                                  "" if copo else " + mean-field critic-obs fusion + central value head (184-256-256-1)"),
                   "name": args.config, "map": c["map"], "agents_per_scene": A, "scenes_per_gpu": S, "obs_dim": D,
                   "actions": "sampled from the randomly initialised policy (normc init, seed %d)" % args.seed,
                   "l2": "flushed between steps with a 256 MiB write, outside the timed events",
                   "slots_per_step": N * world, "counted": "agents that received an action (valid slots)"},
        # `roofline`: the scene step (state + lidar kernels) - the HBM-class kernel the north star names, by SURVEY
        # 8d's bytes.  `roofline_other`: the other heavy kernels of the step.
        "roofline": env_roof,
        "roofline_other": others,
        "kernel_ms": {"env_step": env_ms, "env_lidar_kernel": lidar_ms,
                      "env_state_kernel": (env_ms - lidar_ms) if lidar_ms is not None else None,
                      "tc_mlp2_policy_forward_and_sample": mlp_ms,
                      "tc_linear_layer1": l1_ms, "tc_linear_layer2_with_logits_and_sample": l2_ms,
                      "cc_obs_fuse_mf": fuse_ms},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step * world,
                "d2h_bytes_per_step": (env.d2h_bytes_per_step + env.h2d_bytes_per_step) * world, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "output_copy_alone_ms": pcie_ms,
                "output_copy_GBps": env.d2h_bytes_per_step / (pcie_ms * 1e-3) / 1e9,
                "api": "host-in-the-loop rollout step: BatchedDrivingEnv.step_host (pinned host actions in, every env "
                       "output back to pinned host as one arena copy on a copy stream) + forward_sample on "
                       "the device meanwhile, sampled actions back to pinned host; the step ends when both have "
                       "landed"},
        "gpu_launches": launches,
        "clocks": clk,
        "train": train,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_training(args, dev, world, rank):
    """Whole training iterations (rollout + postprocess + SGD epochs + meta update, gradient all-reduce per minibatch)
    at the configuration's per-GPU scene count on a short fragment."""
    import torch
    from copo_b200 import trainer as T
    from copo_b200.batched_env import MAP_OF_ENV
    c = CONFIGS[args.config]
    env_name = {v: k for k, v in MAP_OF_ENV.items()}[c["map"]]
    scenes = args.train_scenes or min(c["scenes"], 1024)
    cls = T.CoPOTrainer if c["algo"] == "copo" else T.CCPPOTrainer
    # weak scaling: the scenes AND the SGD minibatch per GPU are fixed (sgd_minibatch_size is the GLOBAL minibatch, so
    # it grows with the world size: 65 536 rows per rank and gradient all-reduce)
    mb = 65536 * world
    tr = cls(dict(env=env_name, num_scenes=scenes, rollout_fragment_length=args.train_fragment,
                  sgd_minibatch_size=mb, num_sgd_iter=5, lcf_num_iters=5, env_config={"num_agents": c["slots"]},
                  seed=args.seed), device=dev)
    # pre-warm the caching allocator: the valid-row count (and with it the ragged last minibatch) changes from iteration
    # to iteration, and a request that fits no cached block is a cudaMalloc in the middle of an iteration (1-2 per
    # iteration observed, sporadically 35-100 ms each on these boxes); one 4 GiB segment, freed again, serves them all
    del_me = torch.empty(4 << 30, dtype=torch.uint8, device=dev)
    del del_me
    for _ in range(2):        # warm-up: the first iteration captures the minibatch graph, the second is the first with a
        tr.train()            # ragged last minibatch (eager path: its buffers come from the caching allocator once)
    import gc
    gc.collect()              # a full collection of this long-lived process now, and everything alive so far out of the
    gc.freeze()               # collector's sight: a gen-2 pass over the bench's heap cost 20-60 ms inside a timed iteration
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 0
    sample_ms, learn_ms, ar_ms, iter_ms, dev_allocs = [], [], [], [], []
    for _ in range(args.train_iters):
        ti = time.perf_counter()
        n_alloc = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        res = tr.train()
        iter_ms.append((time.perf_counter() - ti) * 1e3)
        dev_allocs.append(torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - n_alloc)
        steps += res["custom_metrics"]["agent_steps"]
        sample_ms.append(tr._timers["sample_time_ms"])
        learn_ms.append(tr._timers["learn_time_ms"])
        ar_ms.append(tr._timers.get("allreduce_ms"))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gc.unfreeze()
    out = {"agent_env_steps_per_s_per_gpu": steps / dt, "iterations": args.train_iters, "seconds": dt,
           "scenes_per_gpu": scenes, "fragment": args.train_fragment, "sgd_minibatch_size": mb,
           "sgd_minibatch_rows_per_gpu": 65536,
           "num_sgd_iter": 5, "lcf_num_iters": 5, "iteration_ms": iter_ms, "cuda_mallocs_per_iteration": dev_allocs, "sample_ms": sample_ms, "learn_ms": learn_ms,
           "allreduce_ms_per_iteration": ar_ms, "allreduces_per_iteration": tr._timers.get("allreduces"),
           "what": "full %s.training_step iterations, wall clock, this rank; allreduce_ms = CUDA-event time inside "
                   "the gradient / statistics all-reduces of one iteration" % cls.__name__}
    tr.stop()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scenes", type=int, default=0, help="scenes per GPU (default: the configuration's)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-iters", type=int, default=3, help="timed full training iterations (0: skip)")
    ap.add_argument("--train-scenes", type=int, default=0)
    ap.add_argument("--train-fragment", type=int, default=16)
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries write to file descriptor 1 while the bench runs (NCCL's
    # version banner / debug log when NCCL_DEBUG is set, torch warnings) is sent to stderr instead; the JSON line goes
    # to the original stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
