#!/usr/bin/env python
"""Headline benchmark: agent-env-steps/s of the batched multi-agent driving step (BASELINE.json metric) on
Intersection, 40 agents x 4096 scenes per GPU (configs[1]); scenes shard across GPUs without a data-path
collective (weak scaling).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...            the CPU restatement (oracle/) on every host core

One JSON line on stdout (rank 0).  Timing: CUDA events on the launching stream around every step; L2 is flushed
(256 MiB write) between steps, outside the timed events; max over ranks.  `e2e` goes through the host-buffer API
(pinned host actions -> device, step, every output -> pinned host) with the copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent_env_steps_per_s"
UNIT = "agent-env-steps/s"
MAP, SLOTS, SCENES_PER_GPU = "intersection", 40, 4096


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
def _oracle_worker(args):
    scenes, steps, warmup, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"          # the reference sets this for its workers (utils/utils.py:183)
    import numpy as np
    from copo_b200.maps import build_map
    from oracle import sim as osim
    cfg = osim.SimConfig(seed=seed)
    cfg.num_agents = SLOTS
    sim = osim.OracleSim(build_map(MAP), scenes, SLOTS, cfg, scene_offset=seed * 1000)
    sim.reset()
    rng = np.random.default_rng(seed)
    acts = [rng.uniform(-1, 1, (scenes, SLOTS, 2)).astype(np.float32) for _ in range(4)]   # recoder.py:385
    for t in range(warmup):
        sim.step(acts[t % 4])
    n0 = int(sim.agent_steps.sum())
    t0 = time.perf_counter()
    for t in range(steps):
        sim.step(acts[t % 4])
    dt = time.perf_counter() - t0
    return int(sim.agent_steps.sum()) - n0, dt


def cpu_oracle_throughput(procs, scenes_per_proc, steps, warmup):
    """Runs `procs` oracle processes side by side; returns (agent-steps/s aggregate, seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if procs == 1:
        res = [_oracle_worker((scenes_per_proc, steps, warmup, 0))]
    else:
        with ctx.Pool(procs) as pool:
            res = pool.map(_oracle_worker, [(scenes_per_proc, steps, warmup, k) for k in range(procs)])
    wall = time.perf_counter() - t0
    total = sum(n for n, _ in res)
    slowest = max(dt for _, dt in res)
    return total / slowest, slowest, wall, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64))
    scenes = 16
    # keep the whole run to a couple of minutes: ~80 ms per step of 16 scenes x 40 agents per process
    steps = max(1, min(args.steps, 200))
    warmup = max(1, min(args.warmup, 20))
    value, slowest, wall, total = cpu_oracle_throughput(procs, scenes, steps, warmup)
    sample = "%d processes x %d scenes x %d agents x %d steps of the numpy oracle (U(-1,1)^2 actions)" % (
        procs, scenes, SLOTS, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": slowest / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CoPO Intersection 40 agents: env step + neighbour/LCF bookkeeping, CPU restatement "
                               "(oracle/sim.py; MetaDrive itself is not installable here)", "map": MAP,
                   "agents_per_scene": SLOTS, "scenes": procs * scenes},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
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
    from copo_b200.batched_env import BatchedDrivingEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, A = args.scenes, SLOTS
    env = BatchedDrivingEnv(MAP, num_scenes=S, num_slots=A, num_agents=A, seed=args.seed, scene_offset=rank * S,
                            device=dev)
    D = env.D
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = [torch.rand((S, A, 2), device=dev, generator=gen) * 2 - 1 for _ in range(8)]     # recoder.py:385
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(args.warmup):
        env.step(acts[t % 8])
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = env.agent_steps()
    barrier()
    launches = 0
    for t in range(args.steps):
        flush.fill_(t & 0xff)                        # L2 flush, outside the timed events
        ev[t][0].record(stream)
        env.step(acts[t % 8])
        launches += 1
        ev[t][1].record(stream)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    n1 = env.agent_steps()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    agent_steps = n1 - n0

    # ---- end to end through the host-buffer API ------------------------------------------------------------
    host_acts = [a.cpu().pin_memory() for a in acts[:4]]
    for t in range(3):
        env.step_host(host_acts[t % 4])
    e2e_steps = max(3, min(args.steps, 50))
    m0 = env.agent_steps()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(e2e_steps):
        env.step_host(host_acts[t % 4])
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_agent_steps = env.agent_steps() - m0

    stats = torch.tensor([total_ms, e2e_ms, float(agent_steps), float(e2e_agent_steps)], dtype=torch.float64,
                         device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = float(mx[0]), float(mx[1])
        agent_steps, e2e_agent_steps = float(sm[2]), float(sm[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = agent_steps / (total_ms * 1e-3)
    e2e_value = e2e_agent_steps / (e2e_ms * 1e-3)
    peak, peak_src = _peaks()
    algo_bytes = S * A * (4 * D + 153)               # SURVEY.md 8d / DESIGN.md: per agent slot 4*D + 153 bytes
    kernel_ms = float(np.mean(step_ms))              # one launch per step: the step time IS the kernel time
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "env_step_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, slowest, wall, total = cpu_oracle_throughput(1, 16, 150, 3)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "numpy oracle, 16 scenes x 40 agents x 150 steps on one core (%.1f s)" % slowest}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CoPO Intersection 40 agents x %d scenes per GPU: fused scene step (dynamics, "
                               "crash/out/arrive, respawn, neighbours + nei/global reward, 72-laser lidar obs + LCF)"
                               % S, "map": MAP, "agents_per_scene": A, "scenes_per_gpu": S, "obs_dim": D,
                   "actions": "iid U(-1,1)^2 (reference FPS harness, eval/recoder.py:385)",
                   "l2": "flushed between steps with a 256 MiB write, outside the timed events",
                   "slots_per_step": S * A * world, "counted": "agents that received an action (valid slots)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "env_step_kernel",
                     "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": env.h2d_bytes_per_step * world,
                "d2h_bytes_per_step": env.d2h_bytes_per_step * world, "steps": e2e_steps},
        "gpu_launches": launches,
        "clocks": clk,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenes", type=int, default=SCENES_PER_GPU, help="scenes per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
