"""Host-link ceiling of the end-to-end (host buffers) rollout step: every rank copies a 64 MiB device buffer to its own
pinned host buffer (and back), first all ranks at once, then one rank at a time; prints one JSON line from rank 0.
Run under torchrun (one process per GPU).  The e2e step moves 66 MB per GPU and step over this link."""
import json, os, sys
import torch, torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = 64 << 20
d = torch.empty(NB, dtype=torch.uint8, device=dev)
h = torch.empty(NB, dtype=torch.uint8).pin_memory()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return NB * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def gather(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x) for x in out]
    return [v]


res = {"n_gpus": world, "buffer_bytes": NB}
res["d2h_all_ranks_GBps"] = gather(timed(lambda: h.copy_(d, non_blocking=True)))
res["h2d_all_ranks_GBps"] = gather(timed(lambda: d.copy_(h, non_blocking=True)))
alone = []
for r in range(world):
    sync()
    v = timed(lambda: h.copy_(d, non_blocking=True)) if r == rank else 0.0
    sync()
    alone.append(v)
res["d2h_one_rank_at_a_time_GBps"] = [max(x) for x in zip(*[gather(a) for a in alone])] if world > 1 else alone
res["d2h_aggregate_GBps"] = sum(res["d2h_all_ranks_GBps"])
try:
    res["cpu_affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
    nodes = [n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")]
    res["numa_nodes"] = len(nodes)
except Exception as e:
    res["numa_nodes"] = str(e)
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
