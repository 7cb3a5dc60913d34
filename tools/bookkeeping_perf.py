"""HBM-bound bookkeeping kernels at C2 fragment size: algorithmic GB/s against the measured HBM peak."""
import json, os, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = None

def timeit(fn, n=10):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for _ in range(n + 2):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms[2:])[len(ms[2:]) // 2]

def main(T=64, S=4096, A=40, D=92):
    N = S * A
    R = T * N
    out = {}
    flags = (torch.rand(T, N, device="cuda") < 0.85).to(torch.uint8)
    flags |= ((torch.rand(T, N, device="cuda") < 0.02).to(torch.uint8) * 2) & (flags * 2)
    r = lambda *s: torch.randn(*s, device="cuda")
    rew, val, nrew, nval, gval = r(T, N), r(T, N), r(T, N), r(T, N), r(T, N)
    grew = r(T, S)
    ms = timeit(lambda: ops.gae3(flags, [rew, nrew, grew], [val, nval, gval], 0.99, 0.95, global_reward_per_scene=A))
    b = R * 49
    out["gae3"] = dict(ms=ms, algo_bytes=b, gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK)
    adv, nei, lcf, gadv = r(R), r(R), torch.rand(R, device="cuda") * 2 - 1, r(R)
    f = flags.reshape(-1)
    ms = timeit(lambda: ops.lcf_mix_stats(f, adv, nei, lcf, gadv))
    b = R * 17
    out["lcf_mix_stats"] = dict(ms=ms, algo_bytes=b, gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK)
    ms = timeit(lambda: ops.lcf_mix_apply(f, adv, nei, lcf, gadv, 0.0, 1.0, 0.0, 1.0))
    b = R * 25
    out["lcf_mix_apply"] = dict(ms=ms, algo_bytes=b, gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK)
    T2 = 8
    R2 = T2 * N
    obs, act = torch.rand(R2, 91, device="cuda"), r(R2, 2)
    mf = torch.randint(0, 2 ** 40, (R2,), device="cuda", dtype=torch.int64) & torch.randint(0, 2 ** 40, (R2,), device="cuda", dtype=torch.int64) & torch.randint(0, 2 ** 40, (R2,), device="cuda", dtype=torch.int64)
    fl2 = flags[:T2].reshape(-1).contiguous()
    ms = timeit(lambda: ops.cc_obs_fuse(obs, act, fl2, mf, None, A, "mf", True))
    b = R2 * 1100
    out["cc_obs_fuse_mf"] = dict(ms=ms, algo_bytes=b, gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK)
    idx = torch.randperm(R2, device="cuda")[:65536 * 4].contiguous()
    ms = timeit(lambda: ops.gather_rows(obs, idx))
    b = idx.numel() * 91 * 4 * 2
    out["gather_rows_obs"] = dict(ms=ms, algo_bytes=b, gbs=b / ms / 1e6, frac=b / ms / 1e6 / PEAK)
    n = 360199
    p, g, m, v = r(n), r(n), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ms = timeit(lambda: ops.adam_step(p, g, m, v, 3e-4, 1))
    out["adam_360199"] = dict(ms=ms, note="latency-bound: 1.44 MB of parameters")
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
