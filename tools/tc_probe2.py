"""tc_linear: fixed cost vs per-tile cost (three row counts), with the epilogue reduced to its tcgen05.ld
(B2C_TC_PROBE=1, set by the caller) or complete."""
import json, os, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=15):
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


out = {"probe": os.environ.get("B2C_TC_PROBE", "0")}
for K in (64, 92, 256):
    for M in (81920, 163840, 327680):
        x = torch.rand(M, K, device=dev)
        W = torch.randn(256, K, device=dev) / K ** 0.5
        b = torch.zeros(256, device=dev)
        a = ops.tc_split_rows(x)
        w = ops.tc_prep_weight(W)
        sp = torch.empty((M, 512), dtype=torch.bfloat16, device=dev)
        hw, hb = torch.randn(4, 256, device=dev) * 0.01, torch.zeros(4, device=dev)
        r = {}
        r["act0,head4"] = timed(lambda: ops.tc_linear_head(a, w, b, hw, hb, act=0))
        r["tanh,head4+sample"] = timed(lambda: ops.tc_linear_head(a, w, b, hw, hb, act=1, sample=(1, 1)))
        r["tanh,split"] = timed(lambda: ops.tc_linear(a, w, b, act=1, want_f32=False, out_split=sp, want_split=True))
        out["K=%d,M=%d" % (K, M)] = r
print(json.dumps(out, indent=1))
