"""Where does a tc_linear launch spend its time?  Times the rollout shapes (M = 163 840 rows; K = 92 and K = 256) with
the epilogue's pieces switched on one by one: split output without / with tanh, fp32 output, both, and the fused head
(4 outputs per row: next to no stores) without / with tanh.  L2 is flushed before every launch."""
import json, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops

M = 163840
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


out = {}
for K in (64, 92, 256):
    x = torch.rand(M, K, device=dev)
    W = torch.randn(256, K, device=dev) / K ** 0.5
    b = torch.zeros(256, device=dev)
    a = ops.tc_split_rows(x)
    w = ops.tc_prep_weight(W)
    f32 = torch.empty((M, 256), device=dev)
    sp = torch.empty((M, 512), dtype=torch.bfloat16, device=dev)
    r = {}
    r["act0,split"] = timed(lambda: ops.tc_linear(a, w, b, act=0, want_f32=False, out_split=sp, want_split=True))
    r["tanh,split"] = timed(lambda: ops.tc_linear(a, w, b, act=1, want_f32=False, out_split=sp, want_split=True))
    r["tanh,f32"] = timed(lambda: ops.tc_linear(a, w, b, act=1, out_f32=f32, want_f32=True))
    r["tanh,f32,split"] = timed(lambda: ops.tc_linear(a, w, b, act=1, out_f32=f32, out_split=sp, want_split=True))
    hw, hb = torch.randn(4, 256, device=dev) * 0.01, torch.zeros(4, device=dev)
    r["tanh,head4+sample"] = timed(lambda: ops.tc_linear_head(a, w, b, hw, hb, act=1, sample=(1, 1)))
    r["act0,head4"] = timed(lambda: ops.tc_linear_head(a, w, b, hw, hb, act=0))
    out["K=%d" % K] = r
print(json.dumps(out, indent=1))
