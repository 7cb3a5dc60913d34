"""Times the one-kernel network (mlp_fused.cu) against the two layer kernels it replaces at the rollout shape
(M = 163 840 rows), L2 flushed before every launch; K = the observation widths of the BASELINE configurations."""
import json, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
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
for K in (92, 157):
    x = torch.rand(M, K, device=dev)
    W1, b1 = torch.randn(256, K, device=dev) / K ** 0.5, torch.zeros(256, device=dev)
    W2, b2 = torch.randn(256, 256, device=dev) / 16, torch.zeros(256, device=dev)
    hw, hb = torch.randn(4, 256, device=dev) * 0.01, torch.zeros(4, device=dev)
    a, w1, w2 = ops.tc_split_rows(x), ops.tc_prep_weight(W1), ops.tc_prep_weight(W2)
    sp = torch.empty((M, 512), dtype=torch.bfloat16, device=dev)

    def two():
        ops.tc_linear(a, w1, b1, act=1, want_f32=False, out_split=sp, want_split=True)
        ops.tc_linear_head(sp, w2, b2, hw, hb, act=1, sample=(1, 1))
    r = {"two_kernels": timed(two), "one_kernel": timed(lambda: ops.tc_mlp2_head(a, w1, b1, w2, b2, hw, hb, sample=(1, 1))),
         "one_kernel_value_head": timed(lambda: ops.tc_mlp2_head(a, w1, b1, w2, b2, hw[:1].contiguous(), hb[:1].contiguous()))}
    flop = 2.0 * M * (K * 256 + 256 * 256 + 4 * 256)
    r["one_kernel_fp32_equiv_TFLOPs"] = flop / r["one_kernel"] / 1e9
    out["K=%d" % K] = r
print(json.dumps(out, indent=1))
