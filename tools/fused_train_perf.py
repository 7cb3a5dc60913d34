"""Times the learner's forward pass of one network at M rows: the one-kernel form (b2c_tc_mlp2_train) against the two
layer kernels, and the inference form of the same kernel (no hidden-layer outputs) for reference.  CUDA events, back to
back launches after warm-up; outputs are re-used buffers (no allocation inside the timed region is avoidable through
the Python wrapper, so the allocator's cached blocks are what is timed)."""
import json, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 92
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(M, K, device="cuda", generator=g) * 2 - 1
W1, b1 = torch.randn(256, K, device="cuda", generator=g) / K ** 0.5, 0.1 * torch.randn(256, device="cuda", generator=g)
W2, b2 = torch.randn(256, 256, device="cuda", generator=g) / 16, 0.1 * torch.randn(256, device="cuda", generator=g)
out = {"M": M, "K": K}
for n in (4, 1):
    W3, b3 = torch.randn(n, 256, device="cuda", generator=g) / 16, 0.1 * torch.randn(n, device="cuda", generator=g)
    a, w1, w2 = ops.tc_split_rows(x), ops.tc_prep_weight(W1), ops.tc_prep_weight(W2)

    def timed(fn, reps=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / reps * 1e3, 1)

    def layers():
        _, s1 = ops.tc_linear(a, w1, b1, act=1, want_f32=False, want_split=True)
        ops.tc_linear_head(s1, w2, b2, W3, b3, act=1, want_f32=True)

    out["n%d" % n] = {"layer_kernels_us": timed(layers),
                      "one_kernel_train_us": timed(lambda: ops.tc_mlp2_head(a, w1, b1, w2, b2, W3, b3, train=True)),
                      "one_kernel_inference_us": timed(lambda: ops.tc_mlp2_head(a, w1, b1, w2, b2, W3, b3))}
print(json.dumps(out))
