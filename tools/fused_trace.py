"""Pipeline timeline of the one-kernel network (mlp_fused.cu) on one SM: clock64() stamps of CTA 0 at 16 events per
tile, printed in microseconds relative to the first event (SM clock taken as 1.965 GHz)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops, _lib

M, K = 163840, int(sys.argv[1]) if len(sys.argv) > 1 else 92
dev = "cuda"
x = torch.rand(M, K, device=dev)
W1, b1 = torch.randn(256, K, device=dev) / K ** 0.5, torch.zeros(256, device=dev)
W2, b2 = torch.randn(256, 256, device=dev) / 16, torch.zeros(256, device=dev)
hw, hb = torch.randn(4, 256, device=dev) * 0.01, torch.zeros(4, device=dev)
a, w1, w2 = ops.tc_split_rows(x), ops.tc_prep_weight(W1), ops.tc_prep_weight(W2)
for _ in range(3):
    ops.tc_mlp2_head(a, w1, b1, w2, b2, hw, hb, sample=(1, 1))
tiles = (M // 128 + 147) // 148
trace = torch.zeros(tiles * 16, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.b2c_tc_mlp2_set_trace(ctypes.c_void_p(trace.data_ptr()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.fill_(1)
ops.tc_mlp2_head(a, w1, b1, w2, b2, hw, hb, sample=(1, 1))
torch.cuda.synchronize()
lib.b2c_tc_mlp2_set_trace(None)
t = trace.cpu().view(tiles, 16).double()
t0 = t[t > 0].min()
us = (t - t0) / 1965.0
names = ["mma:L1a ready", "mma:L1b ready", "mma:D2 free", "mma:L2-0 ready", "mma:L2-1 ready", "mma:L2-2 ready", "mma:L2-3 ready",
         "epi1:D1 full", "epi1:chunk0 stored", "epi1:chunk1 stored", "epi2:D2 full", "epi2:D2 released", "finaliser:tile done",
         "mma:D1 free", "tma:L1a slot free", "tma:L2-0 slot free"]
order = [14, 13, 0, 1, 7, 15, 8, 2, 3, 4, 9, 5, 6, 10, 11, 12]
for i in range(tiles):
    print("tile %d: " % i + "  ".join("%s %.2f" % (names[k], us[i, k]) for k in order))
