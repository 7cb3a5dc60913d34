import sys, torch
sys.path.insert(0, ".")
from copo_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for M, K in ((129, 92), (300, 157), (1, 92)):
    x = torch.rand(M, K, device="cuda", generator=g)
    W1, b1 = torch.randn(256, K, device="cuda", generator=g) / 10, torch.zeros(256, device="cuda")
    W2, b2 = torch.randn(256, 256, device="cuda", generator=g) / 16, torch.zeros(256, device="cuda")
    W3, b3 = torch.randn(4, 256, device="cuda", generator=g) / 16, torch.zeros(4, device="cuda")
    a, w1, w2 = ops.tc_split_rows(x), ops.tc_prep_weight(W1), ops.tc_prep_weight(W2)
    out, s1, h2 = ops.tc_mlp2_head(a, w1, b1, w2, b2, W3, b3, train=True)
    torch.cuda.synchronize()
    print(M, K, "ok", float(out.sum()), float(h2.sum()))
