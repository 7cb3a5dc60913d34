"""tcgen05 path of the 256-wide layers against float64 references: split-bf16 accuracy must be far inside
the 1e-4 budget of the value predictions; covers ragged M (tile tails), every padded K the nets use, the input-gradient
form and the chained [hi | lo] operand hand-over between layers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, W, b, act):
    y = x.double() @ W.double().T + (b.double() if b is not None else 0)
    return torch.tanh(y) if act else y


@pytest.mark.parametrize("M,K", [(128, 64), (1000, 92), (4096 + 37, 256), (300, 157), (129, 184), (77, 463), (1, 92),
                                 (40000, 92), (148 * 128 * 3 + 77, 92), (148 * 128 * 5, 40)])
def test_tc_linear_forward(M, K):
    """Reduction lengths up to 128 run with the weights resident in shared memory (several tiles per CTA in the two
    largest cases: the A ring wraps), longer ones stream the weights with the activations."""
    from copo_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + K)
    x = torch.rand(M, K, generator=g)
    W = torch.randn(256, K, generator=g) / K ** 0.5
    b = 0.1 * torch.randn(256, generator=g)
    a = ops.tc_split_rows(x.cuda())
    Kp = ops.tc_padded_k(K)
    assert a.shape == (M, 2 * Kp)
    # the split itself: hi + lo reproduces x to ~2^-17 relative
    rec = a[:, :K].float() + a[:, Kp:Kp + K].float()
    assert float((rec.cpu() - x).abs().max()) < 2e-5
    assert float(a[:, K:Kp].float().abs().max()) == 0.0 if Kp > K else True
    w = ops.tc_prep_weight(W.cuda())
    for act in (0, 1):
        y, ys = ops.tc_linear(a, w, b.cuda(), act=act, want_f32=True, want_split=True)
        want = _ref(x, W, b, act)
        err = float((y.cpu().double() - want).abs().max())
        assert err < 3e-5, (act, err)
        rec = ys[:, :256].float() + ys[:, 256:].float()
        assert float((rec - y).abs().max()) < 2e-5


def test_tc_matches_fp32_kernels_and_chains():
    """Three-layer policy forward through the tensor-core path vs the exact fp32 kernels."""
    from copo_b200 import ops
    g = torch.Generator().manual_seed(0)
    M = 5000
    x = torch.rand(M, 92, generator=g).cuda()
    W1, W2 = (torch.randn(256, 92, generator=g) / 92 ** 0.5).cuda(), (torch.randn(256, 256, generator=g) / 16).cuda()
    b1, b2 = (0.1 * torch.randn(256, generator=g)).cuda(), (0.1 * torch.randn(256, generator=g)).cuda()
    h1 = ops.linear_forward(x, W1, b1, 1)
    h2 = ops.linear_forward(h1, W2, b2, 1)
    _, s1 = ops.tc_linear(ops.tc_split_rows(x), ops.tc_prep_weight(W1), b1, act=1, want_f32=False, want_split=True)
    t2, _ = ops.tc_linear(s1, ops.tc_prep_weight(W2), b2, act=1)
    assert float((t2 - h2).abs().max()) < 5e-5


def test_tc_input_gradient():
    from copo_b200 import ops
    g = torch.Generator().manual_seed(1)
    M = 3000
    dy = torch.randn(M, 256, generator=g)
    W = torch.randn(256, 256, generator=g) / 16
    h = torch.tanh(torch.randn(M, 256, generator=g))
    want = (dy.double() @ W.double()) * (1 - h.double() ** 2)
    a = ops.tc_split_rows(dy.cuda())
    wt = ops.tc_prep_weight(W.cuda(), transpose=True)
    dx, _ = ops.tc_linear(a, wt, None, act=0, dtanh_src=h.cuda())
    assert float((dx.cpu().double() - want).abs().max()) < 1e-4 * float(want.abs().max())


def test_tc_tanh_is_accurate_near_zero():
    """The epilogue's tanh keeps relative accuracy for tiny pre-activations (identity weights, scaled inputs); what is
    left is the 2^-17 relative rounding of the [hi | lo] operand split."""
    from copo_b200 import ops
    M = 512
    x = torch.zeros(M, 256)
    vals = torch.logspace(-6, 1, M)
    for j in range(256):
        x[:, j] = vals * (1 if j % 2 == 0 else -1)
    W = torch.eye(256)
    y, _ = ops.tc_linear(ops.tc_split_rows(x.cuda()), ops.tc_prep_weight(W.cuda()), None, act=1)
    want = torch.tanh(x.double())
    rel = ((y.cpu().double() - want).abs() / want.abs()).max()
    assert float(rel) < 2e-5, float(rel)


@pytest.mark.parametrize("M,K", [(4096, 256), (1000, 92), (148 * 32 * 3 + 17, 157), (65536, 256), (31, 64)])
def test_tc_weight_gradient(M, K):
    from copo_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    dz = torch.randn(M, 256, generator=g) * 0.1
    x = torch.tanh(torch.randn(M, K, generator=g))
    want = dz.double().T @ x.double()
    dW = torch.full((256, K), 0.5, device="cuda")          # accumulates on top of what is there
    ops.tc_wgrad(ops.tc_split_rows(dz.cuda()), ops.tc_split_rows(x.cuda()), dW)
    got = dW.cpu().double() - 0.5
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) < 1e-4 * scale + 1e-5, float((got - want).abs().max()) / scale
    dW2 = torch.full((256, K), 0.5, device="cuda")
    ops.tc_wgrad(ops.tc_split_rows(dz.cuda()), ops.tc_split_rows(x.cuda()), dW2)
    assert torch.equal(dW, dW2)                            # fixed summation order: bitwise reproducible


@pytest.mark.parametrize("M", [1, 129, 5000])
def test_tc_fused_head_and_sample(M):
    """Layer 2 with the logits / value layer (and the Gaussian sample) fused into its epilogue equals the unfused
    kernels; sampled actions are consistent with their log-probabilities."""
    from copo_b200 import ops
    from oracle import models as om
    g = torch.Generator().manual_seed(M)
    x = torch.rand(M, 256, generator=g).cuda()
    W2, b2 = (torch.randn(256, 256, generator=g) / 16).cuda(), (0.1 * torch.randn(256, generator=g)).cuda()
    W3, b3 = (torch.randn(4, 256, generator=g) / 16).cuda(), (0.1 * torch.randn(4, generator=g)).cuda()
    a, w = ops.tc_split_rows(x), ops.tc_prep_weight(W2)
    h2, _ = ops.tc_linear(a, w, b2, act=1)
    want = ops.linear_forward(h2, W3, b3, 0)
    logits, actions, logp, h = ops.tc_linear_head(a, w, b2, W3, b3, act=1, sample=(5, 9), want_f32=True)
    assert float((logits - want).abs().max()) < 2e-5 and torch.equal(h, h2)
    d = om.DiagGaussian(logits.cpu())
    assert float((d.logp(actions.cpu()) - logp.cpu()).abs().max()) < 1e-4
    z = (actions.cpu() - d.mean) / d.std
    if M >= 5000:
        assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1) < 0.05
    l2, a2, p2, _ = ops.tc_linear_head(a, w, b2, W3, b3, act=1, sample=(5, 9))
    assert torch.equal(a2, actions) and torch.equal(l2, logits)
    v, _, _, _ = ops.tc_linear_head(a, w, b2, W3[:1].contiguous(), b3[:1].contiguous(), act=1)
    assert float((v[:, 0] - want[:, 0]).abs().max()) < 2e-5


@pytest.mark.parametrize("M,K", [(1, 92), (129, 92), (5000, 157), (148 * 128 * 2 + 77, 92), (4096, 40), (3000, 184)])
def test_tc_one_kernel_network_matches_the_layer_kernels(M, K):
    """mlp_fused.cu: obs -> 256 tanh -> 256 tanh -> logits (+ sample) in one kernel, the hidden operand handed over in
    shared memory, against tc_linear + tc_linear_head (same MMAs in the same order, same epilogue functions; only the
    output layer's 256-term sum is split over two column groups instead of four, so the last bits can differ); and
    against float64 within the split-bf16 budget.  Covers ragged tile tails, 1..3 first-layer reduction blocks (odd and
    even stage parity of the ring across tiles) and both output widths."""
    from copo_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    x = (torch.rand(M, K, generator=g) * 2 - 1).cuda()
    W1, b1 = (torch.randn(256, K, generator=g) / K ** 0.5).cuda(), (0.1 * torch.randn(256, generator=g)).cuda()
    W2, b2 = (torch.randn(256, 256, generator=g) / 16).cuda(), (0.1 * torch.randn(256, generator=g)).cuda()
    W3, b3 = (torch.randn(4, 256, generator=g) / 16).cuda(), (0.1 * torch.randn(4, generator=g)).cuda()
    a, w1, w2 = ops.tc_split_rows(x), ops.tc_prep_weight(W1), ops.tc_prep_weight(W2)
    _, s1 = ops.tc_linear(a, w1, b1, act=1, want_f32=False, want_split=True)
    want, wact, wlogp, _ = ops.tc_linear_head(s1, w2, b2, W3, b3, act=1, sample=(7, 3))
    first = None
    for rep in range(3):                                    # same bits on every launch
        got, act, logp = ops.tc_mlp2_head(a, w1, b1, w2, b2, W3, b3, sample=(7, 3))
        assert float((got - want).abs().max()) < 2e-6 and float((act - wact).abs().max()) < 1e-5
        assert float((logp - wlogp).abs().max()) < 1e-4
        if first is None:
            first = (got.clone(), act.clone(), logp.clone())
        assert torch.equal(got, first[0]) and torch.equal(act, first[1]) and torch.equal(logp, first[2])
    ref = _ref(_ref(_ref(x.cpu(), W1.cpu(), b1.cpu(), 1), W2.cpu(), b2.cpu(), 1), W3.cpu(), b3.cpu(), 0)
    assert float((got.cpu().double() - ref).abs().max()) < 2e-5
    v, _, _ = ops.tc_mlp2_head(a, w1, b1, w2, b2, W3[:1].contiguous(), b3[:1].contiguous())
    wv, _, _, wh2 = ops.tc_linear_head(s1, w2, b2, W3[:1].contiguous(), b3[:1].contiguous(), act=1, want_f32=True)
    assert float((v - wv).abs().max()) < 2e-6
    # the learner's forward (b2c_tc_mlp2_train): the same kernel also leaves h1's [hi | lo] operand and h2 - the bits the
    # layer kernels write, every row of a ragged tile tail included
    for W, b in ((W3, b3), (W3[:1].contiguous(), b3[:1].contiguous())):
        out, t1, t2 = ops.tc_mlp2_head(a, w1, b1, w2, b2, W, b, train=True)
        assert t1.shape == (M, 512) and t2.shape == (M, 256)
        assert torch.equal(t1.view(torch.int16), s1.view(torch.int16)) and torch.equal(t2, wh2)
        assert torch.equal(out, got if W.shape[0] == 4 else v)
