"""Times the MLP pieces (CUDA events): policy forward on both paths, and one CoPO loss forward+backward."""
import json, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import ops, policy as P

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main(M=163840, D=92):
    out = {}
    for prec in ("bf16_split", "fp32"):
        pol = P.CoPOPolicy(D, 2, P.copo_config(precision=prec))
        obs = torch.rand(M, D, device="cuda")
        ms = timeit(lambda: pol.model.forward(obs))
        flop = 2.0 * M * (D * 256 + 256 * 256 + 256 * 4)
        out["policy_forward_%s" % prec] = dict(ms=ms, tflops=flop / ms / 1e9)
        ms = timeit(lambda: pol.model.get_nei_value(obs))
        out["value_forward_%s" % prec] = dict(ms=ms)
        B = 65536
        g = torch.Generator(device="cuda").manual_seed(0)
        r = lambda *s: torch.randn(*s, device="cuda", generator=g)
        batch = dict(obs=obs[:B], centralized_critic_obs=obs[:B], actions=0.5 * r(B, 2), action_logp=-1.5 + 0.3 * r(B),
                     action_dist_inputs=0.3 * r(B, 4), advantages=r(B), normalized_advantages=r(B), vf_preds=r(B),
                     value_targets=r(B), nei_values=r(B), nei_target=r(B), global_values=r(B), global_target=r(B),
                     nei_advantage=r(B), global_advantages=r(B))
        ms = timeit(lambda: pol.learn_on_batch(batch), n=10)
        flop = 3.0 * 2.0 * B * (4 * (D * 256 + 256 * 256) + 256 * 4 + 3 * 256)
        out["copo_learn_on_batch_%s_B%d" % (prec, B)] = dict(ms=ms, tflops=flop / ms / 1e9, rows_per_s=B / ms * 1e3)
    # raw layer GEMMs
    x = torch.rand(M, 256, device="cuda"); W = torch.randn(256, 256, device="cuda") / 16; b = torch.zeros(256, device="cuda")
    a, w = ops.tc_split_rows(x), ops.tc_prep_weight(W)
    ms = timeit(lambda: ops.tc_linear(a, w, b, act=1, want_f32=True, want_split=True))
    out["tc_linear_256x256"] = dict(ms=ms, eff_tflops=2.0 * M * 256 * 256 / ms / 1e9, bf16_tflops=3 * 2.0 * M * 256 * 256 / ms / 1e9)
    ms = timeit(lambda: ops.linear_forward(x, W, b, 1))
    out["simt_linear_256x256"] = dict(ms=ms, tflops=2.0 * M * 256 * 256 / ms / 1e9)
    ms = timeit(lambda: ops.tc_split_rows(x))
    out["split_rows_256"] = dict(ms=ms)
    dW = torch.zeros(256, 256, device="cuda"); db = torch.zeros(256, device="cuda")
    ms = timeit(lambda: ops.linear_backward(x, x, W, dW, db, True, need_dx=False))
    out["simt_wgrad_256x256"] = dict(ms=ms, tflops=2.0 * M * 256 * 256 / ms / 1e9)
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
