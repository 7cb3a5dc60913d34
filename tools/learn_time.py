"""CUDA-event time of CoPOPolicy.learn_on_batch (B rows, four networks) and of meta_update, back to back."""
import json, sys
import torch
sys.path.insert(0, ".")
from copo_b200 import policy as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pol = P.CoPOPolicy(92, 2, P.copo_config())
pol.graph_batch_rows = B
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
obs = torch.rand(B, 92, device="cuda", generator=g)
batch = dict(obs=obs, centralized_critic_obs=obs, actions=0.5 * r(B, 2), action_logp=-1.5 + 0.3 * r(B),
             action_dist_inputs=0.3 * r(B, 4), advantages=r(B), normalized_advantages=r(B), vf_preds=r(B),
             value_targets=r(B), nei_values=r(B), nei_target=r(B), global_values=r(B), global_target=r(B),
             nei_advantage=r(B), global_advantages=r(B), step_lcf=0.1 * r(B))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {"rows": B, "learn_on_batch_ms": timed(lambda: pol.learn_on_batch(batch))}
try:
    pol.sync_stats = False
    out["meta_update_ms"] = timed(lambda: pol.meta_update(batch))
except Exception as e:
    out["meta_update_ms"] = repr(e)[:200]
print(json.dumps(out))
