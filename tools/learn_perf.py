"""One CoPO learn_on_batch (B rows) a few times - to be run under `ncu --metrics gpu__time_duration.sum`."""
import sys
import torch
sys.path.insert(0, ".")
from copo_b200 import policy as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pol = P.CoPOPolicy(92, 2, P.copo_config())
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
obs = torch.rand(B, 92, device="cuda", generator=g)
batch = dict(obs=obs, centralized_critic_obs=obs, actions=0.5 * r(B, 2), action_logp=-1.5 + 0.3 * r(B),
             action_dist_inputs=0.3 * r(B, 4), advantages=r(B), normalized_advantages=r(B), vf_preds=r(B),
             value_targets=r(B), nei_values=r(B), nei_target=r(B), global_values=r(B), global_target=r(B),
             nei_advantage=r(B), global_advantages=r(B))
for _ in range(3):
    pol.learn_on_batch(batch)
torch.cuda.synchronize()
