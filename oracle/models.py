"""TEST INFRASTRUCTURE (oracle) - torch-CPU restatement of the reference's networks, action distribution and losses.

Follows (paths under copo_code/copo/torch_copo/):
  CCModel                 algo_ccppo.py:74-219   (policy trunk + separate value branch fed the critic observation)
  CoPOModel               algo_copo.py:96-182    (+ neighbourhood / global value nets, lcf_parameters, lcf_mean/std,
                                                  compute_coordinated)
  get_centralized_critic_obs_dim   algo_ccppo.py:55-71
  IPPOPolicy.loss         algo_ippo.py:79-172
  CCPPOPolicy.loss        algo_ccppo.py:376-472
  CoPOPolicy.loss         algo_copo.py:311-424
  CoPOPolicy.meta_update  algo_copo.py:228-309
and, from ray 2.2.0 (not vendored; restated from the published implementation):
  SlimFC / normc_initializer   rllib/models/torch/misc.py   (Linear [+ Tanh], weight ~ N(0,1) rows scaled to norm std,
                                                             bias 0)
  TorchDiagGaussian            rllib/models/torch/torch_action_dist.py (mean | log_std split, logp, kl, entropy)
  update_kl                    rllib/algorithms/ppo/ppo_torch_policy.py (x1.5 above 2*target, x0.5 below 0.5*target)

PINNING: the policy forward is pinned against the reference's own numpy forward on the shipped weights
(tests/golden/mlp_golden.npz, made by tests/golden/make_mlp_golden.py from copo/eval/get_policy_function.py:54-98).
The losses and the meta update are pinned against the reference's OWN code executed in the build container: the bodies of
IPPOPolicy.loss / CCPPOPolicy.loss / CoPOPolicy.loss / CoPOPolicy.meta_update (and CoPOModel.compute_coordinated /
lcf_mean / lcf_std) are extracted from /root/reference with `ast` and run on seeded batches; total loss, statistics,
full gradient and the LCF parameters after the Adam step are stored in tests/golden/ref_golden.npz
(tests/golden/make_ref_golden.py; checked in tests/test_ref_golden_cpu.py to 1e-6).
Parameter names equal RLlib's state_dict names so the shipped ccppo_*.npz load directly.
Only tests/, __graft_entry__.smoke() and bench.py may import this.
"""
import math

import numpy as np
import torch
from torch import nn


def normc_(weight, std=1.0):
    weight.data.normal_(0, 1)
    weight.data *= std / torch.sqrt(weight.data.pow(2).sum(1, keepdim=True))


class SlimFC(nn.Module):
    def __init__(self, in_size, out_size, std, activation):
        super().__init__()
        linear = nn.Linear(in_size, out_size)
        normc_(linear.weight, std)
        nn.init.constant_(linear.bias, 0.0)
        layers = [linear]
        if activation == "tanh":
            layers.append(nn.Tanh())
        self._model = nn.Sequential(*layers)

    def forward(self, x):
        return self._model(x)


def centralized_critic_obs_dim(odim, adim, counterfactual=True, num_neighbours=4, fuse_mode="mf"):
    n = {"concat": num_neighbours, "mf": 1, "none": 0}[fuse_mode] + 1
    d = n * odim
    if counterfactual:
        d += (n - 1) * adim
    return d


def _value_net(in_size, hiddens):
    layers, prev = [], in_size
    for h in hiddens:
        layers.append(SlimFC(prev, h, 1.0, "tanh"))
        prev = h
    layers.append(SlimFC(prev, 1, 0.01, None))
    return nn.Sequential(*layers)


class CCModel(nn.Module):
    def __init__(self, odim, adim=2, hiddens=(256, 256), cdim=None):
        super().__init__()
        self.odim, self.adim = odim, adim
        self.cdim = cdim if cdim is not None else odim
        layers, prev = [], odim
        for h in hiddens:
            layers.append(SlimFC(prev, h, 1.0, "tanh"))
            prev = h
        self._hidden_layers = nn.Sequential(*layers)
        self._logits = SlimFC(prev, 2 * adim, 0.01, None)
        vf, prev = [], self.cdim
        for h in hiddens:
            vf.append(SlimFC(prev, h, 1.0, "tanh"))
            prev = h
        self._value_branch_separate = nn.Sequential(*vf)
        self._value_branch = SlimFC(prev, 1, 0.01, None)

    def forward(self, obs):
        return self._logits(self._hidden_layers(obs))

    def value_function(self):
        raise ValueError("Centralized Value Function should not be called directly! "
                         "Call central_value_function(cobs) instead!")

    def central_value_function(self, cobs):
        return torch.reshape(self._value_branch(self._value_branch_separate(cobs)), [-1])


class CoPOModel(CCModel):
    def __init__(self, odim, adim=2, hiddens=(256, 256), cdim=None, initial_lcf_std=0.1):
        super().__init__(odim, adim, hiddens, cdim)
        self.nei_value_network = _value_net(self.cdim, hiddens)
        self.global_value_network = _value_net(self.cdim, hiddens)
        self.lcf_parameters = nn.Parameter(torch.as_tensor([0.0, np.log(initial_lcf_std)], dtype=torch.float32))

    def get_nei_value(self, cobs):
        return torch.reshape(self.nei_value_network(cobs), [-1])

    def get_global_value(self, cobs):
        return torch.reshape(self.global_value_network(cobs), [-1])

    @property
    def lcf_mean(self):
        return torch.clamp(torch.tanh(self.lcf_parameters[0]), -1 + 1e-6, 1 - 1e-6)

    @property
    def lcf_std(self):
        return torch.exp(torch.clamp(self.lcf_parameters[1], -20, 2))

    def compute_coordinated(self, ego, neighbor, eps):
        """eps: the standard-normal draws of `lcf_dist.rsample(ego.size())` (algo_copo.py:158), injected."""
        lcf_rad = (self.lcf_mean + self.lcf_std * eps) * np.pi / 2
        return torch.cos(lcf_rad) * ego + torch.sin(lcf_rad) * neighbor


class DiagGaussian:
    def __init__(self, inputs):
        self.mean, self.log_std = torch.chunk(inputs, 2, dim=1)
        self.std = torch.exp(self.log_std)

    def logp(self, x):
        return (-0.5 * torch.sum(torch.pow((x - self.mean) / self.std, 2.0), dim=1)
                - 0.5 * np.log(2.0 * np.pi) * x.shape[1] - torch.sum(self.log_std, dim=1))

    def kl(self, other):
        return torch.sum(other.log_std - self.log_std
                         + (torch.pow(self.std, 2.0) + torch.pow(self.mean - other.mean, 2.0))
                         / (2.0 * torch.pow(other.std, 2.0)) - 0.5, dim=1)

    def entropy(self):
        return torch.sum(self.log_std + 0.5 * np.log(2.0 * np.pi * np.e), dim=1)

    def sample(self, eps):
        return self.mean + self.std * eps


DEFAULT_CFG = dict(clip_param=0.2, vf_clip_param=100.0, vf_loss_coeff=1.0, entropy_coeff=0.0, kl_coeff=0.2,
                   old_value_loss=True, gamma=0.99, lambda_=0.95, lr=3e-4, lcf_lr=1e-4, kl_target=0.01)


def _value_loss(cfg, current_vf, prev_vf, target):
    if cfg["old_value_loss"]:
        l1 = torch.pow(current_vf - target, 2.0)
        clipped = prev_vf + torch.clamp(current_vf - prev_vf, -cfg["vf_clip_param"], cfg["vf_clip_param"])
        l2 = torch.pow(clipped - target, 2.0)
        return torch.max(l1, l2)
    return torch.clamp(torch.pow(current_vf - target, 2.0), 0, cfg["vf_clip_param"])


def ppo_loss(model, batch, cfg, algo="copo"):
    """batch: dict of torch tensors: obs, actions, action_logp, action_dist_inputs, advantages (ippo/ccppo) or
    normalized_advantages (copo), vf_preds, value_targets, centralized_critic_obs (ccppo/copo), and for copo
    nei_values/nei_target/global_values/global_target.  Returns (total_loss, stats dict)."""
    logits = model(batch["obs"])
    dist = DiagGaussian(logits)
    prev = DiagGaussian(batch["action_dist_inputs"])
    ratio = torch.exp(dist.logp(batch["actions"]) - batch["action_logp"])
    if cfg["kl_coeff"] > 0.0:
        mean_kl = torch.mean(prev.kl(dist))
    else:
        mean_kl = torch.tensor(0.0)
    ent = dist.entropy()
    adv = batch["normalized_advantages"] if algo == "copo" else batch["advantages"]
    surr = torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cfg["clip_param"], 1 + cfg["clip_param"]))
    if algo == "ippo":
        vf = torch.reshape(model._value_branch(model._value_branch_separate(batch["obs"])), [-1])
    else:
        vf = model.central_value_function(batch["centralized_critic_obs"])
    vloss = _value_loss(cfg, vf, batch["vf_preds"], batch["value_targets"])
    stats = {}
    per = -surr + cfg["vf_loss_coeff"] * vloss
    if algo == "copo":
        nloss = _value_loss(cfg, model.get_nei_value(batch["centralized_critic_obs"]), batch["nei_values"],
                            batch["nei_target"])
        gloss = _value_loss(cfg, model.get_global_value(batch["centralized_critic_obs"]), batch["global_values"],
                            batch["global_target"])
        per = per + cfg["vf_loss_coeff"] * nloss + cfg["vf_loss_coeff"] * gloss
        stats["mean_nei_vf_loss"] = torch.mean(nloss)
        stats["mean_global_vf_loss"] = torch.mean(gloss)
    total = torch.mean(per - cfg["entropy_coeff"] * ent)
    if cfg["kl_coeff"] > 0.0:
        total = total + cfg["kl_coeff"] * mean_kl
    stats.update(total_loss=total, mean_policy_loss=torch.mean(-surr), mean_vf_loss=torch.mean(vloss),
                 mean_entropy=torch.mean(ent), mean_kl_loss=mean_kl)
    return total, stats


def policy_params(model):
    """Parameters that receive a gradient from a policy-only loss (trunk + logits), in module order."""
    return list(model._hidden_layers.parameters()) + list(model._logits.parameters())


def meta_gradient(model, target_model, batch, cfg, raw_mean, raw_std, eps):
    """algo_copo.py:228-309 up to (not including) the optimizer step.
    Returns (lcf_final_loss, grad wrt lcf_parameters, stats)."""
    dist = DiagGaussian(model(batch["obs"]))
    ratio = torch.exp(dist.logp(batch["actions"]) - batch["action_logp"])
    adv = batch["global_advantages"]
    surr = torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cfg["clip_param"], 1 + cfg["clip_param"]))
    new_loss = torch.mean(-surr)
    g_new = torch.autograd.grad(new_loss, policy_params(model))
    old_logp = DiagGaussian(target_model(batch["obs"])).logp(batch["actions"])
    old_loss = torch.mean(old_logp)
    g_old = torch.autograd.grad(old_loss, policy_params(target_model))
    grad_value = sum((a * b).sum() for a, b in zip(g_new, g_old))
    coordinated = model.compute_coordinated(batch["advantages"], batch["nei_advantage"], eps)
    lcf_adv = (coordinated - raw_mean) / raw_std
    lcf_adv_loss = torch.mean(lcf_adv)
    final = grad_value * lcf_adv_loss
    (g_lcf,) = torch.autograd.grad(final, [model.lcf_parameters])
    stats = dict(new_policy_ego_loss=new_loss.item(), old_policy_logp_loss=old_loss.item(),
                 lcf_lcf_adv_loss=lcf_adv_loss.item(), lcf_final_loss=final.item(), grad_value=grad_value.item(),
                 coordinated_adv=coordinated.mean().item(), global_adv=adv.mean().item())
    return final, g_lcf, stats, (g_new, g_old)


def update_kl(kl_coeff, sampled_kl, kl_target=0.01):
    if sampled_kl > 2.0 * kl_target:
        kl_coeff *= 1.5
    elif sampled_kl < 0.5 * kl_target:
        kl_coeff *= 0.5
    return kl_coeff


def adam_step(param, grad, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) on numpy/torch tensors; returns (param, m, v)."""
    b1, b2 = betas
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    return param - (lr / bc1) * (m / denom), m, v


# ---- numpy forward with either checkpoint naming (what the fixtures are made with upstream) -----------------------
def load_policy_npz(path):
    """-> list of (W [in, out], b) for the three policy layers, from TF-era or torch-era key names."""
    w = np.load(path)
    keys = list(w.files)
    if "_hidden_layers.0._model.0.weight" in keys:
        names = ["_hidden_layers.0._model.0", "_hidden_layers.1._model.0", "_logits._model.0"]
        return [(w[n + ".weight"].T.copy(), w[n + ".bias"].copy()) for n in names]
    suffix = "_1" if any(k.endswith("fc_1_1/kernel") for k in keys) else ""
    out = []
    for layer in ("fc_1", "fc_2", "fc_out"):
        k = [x for x in keys if x.endswith("/%s%s/kernel" % (layer, suffix))][0]
        out.append((w[k].copy(), w[k.replace("kernel", "bias")].copy()))
    return out


def mlp_forward_np(layers, x):
    x = np.asarray(x, np.float32)
    for n, (W, b) in enumerate(layers):
        x = x @ W + b
        if n < len(layers) - 1:
            x = np.tanh(x)
    return x
