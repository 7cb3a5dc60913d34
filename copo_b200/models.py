"""Policy / value networks of IPPO, CCPPO and CoPO on the CUDA kernels of libcopo_b200.so.

Mirrors the reference's model interface (torch_copo/algo_ccppo.py:74-219 `CCModel`, algo_copo.py:96-182 `CoPOModel`):
`forward(obs) -> logits`, `central_value_function(cobs)`, `get_nei_value`, `get_global_value`, `value_function()`
raising, `compute_coordinated`, `lcf_mean / lcf_std / lcf_parameters`, and a `state_dict()` whose keys are RLlib's
(`_hidden_layers.0._model.0.weight`, `_logits._model.0.bias`, `_value_branch_separate.1._model.0.weight`,
`nei_value_network.2._model.0.weight`, ...), so the shipped `best_checkpoints/ccppo_*.npz` load directly and the TF-era
`ippo_* / cl_* / copo_*` files load through `load_policy_npz` (copo/eval/get_policy_function.py:54-98 naming).

All network parameters live in one flat float32 device buffer (one Adam launch, one gradient all-reduce); a second
flat buffer holds the gradients.  There is no autograd in here: backward is the hand-written kernels.
"""
import math

import numpy as np
import torch

import os

from . import ops

# inference passes (rollout actions, value predictions) run the one-kernel network (mlp_fused.cu); B2C_TC_FUSED=0 falls
# back to the layer-by-layer kernels (same bits) for A/B measurements
FUSED_INFERENCE = os.environ.get("B2C_TC_FUSED", "1") != "0"
FUSED_TRAINING = os.environ.get("B2C_TC_FUSED_TRAIN", "1") != "0"      # the learner's forward in the same kernel


def centralized_critic_obs_dim(obs_dim, act_dim, counterfactual=True, num_neighbours=4, fuse_mode="mf"):
    """algo_ccppo.py:55-71."""
    if fuse_mode not in ("concat", "mf", "none"):
        raise ValueError("Unknown fuse mode: %s" % fuse_mode)
    n = {"concat": num_neighbours, "mf": 1, "none": 0}[fuse_mode] + 1
    d = n * obs_dim
    if counterfactual:
        d += (n - 1) * act_dim
    return d


class _Net:
    """One 3-layer tanh MLP: views into the model's flat parameter / gradient buffers."""

    def __init__(self, names, in_dim, hiddens, out_dim):
        self.names, self.in_dim, self.hiddens, self.out_dim = names, in_dim, tuple(hiddens), out_dim
        dims = [in_dim] + list(hiddens) + [out_dim]
        self.shapes = [(dims[k + 1], dims[k]) for k in range(len(dims) - 1)]
        self.W, self.b, self.dW, self.db = [], [], [], []

    def numel(self):
        return sum(o * i + o for o, i in self.shapes)

    def bind(self, flat, gflat, offset):
        self.offset = offset
        for o, i in self.shapes:
            self.W.append(flat[offset:offset + o * i].view(o, i))
            self.dW.append(gflat[offset:offset + o * i].view(o, i))
            offset += o * i
            self.b.append(flat[offset:offset + o])
            self.db.append(gflat[offset:offset + o])
            offset += o
        self.end = offset
        return offset

    def init(self, gen):
        n_layers = len(self.shapes)
        for k, (o, i) in enumerate(self.shapes):
            std = 1.0 if k < n_layers - 1 else 0.01                      # normc_initializer(1.0) / (0.01)
            w = torch.randn(o, i, generator=gen)
            w *= std / torch.sqrt(w.pow(2).sum(1, keepdim=True))
            self.W[k].copy_(w)
            self.b[k].zero_()

    # ---- tensor-core operands: [hi | lo] bf16 copies of the 256-wide layers, rebuilt when weights change ----
    def tc_ok(self):
        return len(self.shapes) == 3 and self.hiddens == (256, 256)

    def tc_weights(self, version):
        if getattr(self, "_tc_version", None) != version:
            self._tc_w = [ops.tc_prep_weight(self.W[0]), ops.tc_prep_weight(self.W[1]),
                          ops.tc_prep_weight(self.W[1], transpose=True)]
            self._tc_version = version
        return self._tc_w

    # inference: no activations kept
    def forward(self, x, tc_version=None, x_split=None, sample=None, out=None, actions=None, logp=None):
        if tc_version is not None and self.tc_ok():
            w1, w2, _ = self.tc_weights(tc_version)
            s0 = x_split if x_split is not None else ops.tc_split_rows(x)
            if self.out_dim in (1, 4) and FUSED_INFERENCE:       # one kernel, the hidden layers never leave the SM
                out, actions, logp = ops.tc_mlp2_head(s0, w1, self.b[0], w2, self.b[1], self.W[2], self.b[2], sample=sample,
                                                      out=out, actions=actions, logp=logp)
                return out if sample is None else (out, actions, logp)
            _, s1 = ops.tc_linear(s0, w1, self.b[0], act=1, want_f32=False, want_split=True)
            if self.out_dim in (1, 4):       # the narrow output layer rides in the layer-2 epilogue
                out, actions, logp, _ = ops.tc_linear_head(s1, w2, self.b[1], self.W[2], self.b[2], act=1, sample=sample,
                                                           out=out, actions=actions, logp=logp)
                return out if sample is None else (out, actions, logp)
            h2, _ = ops.tc_linear(s1, w2, self.b[1], act=1)
            return ops.linear_forward(h2, self.W[2], self.b[2], 0)
        h = x
        for k in range(len(self.shapes)):
            last = k == len(self.shapes) - 1
            h = ops.linear_forward(h, self.W[k], self.b[k], 0 if last else 1)
        return h

    def forward_train(self, x, tc_version=None, x_split=None):
        if tc_version is not None and self.tc_ok():
            w1, w2, _ = self.tc_weights(tc_version)
            # the first layer's operand carries a ones column when it has a padding column to spare: its weight-gradient
            # launch then returns the bias gradient too (backward below)
            ones = ops.tc_has_ones_col(self.in_dim) and self.in_dim <= 256
            s0 = x_split if x_split is not None else ops.tc_split_rows(x, ones_col=ones)
            # h1 only leaves as the [hi | lo] operand of layer 2: the backward pass takes 1 - h1^2 from it
            if self.out_dim in (1, 4) and FUSED_TRAINING:        # one kernel: h1's operand and h2 leave from its epilogues
                out, s1, h2 = ops.tc_mlp2_head(s0, w1, self.b[0], w2, self.b[1], self.W[2], self.b[2], train=True)
                return [x, None, h2, out, s0, s1, ones]
            _, s1 = ops.tc_linear(s0, w1, self.b[0], act=1, want_f32=False, want_split=True)
            if self.out_dim in (1, 4):   # output layer in the layer-2 epilogue; h2 is kept for the backward pass
                out, _, _, h2 = ops.tc_linear_head(s1, w2, self.b[1], self.W[2], self.b[2], act=1, want_f32=True)
            else:
                h2, _ = ops.tc_linear(s1, w2, self.b[1], act=1)
                out = ops.linear_forward(h2, self.W[2], self.b[2], 0)
            # the [hi | lo] operands are kept: the weight gradients read them again
            return [x, None, h2, out, s0, s1, ones]
        acts = [x]
        for k in range(len(self.shapes)):
            last = k == len(self.shapes) - 1
            acts.append(ops.linear_forward(acts[-1], self.W[k], self.b[k], 0 if last else 1))
        return acts                                                     # [x, h1, h2, out]

    def backward(self, acts, dout, tc_version=None):
        """Accumulates dW / db of every layer given d(loss)/d(out)."""
        if tc_version is not None and self.tc_ok():
            x, _, h2, _, s0, s1, ones = acts
            # output layer backward: dz2 leaves only as the tensor-core operand; its column sums are the bias gradient of
            # layer 2 (the kernel's threads own columns)
            _, dz2s = ops.head_backward_split(dout, h2, self.W[2], self.dW[2], self.db[2], db_hidden=self.db[1],
                                              want_f32=False)
            # layer 2: input gradient ((dz2 W2) * (1 - h1^2), h1 from its own operand) and weight gradient on the tensor
            # cores from the same [hi | lo] operand
            narrow = s0.shape[1] <= 512
            dz1, dz1s = ops.tc_linear_dgrad(dz2s, self.tc_weights(tc_version)[2], s1, want_f32=not (narrow and ones))
            ops.tc_wgrad(dz2s, s1, self.dW[1])
            if narrow and ones:
                ops.tc_wgrad(dz1s, s0, self.dW[0], db=self.db[0])       # bias gradient from the ones column
            elif narrow:
                ops.tc_wgrad(dz1s, s0, self.dW[0])
                ops.colsum(dz1, self.db[0])
            else:                                    # very wide critic inputs (concat fusion): fp32 path
                ops.linear_backward(dz1, x, self.W[0], self.dW[0], self.db[0], h_prev_is_tanh=False, need_dx=False)
            return
        d = dout
        for k in reversed(range(len(self.shapes))):
            d = ops.linear_backward(d, acts[k], self.W[k], self.dW[k], self.db[k], h_prev_is_tanh=(k > 0),
                                    need_dx=(k > 0))


class CCModel:
    POLICY = ("_hidden_layers.0._model.0", "_hidden_layers.1._model.0", "_logits._model.0")
    VALUE = ("_value_branch_separate.0._model.0", "_value_branch_separate.1._model.0", "_value_branch._model.0")

    def __init__(self, obs_dim, act_dim=2, hiddens=(256, 256), fuse_mode="none", counterfactual=True, num_neighbours=4,
                 device=None, seed=0, critic_obs_dim=None, precision="bf16_split"):
        self.obs_dim, self.act_dim, self.num_outputs = int(obs_dim), int(act_dim), 2 * int(act_dim)
        self.custom = dict(fuse_mode=fuse_mode, counterfactual=counterfactual, num_neighbours=num_neighbours)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.cobs_dim = critic_obs_dim if critic_obs_dim is not None else self.get_centralized_critic_obs_dim()
        self.nets = {"policy": _Net(self.POLICY, self.obs_dim, hiddens, self.num_outputs),
                     "value": _Net(self.VALUE, self.cobs_dim, hiddens, 1)}
        self._extra_nets(hiddens)
        self._allocate(seed)
        self.tower_stats = {}
        # "bf16_split": 256-wide layers on tcgen05 tensor cores with split-bf16 operands (fp32-grade accuracy);
        # "fp32": CUDA-core kernels, exact fp32 (used by the tight parity tests)
        self.precision = precision
        self.weights_version = 0

    def _tc(self):
        return self.weights_version if self.precision == "bf16_split" else None

    def mark_weights_changed(self):
        self.weights_version += 1

    def _extra_nets(self, hiddens):
        pass

    def get_centralized_critic_obs_dim(self):
        return centralized_critic_obs_dim(self.obs_dim, self.act_dim, self.custom["counterfactual"],
                                          self.custom["num_neighbours"], self.custom["fuse_mode"])

    def _allocate(self, seed):
        n = sum(net.numel() for net in self.nets.values())
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(n, dtype=torch.float32, device=self.device)
        off = 0
        for net in self.nets.values():
            off = net.bind(self.flat, self.grad, off)
        gen = torch.Generator().manual_seed(seed)
        for net in self.nets.values():
            net.init(gen)

    # ---- reference interface -------------------------------------------------------------------------------
    def forward(self, obs, state=None, seq_lens=None, obs_split=None):
        """obs_split: the env's `obs_split` output for the same rows (skips the operand split pass)."""
        if isinstance(obs, dict):
            obs = obs.get("obs_flat", obs.get("obs"))
        obs = obs.reshape(obs.shape[0], -1)
        logits = self.nets["policy"].forward(obs, self._tc(), obs_split)
        return logits if state is None else (logits, state)

    __call__ = forward

    def forward_sample(self, obs, seed, step, obs_split=None, out=None, actions=None, logp=None):
        """Rollout step of the policy: logits, sampled actions and their log-probabilities.  On the tensor-core path
        this is two kernels (layer 1; layer 2 + logits + Gaussian sample in its epilogue)."""
        net = self.nets["policy"]
        if self._tc() is not None and net.tc_ok():
            return net.forward(obs, self._tc(), obs_split, sample=(seed, step), out=out, actions=actions, logp=logp)
        logits = net.forward(obs, None)
        a, lp = ops.gaussian_sample(logits, seed=seed, step=step)
        if out is not None:
            out.copy_(logits); logits = out
        if actions is not None:
            actions.copy_(a); a = actions
        if logp is not None:
            logp.copy_(lp); lp = logp
        return logits, a, lp

    def value_function(self):
        raise ValueError("Centralized Value Function should not be called directly! "
                         "Call central_value_function(cobs) instead!")

    def central_value_function(self, cobs, cobs_split=None):
        """cobs_split: the rows as the [hi | lo] operand (ops.cc_obs_fuse(..., want_split=True)); skips the split pass."""
        return self.nets["value"].forward(cobs, self._tc(), cobs_split).reshape(-1)

    def zero_grad(self):
        self.grad.zero_()

    def parameters(self):
        return [self.flat]

    def policy_slice(self):
        p = self.nets["policy"]
        return slice(p.offset, p.end)

    def num_parameters(self):
        return self.flat.numel()

    # ---- checkpoints (RLlib state_dict names) -------------------------------------------------
    def state_dict(self):
        out = {}
        for net in self.nets.values():
            for name, W, b in zip(net.names, net.W, net.b):
                out[name + ".weight"] = W.detach().clone()
                out[name + ".bias"] = b.detach().clone()
        return out

    def load_state_dict(self, sd, strict=True):
        self.mark_weights_changed()
        for net in self.nets.values():
            for name, W, b in zip(net.names, net.W, net.b):
                if name + ".weight" not in sd:
                    if strict:
                        raise KeyError(name + ".weight")
                    continue
                W.copy_(torch.as_tensor(np.asarray(sd[name + ".weight"]), dtype=torch.float32))
                b.copy_(torch.as_tensor(np.asarray(sd[name + ".bias"]), dtype=torch.float32))

    def load_policy_npz(self, path_or_dict):
        """Policy-only weights in either naming of best_checkpoints/*.npz (get_policy_function.py:54-98)."""
        w = path_or_dict if isinstance(path_or_dict, dict) else dict(np.load(path_or_dict))
        self.mark_weights_changed()
        keys = list(w.keys())
        net = self.nets["policy"]
        if self.POLICY[0] + ".weight" in keys:
            for name, W, b in zip(net.names, net.W, net.b):
                W.copy_(torch.as_tensor(w[name + ".weight"]))
                b.copy_(torch.as_tensor(w[name + ".bias"]))
            return
        suffix = "_1" if any(k.endswith("fc_1_1/kernel") for k in keys) else ""
        for layer, W, b in zip(("fc_1", "fc_2", "fc_out"), net.W, net.b):
            k = [x for x in keys if x.endswith("/%s%s/kernel" % (layer, suffix))][0]
            W.copy_(torch.as_tensor(np.ascontiguousarray(w[k].T)))           # TF kernels are stored [in, out]
            b.copy_(torch.as_tensor(w[k.replace("kernel", "bias")]))


class CoPOModel(CCModel):
    NEI = ("nei_value_network.0._model.0", "nei_value_network.1._model.0", "nei_value_network.2._model.0")
    GLOBAL = ("global_value_network.0._model.0", "global_value_network.1._model.0", "global_value_network.2._model.0")

    def __init__(self, obs_dim, act_dim=2, hiddens=(256, 256), fuse_mode="none", counterfactual=True, num_neighbours=4,
                 initial_lcf_std=0.1, use_distributional_lcf=True, device=None, seed=0, precision="bf16_split"):
        self.use_distributional_lcf = use_distributional_lcf
        super().__init__(obs_dim, act_dim, hiddens, fuse_mode, counterfactual, num_neighbours, device, seed,
                         precision=precision)
        init = [0.0, math.log(initial_lcf_std)] if use_distributional_lcf else [0.0]
        self.lcf_parameters = torch.tensor(init, dtype=torch.float32, device=self.device)     # algo_copo.py:120-124
        self.lcf_grad = torch.zeros_like(self.lcf_parameters)

    def _extra_nets(self, hiddens):
        self.nets["nei"] = _Net(self.NEI, self.cobs_dim, hiddens, 1)
        self.nets["global"] = _Net(self.GLOBAL, self.cobs_dim, hiddens, 1)

    def get_nei_value(self, cobs, cobs_split=None):
        return self.nets["nei"].forward(cobs, self._tc(), cobs_split).reshape(-1)

    def get_global_value(self, cobs, cobs_split=None):
        return self.nets["global"].forward(cobs, self._tc(), cobs_split).reshape(-1)

    @property
    def lcf_mean(self):
        return torch.clamp(torch.tanh(self.lcf_parameters[0]), -1 + 1e-6, 1 - 1e-6)          # algo_copo.py:171-172

    @property
    def lcf_std(self):
        if not self.use_distributional_lcf:
            return None
        return torch.exp(torch.clamp(self.lcf_parameters[1], -20, 2))                         # algo_copo.py:175-177

    @property
    def lcf_dist(self):
        if not self.use_distributional_lcf:
            return None
        return torch.distributions.normal.Normal(self.lcf_mean, self.lcf_std)

    def compute_coordinated(self, ego, neighbor, eps=None):
        if self.use_distributional_lcf:
            if eps is None:
                eps = torch.randn_like(ego)
            lcf_rad = (self.lcf_mean + self.lcf_std * eps) * np.pi / 2
        else:
            lcf_rad = self.lcf_mean * np.pi / 2
        return torch.cos(lcf_rad) * ego + torch.sin(lcf_rad) * neighbor

    def state_dict(self):
        out = super().state_dict()
        out["lcf_parameters"] = self.lcf_parameters.detach().clone()
        return out

    def load_state_dict(self, sd, strict=True):
        super().load_state_dict(sd, strict)
        if "lcf_parameters" in sd:
            self.lcf_parameters.copy_(torch.as_tensor(np.asarray(sd["lcf_parameters"]), dtype=torch.float32))
