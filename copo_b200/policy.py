"""IPPO / CCPPO / CoPO policies on the CUDA kernels: the reference's hook names with batched semantics.

  IPPOPolicy.loss          torch_copo/algo_ippo.py:79-172
  CCPPOPolicy.loss         torch_copo/algo_ccppo.py:376-472      postprocess: algo_ccppo.py:322-374
  CoPOPolicy.loss          torch_copo/algo_copo.py:311-424       postprocess: algo_copo.py:473-502
  CoPOPolicy.meta_update   torch_copo/algo_copo.py:228-309       assign_lcf / update_old_policy: :442-471

`train_batch` is a dict of device tensors under RLlib's SampleBatch column names.  `loss()` runs forward AND backward
through the hand-written kernels (gradients land in `model.grad`) and returns the scalar loss tensor; `learn_on_batch`
adds the data-parallel gradient all-reduce and the Adam step.  `postprocess_rollout` is the batched form of
`postprocess_trajectory`: it works on whole [T, N] rollout columns instead of one agent's SampleBatch.
"""
import math

import numpy as np
import os

import torch

from . import ops
from . import parallel
from .models import CCModel, CoPOModel

# SampleBatch / Postprocessing column names (rllib) and CoPO's own (algo_copo.py:48-60)
OBS, ACTIONS, ACTION_LOGP, ACTION_DIST_INPUTS = "obs", "actions", "action_logp", "action_dist_inputs"
REWARDS, DONES, VF_PREDS, ADVANTAGES, VALUE_TARGETS = "rewards", "dones", "vf_preds", "advantages", "value_targets"
CENTRALIZED_CRITIC_OBS = "centralized_critic_obs"
NEI_REWARDS, NEI_VALUES, NEI_ADVANTAGE, NEI_TARGET = "nei_rewards", "nei_values", "nei_advantage", "nei_target"
GLOBAL_REWARDS, GLOBAL_VALUES, GLOBAL_ADVANTAGES, GLOBAL_TARGET = ("global_rewards", "global_values",
                                                                   "global_advantages", "global_target")


class AlgoConfig(dict):
    """Attribute- and item-indexable config (the reference reads both `config["lcf_lr"]` and `config.lr`)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def update_from_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].update(v)
            else:
                self[k] = v
        return self


def ippo_config(**over):
    """IPPOConfig (algo_ippo.py:17-42) on top of the rllib 2.2.0 PPO defaults."""
    c = AlgoConfig(gamma=0.99, lambda_=0.95, kl_coeff=0.2, kl_target=0.01, vf_loss_coeff=1.0, entropy_coeff=0.0,
                   clip_param=0.2, vf_clip_param=100.0, old_value_loss=True, use_gae=True, use_critic=True, lr=3e-4,
                   sgd_minibatch_size=512, rollout_fragment_length=200, train_batch_size=2000, num_sgd_iter=5,
                   fcnet_hiddens=(256, 256), env_config={}, seed=0, precision="bf16_split")
    c["lambda"] = c["lambda_"]
    return c.update_from_dict(over)


def ccppo_config(**over):
    """CCPPOConfig (algo_ccppo.py:37-45)."""
    c = ippo_config(counterfactual=True, num_neighbours=4, fuse_mode="mf", mf_nei_distance=10)
    return c.update_from_dict(over)


def copo_config(**over):
    """CoPOConfig (algo_copo.py:63-92)."""
    c = ccppo_config(initial_lcf_std=0.1, lcf_sgd_minibatch_size=None, lcf_num_iters=5, lcf_lr=1e-4,
                     use_distributional_lcf=True, use_centralized_critic=False, fuse_mode="none")
    c.update_from_dict(over)
    assert c["use_distributional_lcf"]
    c["env_config"].update(return_native_reward=True, lcf_dist="normal", lcf_normal_std=c["initial_lcf_std"])
    return c


class _Adam:
    def __init__(self, param, lr):
        self.param, self.lr, self.step = param, lr, 0
        self.m, self.v = torch.zeros_like(param), torch.zeros_like(param)

    def apply(self, grad, grad_scale=1.0):
        self.step += 1
        ops.adam_step(self.param, grad, self.m, self.v, self.lr, self.step, grad_scale=grad_scale)


class IPPOPolicy:
    algo = "ippo"
    model_cls = CCModel

    def __init__(self, obs_dim, act_dim=2, config=None, device=None, dist=None):
        self.config = config if config is not None else self.default_config()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.model = self._build_model(self.config.get("seed", 0))
        self.kl_coeff = float(self.config["kl_coeff"])
        self.entropy_coeff = float(self.config["entropy_coeff"])
        # the two coefficients that change while training, mirrored on the device: the loss kernels read them there, so a
        # captured minibatch step (below) stays valid across KL-coefficient updates
        self._coeffs_dev = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._coeffs_host = None
        self._graphs = {}                                # captured minibatch steps, by batch signature
        self._graph_seen = {}
        self.graph_batch_rows = None                     # set by the trainer: this rank's rows in a full minibatch
        self._graph_off = os.environ.get("B2C_LEARN_GRAPH", "1") == "0"
        self._optimizer = _Adam(self.model.flat, self.config["lr"])
        self.dist = dist                                 # torch.distributed module when data parallel, else None
        self.ar_timer = None                             # parallel.AllReduceTimer when the trainer wants the share
        self.num_grad_updates = 0
        if self.config.get("num_neighbours", 4) != 4 and self.config.get("fuse_mode") == "concat":
            raise ValueError("fuse_mode='concat' is built for num_neighbours=4 (the reference default)")
        if not self.config.get("use_gae", True) or not self.config.get("use_critic", True):
            raise ValueError("use_gae=False / use_critic=False are not supported (the reference asserts use_critic)")

    @classmethod
    def default_config(cls):
        return ippo_config()

    def _build_model(self, seed):
        return CCModel(self.obs_dim, self.act_dim, self.config["fcnet_hiddens"], fuse_mode="none", device=self.device,
                       seed=seed, precision=self.config.get("precision", "bf16_split"))

    # ---- acting (a7) ---------------------------------------------------------------------------------------
    def compute_actions(self, obs, eps=None, step=0, deterministic=False):
        """obs [M, D] -> (actions [M, 2], logp [M], logits [M, 4]).  No squashing / clipping (the env clips)."""
        logits = self.model.forward(obs)
        actions, logp = ops.gaussian_sample(logits, eps=eps, seed=self.config.get("seed", 0), step=step,
                                            deterministic=deterministic)
        return actions, logp, logits

    def extra_action_out(self, *a, **k):
        return {}

    # ---- loss (a16 / a17) ----------------------------------------------------------------------------------
    def _heads(self, train_batch):
        """[(net name, input column, VF_PREDS-like, target)]"""
        return [("value", OBS, VF_PREDS, VALUE_TARGETS)]

    def _adv_column(self):
        return ADVANTAGES

    def loss(self, model, dist_class, train_batch, global_rows=None):
        """Forward + backward through the kernels; gradients land in `model.grad`.  `global_rows`: the row count the
        means are taken over (data parallel: the GLOBAL minibatch size, so that summing the ranks' gradients and
        losses gives the whole-minibatch mean; default: this batch's rows)."""
        cfg = dict(clip_param=self.config["clip_param"], vf_clip_param=self.config["vf_clip_param"],
                   vf_loss_coeff=self.config["vf_loss_coeff"], entropy_coeff=self.entropy_coeff,
                   kl_coeff=self.kl_coeff, old_value_loss=self.config.get("old_value_loss", True),
                   dyn_coeffs=self._sync_coeffs())
        B = train_batch[OBS].shape[0]
        rows = int(global_rows) if global_rows else B
        if B == 0:                                   # a rank without rows in this minibatch contributes nothing
            m = torch.zeros(8, dtype=torch.float64, device=self.device)
            return self._fill_tower_stats(model, m, cfg, train_batch)
        pol = model.nets["policy"]
        tc = model._tc()
        # one [hi | lo] operand split per distinct input tensor (CoPO: obs is the critic input of all four networks)
        splits = {}

        def split_of(t):
            if tc is None:
                return None
            key = (t.data_ptr(), tuple(t.shape))
            if key not in splits:
                # ones column where the input has a padding column (the first layer's bias gradient rides in its weight
                # gradient, models._Net.backward); the forward weights are zero in the padding
                splits[key] = ops.tc_split_rows(t, ones_col=ops.tc_has_ones_col(t.shape[1]) and t.shape[1] <= 256)
            return splits[key]

        acts_p = pol.forward_train(train_batch[OBS], tc, split_of(train_batch[OBS]))
        head_acts, heads = [], []
        for net_name, in_col, old_col, tgt_col in self._heads(train_batch):
            acts = model.nets[net_name].forward_train(train_batch[in_col], tc, split_of(train_batch[in_col]))
            head_acts.append((net_name, acts))
            heads.append((acts[3].reshape(-1), train_batch[old_col], train_batch[tgt_col]))
        dlogits, dvs, st = ops.ppo_head(acts_p[3], train_batch[ACTIONS], train_batch[ACTION_LOGP],
                                        train_batch[ACTION_DIST_INPUTS], train_batch[self._adv_column()], heads, cfg,
                                        norm_rows=rows)
        pol.backward(acts_p, dlogits, tc)
        for (net_name, acts), dv in zip(head_acts, dvs):
            model.nets[net_name].backward(acts, dv.unsqueeze(1), tc)
        return self._fill_tower_stats(model, st / rows, cfg, train_batch)

    def _sync_coeffs(self):
        """[kl_coeff, entropy_coeff] on the device, refreshed when the host values changed."""
        cur = (self.kl_coeff, self.entropy_coeff)
        if cur != self._coeffs_host:
            self._coeffs_dev.copy_(torch.tensor(cur, dtype=torch.float32))
            self._coeffs_host = cur
        return self._coeffs_dev

    def _fill_tower_stats(self, model, m, cfg, train_batch):
        c = cfg.get("dyn_coeffs")
        klc, entc = (c[0], c[1]) if c is not None else (cfg["kl_coeff"], cfg["entropy_coeff"])
        total = m[0] + cfg["vf_loss_coeff"] * (m[1] + m[2] + m[3]) - entc * m[4] + klc * m[5]
        ts = model.tower_stats
        ts["total_loss"], ts["mean_policy_loss"], ts["mean_vf_loss"] = total, m[0], m[1]
        ts["mean_entropy"], ts["mean_kl_loss"] = m[4], m[5]
        ts["vf_explained_var"] = torch.zeros((), device=self.device)
        self._extra_tower_stats(model, m, train_batch)
        # the same numbers as one device vector (the trainer accumulates these without a host sync per minibatch):
        # total, policy, vf, nei vf, global vf, entropy, kl, mean logp
        self.stats_vector = torch.stack([total, m[0], m[1], m[2], m[3], m[4], m[5], m[6]])
        return total

    compute_loss = loss                                  # BASELINE.json's name for the same hook

    def _extra_tower_stats(self, model, m, train_batch):
        pass

    def _global_rows(self, B, global_rows):
        """Rows of the global minibatch: given by the trainer's plan, else summed over the ranks here."""
        if global_rows:
            return int(global_rows)
        if parallel.active(self.dist):
            t = torch.tensor([B], dtype=torch.int64, device=self.device)
            self.dist.all_reduce(t)
            return int(t.item())
        return B

    def _grad_step(self, train_batch, rows):
        """zero grads -> loss forward + backward.  Device work only (no host reads): this is what gets captured."""
        self.model.zero_grad()
        self.loss(self.model, None, train_batch, global_rows=rows)

    def _graphed_step(self, train_batch, rows):
        """The minibatch's forward + backward as ONE CUDA-graph launch: ~110 kernel launches and
        the tiny statistics kernels replay without per-launch host work.  Captured once per batch signature (row count,
        columns); the minibatch is copied into the graph's static input buffers by one multi-tensor copy.  Returns False
        when this batch runs eagerly (graphs off, data-dependent shapes, capture failed)."""
        B = train_batch[OBS].shape[0]
        if self._graph_off or B == 0 or self.model._tc() is None:
            return False
        tensors = {k: v for k, v in train_batch.items() if torch.is_tensor(v)}
        key = (B, rows) + tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in tensors.items()))
        g = self._graphs.get(key)
        if g is None:
            # what to capture: full-size minibatches repeat from step to step and iteration to iteration
            # (parallel.minibatch_plan_all); the ragged tail of an epoch comes back once per SGD epoch and then never
            # again, so it runs eagerly.  The trainer names the steady size (graph_batch_rows); without that hint a
            # signature is captured once it has shown up more often than any tail would
            if self.graph_batch_rows is not None:
                if B != self.graph_batch_rows:
                    return False
            else:
                seen = self._graph_seen.get(key, 0)
                if seen < 8:
                    if len(self._graph_seen) > 64:
                        self._graph_seen.clear()
                    self._graph_seen[key] = seen + 1
                    return False
            if len(self._graphs) >= 2:
                self._graphs.clear()
            static, by_ptr = {}, {}
            for k, v in tensors.items():                 # columns that alias one tensor keep aliasing one buffer
                if v.data_ptr() not in by_ptr:
                    by_ptr[v.data_ptr()] = torch.empty_like(v, memory_format=torch.contiguous_format)
                static[k] = by_ptr[v.data_ptr()]
            g = dict(static=static, dst=list(by_ptr.values()), graph=None)
            try:
                self._sync_coeffs()
                torch._foreach_copy_(g["dst"], [tensors[k] for k in self._first_keys(tensors)])
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):            # warm-up outside capture: lazy initialisation, workspaces
                    self._grad_step(static, rows)
                torch.cuda.current_stream(self.device).wait_stream(side)
                for net in self.model.nets.values():     # the weight operands are rebuilt INSIDE the graph on every replay
                    net._tc_version = None
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._grad_step(static, rows)
                g["graph"] = graph
                g["out"] = (self.stats_vector, dict(self.model.tower_stats))
            except Exception as e:                       # pragma: no cover - depends on the driver / torch build
                import warnings
                warnings.warn("learn_on_batch: CUDA-graph capture failed (%r); running eagerly" % (e,))
                self._graph_off = True
                return False
            self._graphs[key] = g
        torch._foreach_copy_(g["dst"], [tensors[k] for k in self._first_keys(tensors)])
        self._sync_coeffs()
        g["graph"].replay()
        self.stats_vector, ts = g["out"]
        self.model.tower_stats.update(ts)
        for net in self.model.nets.values():             # eager users must not trust the operands cached at capture time
            net._tc_version = None
        return True

    @staticmethod
    def _first_keys(tensors):
        """One key per distinct tensor, in dict order (the order the static buffers were made in)."""
        seen, keys = set(), []
        for k, v in tensors.items():
            if v.data_ptr() not in seen:
                seen.add(v.data_ptr())
                keys.append(k)
        return keys

    def learn_on_batch(self, train_batch, global_rows=None):
        """zero grads -> loss fwd+bwd -> (all-reduce) -> Adam.  Returns the stats dict of this minibatch.  Data
        parallel: every rank's loss is normalised by the GLOBAL minibatch row count, so the all-reduced SUM of the
        gradients is the gradient of the whole-minibatch mean (ranks may hold different numbers of rows)."""
        rows = self._global_rows(train_batch[OBS].shape[0], global_rows)
        if not self._graphed_step(train_batch, rows):
            self._grad_step(train_batch, rows)
        parallel.allreduce_sum_(self.model.grad, self.dist, self.ar_timer)  # one NCCL all-reduce of the flat gradient
        self._optimizer.apply(self.model.grad)
        self.model.mark_weights_changed()
        self.num_grad_updates += 1
        return dict(self.model.tower_stats)

    def extra_grad_info(self, train_batch=None):
        ts = self.model.tower_stats
        return {"total_loss": float(ts["total_loss"]), "policy_loss": float(ts["mean_policy_loss"]),
                "vf_loss": float(ts["mean_vf_loss"]), "kl": float(ts["mean_kl_loss"]),
                "entropy": float(ts["mean_entropy"]), "cur_kl_coeff": self.kl_coeff, "cur_lr": self.config["lr"],
                "entropy_coeff": self.entropy_coeff, "vf_explained_var": 0.0}

    def update_kl(self, sampled_kl):
        """rllib KLCoeffMixin.update_kl (ppo_torch_policy, ray 2.2.0)."""
        if sampled_kl > 2.0 * self.config["kl_target"]:
            self.kl_coeff *= 1.5
        elif sampled_kl < 0.5 * self.config["kl_target"]:
            self.kl_coeff *= 0.5
        return self.kl_coeff

    # ---- postprocessing (a8, a11) --------------------------------------------------------------------------
    def _critic_obs(self, ro):
        return ro[OBS].reshape(-1, ro[OBS].shape[-1])

    def _critic_obs_operand(self, ro):
        """(critic observations, the same rows as the value network's tensor-core operand or None)."""
        return self._critic_obs(ro), None

    def _bootstrap_values(self, ro):
        """Stock rllib PPO postprocessing (IPPOPolicy overrides only `loss`, algo_ippo.py:78-79): a trajectory the
        fragment end cuts bootstraps with the value of its NEXT observation."""
        nxt = ro.get("next_obs")
        return None if nxt is None else [self.model.central_value_function(nxt).reshape(-1).contiguous()]

    def postprocess_rollout(self, ro):
        """ro: dict of [T, N, ...] rollout columns (obs, actions, rewards, flags, ...; `next_obs` [N, D] = the
        observation after the last row).  Adds vf_preds, advantages, value_targets (GAE)."""
        T, N = ro["flags"].shape
        cobs, cobs_split = self._critic_obs_operand(ro)
        ro[CENTRALIZED_CRITIC_OBS] = cobs.reshape(T, N, -1)
        ro[VF_PREDS] = self.model.central_value_function(cobs, cobs_split).reshape(T, N)
        adv, tgt = ops.gae3(ro["flags"], [ro[REWARDS]], [ro[VF_PREDS]], self.config["gamma"], self.config["lambda_"],
                            bootstrap=self._bootstrap_values(ro))
        ro[ADVANTAGES], ro[VALUE_TARGETS] = adv[0], tgt[0]
        return ro

    # ---- per-trajectory form: the hook RLlib calls (a8, a13) ------------------------------------------------
    def postprocess_trajectory(self, sample_batch, other_agent_batches=None, episode=None):
        """One agent's trajectory as RLlib hands it to the policy (algo_ccppo.py:322-374, algo_copo.py:473-502; IPPO:
        stock PPO postprocessing): `sample_batch` is a mapping of numpy arrays under SampleBatch's column names
        (`obs`, `new_obs`, `actions`, `rewards`, `dones`, `infos`, `t`), `other_agent_batches` maps agent id ->
        (policy, batch) for the agents of the same episode, `episode=None` marks RLlib's initialisation call (no
        fusion, no infos).  Adds the same columns as the reference; the value heads and the advantage scans run on the
        device (the same kernels as `postprocess_rollout`, which processes all trajectories of a rollout at once and is
        what the trainers use), the neighbour fusion is host logic over the few rows of one trajectory."""
        sb = sample_batch
        obs = np.asarray(sb[OBS], np.float32)
        T = obs.shape[0]
        dev = self.device
        cobs = self._trajectory_critic_obs(sb, obs, other_agent_batches, episode)
        if cobs is not None:
            sb[CENTRALIZED_CRITIC_OBS] = cobs
        cin = torch.as_tensor(cobs if cobs is not None else obs, device=dev).contiguous()
        heads = self._trajectory_heads(sb, cin, episode)            # [(value column, reward column, adv, target)]
        done_last = bool(np.asarray(sb[DONES])[-1])
        flags = torch.full((T, 1), 1, dtype=torch.uint8, device=dev)
        if done_last:
            flags[-1, 0] = 3                                        # VALID | DONE: last_r = 0
        rew, val = [], []
        for vcol, rcol, _, _ in heads:
            rew.append(torch.as_tensor(np.asarray(sb[rcol], np.float32), device=dev).reshape(T, 1).contiguous())
            val.append(torch.as_tensor(sb[vcol], device=dev).reshape(T, 1).contiguous())
        boot = None
        if not done_last and self.algo == "ippo":                   # stock rllib: last_r = V(new_obs[-1])
            nxt = torch.as_tensor(np.asarray(sb["new_obs"], np.float32)[-1:], device=dev)
            boot = [self.model.central_value_function(nxt).reshape(1).contiguous()]
        adv, tgt = ops.gae3(flags, rew, val, self.config["gamma"], self.config["lambda_"], bootstrap=boot)
        for k, (_, _, acol, tcol) in enumerate(heads):
            sb[acol] = adv[k].reshape(T).cpu().numpy()
            sb[tcol] = tgt[k].reshape(T).cpu().numpy()
        if "step_lcf" in sb:
            assert sb["step_lcf"].max() == sb["step_lcf"].min()     # algo_copo.py:501
        return sb

    def _trajectory_critic_obs(self, sb, obs, other_agent_batches, episode):
        return None                                                 # IPPO: the critic sees the agent's own observation

    def _trajectory_heads(self, sb, cin, episode):
        sb[VF_PREDS] = self.model.central_value_function(cin).cpu().numpy().astype(np.float32)
        return [(VF_PREDS, REWARDS, ADVANTAGES, VALUE_TARGETS)]

    def standardize_advantages(self, ro):
        """Stock PPO training_step: standardize_fields(["advantages"]) over the whole train batch."""
        st = ops.lcf_mix_stats(ro["flags"].reshape(-1), ro[ADVANTAGES].reshape(-1), None, None, None)
        st = self._allreduce_stats(st)
        s = st.tolist()
        mean, std = ops.stats_to_mean_std(s[0], s[1], s[2])
        ro[ADVANTAGES] = ops.lcf_mix_apply(ro["flags"].reshape(-1), ro[ADVANTAGES].reshape(-1), None, None, None, mean,
                                           std, 0.0, 1.0).reshape(ro[ADVANTAGES].shape)
        return ro

    def _allreduce_stats(self, st):
        return parallel.allreduce_sum_(st, self.dist, self.ar_timer)


class CCPPOPolicy(IPPOPolicy):
    algo = "ccppo"

    @classmethod
    def default_config(cls):
        return ccppo_config()

    def _build_model(self, seed):
        c = self.config
        return CCModel(self.obs_dim, self.act_dim, c["fcnet_hiddens"], c["fuse_mode"], c["counterfactual"],
                       c["num_neighbours"], device=self.device, seed=seed, precision=c.get("precision", "bf16_split"))

    def _heads(self, train_batch):
        return [("value", CENTRALIZED_CRITIC_OBS, VF_PREDS, VALUE_TARGETS)]

    def _bootstrap_values(self, ro):
        return None      # last_r = VF_PREDS[-1]: the value of the LAST OBSERVED state (algo_ccppo.py:362-365)

    def _critic_obs(self, ro):
        T, N, D = ro[OBS].shape
        mode = self.config["fuse_mode"]
        if mode == "none":
            return ro[OBS].reshape(T * N, D)
        return ops.cc_obs_fuse(ro[OBS].reshape(T * N, D), ro[ACTIONS].reshape(T * N, -1), ro["flags"].reshape(-1),
                               ro["mf_mask"].reshape(-1) if mode == "mf" else None,
                               ro["nei_list"].reshape(T * N, 4) if mode == "concat" else None, ro["slots"], mode,
                               self.config["counterfactual"])

    def _critic_obs_operand(self, ro):
        """Mean-field fusion on the tensor-core path: the fusion kernel also emits the rows as the value network's
        [hi | lo] operand (no conversion pass over the fused observations)."""
        T, N, D = ro[OBS].shape
        if self.config["fuse_mode"] != "mf" or self.model._tc() is None or N % int(ro["slots"]) or int(ro["slots"]) > 64:
            return self._critic_obs(ro), None
        return ops.cc_obs_fuse(ro[OBS].reshape(T * N, D), ro[ACTIONS].reshape(T * N, -1), ro["flags"].reshape(-1),
                               ro["mf_mask"].reshape(-1), None, ro["slots"], "mf", self.config["counterfactual"],
                               want_split=True)

    def _trajectory_critic_obs(self, sb, obs, other_agent_batches, episode):
        """The centralized critic observation of one trajectory (algo_ccppo.py:225-311, 328-355): own observation, then
        the neighbours' observations (+ actions when `counterfactual`) taken from the rows of `other_agent_batches`
        with the same environment time step `t` - the nearest `num_neighbours` in the order the env lists them
        (concat), or the mean over those within `mf_nei_distance` (mean field; `>` skips, so `<=` is kept)."""
        T, odim = obs.shape
        if episode is None:                                # RLlib's initialisation call: the column is already there
            return np.asarray(sb[CENTRALIZED_CRITIC_OBS], np.float32)
        mode, cf = self.config["fuse_mode"], self.config["counterfactual"]
        cobs = np.zeros((T, self.model.cobs_dim), np.float32)
        cobs[:, :odim] = obs
        if mode == "none":
            assert odim == cobs.shape[1]
            return cobs
        assert other_agent_batches is not None
        acts = np.asarray(sb[ACTIONS], np.float32)
        adim = acts.shape[1]
        other = odim + (adim if cf else 0)

        def row_of(name, t):
            if name not in other_agent_batches:
                return None
            _, nb = other_agent_batches[name]
            hit = np.where(np.asarray(nb["t"]) == t)[0]
            if len(hit) > 1:
                raise ValueError("two rows of agent %r share the time step %r" % (name, t))
            return (nb[OBS][hit[0]], nb[ACTIONS][hit[0]]) if len(hit) else None

        for i in range(T):
            info, t = sb["infos"][i], sb["t"][i]
            if mode == "concat":
                for cnt, name in enumerate(info["neighbours"]):
                    if cnt >= self.config["num_neighbours"]:
                        break
                    r = row_of(name, t)
                    if r is not None:
                        start = odim + cnt * other
                        cobs[i, start:start + odim] = r[0]
                        if cf:
                            cobs[i, start + odim:start + other] = r[1]
            else:
                obs_list, act_list = [], []
                for name, dist in zip(info["neighbours"], info["neighbours_distance"]):
                    if dist > self.config["mf_nei_distance"]:
                        continue
                    r = row_of(name, t)
                    if r is not None:
                        obs_list.append(r[0])
                        act_list.append(r[1])
                if obs_list:
                    cobs[i, odim:2 * odim] = np.mean(np.asarray(obs_list, np.float32), axis=0)
                    if cf:
                        cobs[i, 2 * odim:2 * odim + adim] = np.mean(np.asarray(act_list, np.float32), axis=0)
        return cobs


class CoPOPolicy(CCPPOPolicy):
    algo = "copo"

    def __init__(self, obs_dim, act_dim=2, config=None, device=None, dist=None):
        super().__init__(obs_dim, act_dim, config, device, dist)
        self.target_model = self._build_model(self.config.get("seed", 0))        # algo_copo.py:210-221
        self.update_old_policy()
        self._lcf_optimizer = _Adam(self.model.lcf_parameters, self.config["lcf_lr"])
        self._raw_lcf_adv_mean, self._raw_lcf_adv_std = 0.0, 1.0

    @classmethod
    def default_config(cls):
        return copo_config()

    def _build_model(self, seed):
        c = self.config
        return CoPOModel(self.obs_dim, self.act_dim, c["fcnet_hiddens"], c["fuse_mode"], c["counterfactual"],
                         c["num_neighbours"], c["initial_lcf_std"], c["use_distributional_lcf"], device=self.device,
                         seed=seed, precision=c.get("precision", "bf16_split"))

    def _heads(self, train_batch):
        return [("value", CENTRALIZED_CRITIC_OBS, VF_PREDS, VALUE_TARGETS),
                ("nei", CENTRALIZED_CRITIC_OBS, NEI_VALUES, NEI_TARGET),
                ("global", CENTRALIZED_CRITIC_OBS, GLOBAL_VALUES, GLOBAL_TARGET)]

    def _adv_column(self):
        return "normalized_advantages"

    def _extra_tower_stats(self, model, m, train_batch):
        ts = model.tower_stats
        ts["lcf"], ts["lcf_std"] = model.lcf_mean, model.lcf_std
        ts["mean_nei_vf_loss"], ts["mean_global_vf_loss"] = m[2], m[3]
        ts["normalized_advantages"] = train_batch["normalized_advantages"].mean()

    def extra_grad_info(self, train_batch=None):
        ret = super().extra_grad_info(train_batch)
        ts = self.model.tower_stats
        ret.update(lcf=float(ts["lcf"]), lcf_std=float(ts["lcf_std"]), mean_nei_vf_loss=float(ts["mean_nei_vf_loss"]),
                   mean_global_vf_loss=float(ts["mean_global_vf_loss"]),
                   normalized_advantages=float(ts["normalized_advantages"]))
        return ret

    # ---- postprocessing (a13, a12, a15) --------------------------------------------------------------------
    def _trajectory_heads(self, sb, cin, episode):
        """algo_copo.py:473-500: three value heads; neighbourhood / global rewards and the step LCF come out of the
        infos the LCF environment filled (env_wrappers.py:313-357)."""
        heads = super()._trajectory_heads(sb, cin, episode)
        sb[NEI_VALUES] = self.model.get_nei_value(cin).cpu().numpy().astype(np.float32)
        sb[GLOBAL_VALUES] = self.model.get_global_value(cin).cpu().numpy().astype(np.float32)
        if episode is not None:
            infos = sb["infos"]
            assert isinstance(infos[0], dict)
            sb[NEI_REWARDS] = np.array([info[NEI_REWARDS] for info in infos]).astype(np.float32)
            sb[GLOBAL_REWARDS] = np.array([info[GLOBAL_REWARDS] for info in infos]).astype(np.float32)
            if self.config["use_distributional_lcf"]:
                sb["step_lcf"] = np.array([info["lcf"] for info in infos]).astype(np.float32)
        return heads + [(NEI_VALUES, NEI_REWARDS, NEI_ADVANTAGE, NEI_TARGET),
                        (GLOBAL_VALUES, GLOBAL_REWARDS, GLOBAL_ADVANTAGES, GLOBAL_TARGET)]

    def postprocess_rollout(self, ro):
        T, N = ro["flags"].shape
        cobs, sp = self._critic_obs_operand(ro)
        if sp is None and self.model._tc() is not None:
            sp = ops.tc_split_rows(cobs)                 # one [hi | lo] operand for the three value networks
        ro[CENTRALIZED_CRITIC_OBS] = cobs.reshape(T, N, -1)
        ro[VF_PREDS] = self.model.central_value_function(cobs, sp).reshape(T, N)
        ro[NEI_VALUES] = self.model.get_nei_value(cobs, sp).reshape(T, N)
        ro[GLOBAL_VALUES] = self.model.get_global_value(cobs, sp).reshape(T, N)
        per_scene = ro["slots"] if ro[GLOBAL_REWARDS].shape[1] != N else 0
        adv, tgt = ops.gae3(ro["flags"], [ro[REWARDS], ro[NEI_REWARDS], ro[GLOBAL_REWARDS]],
                            [ro[VF_PREDS], ro[NEI_VALUES], ro[GLOBAL_VALUES]], self.config["gamma"],
                            self.config["lambda_"], global_reward_per_scene=per_scene)
        ro[ADVANTAGES], ro[NEI_ADVANTAGE], ro[GLOBAL_ADVANTAGES] = adv
        ro[VALUE_TARGETS], ro[NEI_TARGET], ro[GLOBAL_TARGET] = tgt
        return ro

    def standardize_advantages(self, ro):
        """CoPOTrainer.training_step:539-551: LCF mix with the per-row step_lcf, whole-batch standardisation of the
        mixed and of the global advantage; the raw mean / std are kept for the meta update."""
        f = ro["flags"].reshape(-1)
        a, n, l = ro[ADVANTAGES].reshape(-1), ro[NEI_ADVANTAGE].reshape(-1), ro["step_lcf"].reshape(-1)
        g = ro[GLOBAL_ADVANTAGES].reshape(-1)
        s = self._allreduce_stats(ops.lcf_mix_stats(f, a, n, l, g)).tolist()
        mean, std = ops.stats_to_mean_std(s[0], s[1], s[2])
        gmean, gstd = ops.stats_to_mean_std(s[3], s[4], s[2])
        self._raw_lcf_adv_mean, self._raw_lcf_adv_std = mean, max(1e-4, std)
        ro["normalized_advantages"] = ops.lcf_mix_apply(f, a, n, l, g, mean, std, gmean, gstd).reshape(
            ro[ADVANTAGES].shape)
        return ro

    # ---- meta-gradient (a18) -------------------------------------------------------------------------------
    def _policy_grad(self, model, batch, mode, adv, rows, out, x_split=None, st=None):
        """Gradient of the policy network only, into `out` (this rank's share of the global-minibatch mean).  st: zeroed
        float64 [8] the loss kernel accumulates its statistics in."""
        cfg = dict(clip_param=self.config["clip_param"], vf_clip_param=0.0, vf_loss_coeff=0.0, entropy_coeff=0.0,
                   kl_coeff=0.0)
        pol = model.nets["policy"]
        model.grad[model.policy_slice()].zero_()
        if st is None:
            st = torch.zeros(8, dtype=torch.float64, device=self.device)
        if batch[OBS].shape[0] > 0:
            acts = pol.forward_train(batch[OBS], model._tc(), x_split)
            dlogits, _, st = ops.ppo_head(acts[3], batch[ACTIONS], batch[ACTION_LOGP], None, adv, [], cfg, mode=mode,
                                          norm_rows=rows, stats=st)
            pol.backward(acts, dlogits, model._tc())
        out.copy_(model.grad[model.policy_slice()])
        return st

    def _meta_grads(self, batch, rows, g_new, g_old, st_new=None, st_old=None):
        """Both policy-gradient vectors of the meta step (algo_copo.py:236-262): the surrogate of the CURRENT policy on
        the global advantage, the log-probability gradient of the OLD policy.  Device work only."""
        sp = None                                        # both networks read the same [hi | lo] operand of the observations
        if self.model._tc() is not None and batch[OBS].shape[0] > 0:
            d = self.model.nets["policy"].in_dim
            sp = ops.tc_split_rows(batch[OBS], ones_col=ops.tc_has_ones_col(d) and d <= 256)
        st_new = self._policy_grad(self.model, batch, 0, batch[GLOBAL_ADVANTAGES], rows, g_new, sp, st_new)
        st_old = self._policy_grad(self.target_model, batch, 1, None, rows, g_old, sp, st_old)
        return st_new, st_old

    META_KEYS = ("new_policy_ego_loss", "old_policy_logp_loss", "lcf_lcf_adv_loss", "lcf_final_loss", "grad_value", "lcf",
                 "lcf_deg", "lcf_param", "coordinated_adv", "global_adv", "lcf_std", "lcf_std_deg", "lcf_std_param")

    def meta_update(self, train_batch, eps=None, global_rows=None):
        """algo_copo.py:228-309.  Device work: two policy gradients (new policy on the global advantage, old policy's
        log-probability), their dot product, one pass over the advantage columns for the LCF sums, and ONE kernel for
        everything after that (LCF loss and gradient, Adam on lcf_parameters, the logged statistics) - no host read
        unless `sync_stats`."""
        B = train_batch[OBS].shape[0]
        rows = self._global_rows(B, global_rows)
        n = self.model.policy_slice().stop - self.model.policy_slice().start
        if getattr(self, "_meta_g", None) is None:
            self._meta_g = torch.empty(2 * n, dtype=torch.float32, device=self.device)
            # [0:8] statistics of the new policy's loss, [8:16] of the old policy's, [16:20] LCF sums, [20] the dot product
            self._meta_buf = torch.zeros(21, dtype=torch.float64, device=self.device)
        g_new, g_old = self._meta_g[:n], self._meta_g[n:]
        buf = self._meta_buf
        buf.zero_()
        st_new, st_old, sums, grad_value = buf[0:8], buf[8:16], buf[16:20], buf[20:21]
        self._meta_grads(train_batch, rows, g_new, g_old, st_new, st_old)
        # reduce BOTH gradient vectors (one all-reduce) BEFORE the dot product: it is bilinear
        parallel.allreduce_sum_(self._meta_g, self.dist, self.ar_timer)
        if eps is None:
            eps = torch.randn(B, dtype=torch.float32, device=self.device)
        if B > 0:
            ops.lcf_meta_sums(train_batch[ADVANTAGES], train_batch[NEI_ADVANTAGE], eps, train_batch[GLOBAL_ADVANTAGES],
                              self.model.lcf_parameters, sums)
        # ranks can hold minibatches of different sizes: the sums are reduced, the means taken over the global row count
        if parallel.active(self.dist):
            self._allreduce_stats(buf[:20])
        ops.dot(g_new, g_old, out=grad_value)            # after the reduction above: identical on every rank already
        opt = self._lcf_optimizer
        opt.step += 1
        vec = torch.empty(13, dtype=torch.float64, device=self.device)
        ops.lcf_meta_finish(grad_value, st_new, st_old, sums, rows, self._raw_lcf_adv_mean, self._raw_lcf_adv_std,
                            self.model.lcf_parameters, opt.m, opt.v, self.model.lcf_grad, vec, opt.lr, opt.step)
        self.meta_stats_keys, self.meta_stats_vector = self.META_KEYS, vec   # device copy (no host sync) for the trainer
        if not self.sync_stats:
            return {}
        return dict(zip(self.META_KEYS, vec.tolist()))                    # one device -> host read

    sync_stats = True      # False: meta_update leaves its statistics on the device (meta_stats_vector) and returns {}

    def update_old_policy(self):
        self.target_model.flat.copy_(self.model.flat)
        self.target_model.mark_weights_changed()
        self.target_model.lcf_parameters.copy_(self.model.lcf_parameters)

    def assign_lcf(self, lcf_parameters, lcf_mean, lcf_std=None, my_name=None):
        """algo_copo.py:446-471 (same assertions)."""
        lcf_parameters = torch.as_tensor(lcf_parameters, dtype=torch.float32).to(self.device)
        assert self.model.lcf_parameters.size() == lcf_parameters.size()
        self.model.lcf_parameters.copy_(lcf_parameters)
        assert abs(float(self.model.lcf_mean) - lcf_mean) < 1e-5
        if lcf_std is not None:
            assert abs(float(self.model.lcf_std) - lcf_std) < 1e-5
