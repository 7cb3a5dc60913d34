"""IPPO / CCPPO / CoPO trainers over device-resident rollouts.

Mirrors torch_copo/algo_{ippo,ccppo,copo}.py `*Trainer`: `get_default_config()`, `get_default_policy_class()`,
`training_step()` (CoPO's: algo_copo.py:516-661) and `train()`.  One process drives one GPU; every rank holds
`num_scenes` scenes (scene sharding, no data-path collective during the rollout).  Per iteration:

  sample      T env steps: policy forward + Gaussian sample + fused scene step, the env writes straight into the
              [T, N] rollout columns                                            (synchronous_parallel_sample, :518-525)
  postprocess critic obs, value heads, GAE x3                                   (postprocess_trajectory, :473-502)
  mix         LCF-mixed advantage, whole-batch standardisation (all-reduce of 5 sums)                 (:539-551)
  sgd         num_sgd_iter epochs of shuffled minibatches, one gradient all-reduce + Adam each (train_one_step, :555)
  meta        lcf_num_iters epochs of meta_update minibatches                                         (:582-589)
  hand-over   theta_old <- theta, env.set_lcf_dist(mean, std), KL coefficient update                  (:596-632)
"""
import math
import time

import torch

from . import ops
from . import parallel
from . import policy as P
from .batched_env import (BatchedDrivingEnv, FLAG_ARRIVE, FLAG_CRASH, FLAG_DONE, FLAG_MAXSTEP, FLAG_OUT, FLAG_VALID,
                          MAP_OF_ENV)

SCALAR_COLUMNS = (P.ACTION_LOGP, P.ADVANTAGES, P.VALUE_TARGETS, P.VF_PREDS, "normalized_advantages", P.NEI_VALUES,
                  P.NEI_TARGET, P.NEI_ADVANTAGE, P.GLOBAL_VALUES, P.GLOBAL_TARGET, P.GLOBAL_ADVANTAGES)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def resolve_env(env):
    """`config["env"]` as the reference's scripts pass it - a name registered through `get_rllib_compatible_env`
    (torch_copo/utils/env_wrappers.py:559-597), an env class, a `MultiAgent*Env` class name or a map name - to
    (map name, appends the LCF to the observation or None if the name does not say, default env config)."""
    from . import envs as E
    cls = env if isinstance(env, type) else E._REGISTRY.get(env)
    if cls is None and isinstance(env, str) and hasattr(E, env) and isinstance(getattr(E, env), type):
        cls = getattr(E, env)
    if cls is not None:
        in_registry = isinstance(env, type) or env in E._REGISTRY
        return cls.MAP, (bool(cls.APPEND_LCF) if in_registry else None), cls.default_config()
    return MAP_OF_ENV.get(env, env), None, {}


class IPPOTrainer:
    policy_cls = P.IPPOPolicy

    @classmethod
    def get_default_config(cls):
        c = cls.policy_cls.default_config()
        c.update_from_dict(dict(env="MultiAgentIntersectionEnv", num_scenes=None, sgd_minibatch_size=512,
                                rollout_fragment_length=200))
        return c

    def get_default_policy_class(self, config=None):
        return self.policy_cls

    def __init__(self, config=None, env=None, device=None, logger_creator=None, **_unused):
        """`config` may be an RLlib-style dict as the reference's scripts build it (train_copo.py:19-48): `env` is a
        registered env name, `env_config` its overrides, `seed` may be None, `callbacks` a class with
        `on_train_result`; RLlib resource keys (`num_gpus`, `num_cpus_per_worker`, ...) and the dead legacy keys the
        scripts still pass (`initial_svo_std`, `svo_lr`, ...) are carried along and ignored.  Scenes per GPU:
        `num_scenes`, else $B2C_NUM_SCENES, else train_batch_size / rollout_fragment_length (the number of concurrent
        envs of the reference: 5 workers x 200-step fragments of a 2000-step batch -> 10)."""
        import os
        cfg = self.get_default_config()
        if config:
            cfg.update_from_dict(dict(config))
        if cfg.get("seed") is None:
            cfg["seed"] = 0
        self.config = cfg
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        cb = cfg.get("callbacks")
        self.callbacks = cb() if isinstance(cb, type) else (cb or None)
        if env is None:
            map_name, env_lcf, ec = resolve_env(cfg["env"])
            ec = dict(ec, **dict(cfg.get("env_config") or {}))
            if not cfg.get("num_scenes"):
                cfg["num_scenes"] = int(os.environ.get("B2C_NUM_SCENES", 0)) or max(
                    1, int(cfg["train_batch_size"]) // int(cfg["rollout_fragment_length"]))
            S = int(cfg["num_scenes"])
            append_lcf = env_lcf if env_lcf is not None else self.policy_cls.algo == "copo"
            if self.policy_cls.algo == "copo" and not append_lcf:
                raise ValueError("CoPO needs an LCF environment: wrap the env class with get_lcf_env "
                                 "(torch_copo/train_copo.py:30)")
            env = BatchedDrivingEnv(map_name, num_scenes=S, num_agents=ec.get("num_agents"),
                                    num_slots=ec.get("num_agents"), seed=int(ec.get("start_seed", cfg.get("seed", 0))),
                                    scene_offset=parallel.scene_offset(self.rank, S), append_lcf=append_lcf,
                                    neighbours_distance=float(ec.get("neighbours_distance", 40.0)),
                                    mf_nei_distance=float(cfg.get("mf_nei_distance", 10.0)),
                                    lcf_std=float(ec.get("lcf_normal_std", 0.1)), horizon=int(ec.get("horizon", 1000)),
                                    delay_done=int(ec.get("delay_done", 25)),
                                    lcf_uniform=(ec.get("lcf_dist") == "uniform"),
                                    force_lcf=float(ec.get("force_lcf", -100.0)), map_kwargs=ec.get("map_config"),
                                    device=self.device)
        self.env = env
        self.policy = self.policy_cls(env.D, 2, cfg, device=self.device, dist=dist)
        if dist is not None and self.world > 1:
            self.policy.ar_timer = parallel.AllReduceTimer()
        self._counters = {"num_env_steps_sampled": 0, "num_agent_steps_sampled": 0}
        self._timers = {}
        self._iteration = 0
        self._step_counter = 0
        self._alloc_rollout()
        # the env also emits each observation as the policy's [hi | lo] bf16 operand (two buffers, ping-pong)
        self._split = [self.env.alloc_obs_split() for _ in range(2)] if self.policy.model.precision == "bf16_split" else None
        first = dict(self.env.out)
        if self._split:
            first["obs_split"] = self._split[0]
        self.env.reset(out=first)
        self.ro[P.OBS][0].copy_(self.env.out["obs"].reshape(self.N, -1))

    def reset_scenes(self):
        """Restarts every scene (new episode) - used by the curriculum after a population change."""
        first = dict(self.env.out)
        if self._split:
            first["obs_split"] = self._split[self._step_counter % 2]
        self.env.reset(out=first, new_episode=True)
        slot = self.T if self._iteration > 0 else 0
        self.ro[P.OBS][slot].copy_(self.env.out["obs"].reshape(self.N, -1))

    def get_policy(self, policy_id="default"):
        return self.policy

    # ---- rollout storage: time-major columns the env and the policy write into directly -----------------------
    def _alloc_rollout(self):
        T = int(self.config["rollout_fragment_length"])
        S, A, D = self.env.S, self.env.A, self.env.D
        N, dev = S * A, self.device
        self.T, self.N = T, N
        z = lambda shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=dev)
        self.ro = {P.OBS: z((T + 1, N, D)), P.ACTIONS: z((T, N, 2)), P.ACTION_LOGP: z((T, N)),
                   P.ACTION_DIST_INPUTS: z((T, N, 4)), P.REWARDS: z((T, N)), "flags": z((T, N), torch.uint8),
                   P.NEI_REWARDS: z((T, N)), P.GLOBAL_REWARDS: z((T, S)), "step_lcf": z((T, N)),
                   "mf_mask": z((T, N), torch.int64), "nei_mask": z((T, N), torch.int64),
                   "nei_list": z((T, N, 4), torch.int8), "agent_id": z((T, N), torch.int32),
                   "scene_done": z((T, S), torch.uint8)}
        # optional outputs the scene step only computes when asked: the mean-field mask and the nearest-neighbour list
        # feed the critic-obs fusion of CCPPO (algo_ccppo.py:225-311) and nothing else
        fuse = self.config.get("fuse_mode", "none") if self.policy_cls.algo != "ippo" else "none"
        self._step_out = []
        for t in range(T):
            r = self.ro
            self._step_out.append(dict(
                obs=r[P.OBS][t + 1].view(S, A, D), reward=r[P.REWARDS][t].view(S, A), flags=r["flags"][t].view(S, A),
                nei_mask=r["nei_mask"][t].view(S, A), mf_mask=r["mf_mask"][t].view(S, A) if fuse == "mf" else None,
                nei_reward=r[P.NEI_REWARDS][t].view(S, A), global_reward=r[P.GLOBAL_REWARDS][t],
                nei_list=r["nei_list"][t].view(S, A, 4) if fuse == "concat" else None,
                agent_id=r["agent_id"][t].view(S, A),
                lcf=r["step_lcf"][t].view(S, A), scene_done=r["scene_done"][t]))

    def sample(self):
        """One rollout fragment of T env steps for every scene of this rank."""
        ro, pol = self.ro, self.policy
        if self._iteration > 0:
            ro[P.OBS][0].copy_(ro[P.OBS][self.T])        # the fragment continues where the last one stopped
        for t in range(self.T):
            g = self._step_counter
            out = self._step_out[t]
            sp = None
            if self._split:
                sp = self._split[g % 2].view(self.N, -1)
                out["obs_split"] = self._split[(g + 1) % 2]
            # policy forward + Gaussian sample write straight into the rollout columns
            pol.model.forward_sample(ro[P.OBS][t], self.config.get("seed", 0) + self.rank * 7919, g, obs_split=sp,
                                     out=ro[P.ACTION_DIST_INPUTS][t], actions=ro[P.ACTIONS][t],
                                     logp=ro[P.ACTION_LOGP][t])
            self.env.step(ro[P.ACTIONS][t].view(self.env.S, self.env.A, 2), out=out)
            self._step_counter += 1
        view = {k: (v[:self.T] if k == P.OBS else v) for k, v in ro.items()}
        view["next_obs"] = ro[P.OBS][self.T]             # the observation after the fragment's last row
        view["slots"] = self.env.A
        return view

    # ---- minibatches -----------------------------------------------------------------------------------------
    def _flatten(self, ro):
        R = self.T * self.N
        cols = [c for c in SCALAR_COLUMNS if c in ro]
        scal = torch.stack([ro[c].reshape(R) for c in cols], dim=1).contiguous()
        wide = {P.OBS: ro[P.OBS].reshape(R, -1), P.ACTIONS: ro[P.ACTIONS].reshape(R, 2),
                P.ACTION_DIST_INPUTS: ro[P.ACTION_DIST_INPUTS].reshape(R, 4)}
        cobs = ro[P.CENTRALIZED_CRITIC_OBS].reshape(R, -1)
        if cobs.data_ptr() != wide[P.OBS].data_ptr():
            wide[P.CENTRALIZED_CRITIC_OBS] = cobs
        valid = torch.nonzero(ro["flags"].reshape(R) & FLAG_VALID).reshape(-1)
        return cols, scal, wide, valid

    def _minibatches(self, cols, scal, wide, valid, mb):
        """Shuffled minibatches over the valid rows (rllib.utils.sgd.minibatches).  Yields (batch, global_rows): with
        more than one rank the GLOBAL train batch is cut into ceil(total / mb) minibatches and every rank contributes
        a near-equal slice of its own shuffled rows to each (parallel.minibatch_plan)."""
        bounds, sizes = parallel.minibatch_plan(self._row_counts, self.rank, mb)
        perm = valid[torch.randperm(valid.numel(), device=self.device)]
        for (lo, hi), rows in zip(bounds, sizes):
            idx = perm[lo:hi].contiguous()
            batch = {c: ops.gather_rows(w, idx) for c, w in wide.items()}
            if P.CENTRALIZED_CRITIC_OBS not in batch:
                batch[P.CENTRALIZED_CRITIC_OBS] = batch[P.OBS]
            s = ops.gather_cols(scal, idx)               # [columns, rows]: every scalar column a contiguous vector
            for n, c in enumerate(cols):
                batch[c] = s[n]
            yield batch, rows

    # ---- one training iteration ------------------------------------------------------------------------------
    LEARNER_KEYS = ("total_loss", "policy_loss", "vf_loss", "mean_nei_vf_loss", "mean_global_vf_loss", "entropy", "kl",
                    "mean_logp")

    def _learn(self, ro):
        cols, scal, wide, valid = self._flatten(ro)
        self._row_counts = parallel.gather_row_counts(valid.numel(), _dist(), self.device)
        self.policy.graph_batch_rows = max(1, int(self.config["sgd_minibatch_size"]) // max(1, len(self._row_counts)))
        acc, n = torch.zeros(len(self.LEARNER_KEYS), dtype=torch.float64, device=self.device), 0
        for _ in range(int(self.config["num_sgd_iter"])):
            for batch, rows in self._minibatches(cols, scal, wide, valid, int(self.config["sgd_minibatch_size"])):
                self.policy.learn_on_batch(batch, global_rows=rows)
                acc += self.policy.stats_vector          # stays on the device: no host sync per minibatch
                n += 1
        # every rank holds its share of each global-minibatch mean: the all-reduced sum is the learner statistic of
        # the whole batch, identical on every rank (so is the KL coefficient derived from it)
        acc = parallel.allreduce_sum_(acc / max(n, 1), _dist(), self.policy.ar_timer)
        stats = dict(zip(self.LEARNER_KEYS, acc.tolist()))
        pol = self.policy
        stats.update(cur_kl_coeff=pol.kl_coeff, cur_lr=pol.config["lr"], entropy_coeff=pol.entropy_coeff,
                     vf_explained_var=0.0)
        if pol.algo == "copo":
            stats.update(lcf=float(pol.model.lcf_mean), lcf_std=float(pol.model.lcf_std))
        else:
            stats.pop("mean_nei_vf_loss"), stats.pop("mean_global_vf_loss")
        return stats, (cols, scal, wide, valid)

    def _episode_metrics(self, ro):
        f = ro["flags"]
        done = (f & FLAG_DONE) > 0
        n = max(int(done.sum()), 1)
        rate = lambda bit: float(((f & bit) > 0)[done].sum()) / n
        valid = (f & FLAG_VALID) > 0
        steps, eps = int(valid.sum()), int(done.sum())
        rsum = float((ro[P.REWARDS] * valid).sum())
        m = dict(success_rate=rate(FLAG_ARRIVE), crash_rate=rate(FLAG_CRASH), out_of_road_rate=rate(FLAG_OUT),
                 max_step_rate=rate(FLAG_MAXSTEP), episodes=eps, step_reward_mean=rsum / max(steps, 1),
                 agent_steps=steps,
                 # per-agent episode length / reward / cost of this fragment in steady state: totals over the
                 # fragment divided by the episodes that ended in it
                 episode_length=steps / n, episode_reward=rsum / n,
                 episode_cost=float((((f & FLAG_CRASH) > 0) & valid).sum()) / n)
        # the names RLlib gives the means of MultiAgentDrivingCallbacks' custom metrics (callbacks.py:48-147)
        for k in ("success_rate", "crash_rate", "out_of_road_rate", "max_step_rate", "episode_length", "episode_reward",
                  "episode_cost", "step_reward"):
            m[k + "_mean"] = m.get(k, m.get(k + "_mean"))
        return m

    def training_step(self):
        t0 = time.perf_counter()
        ro = self.sample()
        torch.cuda.synchronize(self.device)
        t1 = time.perf_counter()
        ro = self.policy.postprocess_rollout(ro)
        ro = self.policy.standardize_advantages(ro)
        learner_stats, flat = self._learn(ro)
        torch.cuda.synchronize(self.device)
        t2 = time.perf_counter()
        results = {"default": {"learner_stats": learner_stats, "custom_metrics": {}}}
        self._after_sgd(ro, flat, results)
        self.policy.update_kl(learner_stats["kl"])       # the all-reduced mean KL: the same coefficient on every rank
        metrics = self._episode_metrics(ro)
        self._counters["num_env_steps_sampled"] += self.T * self.env.S
        self._counters["num_agent_steps_sampled"] += metrics["agent_steps"]
        self._timers = {"sample_time_ms": (t1 - t0) * 1e3, "learn_time_ms": (t2 - t1) * 1e3,
                        "sample_throughput": metrics["agent_steps"] / max(t1 - t0, 1e-9)}
        if self.policy.ar_timer is not None:
            ms, cnt = self.policy.ar_timer.total_ms()
            self._timers.update(allreduce_ms=ms, allreduces=cnt)
        results["default"]["custom_metrics"].update(metrics)
        self._iteration += 1
        return results

    def _after_sgd(self, ro, flat, results):
        pass

    def train(self):
        t0 = time.perf_counter()
        res = self.training_step()
        info = res["default"]
        cm = info["custom_metrics"]
        return {"training_iteration": self._iteration, "time_this_iter_s": time.perf_counter() - t0,
                "timesteps_total": self._counters["num_env_steps_sampled"] * self.world,
                "agent_timesteps_total": self._counters["num_agent_steps_sampled"] * self.world,
                "episodes_this_iter": cm.get("episodes", 0), "episode_reward_mean": cm.get("episode_reward", 0.0),
                "episode_len_mean": cm.get("episode_length", 0.0), "policy_reward_mean": {},
                "info": {"learner": {"default": info}}, "timers": dict(self._timers),
                "custom_metrics": cm, "success": cm.get("success_rate", 0.0)}

    def save(self, checkpoint_dir="."):
        """Writes `checkpoint-<iteration>` in the trial-checkpoint layout the reference's evaluator reads
        (copo/eval/get_policy_function_from_checkpoint.py:12-50); returns its path."""
        import os
        from . import checkpoint as C
        d = os.path.join(checkpoint_dir, "checkpoint_%06d" % self._iteration)
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "checkpoint-%d" % self._iteration)
        C.save_rllib_checkpoint(path, self.policy.model.state_dict())
        return path

    def stop(self):
        self.env.close()

    cleanup = stop


class CCPPOTrainer(IPPOTrainer):
    policy_cls = P.CCPPOPolicy


class CoPOTrainer(CCPPOTrainer):
    policy_cls = P.CoPOPolicy

    def _after_sgd(self, ro, flat, results):
        """LCF meta update + hand-over (algo_copo.py:579-624)."""
        cols, scal, wide, valid = flat
        pol = self.policy
        mb = int(self.config["lcf_sgd_minibatch_size"] or self.config["sgd_minibatch_size"])
        acc, n = None, 0
        pol.sync_stats = False                           # meta_update leaves its statistics on the device
        try:
            for _ in range(int(self.config["lcf_num_iters"])):
                for batch, rows in self._minibatches(cols, scal, wide, valid, mb):
                    pol.meta_update(batch, global_rows=rows)
                    acc = pol.meta_stats_vector.clone() if acc is None else acc + pol.meta_stats_vector
                    n += 1
        finally:
            pol.sync_stats = True
        avg, last = {}, {}
        if n:
            mean_v, last_v = (acc / n).tolist(), pol.meta_stats_vector.tolist()
            avg = {k: v for k, v in zip(pol.meta_stats_keys, mean_v) if k not in ("lcf", "lcf_std")}
            last = {k: v for k, v in zip(pol.meta_stats_keys, last_v) if k in ("lcf", "lcf_std")}
        lcf_mean, lcf_std = float(pol.model.lcf_mean), float(pol.model.lcf_std)
        pol.assign_lcf(pol.model.lcf_parameters.clone(), lcf_mean, lcf_std)
        pol.update_old_policy()
        self.env.set_lcf_dist(mean=lcf_mean, std=lcf_std)
        fetches = {"raw_lcf_adv_mean_value": pol._raw_lcf_adv_mean, "raw_lcf_adv_std_value": pol._raw_lcf_adv_std}
        fetches.update(avg)
        fetches.update(last)
        results["default"]["custom_metrics"]["meta_update"] = fetches
