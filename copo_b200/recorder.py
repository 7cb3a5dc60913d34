"""Episode report of the reference's evaluation wrapper over the dict-API environments.

The reference evaluates a policy by wrapping one environment in `RecorderEnv` (copo/eval/recoder.py:73-349) and calling
`get_episode_result()` when the scene ends; the result is one row of the CSV files its plotting notebooks read
(eval.py:205-232).  This is that wrapper for `copo_b200.envs` - same constructor, `reset` / `step`, `get_step_result`,
`get_episode_result`, same 30 keys and the same quirks (an agent without neighbours counts its own reward as the
neighbourhood reward, recoder.py:62-63; rates are over the agents that FINISHED; neighbours are taken from the vehicles
still on the road within `eval_config["neighbours_distance"]` = 20 m).  Instead of the reference's nested per-step
dictionaries it keeps running per-agent sums and per-step means, so a 1000-step scene costs O(agents) memory.

The tensor path for many scenes at once is `copo_b200.evaluate.evaluate`; this class is the drop-in for scripts written
against the reference's `RecorderEnv` (eval.py, new_vis.py).
"""
import math

import numpy as np


class RecorderEnv:
    _default_eval_config = dict(neighbours_distance=20)

    def __init__(self, env, eval_config=None):
        self.env = env
        self.eval_config = dict(self._default_eval_config, **(eval_config or {}))
        self.episode_step = 0
        self._begin()

    # gym.Wrapper behaviour: everything else is the wrapped environment's
    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env

    def reset(self, *a, **k):
        o = self.env.reset(*a, **k)
        self.episode_step = 0
        return o

    def step(self, actions):
        o, r, d, i = self.env.step(actions)
        if self.episode_step == 0:
            self._begin()
        self._record_step(r, i)
        for k, done in d.items():
            if k != "__all__" and done:
                self._record_end(k, i[k])
        self.episode_step += 1
        return o, r, d, i

    # ---- bookkeeping ------------------------------------------------------------------------------------------
    def _begin(self):
        self._steps = 0
        self._vel_means, self._energy_means, self._nei_means = [], [], []
        self._own, self._nei = {}, {}                     # agent -> summed own / neighbourhood reward
        self._cost, self._ep_reward, self._ep_len = {}, {}, {}
        self._ended = {}                                   # agent -> (success, crash, out, max_step, episode_energy)
        self._last = {}                                    # what get_step_result reports
        self._last_ep_reward_mean = None

    def _neighbours(self, rewards):
        """agent -> ordered neighbour names among the vehicles still on the road (recoder.py:21-45: strict `<`,
        ascending distance, ties in dictionary order)."""
        names = list(self.env.vehicles.keys())
        out = {k: [] for k in rewards}
        if len(names) < 2 or self.eval_config["neighbours_distance"] <= 0:
            return out
        pos = np.array([self.env.vehicles[k].position for k in names], dtype=np.float64)
        dx = pos[:, 0][:, None] - pos[:, 0][None, :]
        dy = pos[:, 1][:, None] - pos[:, 1][None, :]
        dist = np.sqrt(dx * dx + dy * dy)
        for a, k in enumerate(names):
            if k not in out:
                continue
            others = [b for b in range(len(names)) if b != a]
            order = sorted(others, key=lambda b: dist[a, b])               # stable, like the reference's sorted()
            out[k] = [names[b] for b in order if dist[a, b] < self.eval_config["neighbours_distance"]]
        return out

    def _record_step(self, r, i):
        nb = self._neighbours(r)
        vel, energy, nnei, step = [], [], [], {}
        for k, own in r.items():
            others = [r[n] for n in nb[k]]
            nei = float(np.mean(others)) if others else own             # recoder.py:59-63
            self._own[k] = self._own.get(k, 0.0) + own
            self._nei[k] = self._nei.get(k, 0.0) + nei
            nnei.append(len(nb[k]))
            info = i[k]
            if "step_reward" in info:
                vel.append(info["velocity"])
                energy.append(info["step_energy"])
                self._cost[k] = self._cost.get(k, 0.0) + info["cost"]
                self._ep_reward[k] = info["episode_reward"]
                self._ep_len[k] = info["episode_length"]
                for name in ("velocity", "steering", "step_reward", "acceleration", "cost", "episode_length",
                             "episode_reward"):
                    step.setdefault(name, []).append(info[name])
                step.setdefault("energy", []).append(info["step_energy"])
                step.setdefault("raw_action0_l2", []).append(info["raw_action"][0] ** 2)
                step.setdefault("raw_action1_l2", []).append(info["raw_action"][1] ** 2)
            else:                                                         # recoder.py:246,297: `.get(kkk, 0)`
                self._ep_reward[k] = 0
                self._ep_len[k] = 0
            step.setdefault("own_reward", []).append(own)
            step.setdefault("nei_reward", []).append(nei)
            step.setdefault("num_neighbours", []).append(len(nb[k]))
        if vel:
            self._vel_means.append(float(np.mean(vel)))
            self._energy_means.append(float(np.mean(energy)))
        if nnei:
            self._nei_means.append(float(np.mean(nnei)))
        self._steps += 1
        self._last = {name: float(np.mean(v)) for name, v in step.items()}
        if "episode_reward" in step:
            self._last_ep_reward_mean = self._last["episode_reward"]

    def _record_end(self, k, info):
        arrive, crash, out = info.get("arrive_dest", False), info.get("crash", False), info.get("out_of_road", False)
        self._ended[k] = (bool(arrive), bool(crash), bool(out), not (arrive or crash or out), info["episode_energy"])

    # ---- reports ----------------------------------------------------------------------------------------------
    def get_step_result(self):
        ret = dict(self._last)
        if self._last_ep_reward_mean is not None:                         # recoder.py:158-159: the last recorded step's
            ret["episode_reward_mean"] = self._last_ep_reward_mean
        if self._cost:
            ret["episode_cost_mean"] = float(np.mean(list(self._cost.values())))
            ret["episode_cost_sum"] = float(np.sum(list(self._cost.values())))
        return ret

    def get_episode_result(self):
        ret = {}
        ret["velocity_step_mean_episode_min"] = np.min(self._vel_means)
        ret["velocity_step_mean_episode_mean"] = np.mean(self._vel_means)
        ret["velocity_step_mean_episode_max"] = np.max(self._vel_means)
        ret["energy_step_mean_episode_min"] = np.min(self._energy_means)
        ret["energy_step_mean_episode_mean"] = np.mean(self._energy_means)
        ret["energy_step_mean_episode_max"] = np.max(self._energy_means)
        ret["num_neighbours_mean_episode_mean"] = np.mean(self._nei_means)
        ret["num_neighbours_mean_episode_max"] = np.max(self._nei_means)
        success = [e[0] for e in self._ended.values()]
        crash = [e[1] for e in self._ended.values()]
        out = [e[2] for e in self._ended.values()]
        n = len(success)
        ret["num_agents_total"] = n
        ret["num_agents_total_per_300_steps"] = n / self._steps * 300
        ret["success_rate"] = sum(success) / n
        ret["num_agents_success"] = sum(success)
        ret["num_agents_success_per_300_steps"] = sum(success) / self._steps * 300
        ret["num_agents_failed_per_300_steps"] = sum(crash) / self._steps * 300
        rewards = list(self._ep_reward.values())
        ret["episode_reward_mean"], ret["episode_reward_min"] = np.mean(rewards), np.min(rewards)
        ret["episode_reward_max"] = np.max(rewards)
        costs = list(self._cost.values())
        ret["episode_cost_mean"], ret["episode_cost_min"] = np.mean(costs), np.min(costs)
        ret["episode_cost_max"], ret["episode_cost_sum"] = np.max(costs), np.sum(costs)
        ret["crash_rate"], ret["num_agents_crash"] = sum(crash) / n, sum(crash)
        ret["out_rate"], ret["num_agents_out"] = sum(out) / n, sum(out)
        ret["episode_length_mean"] = np.mean(list(self._ep_len.values()))
        won = [v for k, v in self._ep_len.items() if self._ended.get(k, (False,))[0]]
        ret["success_episode_length_mean"] = np.mean(won) if won else 0
        svos, svo_rewards = [], []
        for k, own in self._own.items():                                  # recoder.py:316-343
            nei = self._nei[k]
            alpha = np.rad2deg(math.atan2(nei, own))
            svo = min(max(0, alpha), 90)
            svos.append(svo)
            svo_rewards.append(math.sqrt(nei ** 2 + own ** 2) * math.cos(np.deg2rad(svo) - np.deg2rad(alpha)))
        ret["svo_estimate_deg_mean"], ret["svo_estimate_deg_min"] = np.mean(svos), np.min(svos)
        ret["svo_estimate_deg_max"] = np.max(svos)
        ret["svo_reward"] = np.sum(svo_rewards) / n
        return ret
