"""TEST INFRASTRUCTURE (oracle) - dict-keyed restatement of the reference's evaluation recorder.

  DistanceMap.update_distance_map / find_in_range / get_rewards   copo/eval/recoder.py:16-70
  RecorderEnv.on_episode_step / on_episode_end                    recoder.py:102-150
  RecorderEnv.get_step_result                                     recoder.py:152-175
  RecorderEnv.get_episode_result                                  recoder.py:177-349

It keeps the reference's data structure (stat -> step -> agent -> value) and is fed the (rewards, dones, infos) stream
of a dict-API environment plus the positions of the vehicles still on the road after the step.  Only tests/ may import
this."""
import math
from collections import defaultdict

import numpy as np

END = -1


class Recorder:
    def __init__(self, neighbours_distance=20):
        self.distance = neighbours_distance
        self.episode_step = 0

    def start(self):                                                       # :102-106
        self.user_data = defaultdict(lambda: defaultdict(dict))
        self.step_active_agents = {}
        self.episode_step = 0

    @staticmethod
    def _distance_map(positions):                                          # :33-44
        dm = defaultdict(lambda: defaultdict(lambda: float("inf")))
        keys = list(positions.keys())
        for c1 in range(0, len(keys) - 1):
            for c2 in range(c1 + 1, len(keys)):
                k1, k2 = keys[c1], keys[c2]
                p1, p2 = positions[k1], positions[k2]
                d = math.sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)
                dm[k1][k2] = d
                dm[k2][k1] = d
        return dm

    def _rewards(self, dm, reward_dict):                                   # :21-31, :49-70
        own_r, nei_r, num = {}, {}, {}
        for k, own in reward_dict.items():
            to_others = dm[k]
            order = sorted(to_others, key=lambda n: to_others[n]) if self.distance > 0 else []
            neighbours = [n for n in order if to_others[n] < self.distance]
            others = [reward_dict[n] for n in neighbours]
            own_r[k] = own
            nei_r[k] = own if len(others) == 0 else np.mean(others)
            num[k] = len(neighbours)
        return own_r, nei_r, num

    def step(self, positions, r, d, i):                                    # :91-99, :108-137
        if self.episode_step == 0:
            self.start()
        t = self.episode_step
        own_r, nei_r, num = self._rewards(self._distance_map(positions), r)
        for k in own_r:
            self.user_data["own_reward"][t][k] = own_r[k]
            self.user_data["num_neighbours"][t][k] = num[k]
            self.user_data["nei_reward"][t][k] = nei_r[k]
        self.step_active_agents[t] = set(r.keys())
        for k in r:
            info = i[k]
            if "step_reward" in info:
                for name in ("velocity", "steering", "step_reward", "acceleration", "cost", "episode_length",
                             "episode_reward"):
                    self.user_data[name][t][k] = info[name]
                self.user_data["energy"][t][k] = info["step_energy"]
                self.user_data["raw_action0_l2"][t][k] = info["raw_action"][0] ** 2
                self.user_data["raw_action1_l2"][t][k] = info["raw_action"][1] ** 2
        self.episode_step += 1
        for k, done in d.items():                                          # :95-98, :139-150
            if k != "__all__" and done:
                info = i[k]
                arrive, crash, out = info.get("arrive_dest", False), info.get("crash", False), info.get("out_of_road", False)
                self.user_data["success"][END][k] = arrive
                self.user_data["crash"][END][k] = crash
                self.user_data["max_step"][END][k] = not (arrive or crash or out)
                self.user_data["out"][END][k] = out
                self.user_data["episode_energy"][END][k] = info["episode_energy"]

    def _step_means(self, stat):
        out = []
        for t, active in self.step_active_agents.items():
            vals = [self.user_data[stat][t].get(k, None) for k in active]
            vals = [v for v in vals if v is not None]
            if len(vals) > 0:
                out.append(np.mean(vals))
        return out

    def step_result(self):                                                 # :152-175
        ret = {}
        t, active = list(self.step_active_agents.items())[-1]
        for stat in list(self.user_data.keys()):
            vals = [self.user_data[stat][t].get(k, None) for k in active]
            vals = [v for v in vals if v is not None]
            if len(vals) > 0:
                ret[stat] = np.mean(vals)
        ret["episode_reward_mean"] = np.mean(list(list(self.user_data["episode_reward"].values())[-1].values()))
        cost = self._agent_sums("cost")
        ret["episode_cost_mean"] = np.mean(list(cost.values()))
        ret["episode_cost_sum"] = np.sum(list(cost.values()))
        return ret

    def _agent_sums(self, stat):
        acc = defaultdict(float)
        for t, active in self.step_active_agents.items():
            for k in active:
                v = self.user_data[stat][t].get(k, None)
                if v is not None:
                    acc[k] += v
        return acc

    def _agent_last(self, stat):
        acc = defaultdict(float)
        for t in sorted(self.step_active_agents):
            if t == END:
                continue
            for k in self.step_active_agents[t]:
                acc[k] = self.user_data[stat][t].get(k, 0)
        return acc

    def episode_result(self):                                              # :177-349
        ret = {}
        v = self._step_means("velocity")
        ret["velocity_step_mean_episode_min"], ret["velocity_step_mean_episode_mean"] = np.min(v), np.mean(v)
        ret["velocity_step_mean_episode_max"] = np.max(v)
        e = self._step_means("energy")
        ret["energy_step_mean_episode_min"], ret["energy_step_mean_episode_mean"] = np.min(e), np.mean(e)
        ret["energy_step_mean_episode_max"] = np.max(e)
        n = self._step_means("num_neighbours")
        ret["num_neighbours_mean_episode_mean"], ret["num_neighbours_mean_episode_max"] = np.mean(n), np.max(n)
        ep_len = len(self.step_active_agents)
        success = list(self.user_data["success"][END].values())
        crash = list(self.user_data["crash"][END].values())
        out = list(self.user_data["out"][END].values())
        num = len(success)
        ret["num_agents_total"] = num
        ret["num_agents_total_per_300_steps"] = num / ep_len * 300
        ret["success_rate"] = sum(success) / num
        ret["num_agents_success"] = sum(success)
        ret["num_agents_success_per_300_steps"] = sum(success) / ep_len * 300
        ret["num_agents_failed_per_300_steps"] = sum(crash) / ep_len * 300
        rew = list(self._agent_last("episode_reward").values())
        ret["episode_reward_mean"], ret["episode_reward_min"], ret["episode_reward_max"] = np.mean(rew), np.min(rew), np.max(rew)
        cost = list(self._agent_sums("cost").values())
        ret["episode_cost_mean"], ret["episode_cost_min"] = np.mean(cost), np.min(cost)
        ret["episode_cost_max"], ret["episode_cost_sum"] = np.max(cost), np.sum(cost)
        ret["crash_rate"], ret["num_agents_crash"] = sum(crash) / num, sum(crash)
        ret["out_rate"], ret["num_agents_out"] = sum(out) / num, sum(out)
        lens = self._agent_last("episode_length")
        ret["episode_length_mean"] = np.mean(list(lens.values()))
        won = [val for k, val in lens.items() if self.user_data["success"][END][k]]
        ret["success_episode_length_mean"] = np.mean(won) if won else 0
        own, nei = defaultdict(float), defaultdict(float)
        for step_dict in self.user_data["own_reward"].values():
            for k, val in step_dict.items():
                own[k] += val
        for step_dict in self.user_data["nei_reward"].values():
            for k, val in step_dict.items():
                nei[k] += val
        svos, svo_rewards = [], []
        for k, o in own.items():
            alpha = np.rad2deg(math.atan2(nei[k], o))
            svo = min(max(0, alpha), 90)
            svos.append(svo)
            svo_rewards.append(math.sqrt(nei[k] ** 2 + o ** 2) * math.cos(np.deg2rad(svo) - np.deg2rad(alpha)))
        ret["svo_estimate_deg_mean"], ret["svo_estimate_deg_min"], ret["svo_estimate_deg_max"] = np.mean(svos), np.min(svos), np.max(svos)
        ret["svo_reward"] = np.sum(svo_rewards) / num
        return ret
