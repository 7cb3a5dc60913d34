"""RLlib-style dict API of the reference's environments over the batched CUDA scene step.

The reference wraps MetaDrive's `MultiAgent{Intersection,Roundabout,Tollgate,Bottleneck,ParkingLot}Env`
(`from metadrive.envs.marl_envs import ...`, torch_copo/train_copo.py:1-2) with `CCEnv` / `LCFEnv`
(torch_copo/utils/env_wrappers.py:30-430) and `get_rllib_compatible_env` (:559-597).  These classes keep that surface:
`reset() -> {agent: obs}`, `step({agent: action}) -> (obs, reward, done, info)` dicts with `done["__all__"]`, the
`info` keys the wrappers and callbacks read, `default_config()`, `set_lcf_dist`, `set_force_lcf`,
`close_and_reset_num_agents`, `observation_space` / `action_space` dict spaces, `vehicles`.  One instance is one scene
of a `BatchedDrivingEnv` (S = 1); the arithmetic of a step is the CUDA kernel's, this file only re-keys its outputs by
agent name ("agent{id}", fresh ids on respawn as in MetaDrive's agent manager).  Training at scale uses the tensor
API (copo_b200/batched_env.py, copo_b200/trainer.py) instead.
"""
import numpy as np
import torch

from .batched_env import (BatchedDrivingEnv, DEFAULT_NUM_AGENTS, FLAG_ARRIVE, FLAG_CRASH, FLAG_DONE, FLAG_MAXSTEP,
                          FLAG_OUT, FLAG_SPAWNED, FLAG_VALID)

COMM_CURRENT_OBS, COMM_METHOD, NEI_OBS = "comm_current_obs", "comm_method", "nei_obs"     # env_wrappers.py:17,24,26

F_X, F_Y, F_H, F_V, F_STEER, F_THR, F_S, F_DONE_LEN, F_ROUTE, F_SEG, F_EPLEN, F_EPREW, F_LCF, F_STATUS, F_ID, F_YAW = range(16)
VMAX = 22.22222137451172


class Box:
    """Minimal stand-in for gym.spaces.Box (gym is not a dependency of the hot path)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low = np.full(shape, low, dtype)
        self.high = np.full(shape, high, dtype)
        self.shape, self.dtype = tuple(shape), dtype

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class DictSpace:
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def keys(self):
        return self.spaces.keys()

    def __getitem__(self, k):
        return self.spaces[k]

    def __contains__(self, k):
        return k in self.spaces

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


class _Vehicle:
    def __init__(self, x, y, speed, heading_theta=0.0):
        self.position = np.array([x, y], dtype=np.float64)
        self.speed = speed
        self.heading_theta = float(heading_theta)

    @property
    def heading(self):
        return np.array([np.cos(self.heading_theta), np.sin(self.heading_theta)])

    def projection(self, vector):
        """(forward, leftward) components of `vector` in the vehicle frame (MetaDrive's BaseVehicle.projection as
        recalled - MetaDrive is not available to check; only the disabled-by-default `add_pos_in_comm` branch uses it)."""
        h = self.heading
        return np.array([h[0] * vector[0] + h[1] * vector[1], -h[1] * vector[0] + h[0] * vector[1]])


class MultiAgentDrivingEnv:
    """One scene with the reference's multi-agent dict API (native MetaDrive-level view, no CoPO wrappers)."""
    MAP = "intersection"
    APPEND_LCF = False
    SIM_FACTORY = None          # tests substitute a host simulator with BatchedDrivingEnv's interface here

    @classmethod
    def default_config(cls):
        return dict(num_agents=DEFAULT_NUM_AGENTS[cls.MAP], horizon=1000, delay_done=25, allow_respawn=True,
                    start_seed=0, neighbours_distance=40, crash_done=True, out_of_road_done=True)

    def __init__(self, config=None):
        self.config = self.default_config()
        for k, v in (config or {}).items():                  # nested option groups ("communication") merge key-wise
            if isinstance(v, dict) and isinstance(self.config.get(k), dict):
                self.config[k] = dict(self.config[k], **v)
            else:
                self.config[k] = v
        self._build()

    def _build(self):
        c = self.config
        make = type(self).SIM_FACTORY or BatchedDrivingEnv
        self._sim = make(self.MAP, num_scenes=1, num_slots=int(c["num_agents"]),
                                      num_agents=int(c["num_agents"]), delay_done=int(c["delay_done"]),
                                      horizon=int(c["horizon"]), neighbours_distance=float(c["neighbours_distance"]),
                                      allow_respawn=bool(c["allow_respawn"]), auto_reset=False,
                                      append_lcf=self.APPEND_LCF, lcf_uniform=(c.get("lcf_dist") == "uniform"),
                                      seed=int(c["start_seed"]), lcf_std=float(c.get("lcf_normal_std", 0.1)),
                                      force_lcf=float(c.get("force_lcf", -100)), map_kwargs=c.get("map_config"))
        self.A, self.D = self._sim.A, self._sim.D
        self._slot_of, self.vehicles, self.vehicles_including_just_terminated = {}, {}, {}
        self._agent_ids = set(["agent{}".format(i) for i in range(100)] + ["{}".format(i) for i in range(10000)])
        self.episode_step = 0

    # ---- spaces -----------------------------------------------------------------------------------------------
    @property
    def observation_space(self):
        names = list(self.vehicles.keys()) or ["agent%d" % i for i in range(self.A)]
        low = -1.0 if self.APPEND_LCF else 0.0                     # LCFObs space: Box(-1, 1) (env_wrappers.py:227-246)
        return DictSpace({k: Box(low, 1.0, (self.D,)) for k in names})

    @property
    def action_space(self):
        names = list(self.vehicles.keys()) or ["agent%d" % i for i in range(self.A)]
        return DictSpace({k: Box(-1.0, 1.0, (2,)) for k in names})

    def action_space_sample(self, agent_ids=None):
        return {k: v for k, v in self.action_space.sample().items() if agent_ids is None or k in agent_ids}

    # ---- stepping ---------------------------------------------------------------------------------------------
    def _harvest(self, out, first):
        st = self._sim.get_state()[0]
        AP = (self.A + 3) // 4 * 4
        fld = st[:16 * AP].reshape(16, AP)[:, :self.A]
        ff = fld.view(np.float32)
        flags = out["flags"][0].cpu().numpy()
        obs, rew = out["obs"][0].cpu().numpy(), out["reward"][0].cpu().numpy()
        ids = out["agent_id"][0].cpu().numpy()
        nei_mask = out["nei_mask"][0].cpu().numpy().view(np.uint64)
        o, r, d, info = {}, {}, {}, {}
        part = [i for i in range(self.A) if flags[i] & (FLAG_VALID | FLAG_SPAWNED)]
        # The agent that ACTED in slot i this step is whoever occupied it after the previous step; the kernel's agent_id
        # is the occupant after this step.  They differ when the slot's agent terminated and the slot was refilled in
        # the same step (delay_done = 0, or the scene's horizon restart): such a row carries two agents - the terminal
        # reward / flags of the one that acted, the first observation of the one that spawned.
        occupant = getattr(self, "_occupant", {})
        new_name = lambda i: "agent%d" % int(ids[i])
        reused = {i for i in part if not first and (flags[i] & FLAG_VALID) and (flags[i] & FLAG_SPAWNED) and
                  i in occupant and occupant[i] != new_name(i)}
        name = lambda i: occupant[i] if i in reused else new_name(i)        # the agent this row's step data belongs to
        self.vehicles_including_just_terminated = {name(i): _Vehicle(ff[F_X, i], ff[F_Y, i], ff[F_V, i], ff[F_H, i])
                                                   for i in part}
        self.vehicles = {name(i): self.vehicles_including_just_terminated[name(i)] for i in part
                         if not (flags[i] & FLAG_DONE)}
        self._slot_of = {name(i): i for i in part if not (flags[i] & FLAG_DONE)}
        route_len = self._sim.tables.route_len
        for i in part:
            k = name(i)
            o[k] = self._last_obs_of.get(k, obs[i]).copy() if i in reused else obs[i].copy()
            if first:
                continue
            f = int(flags[i])
            acted = bool(f & FLAG_VALID)
            r[k] = float(rew[i]) if acted else 0.0
            d[k] = bool(f & FLAG_DONE)
            others = [j for j in part if j != i and (int(nei_mask[i]) >> j) & 1]
            # distances as the kernel takes them (float32 squared distance, sim_core.cuh phase_neighbours), square root
            # in float64: `dist > r` here decides exactly like `d2 > r * r` there, so the masks the kernel emits
            # (nei_mask, mf_mask, nei_list) and the lists a host-side consumer derives from these infos agree, ties on
            # the spawn grid included
            dx, dy = ff[F_X, i] - ff[F_X, others], ff[F_Y, i] - ff[F_Y, others]
            d2 = (dx * dx + dy * dy).astype(np.float32)
            dist = {j: float(np.sqrt(np.float64(d2[n]))) for n, j in enumerate(others)}
            order = sorted(others, key=lambda j: dist[j])
            info[k] = dict(all_agents=[name(j) for j in part], neighbours=[name(j) for j in order],
                           neighbours_distance=[dist[j] for j in order])
            if not acted:
                # a freshly spawned agent has not stepped yet: MetaDrive gives it no step_reward / velocity / cost /
                # episode_* entries (the reference's recorder keys on that: eval/recoder.py:128), only what the wrappers add
                continue
            total = float(route_len[int(fld[F_ROUTE, i])])
            cur = float(ff[F_DONE_LEN, i] + ff[F_S, i])
            info[k].update(velocity=float(ff[F_V, i]) * 3.6, steering=-float(ff[F_STEER, i]) + 0.0,
                           acceleration=float(ff[F_THR, i]), step_reward=r[k], cost=1.0 if f & FLAG_CRASH else 0.0,
                           episode_length=int(fld[F_EPLEN, i]), episode_reward=float(ff[F_EPREW, i]),
                           arrive_dest=bool(f & FLAG_ARRIVE), crash=bool(f & FLAG_CRASH),
                           crash_vehicle=bool(f & FLAG_CRASH), out_of_road=bool(f & FLAG_OUT),
                           max_step=bool(f & FLAG_MAXSTEP), route_completion=cur / total, track_length=total,
                           current_distance=cur, step_energy=0.0, episode_energy=0.0,
                           raw_action=(float(ff[F_STEER, i]), float(ff[F_THR, i])))
        if not first:
            d["__all__"] = bool(out["scene_done"][0]) or (self.episode_step >= self.config["horizon"])
        self._last_out = out
        self._slot_now = {name(i): i for i in part}
        # the new occupants of reused slots: first observation under their own name, acting from the next step on
        for i in reused:
            k2 = new_name(i)
            o[k2] = obs[i].copy()
            r[k2], d[k2] = 0.0, False
            info[k2] = dict(all_agents=[name(j) for j in part], neighbours=[], neighbours_distance=[])
            self._slot_of[k2] = i
            self._slot_now[k2] = i
            self.vehicles[k2] = self.vehicles_including_just_terminated[k2] = _Vehicle(ff[F_X, i], ff[F_Y, i], ff[F_V, i],
                                                                                        ff[F_H, i])
        self._occupant = {i: new_name(i) for i in part}
        self._last_obs_of = dict(o)
        return o, r, d, info

    def reset(self, force_seed=None):
        self.episode_step = 0
        out = self._sim.reset(new_episode=True)
        return self._harvest(out, first=True)[0]

    def step(self, actions):
        act = torch.zeros((1, self.A, 2), dtype=torch.float32)
        for k, a in actions.items():
            if k in self._slot_of:
                act[0, self._slot_of[k]] = torch.as_tensor(np.asarray(a, np.float32)[:2])
        self.episode_step += 1
        out = self._sim.step(act.to(self._sim.device))
        return self._harvest(out, first=False)

    def close(self):
        self._sim.close()

    # ---- attributes the reference's callers read off a MetaDrive env ------------------------------
    def _header(self, k):
        AP = (self.A + 3) // 4 * 4
        return int(self._sim.get_state()[0][16 * AP + k])

    @property
    def engine(self):
        """`env.engine.global_seed`, `env.engine.current_map.road_network.get_bounding_box()` (env_wrappers.py:268-272;
        legacy svo_env.py).  The seed of the running episode is the start seed plus the scene's episode counter."""
        env = self

        class _RoadNetwork:
            def get_bounding_box(self):
                return env._sim.tables.bounding_box()

        class _Map:
            road_network = _RoadNetwork()

        class _Engine:
            current_map = _Map()

            @property
            def global_seed(self):
                return int(env.config["start_seed"]) + max(env._header(2) - 1, 0)       # header word 2: episode counter

            episode_step = property(lambda self: env.episode_step)

        return _Engine()

    @property
    def agent_manager(self):
        """`env.agent_manager.next_agent_count` (legacy svo_env.py:256): how many agents the scene has named so far."""
        env = self

        class _AgentManager:
            next_agent_count = property(lambda self: env._header(1))                     # header word 1: next agent id
            active_agents = property(lambda self: env.vehicles)

        return _AgentManager()

    def close_and_reset_num_agents(self, num_agents):
        """ChangeNEnv (env_wrappers.py:444-460): population change = slot masking, no re-allocation."""
        self.config["num_agents"] = num_agents
        self._sim.set_num_agents(num_agents)


def _named(map_name, cls_name):
    return type(cls_name, (MultiAgentDrivingEnv,), {"MAP": map_name})


MultiAgentIntersectionEnv = _named("intersection", "MultiAgentIntersectionEnv")
MultiAgentRoundaboutEnv = _named("roundabout", "MultiAgentRoundaboutEnv")
MultiAgentTollgateEnv = _named("tollgate", "MultiAgentTollgateEnv")
MultiAgentBottleneckEnv = _named("bottleneck", "MultiAgentBottleneckEnv")
MultiAgentParkingLotEnv = _named("parking_lot", "MultiAgentParkingLotEnv")
MultiAgentMetaDrive = _named("pg", "MultiAgentMetaDrive")      # procedurally generated maps: config["map_config"]


def get_ccenv(env_class):
    """CCEnv (env_wrappers.py:30-158, 433-441): `all_agents`, `neighbours`, `neighbours_distance` in every info."""
    name = env_class.__name__
    assert name.startswith("MultiAgent")

    class TMP(env_class):
        @classmethod
        def default_config(cls):
            c = super().default_config()
            c["neighbours_distance"] = 40
            c.update(communication=dict(comm_method="none", comm_size=4, comm_neighbours=4, add_pos_in_comm=False),
                     add_traffic_light=False, traffic_light_interval=30)          # env_wrappers.py:42-48
            return c

        def __init__(self, config=None):
            super().__init__(config)
            comm = self.config["communication"]
            self._comm_on = comm[COMM_METHOD] != "none"
            self._comm_dim = comm["comm_size"] + (3 if comm["add_pos_in_comm"] else 0)      # :58-61

        @property
        def action_space(self):
            old = super().action_space
            if not self._comm_on:
                return old
            n = 2 + self.config["communication"]["comm_size"]                      # :70-87 (not _comm_dim)
            return DictSpace({k: Box(-1.0, 1.0, (n,)) for k in old.keys()})

        def step(self, actions):
            """The message channel of CCEnv.step (env_wrappers.py:89-121; off unless `comm_method` is set): the
            action carries `comm_size` extra entries that are handed to the nearest neighbours as observations."""
            if not self._comm_on:
                return super().step(actions)
            comm = self.config["communication"]
            comm_actions = {k: np.asarray(v)[2:] for k, v in actions.items()}
            o, r, d, i = super().step({k: np.asarray(v)[:2] for k, v in actions.items()})
            veh = self.vehicles_including_just_terminated
            for k, inf in i.items():
                cur = []
                for n in inf["neighbours"][:comm["comm_neighbours"]]:
                    if n not in comm_actions:
                        cur.append(np.zeros((self._comm_dim,)))
                    elif comm["add_pos_in_comm"]:
                        rel = veh[k].projection(veh[n].position - veh[k].position)
                        dis = np.linalg.norm(rel)
                        extra = [dis / 20, ((rel[0] / dis) + 1) / 2, ((rel[1] / dis) + 1) / 2]
                        cur.append(np.concatenate([comm_actions[n], np.clip(np.asarray(extra), 0, 1)]))
                    else:
                        cur.append(comm_actions[n])
                inf[COMM_CURRENT_OBS] = cur
            return o, r, d, i

    TMP.__name__ = TMP.__qualname__ = "CC" + name
    return TMP


def get_lcf_env(env_class):
    """LCFEnv (env_wrappers.py:161-430, 463-471): LCF appended to the obs, nei / global / coordinated rewards."""
    name = env_class.__name__
    base = get_ccenv(env_class)

    class TMP(base):
        APPEND_LCF = True

        @classmethod
        def default_config(cls):
            c = super().default_config()
            c.update(neighbours_distance=40, lcf_mode="angle", lcf_dist="normal", lcf_normal_std=0.1,
                     return_native_reward=True, force_lcf=-100, enable_copo=True)
            return c

        def __init__(self, config=None):
            super().__init__(config)
            assert self.config["lcf_mode"] in ["linear", "angle"]
            assert self.config["lcf_dist"] in ["uniform", "normal"]
            assert self.config["lcf_normal_std"] > 0.0
            self.force_lcf = self.config["force_lcf"]
            self.current_lcf_mean, self.current_lcf_std = 0.0, self.config["lcf_normal_std"]
            self._last_obs = None
            self._traffic_light_counter = 0

        @property
        def enable_copo(self):
            return self.config["enable_copo"]

        # ---- disabled-by-default branches of LCFEnv: traffic-light message, message channel ---------------------
        @property
        def _extra_obs(self):
            comm = self.config["communication"]
            return (3 if self.config["add_traffic_light"] else 0) + \
                (self._comm_dim * comm["comm_neighbours"] if self._comm_on else 0)

        @property
        def observation_space(self):
            sp = super().observation_space
            if not self._extra_obs:
                return sp
            return DictSpace({k: Box(-1.0, 1.0, (self.D + self._extra_obs,)) for k in sp.keys()})   # :227-246

        @property
        def _traffic_light_msg(self):
            fix_interval = self.config["traffic_light_interval"]                   # env_wrappers.py:259-266
            increment = (self._traffic_light_counter % fix_interval) / fix_interval * 0.1
            if ((self._traffic_light_counter // fix_interval) % 2) == 1:
                return 0 + increment
            return 1 - increment

        def get_agent_traffic_light_msg(self, pos):
            b_box = self._sim.tables.bounding_box()                                # :268-272
            pos0 = (pos[0] - b_box[0]) / (b_box[1] - b_box[0])
            pos1 = (pos[1] - b_box[2]) / (b_box[3] - b_box[2])
            return np.clip(np.array([self._traffic_light_msg, pos0, pos1]), 0, 1).astype(np.float32)

        def _with_traffic_light(self, o):
            """[obs | message, x, y | lcf]: the reference appends the message before `_add_lcf` (:277-296, :327-334)."""
            veh = self.vehicles_including_just_terminated
            return {k: np.concatenate([v[:-1], self.get_agent_traffic_light_msg(veh[k].position), v[-1:]])
                    for k, v in o.items()}

        def reset(self, force_seed=None):
            o = super().reset(force_seed)
            if self.config["add_traffic_light"]:
                self._traffic_light_counter = 0
                o = self._with_traffic_light(o)
            if self._comm_on:                                                      # :296-302
                pad = np.zeros((self._comm_dim * self.config["communication"]["comm_neighbours"],))
                o = {k: np.concatenate([v, pad], axis=-1).astype(np.float32) for k, v in o.items()}
            self._last_obs = o
            return o

        def step(self, actions):
            o, r, d, i = super().step(actions)
            assert set(i.keys()) == set(o.keys())
            out = self._last_out
            nei_r, glob = out["nei_reward"][0].cpu().numpy(), float(out["global_reward"][0])
            lcf = out["lcf"][0].cpu().numpy()
            new_r = {}
            for k, inf in i.items():
                s = self._slot_now[k]
                inf["nei_rewards"] = float(nei_r[s])
                inf["global_rewards"] = glob
                agent_lcf = float(lcf[s])
                inf["lcf"], inf["lcf_deg"] = agent_lcf, agent_lcf * 90
                if self.config["lcf_mode"] == "linear":
                    cr = agent_lcf * r[k] + (1 - agent_lcf) * inf["nei_rewards"]
                else:
                    rad = agent_lcf * np.pi / 2
                    cr = np.cos(rad) * r[k] + np.sin(rad) * inf["nei_rewards"]
                inf["coordinated_rewards"], inf["native_rewards"] = cr, r[k]
                new_r[k] = r[k] if self.config["return_native_reward"] else cr
            if self.config["add_traffic_light"]:
                self._traffic_light_counter += 1                                   # :315-316
                o = self._with_traffic_light(o)
            if self._comm_on:                                                      # :363-388
                n_nei = self.config["communication"]["comm_neighbours"]
                new_o = {}
                for k, old_obs in o.items():
                    comm_obs = i[k][COMM_CURRENT_OBS]
                    if len(comm_obs) < n_nei:
                        comm_obs.extend([np.zeros((self._comm_dim,))] * (n_nei - len(comm_obs)))
                    new_o[k] = np.concatenate([old_obs] + comm_obs).astype(np.float32)
                o = new_o
                for k, inf in i.items():
                    nei = inf["neighbours"]
                    inf[NEI_OBS] = [self._last_obs[nei[j]] if j < len(nei) and nei[j] in self._last_obs else None
                                    for j in range(n_nei)]
                    inf[NEI_OBS].append(None)       # the reference's extra None ("to make sure np.array fails")
            self._last_obs = o
            return o, new_r, d, i

        def set_lcf_dist(self, mean, std):
            assert self.enable_copo
            assert self.config["lcf_dist"] == "normal"
            assert std > 0.0
            assert -1.0 <= mean <= 1.0
            self.current_lcf_mean, self.current_lcf_std = mean, std
            self._sim.set_lcf_dist(mean, std)

        def set_force_lcf(self, v):
            assert self.enable_copo
            self.force_lcf = v
            self._sim.set_force_lcf(v)

    TMP.__name__ = TMP.__qualname__ = "LCF" + name
    return TMP


def get_change_n_env(env_class):
    return env_class


_REGISTRY = {}


def get_rllib_compatible_env(env_class, return_class=False):
    """env_wrappers.py:559-597: returns the registered env name (or the class)."""
    env_name = env_class.__name__
    _REGISTRY[env_name] = env_class
    return env_class if return_class else env_name


def make_env(name, config=None):
    return _REGISTRY[name](config)
