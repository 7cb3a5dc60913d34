"""The dict-API environments and wrappers of copo_b200/envs.py (host logic) on the CPU: the simulator behind them is
substituted by the host build of the device phases (tests/hostenv.py), the expected values come from oracle/sim.py and
the float64 restatement of the reference's wrappers (oracle/wrappers.py)."""
import math

import numpy as np
import pytest

import hostenv
from copo_b200 import envs
from copo_b200.maps import build_map
from oracle import sim as osim
from oracle import wrappers as ow


@pytest.fixture(autouse=True)
def host_simulator(monkeypatch):
    monkeypatch.setattr(envs.MultiAgentDrivingEnv, "SIM_FACTORY", staticmethod(hostenv.HostBatchedEnv))


def _oracle(A, seed):
    cfg = osim.SimConfig(seed=seed, auto_reset=False)
    cfg.num_agents = A
    ref = osim.OracleSim(build_map("intersection"), 1, A, cfg)
    ref.episode[:] = 0
    return ref


def test_dict_env_follows_reference_interface():
    cls = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)
    assert cls.__name__ == "LCFMultiAgentIntersectionEnv" and cls.default_config()["neighbours_distance"] == 40
    assert cls.default_config()["communication"]["comm_method"] == "none" and not cls.default_config()["add_traffic_light"]
    name = envs.get_rllib_compatible_env(cls)
    A = 12
    env = envs.make_env(name, {"num_agents": A, "start_seed": 5})
    ref = _oracle(A, 5)
    o = env.reset()
    ref.reset(new_episode=True)
    assert len(o) == A and all(v.shape == (92,) and v.dtype == np.float32 for v in o.values())
    assert set(o.keys()) == set(env.observation_space.keys()) == set("agent%d" % i for i in range(A))
    assert env.observation_space["agent0"].shape == (92,) and env.action_space["agent0"].shape == (2,)
    rng = np.random.default_rng(0)
    seen_done = 0
    for t in range(120):
        acts = {k: np.array([rng.uniform(-0.3, 0.3), rng.uniform(0, 1)], np.float32) for k in env.vehicles}
        arr = np.zeros((1, A, 2), np.float32)
        for k, a in acts.items():
            arr[0, env._slot_of[k]] = a
        o, r, d, i = env.step(acts)
        w = ref.step(arr)
        assert set(o.keys()) == set(i.keys()) == set(r.keys()) and "__all__" in d
        # the reference's own wrapper arithmetic (float64, dict keyed) on this step's positions and native rewards
        pos = {k: env.vehicles_including_just_terminated[k].position for k in i}
        if not r:                                                      # every slot is a lingering wreck this step
            continue
        cc = ow.cc_step(pos, r, 40)
        want_r = ow.lcf_step(dict(r), cc, {k: i[k]["lcf"] for k in i})
        for k in o:
            s = env._slot_now[k]
            assert k == "agent%d" % w["agent_id"][0, s]
            assert np.array_equal(o[k], w["obs"][0, s])
            inf = i[k]
            if w["flags"][0, s] & osim.F_VALID:
                assert r[k] == float(w["reward"][0, s]) and d[k] == bool(w["flags"][0, s] & osim.F_DONE)
            assert inf["lcf"] == float(w["lcf"][0, s]) and 0.0 <= o[k][-1] <= 1.0
            assert set(inf["neighbours"]) == set(cc[k]["neighbours"])
            assert np.allclose(sorted(inf["neighbours_distance"]), cc[k]["neighbours_distance"], rtol=1e-6)
            assert abs(inf["nei_rewards"] - cc[k]["nei_rewards"]) < 1e-5
            assert abs(inf["global_rewards"] - cc[k]["global_rewards"]) < 1e-5
            assert abs(inf["coordinated_rewards"] - (math.cos(inf["lcf"] * math.pi / 2) * r[k] +
                                                     math.sin(inf["lcf"] * math.pi / 2) * inf["nei_rewards"])) < 1e-9
            assert want_r[k] == r[k]                                   # return_native_reward=True
            if w["flags"][0, s] & osim.F_VALID:
                for key in ("velocity", "steering", "acceleration", "step_reward", "cost", "episode_length",
                            "episode_reward", "arrive_dest", "crash", "out_of_road", "route_completion", "all_agents"):
                    assert key in inf
            else:                              # freshly spawned: only what the wrappers add (eval/recoder.py:128 keys on it)
                assert "step_reward" not in inf and "velocity" not in inf and "all_agents" in inf
            seen_done += int(d[k])
    assert seen_done > 0
    with pytest.raises(AssertionError):
        env.set_lcf_dist(0.0, -1.0)
    env.set_lcf_dist(0.3, 0.05)
    assert env._sim.cfg.lcf_mean == np.float32(0.3)


def test_coordinated_reward_can_be_returned():
    cls = envs.get_lcf_env(envs.MultiAgentRoundaboutEnv)
    env = cls({"num_agents": 8, "start_seed": 1, "return_native_reward": False, "lcf_mode": "linear", "force_lcf": 0.5,
               "lcf_dist": "uniform"})
    env.reset()
    for t in range(20):
        o, r, d, i = env.step({k: np.array([0.0, 0.6], np.float32) for k in env.vehicles})
        for k in r:
            assert i[k]["lcf"] == 0.5
            assert r[k] == 0.5 * i[k]["native_rewards"] + 0.5 * i[k]["nei_rewards"]


def test_map_bounding_box_contains_every_spawn_place():
    for name in ("intersection", "roundabout", "tollgate", "bottleneck", "parking_lot"):
        t = build_map(name)
        x0, x1, y0, y1 = t.bounding_box()
        assert x0 < x1 and y0 < y1
        assert (t.spawn_f[:, 0] >= x0).all() and (t.spawn_f[:, 0] <= x1).all()
        assert (t.spawn_f[:, 1] >= y0).all() and (t.spawn_f[:, 1] <= y1).all()
    x0, x1, y0, y1 = build_map("intersection").bounding_box()          # four 60 m arms around a 10 m junction
    assert abs(x0 + x1) < 1e-3 and abs(y0 + y1) < 1e-3 and 130 < x1 - x0 < 150


def test_traffic_light_branch():
    """env_wrappers.py:259-296, 315-334 (off by default): [message, x, y] sits between the simulator's observation and
    the LCF entry; the message is a saw-tooth that flips every `traffic_light_interval` steps."""
    assert [round(ow.traffic_light_msg(c, 4), 3) for c in range(9)] == [1.0, 0.975, 0.95, 0.925, 0.0, 0.025, 0.05, 0.075, 1.0]
    cls = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)
    env = cls({"num_agents": 6, "start_seed": 3, "add_traffic_light": True, "traffic_light_interval": 5})
    plain = cls({"num_agents": 6, "start_seed": 3})
    assert env.observation_space["agent0"].shape == (95,)
    bbox = build_map("intersection").bounding_box()
    o, p = env.reset(), plain.reset()
    for t in range(13):
        for k in o:
            assert o[k].shape == (95,) and o[k].dtype == np.float32
            assert np.array_equal(o[k][:91], p[k][:91]) and o[k][-1] == p[k][-1]
            want = ow.agent_traffic_light_msg(ow.traffic_light_msg(t, 5), env.vehicles_including_just_terminated[k].position, bbox)
            assert np.array_equal(o[k][91:94], want)
        acts = {k: np.array([0.0, 0.5], np.float32) for k in env.vehicles}
        o, r, d, i = env.step(acts)
        p, _, _, _ = plain.step(acts)
    o = env.reset()
    assert all(v[91] == 1.0 for v in o.values())                       # the counter restarts with the episode


def test_message_channel_branch():
    """env_wrappers.py:70-121, 296-302, 363-388 (off by default): the action's extra entries reach the nearest
    neighbours as observations one step later."""
    cls = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)
    cfg = {"num_agents": 10, "start_seed": 2, "communication": {"comm_method": "broadcast"}}
    env, plain = cls(cfg), cls({"num_agents": 10, "start_seed": 2})
    assert env.config["communication"] == dict(comm_method="broadcast", comm_size=4, comm_neighbours=4, add_pos_in_comm=False)
    assert env.action_space["agent0"].shape == (6,) and env.observation_space["agent0"].shape == (92 + 16,)
    o, p = env.reset(), plain.reset()
    assert all(v.shape == (108,) and not v[92:].any() for v in o.values())
    rng = np.random.default_rng(0)
    heard = 0
    for t in range(40):
        last = o
        acts = {k: np.concatenate([[0.0, 0.7], rng.uniform(-1, 1, 4)]).astype(np.float32) for k in env.vehicles}
        o, r, d, i = env.step(acts)
        p, _, _, pi = plain.step({k: v[:2] for k, v in acts.items()})
        for k in o:
            assert np.array_equal(o[k][:92], p[k]) and i[k]["neighbours"] == pi[k]["neighbours"]
            cur = ow.comm_current_obs(i[k]["neighbours"], {n: a[2:] for n, a in acts.items()})
            assert np.array_equal(o[k], ow.lcf_comm_obs(p[k], cur))
            heard += len(cur)
            nei = i[k]["neighbours"]
            assert len(i[k]["nei_obs"]) == 5 and i[k]["nei_obs"][-1] is None
            for j in range(4):
                want = last[nei[j]] if j < len(nei) and nei[j] in last else None
                got = i[k]["nei_obs"][j]
                assert (got is None and want is None) or np.array_equal(got, want)
    assert heard > 0
    # with positions in the message: three more entries per neighbour, all within [0, 1]
    env2 = cls({"num_agents": 10, "start_seed": 2, "communication": {"comm_method": "broadcast", "add_pos_in_comm": True}})
    o = env2.reset()
    assert env2.observation_space["agent0"].shape == (92 + 28,)
    for t in range(15):
        o, r, d, i = env2.step({k: np.concatenate([[0.0, 0.7], np.full(4, 0.25)]).astype(np.float32) for k in env2.vehicles})
    full = [c for inf in i.values() for c in inf["comm_current_obs"] if c.any()]
    assert full and all(c.shape == (7,) and (c[4:] >= 0).all() and (c[4:] <= 1).all() for c in full)


def test_procedural_map_env():
    """`MultiAgentMetaDrive` (the reference's PG-map environment, train_all_cl.py:23): `map_config` picks the map."""
    cls = envs.get_lcf_env(envs.MultiAgentMetaDrive)
    assert cls.__name__ == "LCFMultiAgentMetaDrive" and cls.default_config()["num_agents"] == 15
    a = cls({"start_seed": 1, "map_config": {"seed": 5, "num_blocks": 2}})
    b = cls({"start_seed": 1, "map_config": {"seed": 6, "num_blocks": 2}})
    oa, ob = a.reset(), b.reset()
    assert len(oa) == 15 and all(v.shape == (92,) for v in oa.values())
    assert a._sim.tables.bounding_box() != b._sim.tables.bounding_box()
    for t in range(30):
        oa, r, d, i = a.step({k: np.array([0.0, 0.8], np.float32) for k in a.vehicles})
    assert all("neighbours" in inf and 0.0 <= inf["route_completion"] <= 1.05 for inf in i.values())


def test_attributes_the_reference_reads_off_a_metadrive_env():
    """SURVEY.md 8b "Env attributes read by callers": `engine.global_seed`, `engine.current_map.road_network
    .get_bounding_box()`, `agent_manager.next_agent_count`, besides `vehicles` / `.position`."""
    env = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)({"num_agents": 6, "start_seed": 5000, "delay_done": 0})
    env.reset()
    assert env.engine.global_seed == 5000 and env.agent_manager.next_agent_count == 6
    box = env.engine.current_map.road_network.get_bounding_box()
    assert len(box) == 4 and box[0] < box[1] and box[2] < box[3]
    for k, v in env.vehicles.items():
        assert box[0] - 5 <= v.position[0] <= box[1] + 5 and box[2] - 5 <= v.position[1] <= box[3] + 5
    named = 6
    for t in range(60):                                   # everybody steers off the road: respawns under fresh names
        o, r, d, i = env.step({k: np.array([1.0, 1.0], np.float32) for k in env.vehicles})
        named = max(named, max(int(k[5:]) for k in o) + 1)
    assert env.agent_manager.next_agent_count == named > 6
    env.reset()
    assert env.engine.global_seed == 5001                 # the next episode of the same start seed
    assert set(env.agent_manager.active_agents) == set(env.vehicles)


def test_forced_lcf_with_the_normal_distribution_is_redrawn_every_step():
    """env_wrappers.py:337-342, 398-403: with `force_lcf` set and lcf_dist == "normal" every call of `_add_lcf` - every
    step of every agent - draws a fresh N(force_lcf, std) value (clipped to [-1, 1]) for the observation, info["lcf"] and
    the coordinated reward; the episode value (lcf_map) stays what it was at spawn.  With "uniform" the forced value is
    used as is (previous test)."""
    cls = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)
    env = cls({"num_agents": 6, "start_seed": 2, "return_native_reward": False, "force_lcf": 0.5, "lcf_normal_std": 0.2})
    o = env.reset()
    seen = {k: [] for k in o}
    state_lcf = None
    for t in range(25):
        o, r, d, i = env.step({k: np.array([0.0, 0.3], np.float32) for k in env.vehicles})
        st = env._sim.sim.state()["lcf"][0].copy()
        if state_lcf is not None:
            assert np.array_equal(st[:6], state_lcf[:6])           # the episode value does not move
        state_lcf = st
        for k in r:
            lcf = i[k]["lcf"]
            assert -1.0 <= lcf <= 1.0
            assert o[k][-1] == np.float32((np.float32(lcf) + np.float32(1.0)) * np.float32(0.5))
            rad = lcf * np.pi / 2
            assert abs(r[k] - (np.cos(rad) * i[k]["native_rewards"] + np.sin(rad) * i[k]["nei_rewards"])) < 1e-6
            seen.setdefault(k, []).append(lcf)
    vals = np.concatenate([np.asarray(v) for v in seen.values() if len(v) > 5])
    assert all(len(set(v)) == len(v) for v in seen.values() if len(v) > 5)      # a new draw every step
    assert abs(vals.mean() - 0.5) < 0.1 and 0.1 < vals.std() < 0.3


def test_delay_done_zero_reports_the_terminated_agent_and_its_successor_separately():
    """delay_done = 0: a slot can lose its agent and get a new one in the same step.  The dict API then carries BOTH
    agents, as MetaDrive does: the terminated one under its own name with done = True, its terminal reward and its step
    info; the successor under a fresh name with its first observation, no step info, done = False - and the successor
    can be driven from the next step on."""
    cls = envs.get_lcf_env(envs.MultiAgentIntersectionEnv)
    env = cls({"num_agents": 10, "start_seed": 4, "delay_done": 0})
    o = env.reset()
    born, died, both = set(o), set(), 0
    rng = np.random.default_rng(1)
    for t in range(160):
        acts = {k: np.array([0.5 + 0.3 * rng.standard_normal(), 1.0], np.float32) for k in env.vehicles}
        assert set(acts) <= set(o)                                     # everything that may act has been observed
        o, r, d, i = env.step(acts)
        assert set(o) == set(r) == set(i) and set(d) == set(o) | {"__all__"}
        ended = {k for k in r if d[k]}
        fresh = {k for k in o if k not in born}
        for k in ended:
            assert k in acts and "step_reward" in i[k] and k not in env.vehicles
        for k in fresh:
            assert not d[k] and r[k] == 0.0 and "step_reward" not in i[k] and k in env.vehicles
        slots_of_ended = {env._slot_now[k] for k in ended}
        both += len({env._slot_now[k] for k in fresh} & slots_of_ended)
        assert not (ended & died)                                      # nobody is reported done twice
        born |= fresh
        died |= ended
    assert both > 0 and len(died) > 5                                  # same-step reuse happened and was split
