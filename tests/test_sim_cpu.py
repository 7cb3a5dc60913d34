"""CPU-side checks of the scene step: the device phases compiled for the host (tests/hostsim) against the
numpy spec (oracle/sim.py), the spec's CoPO-owned bookkeeping against the float64 restatement of the reference
wrappers (oracle/wrappers.py), the constants shared by the two, and the geometry tables."""
import os
import re

import numpy as np
import pytest

import simcheck as sc
from copo_b200.maps import build_map
from oracle import sim as osim
from oracle import wrappers as ow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _policy(obs, rng, S, A):
    a0 = -np.clip(-1.5 * (obs[..., 2] - 0.5) * 3.14 - 2.0 * (obs[..., 8] - 0.5) + rng.normal(0, 0.02, (S, A)), -1, 1)
    a1 = np.where(obs[..., 3] < 0.35, 0.6, 0.0) + rng.normal(0, 0.05, (S, A))
    a0[0, ::2] = -0.5                                 # scene 0: every other car steers off the road
    return np.stack([a0, a1], -1).astype(np.float32)


@pytest.mark.parametrize("map_name,S,A,T,kw", [
    ("intersection", 2, 40, 130, dict(horizon=100)),
    ("roundabout", 2, 40, 60, dict()),
    ("parking_lot", 3, 10, 100, dict()),
    ("tollgate", 2, 40, 40, dict()),
    ("bottleneck", 2, 20, 80, dict(delay_done=0)),
    ("intersection", 2, 40, 50, dict(append_lcf=False, num_agents=30, neighbours_distance=10.0)),
    ("intersection", 2, 12, 60, dict(lcf_uniform=True, allow_respawn=False, auto_reset=False, horizon=40)),
    ("intersection", 1, 7, 30, dict(force_lcf=0.5)),
    ("intersection", 2, 64, 25, dict()),
    ("parking_lot", 5, 1, 30, dict()),
    ("roundabout", 1, 3, 30, dict(force_lcf=0.5, delay_done=1)),
    ("pg", 3, 15, 120, dict()),
])
def test_host_phases_match_spec(map_name, S, A, T, kw):
    tables = build_map(map_name)
    cfg = osim.SimConfig(seed=11, **kw)
    if cfg.num_agents is None:
        cfg.num_agents = A
    ref = osim.OracleSim(tables, S, A, cfg)
    hs = sc.HostSim(tables, S, A, cfg)
    r = ref.reset()
    g = hs.reset()
    sc.compare_outputs(r, g, "reset")
    rng = np.random.default_rng(1)
    for t in range(T):
        act = _policy(r["obs"], rng, S, A)
        r = ref.step(act)
        g = hs.step(act)
        sc.compare_outputs(r, g, "%s step %d" % (map_name, t))
        if t % 10 == 0 or t == T - 1:
            sc.compare_state(ref, hs.state(), "%s step %d" % (map_name, t))


def test_spec_events_are_exercised():
    """The scripted driver reaches every terminal kind, so the parity runs above cover them."""
    tables = build_map("intersection")
    S, A = 4, 40
    cfg = osim.SimConfig(seed=3, horizon=150)
    cfg.num_agents = A
    ref = osim.OracleSim(tables, S, A, cfg)
    r = ref.reset()
    rng = np.random.default_rng(0)
    seen = {k: 0 for k in ("arrive", "crash", "out", "spawn", "scene_done")}
    for t in range(320):
        act = _policy(r["obs"], rng, S, A)
        r = ref.step(act)
        f = r["flags"]
        seen["arrive"] += int(((f & osim.F_ARRIVE) > 0).sum())
        seen["crash"] += int(((f & osim.F_CRASH) > 0).sum())
        seen["out"] += int(((f & osim.F_OUT) > 0).sum())
        seen["spawn"] += int(((f & osim.F_SPAWNED) > 0).sum())
        seen["scene_done"] += int(r["scene_done"].sum())
        assert np.isfinite(r["obs"]).all() and r["obs"].min() >= 0 and r["obs"].max() <= 1
    assert all(v > 0 for v in seen.values()), seen


def test_bookkeeping_follows_reference_wrappers():
    """Neighbour lists / masks / nei and global rewards of the float32 spec against the float64 dict
    restatement of env_wrappers.py:125-158, 313-326 on the same positions and rewards."""
    tables = build_map("intersection")
    S, A = 3, 40
    cfg = osim.SimConfig(seed=5)
    cfg.num_agents = A
    ref = osim.OracleSim(tables, S, A, cfg)
    r = ref.reset()
    rng = np.random.default_rng(2)
    checked = 0
    for t in range(60):
        r = ref.step(_policy(r["obs"], rng, S, A))
        part = ((r["flags"] & osim.F_VALID) > 0) | ((r["flags"] & osim.F_SPAWNED) > 0)
        for s in range(S):
            names = [i for i in range(A) if part[s, i]]
            pos = {i: (float(ref.x[s, i]), float(ref.y[s, i])) for i in names}
            rew = {i: float(r["reward"][s, i]) for i in names}
            infos = ow.cc_step(pos, rew, neighbours_distance=40)
            lcf_map = {i: float(r["lcf"][s, i]) for i in names}
            ow.lcf_step(rew, infos, lcf_map)
            # skip agents with a pair distance within 1e-4 of the radius (float32 vs float64 rounding)
            d64 = {i: ow.update_distance_map(pos)[i] for i in names}
            for i in names:
                if any(abs(d - 40.0) < 1e-4 for d in d64[i].values()):
                    continue
                nb = infos[i]["neighbours"]
                mask = 0
                for j in nb:
                    mask |= 1 << j
                assert mask == int(r["nei_mask"][s, i]), (t, s, i)
                assert len(nb) == int(r["nei_count"][s, i])
                dd = infos[i]["neighbours_distance"]
                if all(b - a > 1e-4 for a, b in zip(dd, dd[1:])):      # no near ties: order is well defined
                    want = nb[:4] + [-1] * (4 - min(4, len(nb)))
                    assert list(r["nei_list"][s, i]) == want, (t, s, i)
                assert abs(infos[i]["nei_rewards"] - float(r["nei_reward"][s, i])) < 1e-5
                assert abs(infos[i]["global_rewards"] - float(r["global_reward"][s])) < 1e-5
                assert abs(infos[i]["lcf"]) <= 1.0
                checked += 1
    assert checked > 2000


def test_wrapper_restatement_edge_cases():
    # distance <= 0 -> no neighbours (env_wrappers.py:126-127); strict '<' (env_wrappers.py:133)
    pos = {"a": (0.0, 0.0), "b": (40.0, 0.0), "c": (0.0, 39.999), "d": None, "e": (0.0, 39.999)}
    dm = ow.update_distance_map(pos)
    assert ow.find_in_range(dm, "a", 0) == ([], [])
    names, d = ow.find_in_range(dm, "a", 40)
    assert names == ["c", "e"] and d == sorted(d)          # tie keeps vehicle order, 40.0 itself excluded
    assert "d" not in dm["a"]
    infos = {"a": {"neighbours": names}, "b": {"neighbours": []}}
    out = ow.lcf_step({"a": 1.0, "b": 3.0, "c": 2.0, "e": 4.0}, infos, {"a": 0.5, "b": -1.0})
    assert infos["a"]["nei_rewards"] == 3.0 and infos["b"]["nei_rewards"] == 0.0
    assert infos["a"]["global_rewards"] == 2.5 and out["a"] == 1.0
    assert abs(infos["a"]["coordinated_rewards"] - (np.cos(np.pi / 4) * 1.0 + np.sin(np.pi / 4) * 3.0)) < 1e-12
    lcf, obs = ow.add_lcf(np.zeros(3, np.float32), None, np.random.default_rng(0), mean=0.9, std=5.0)
    assert -1.0 <= lcf <= 1.0 and obs.shape == (4,) and obs.dtype == np.float32 and obs[-1] == np.float32((lcf + 1) / 2)
    assert ow.add_lcf(np.zeros(3), 0.25, None)[0] == 0.25
    assert ow.add_lcf(np.zeros(3), None, None, enable_copo=False)[0] == 0.0


def test_constants_agree_between_spec_and_kernel():
    src = open(os.path.join(ROOT, "copo_b200", "csrc", "sim_core.cuh")).read()
    vals = {m.group(1): np.float32(float(m.group(2).rstrip("f"))) for m in
            re.finditer(r"B2C_F\((\w+),\s*([-+0-9.eE]+f?)\)", src)}
    from oracle import detmath as dm
    alias = {"PI_F": "PI", "TWO_PI_F": "TWO_PI", "HALF_PI_F": "HALF_PI"}
    n = 0
    for name, v in vals.items():
        key = alias.get(name, name)
        if hasattr(osim, key):
            ref = getattr(osim, key)
        elif key in dm.CONSTS:
            ref = dm.CONSTS[key]
        else:
            continue
        assert np.float32(ref).view(np.uint32) == v.view(np.uint32), name
        n += 1
    assert n >= 40
    assert int(re.search(r"NSUB = (\d+)", src).group(1)) == osim.NSUB


def test_det_math_accuracy():
    from oracle.detmath import det_atan2, det_sincos
    x = np.linspace(-20, 20, 20001).astype(np.float32)
    s, c = det_sincos(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 3e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 3e-7
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=5000).astype(np.float32), rng.normal(size=5000).astype(np.float32)
    assert np.abs(det_atan2(a, b) - np.arctan2(a.astype(np.float64), b.astype(np.float64))).max() < 1e-6


@pytest.mark.parametrize("name,odim", [("intersection", 91), ("roundabout", 91), ("parking_lot", 91),
                                       ("bottleneck", 96), ("tollgate", 156), ("pg", 91)])
def test_map_tables(name, odim):
    """Observation widths are the ones the shipped policies pin (first-layer shapes of
    best_checkpoints/*.npz); routes are continuous chains."""
    m = build_map(name)
    assert m.base_obs_dim == odim and 9 + 10 + m.n_ray + m.n_side == odim
    assert m.blob.size % 4 == 0 and m.n_spawn <= 64
    for r in range(m.n_route):
        ids = m.route_seg[r, :m.route_nseg[r]]
        for a, b in zip(ids[:-1], ids[1:]):
            ex, ey = osim.OracleSim._seg_end(m.seg[a])
            assert abs(ex - m.seg[b, 0]) < 1e-3 and abs(ey - m.seg[b, 1]) < 1e-3, (name, r)
    for p in range(m.n_spawn):
        assert 1 <= m.spawn_nroute[p] <= 4
        for k in range(m.spawn_nroute[p]):
            assert m.route_seg[m.spawn_route[p, k], 0] == m.spawn_seg[p]


@pytest.mark.parametrize("seed,num_blocks", [(0, 0), (1, 1), (2, 3), (3, 4), (7, 4), (11, 2)])
def test_procedural_maps(seed, num_blocks):
    """`build_pg`: every seed gives a valid table set (continuous lane chains in both directions, spawn places on the
    entry straights) and the host build of the kernel phases follows the spec on it."""
    m = build_map("pg", seed=seed, num_blocks=num_blocks)
    assert m.base_obs_dim == 91 and m.n_route == 4 and m.n_spawn == 16 and m.blob.size % 4 == 0
    assert (m.route_nseg == num_blocks + 2).all()
    for r in range(m.n_route):
        ids = m.route_seg[r, :m.route_nseg[r]]
        assert m.seg[ids[0], 4] == 0.0 and m.seg[ids[-1], 4] == 0.0
        for a, b in zip(ids[:-1], ids[1:]):
            ex, ey = osim.OracleSim._seg_end(m.seg[a])
            assert abs(ex - m.seg[b, 0]) < 1e-3 and abs(ey - m.seg[b, 1]) < 1e-3, (seed, r)
    # the two carriageways run in opposite directions next to each other: lane 0 of one ends beside lane 0 of the other
    e0 = osim.OracleSim._seg_end(m.seg[m.route_seg[0, m.route_nseg[0] - 1]])
    s2 = m.seg[m.route_seg[2, 0], 0:2]
    assert abs(np.hypot(e0[0] - s2[0], e0[1] - s2[1]) - 3.5) < 1e-2
    assert build_map("pg", seed=seed, num_blocks=num_blocks) is m              # cached per (seed, blocks)
    if num_blocks in (3, 4):
        S, A, cfg = 2, 15, osim.SimConfig(seed=seed)
        cfg.num_agents = A
        ref, hs = osim.OracleSim(m, S, A, cfg), sc.HostSim(m, S, A, cfg)
        r = ref.reset()
        sc.compare_outputs(r, hs.reset(), "reset")
        rng = np.random.default_rng(seed)
        for t in range(60):
            act = _policy(r["obs"], rng, S, A)
            r = ref.step(act)
            sc.compare_outputs(r, hs.step(act), "pg seed %d step %d" % (seed, t))
        sc.compare_state(ref, hs.state(), "pg seed %d" % seed)


def test_laser_ownership_formula():
    """Model of `lidar_spread` (env_step.cu): 32 pairs per warp, each with a window of cnt lasers starting at k0.  Every
    pair's own lane casts its first LIDAR_OWN lasers; the pairs that have more are ranked, pair `rank` owns the warp's
    lasers [excl, excl + rest) and sets bit `excl` of a start bitmap; for the batch of lasers r0 .. r0+31 the owner of
    laser r0 + lane is (start bits before the batch) + popcount(start bits of the batch at positions <= lane) - 1, and
    its laser index is (k0 + LIDAR_OWN - excl) + r wrapped once.  Checked against the plain enumeration: every
    (pair, laser) of every window exactly once."""
    NRAY = 72
    src = open(os.path.join(ROOT, "copo_b200", "csrc", "env_step.cu")).read()
    OWN = int(re.search(r"#define B2C_LIDAR_OWN (\d+)", src).group(1))
    assert 0 <= OWN <= 4
    assert int(re.search(r"SPREAD_BITS_WORDS = (\d+);", src).group(1)) * 32 >= 32 * (NRAY - OWN)
    rng = np.random.default_rng(0)
    for trial in range(300):
        live = rng.random(32) < (1.0 if trial % 5 == 0 else 0.8)
        cnt = np.where(live, rng.integers(1, NRAY + 1, 32), 0)
        if trial % 3 == 0:
            cnt = np.where(live, rng.integers(1, 4, 32), 0)        # short windows: few or no pairs reach the spread
        if trial % 7 == 0:
            cnt = np.where(live, NRAY, 0)                          # every pair the full circle
        k0 = rng.integers(0, NRAY, 32)
        want = sorted((p, (k0[p] + q) % NRAY) for p in range(32) for q in range(cnt[p]))
        got = [(p, (k0[p] + q) % NRAY) for p in range(32) for q in range(min(cnt[p], OWN))]
        rest = np.maximum(cnt - OWN, 0)
        ranked = [p for p in range(32) if rest[p] > 0]
        excl = np.concatenate([[0], np.cumsum(rest)[:-1]])
        total = int(rest.sum())
        bits = [0] * ((total + 31) // 32 + 1)
        table = []
        for p in ranked:
            assert not (bits[excl[p] >> 5] >> (excl[p] & 31)) & 1
            bits[excl[p] >> 5] |= 1 << int(excl[p] & 31)
            code = ((int(k0[p]) + OWN - int(excl[p]) + 4096) << 8) | p
            assert 0 < code < 2 ** 31
            table.append(code)
        before = 0
        for r0 in range(0, total, 32):
            starts = bits[r0 >> 5]
            for lane in range(32):
                r = r0 + lane
                if r >= total:
                    continue
                owner = before + bin(starts & (0xffffffff >> (31 - lane))).count("1") - 1
                code = table[owner]
                k = (code >> 8) - 4096 + r
                assert 0 <= k < 2 * NRAY
                k = min(k, (k - NRAY) & 0xffffffff)
                p = code & 255
                assert excl[p] <= r < excl[p] + rest[p]
                got.append((p, k))
            before += bin(starts).count("1")
        assert sorted(got) == want, trial


def test_side_detector_items_cover_every_slot_and_detector_once():
    """Model of `side_spread` (env_step.cu): with NT threads and n_slots = scenes x slots of the CTA's group, thread t takes
    slot t % n_slots and the detectors part, part + per, ... with part = t / n_slots, per = NT / n_slots (threads with
    part >= per idle).  Every (slot, detector) of a map with n_side detectors exactly once, whenever the kernel turns
    the spreading on (NT >= 2 * n_slots)."""
    src = open(os.path.join(ROOT, "copo_b200", "csrc", "env_step.cu")).read()
    assert "NT >= 2 * ng * A" in src and "tid % n_slots" in src and "kk += per" in src
    for NT, n_slots, n_side in ((128, 40, 65), (128, 30, 65), (128, 64, 65), (256, 40, 65), (128, 20, 16), (96, 40, 65)):
        assert NT >= 2 * n_slots
        per = NT // n_slots
        seen = {}
        for t in range(NT):
            slot, part = t % n_slots, t // n_slots
            if part < per:
                for kk in range(part, n_side, per):
                    seen[(slot, kk)] = seen.get((slot, kk), 0) + 1
        assert len(seen) == n_slots * n_side and set(seen.values()) == {1}, (NT, n_slots, n_side)


def test_sampled_scenes_of_a_large_batch_replay_in_the_oracle():
    """Scenes are independent and keyed by their global index: an oracle built over a few `scene_ids` reproduces
    exactly those scenes of a larger (offset) batch - the mechanism the full-size GPU parity test relies on."""
    name, S, A, offset = "intersection", 40, 40, 3 * 1024
    tables = build_map(name)
    cfg = osim.SimConfig(seed=5, horizon=30)
    cfg.num_agents = A
    pick = [0, 2, 3, 17, 38, 39]
    ref = osim.OracleSim(tables, len(pick), A, cfg, scene_ids=[offset + k for k in pick])
    host = sc.HostSim(tables, S, A, cfg, scene_offset=offset)
    r, g = ref.reset(), host.reset()
    sc.compare_outputs(r, {k: v[pick] for k, v in g.items()}, "reset")
    rng = np.random.default_rng(3)
    done = 0
    for t in range(36):
        act = rng.uniform(-1, 1, (S, A, 2)).astype(np.float32)
        act[..., 0] *= 0.3
        r, g = ref.step(act[pick]), host.step(act)
        sc.compare_outputs(r, {k: v[pick] for k, v in g.items()}, "step %d" % t)
        done += int(r["scene_done"].sum())
    st = host.state()
    sc.compare_state(ref, {k: v[pick] for k, v in st.items()}, "final")
    sc.compare_tiles(st, st)
    assert done == len(pick)
