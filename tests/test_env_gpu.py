"""Parity of the CUDA scene-step kernel (through the C ABI) with oracle/sim.py: bit-exact on every output
and on the full simulator state, for every map, including respawn, lingering wrecks and scene restarts."""
import numpy as np
import pytest
import torch

import simcheck as sc
from copo_b200.maps import build_map
from oracle import sim as osim

pytestmark = pytest.mark.gpu


def _policy(obs, rng, S, A):
    a0 = -np.clip(-1.5 * (obs[..., 2] - 0.5) * 3.14 - 2.0 * (obs[..., 8] - 0.5) + rng.normal(0, 0.02, (S, A)), -1, 1)
    a1 = np.where(obs[..., 3] < 0.35, 0.6, 0.0) + rng.normal(0, 0.05, (S, A))
    a0[0, ::2] = -0.5                                 # scene 0: every other car steers off the road
    return np.stack([a0, a1], -1).astype(np.float32)


def _to_np(out):
    d = {k: v.cpu().numpy() for k, v in out.items()}
    d["nei_mask"] = d["nei_mask"].view(np.uint64)
    d["mf_mask"] = d["mf_mask"].view(np.uint64)
    return d


@pytest.mark.parametrize("map_name,S,A,T,kw", [
    ("intersection", 6, 40, 260, dict(horizon=200)),
    ("roundabout", 5, 40, 200, dict(horizon=150)),
    ("parking_lot", 8, 10, 150, dict()),
    ("tollgate", 3, 40, 120, dict()),
    ("bottleneck", 3, 20, 120, dict(delay_done=0)),
    ("intersection", 3, 40, 80, dict(append_lcf=False, num_agents=30, neighbours_distance=10.0)),
    ("intersection", 2, 40, 60, dict(lcf_uniform=True, allow_respawn=False, auto_reset=False, horizon=40)),
    ("intersection", 5, 64, 50, dict()),                      # maximum slot count (more slots than spawn places)
    ("parking_lot", 33, 1, 40, dict()),                       # single-slot scenes, ragged last CTA group
    ("roundabout", 1, 3, 30, dict(force_lcf=0.5, delay_done=1)),
    ("pg", 4, 15, 150, dict()),                               # procedurally generated map (fused kernel: < 16 slots)
    ("pg", 3, 20, 100, dict()),
])
@pytest.mark.parametrize("split", [0, 1])
def test_env_step_bit_exact(map_name, S, A, T, kw, split, monkeypatch):
    """split = 0: the fused single-kernel step; split = 1: state kernel + lidar kernel (B2C_ENV_SPLIT)."""
    monkeypatch.setenv("B2C_ENV_SPLIT", str(split))
    _run_bit_exact(map_name, S, A, T, kw)


@pytest.mark.parametrize("group", [1, 3])
def test_side_detectors_on_the_slot_thread_and_spread_over_the_cta(group, monkeypatch):
    """Tollgate's 65 side detectors: small batches run one scene per CTA and deal each slot's detectors to three threads
    (the default at this size, group = 1); large batches keep three scenes per CTA with the detectors on the slot's own
    thread (forced here with B2C_ENV_GROUP = 3).  Same bits either way."""
    monkeypatch.setenv("B2C_ENV_SPLIT", "1")
    monkeypatch.setenv("B2C_ENV_GROUP", str(group))
    _run_bit_exact("tollgate", 5, 40, 80, dict())


def _run_bit_exact(map_name, S, A, T, kw):
    from copo_b200.batched_env import BatchedDrivingEnv
    tables = build_map(map_name)
    cfg = osim.SimConfig(seed=11, **kw)
    if cfg.num_agents is None:
        cfg.num_agents = A
    ref = osim.OracleSim(tables, S, A, cfg)
    env = BatchedDrivingEnv(map_name, num_scenes=S, num_slots=A, num_agents=cfg.num_agents, seed=11,
                            delay_done=cfg.delay_done, horizon=cfg.horizon, agent_horizon=cfg.agent_horizon,
                            neighbours_distance=float(cfg.neighbours_distance),
                            mf_nei_distance=float(cfg.mf_nei_distance), allow_respawn=cfg.allow_respawn,
                            auto_reset=cfg.auto_reset, append_lcf=cfg.append_lcf, lcf_uniform=cfg.lcf_uniform,
                            force_lcf=float(cfg.force_lcf))
    r = ref.reset()
    g = _to_np(env.reset())
    sc.compare_outputs(r, g, "reset")
    sc.compare_state(ref, sc.unpack_tiles(env.get_state(), S, A), "reset")
    rng = np.random.default_rng(1)
    seen = dict(arrive=0, crash=0, spawn=0)
    for t in range(T):
        act = _policy(r["obs"], rng, S, A)
        r = ref.step(act)
        g = _to_np(env.step(torch.from_numpy(act).cuda()))
        sc.compare_outputs(r, g, "%s step %d" % (map_name, t))
        if t % 20 == 0 or t == T - 1:
            sc.compare_state(ref, sc.unpack_tiles(env.get_state(), S, A), "%s step %d" % (map_name, t))
        seen["arrive"] += int(((r["flags"] & osim.F_ARRIVE) > 0).sum())
        seen["crash"] += int(((r["flags"] & osim.F_CRASH) > 0).sum())
        seen["spawn"] += int(((r["flags"] & osim.F_SPAWNED) > 0).sum())
    env.close()
    if kw.get("allow_respawn", True) and T >= 200:
        assert seen["crash"] > 0 and seen["spawn"] > 0


@pytest.mark.parametrize("map_name,S,A", [
    ("intersection", 4096, 40),      # C2 (BASELINE.json configs[1])
    ("roundabout", 4096, 40),        # C3
    ("tollgate", 1024, 40),          # C4: 8192 scenes over 8 GPUs
    ("parking_lot", 4096, 10),       # C5: 32768 scenes over 8 GPUs (small-observation path, fused kernel)
])
def test_env_full_size_properties(map_name, S, A):
    """Per-GPU batches of BASELINE.json's configurations: invariants that do not need the oracle."""
    from copo_b200.batched_env import BatchedDrivingEnv, FLAG_VALID, FLAG_DONE, FLAG_SPAWNED, FLAG_ALIVE
    env = BatchedDrivingEnv(map_name, num_scenes=S, num_slots=A, num_agents=A, seed=0)
    o = env.reset()
    assert int(((o["flags"] & FLAG_SPAWNED) > 0).sum()) == S * A
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(60):
        act = torch.rand((S, A, 2), device="cuda", generator=gen) * 2 - 1
        act[..., 0] *= 0.2
        o = env.step(act)
    f = o["flags"]
    obs = o["obs"]
    assert torch.isfinite(obs).all() and obs.min() >= 0 and obs.max() <= 1
    part = ((f & FLAG_VALID) > 0) | ((f & FLAG_SPAWNED) > 0)
    assert (obs[~part] == 0).all()
    # neighbour masks are symmetric and never contain self
    m = o["nei_mask"]
    bit = (m.unsqueeze(-1) >> torch.arange(A, device="cuda")) & 1          # [S, A, A]
    assert torch.equal(bit, bit.transpose(1, 2))
    assert int(torch.diagonal(bit, dim1=1, dim2=2).sum()) == 0
    # the global reward is the mean over participants
    rew = o["reward"] * part
    g = rew.sum(1) / part.sum(1).clamp(min=1)
    assert torch.allclose(g, o["global_reward"], atol=1e-5)
    # identical scenes+seed reproduce bit-identically
    env2 = BatchedDrivingEnv(map_name, num_scenes=S, num_slots=A, num_agents=A, seed=0)
    env2.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(60):
        act = torch.rand((S, A, 2), device="cuda", generator=gen) * 2 - 1
        act[..., 0] *= 0.2
        o2 = env2.step(act)
    for k in o:
        assert torch.equal(o[k], o2[k]), k


def _sampled_scenes(S, A, n_random=6):
    """first / last scenes, the scenes either side of CTA-group edges (the state kernel packs 128 // A scenes per CTA,
    the last group can be ragged), and a few random ones"""
    g = max(1, 128 // A)
    last_group = (S - 1) // g * g
    pick = {0, 1, g - 1, g, 2 * g - 1, S // 2, last_group - 1, last_group, S - 2, S - 1}
    pick |= set(int(k) for k in np.random.default_rng(S + A).integers(0, S, n_random))
    return sorted(k for k in pick if 0 <= k < S)


@pytest.mark.parametrize("map_name,S,A,offset", [
    ("intersection", 4096, 40, 0),        # C2 (BASELINE.json configs[1])
    ("roundabout", 4096, 40, 0),          # C3
    ("tollgate", 1024, 40, 3 * 1024),     # C4: 8192 scenes over 8 GPUs, the shard of rank 3
    ("parking_lot", 4096, 10, 7 * 4096),  # C5: 32768 scenes over 8 GPUs, the shard of rank 7 (fused kernel)
])
def test_env_full_size_oracle_parity(map_name, S, A, offset):
    """Per-GPU batches of BASELINE.json's configurations, bit for bit: sampled scenes of the full-size run are replayed
    by the numpy oracle (scenes are independent and keyed by their global index), and EVERY scene is compared with the
    host build of the phases (tests/hostsim), every output, every step."""
    from copo_b200.batched_env import BatchedDrivingEnv
    T = 70
    tables = build_map(map_name)
    cfg = osim.SimConfig(seed=5, horizon=60)              # scene restarts inside the run
    cfg.num_agents = A
    pick = _sampled_scenes(S, A)
    ref = osim.OracleSim(tables, len(pick), A, cfg, scene_ids=[offset + k for k in pick])
    host = sc.HostSim(tables, S, A, cfg, scene_offset=offset)
    env = BatchedDrivingEnv(map_name, num_scenes=S, num_slots=A, num_agents=A, seed=5, horizon=60, scene_offset=offset)
    r = ref.reset()
    h = host.reset()
    g = _to_np(env.reset())
    sc.compare_outputs(h, g, "%s reset (host build, all scenes)" % map_name)
    sc.compare_outputs(r, {k: v[pick] for k, v in g.items()}, "%s reset (oracle, sampled scenes)" % map_name)
    rng = np.random.default_rng(3)
    seen = dict(crash=0, spawn=0, done=0)
    for t in range(T):
        act = _policy(g["obs"], rng, S, A)
        r = ref.step(act[pick])
        h = host.step(act)
        g = _to_np(env.step(torch.from_numpy(act).cuda()))
        sc.compare_outputs(h, g, "%s step %d (host build, all scenes)" % (map_name, t))
        sc.compare_outputs(r, {k: v[pick] for k, v in g.items()}, "%s step %d (oracle, sampled scenes)" % (map_name, t))
        seen["crash"] += int(((r["flags"] & osim.F_CRASH) > 0).sum())
        seen["spawn"] += int(((r["flags"] & osim.F_SPAWNED) > 0).sum())
        seen["done"] += int(r["scene_done"].sum())
    st = sc.unpack_tiles(env.get_state(), S, A)
    sc.compare_state(ref, {k: v[pick] for k, v in st.items()}, "%s final state (oracle)" % map_name)
    sc.compare_tiles(host.state(), st, "%s final state (host build, all scenes)" % map_name)
    env.close()
    assert seen["spawn"] > 0 and seen["done"] == len(pick), seen       # respawns and one restart per sampled scene


@pytest.mark.parametrize("split", [0, 1])
def test_env_emits_the_policy_operand(split, monkeypatch):
    """`obs_split` equals the [hi | lo] bf16 split of `obs` (bit for bit) - the env saves the policy a pass."""
    monkeypatch.setenv("B2C_ENV_SPLIT", str(split))
    from copo_b200 import ops
    from copo_b200.batched_env import BatchedDrivingEnv
    for name, A in (("intersection", 40), ("tollgate", 40), ("parking_lot", 10)):
        env = BatchedDrivingEnv(name, num_scenes=37, num_slots=A, num_agents=A, seed=2)
        out = dict(env.out)
        out["obs_split"] = env.alloc_obs_split()
        env.reset(out=out)
        gen = torch.Generator(device="cuda").manual_seed(0)
        for t in range(5):
            env.step(torch.rand((37, A, 2), device="cuda", generator=gen) * 2 - 1, out=out)
        want = ops.tc_split_rows(out["obs"].reshape(37 * A, -1))
        assert env.split_width == want.shape[1]
        assert torch.equal(out["obs_split"].reshape(37 * A, -1).view(torch.int16), want.view(torch.int16))
        env.close()


@pytest.mark.parametrize("S,chunks", [(50, 1), (300, 4), (200, 3)])
def test_step_host_matches_device_step(S, chunks):
    """The host-buffer step (scenes stepped in ranges, each range's observations copied on the copy stream while the
    next range is computed; blocking or overlapped with later device work) returns exactly what the one-launch device
    step leaves in HBM."""
    from copo_b200.batched_env import BatchedDrivingEnv
    A = 40
    e_dev = BatchedDrivingEnv("intersection", num_scenes=S, num_slots=A, num_agents=A, seed=4)
    e_host = BatchedDrivingEnv("intersection", num_scenes=S, num_slots=A, num_agents=A, seed=4)
    e_host.host_chunks = chunks
    e_dev.reset()
    e_host.reset()
    rng = np.random.default_rng(0)
    pinned = torch.zeros((S, A, 2)).pin_memory()
    split = e_host.alloc_obs_split()
    for t in range(12):
        act = rng.uniform(-1, 1, (S, A, 2)).astype(np.float32)
        want = e_dev.step(torch.from_numpy(act).cuda())
        if t % 3 == 0:
            got = e_host.step_host(act)                              # numpy in, blocking
        elif t % 3 == 1:
            got = e_host.step_host(torch.from_numpy(act), obs_split=split)     # pageable tensor in, blocking
        else:
            pinned.copy_(torch.from_numpy(act))
            got = e_host.step_host(pinned, wait=False)               # pinned in place, overlapped
            busy = e_host.host_step_out["obs"].sum()                 # device work queued behind the step
            e_host.wait_host()
            assert torch.isfinite(busy)
        assert len(e_host._chunks) == chunks
        for k in BatchedDrivingEnv.HOST_KEYS:
            assert not got[k].is_cuda
            assert torch.equal(got[k], want[k].cpu()), (t, k)
            assert torch.equal(e_host.host_step_out[k], want[k]), (t, k)
        if t % 3 == 1:
            from copo_b200 import ops
            ws = ops.tc_split_rows(want["obs"].reshape(S * A, -1))
            assert torch.equal(split.reshape(S * A, -1).view(torch.int16), ws.view(torch.int16))
    assert e_host.d2h_bytes_per_step == sum(want[k].numel() * want[k].element_size()
                                            for k in BatchedDrivingEnv.HOST_KEYS)
    # a host consumer that only reads observations, rewards and flags: the other outputs stay on the device
    act = rng.uniform(-1, 1, (S, A, 2)).astype(np.float32)
    want = e_dev.step(torch.from_numpy(act).cuda())
    stale = e_host._host["nei_mask"].clone()
    got = e_host.step_host(act, outputs=("obs", "reward", "flags"))
    for k in ("obs", "reward", "flags"):
        assert torch.equal(got[k], want[k].cpu()), k
    assert torch.equal(got["nei_mask"], stale) and torch.equal(e_host.host_step_out["nei_mask"], want["nei_mask"])
    assert e_host.last_d2h_bytes == sum(want[k].numel() * want[k].element_size() for k in ("obs", "reward", "flags"))
    e_dev.close()
    e_host.close()


@pytest.mark.parametrize("name,S,A", [("intersection", 64, 40), ("tollgate", 20, 40), ("parking_lot", 50, 10),
                                      ("intersection", 9, 64)])
def test_optional_outputs_left_out_do_not_change_the_others(name, S, A):
    """The scene step computes the mean-field mask and the nearest-neighbour list only on request (CoPO without a fused
    critic input asks for neither; the kernel then takes the reward sum in a uniform loop instead of the mask walk):
    every other output and the simulator state stay bit-identical."""
    from copo_b200.batched_env import BatchedDrivingEnv
    full = BatchedDrivingEnv(name, num_scenes=S, num_slots=A, num_agents=A, seed=9)
    lean = BatchedDrivingEnv(name, num_scenes=S, num_slots=A, num_agents=A, seed=9)
    out_lean = dict(lean.out)
    out_lean["mf_mask"] = None
    out_lean["nei_list"] = None
    a, b = full.reset(), lean.reset(out=out_lean)
    rng = np.random.default_rng(3)
    for t in range(80):
        act = rng.uniform(-1, 1, (S, A, 2)).astype(np.float32)
        act[..., 0] *= 0.3
        a = full.step(torch.from_numpy(act).cuda())
        b = lean.step(torch.from_numpy(act).cuda(), out=out_lean)
        for k, v in a.items():
            if k in ("mf_mask", "nei_list") or v is None:
                continue
            assert torch.equal(v.view(torch.uint8) if v.dtype.is_floating_point else v,
                               b[k].view(torch.uint8) if b[k].dtype.is_floating_point else b[k]), (t, k)
    assert np.array_equal(full.get_state(), lean.get_state())
    full.close()
    lean.close()
