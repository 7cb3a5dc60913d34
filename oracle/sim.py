"""TEST INFRASTRUCTURE (oracle) - CPU specification of the batched multi-agent driving step.

PARITY UNPINNED for the simulator arithmetic: the reference delegates env.step to MetaDrive 0.2.5
(`super().step` at copo_code/copo/torch_copo/utils/env_wrappers.py:95,309; package pinned at
README.md:41-42), which is neither vendored under /root/reference nor installable here, and the
reference has no tests or golden vectors for it.  This file is therefore the
repo's own written-down spec of a MetaDrive-style step (kinematic bicycle, route following,
72-laser LiDAR against oriented boxes, crash / out-of-road / arrival / horizon, reward, delayed
removal and respawn) and the CUDA kernel copo_b200/csrc/env_step.cu must reproduce it BIT FOR BIT
(every float op below is one binary32 operation in the order written; the kernel is compiled with
-fmad=false).  Two of its conventions ARE anchored in reference-held data: the steering sign and the order of the two
lateral-distance observations were identified with the reference's shipped MetaDrive-trained policies
(tools/metadrive_crosscheck.py, profiles/r02_e_metadrive_crosscheck.md).

The CoPO-owned bookkeeping layered on the step *does* follow reference lines:
  * neighbour search     env_wrappers.py:125-158  (`_update_distance_map`, `_find_in_range`:
                          ascending distance, ties keep vehicle order, strict `<`)
  * nei / global reward  env_wrappers.py:313-326  (mean over neighbours, 0.0 when none; global =
                          mean over every agent returned this step, respawned agents count with 0)
  * LCF draw + obs append env_wrappers.py:393-418  (N(mean, std) clipped to [-1, 1], obs gets (lcf+1)/2)
Numeric type is pinned to float32 (the simulator state is float32); see `reference_wrappers` in
oracle/wrappers.py for the float64 restatement used to bound the difference.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
import numpy as np

from .detmath import (HALF_PI, PI, TWO_PI, det_atan2, det_sincos, f32, rng_u32, u32_to_unit, wrap_pi)

# ---- constants of the spec (mirrored in copo_b200/csrc/sim_consts.h) -----------------------------
DT = f32(0.02)
NSUB = 5
MAX_STEER = f32(0.6981317007977318)      # 40 deg
ACC = f32(3.5)
BRAKE = f32(8.0)
VMAX = f32(22.22222137451172)            # 80 km/h
WHEELBASE = f32(2.5)
HALF_L = f32(2.25)
HALF_W = f32(0.9)
LIDAR_RANGE = f32(40.0)
INV_LIDAR_RANGE = f32(0.025)
LIDAR_CULL = f32(42.5)                   # range + inflated circum-radius (kernel broad phase only)
ARRIVE_DIST = f32(5.0)
BACK_MARGIN = f32(5.0)
END_MARGIN = f32(2.0)
NAVI_SCALE = f32(0.01)                   # 1 / (2 * 50 m)
SPAWN_LONG = f32(7.0)
SPAWN_LAT = f32(2.5)
SUCCESS_REWARD = f32(10.0)
OUT_PENALTY = f32(10.0)
CRASH_PENALTY = f32(10.0)
SPEED_W = f32(0.1)
YAW_SCALE = f32(0.25)
KAPPA_SCALE = f32(5.0)
INV_PI = f32(0.3183098861837907)
ZERO = f32(0.0)
ONE = f32(1.0)
HALF = f32(0.5)

EMPTY, ACTIVE, LINGER, DISABLED = 0, 1, 2, 3
F_VALID, F_DONE, F_ARRIVE, F_CRASH, F_OUT, F_MAXSTEP, F_SPAWNED, F_ALIVE = (1 << k for k in range(8))
EGO_DIM, NAVI_DIM = 9, 10
NEI_K = 4                                 # nearest-neighbour list width (CCPPO concat, algo_ccppo.py:41)


def _clip01(x):
    x = np.where(x < ZERO, ZERO, x)
    x = np.where(x > ONE, ONE, x)
    return x.astype(f32)


class SimConfig:
    def __init__(self, num_agents=None, delay_done=25, horizon=1000, agent_horizon=1000, neighbours_distance=40.0,
                 mf_nei_distance=10.0, allow_respawn=True, auto_reset=True, append_lcf=True, seed=0,
                 lcf_mean=0.0, lcf_std=0.1, force_lcf=-100.0, lcf_uniform=False):
        self.num_agents = num_agents
        self.delay_done = int(delay_done)
        self.horizon = int(horizon)
        self.agent_horizon = int(agent_horizon)
        self.neighbours_distance = f32(neighbours_distance)
        self.mf_nei_distance = f32(mf_nei_distance)
        self.allow_respawn = bool(allow_respawn)
        self.auto_reset = bool(auto_reset)
        self.append_lcf = bool(append_lcf)
        self.seed = int(seed)
        self.lcf_mean = f32(lcf_mean)
        self.lcf_std = f32(lcf_std)
        self.force_lcf = f32(force_lcf)
        self.lcf_uniform = bool(lcf_uniform)


class OracleSim:
    """S scenes x A slots, numpy float32, vectorised over scenes."""

    def __init__(self, tables, num_scenes, num_slots, cfg: SimConfig, scene_offset=0, scene_ids=None):
        """scene_ids: optional explicit global scene indices (replaying sampled scenes of a larger batch: scenes are
        independent and only keyed by their global index through the RNG stream)."""
        self.m = tables
        self.S, self.A = int(num_scenes), int(num_slots)
        self.cfg = cfg
        if cfg.num_agents is None:
            cfg.num_agents = self.A
        S, A = self.S, self.A
        z = lambda dt: np.zeros((S, A), dtype=dt)
        self.x, self.y, self.h, self.v = z(f32), z(f32), z(f32), z(f32)
        self.steer, self.thr, self.seg_s, self.done_len = z(f32), z(f32), z(f32), z(f32)
        self.route, self.seg_k, self.ep_len = z(np.int32), z(np.int32), z(np.int32)
        self.ep_rew, self.lcf, self.yaw = z(f32), z(f32), z(f32)
        self.status, self.linger, self.agent_id = z(np.int32), z(np.int32), z(np.int32)
        self.ep_step = np.zeros(S, np.int32)
        self.next_id = np.zeros(S, np.int32)
        self.episode = np.zeros(S, np.int32)
        self.rng_ctr = np.zeros(S, np.int32)
        self.agent_steps = np.zeros(S, np.int32)          # running count of agent-env-steps (metric counter)
        self.scene_id = (np.arange(S) + scene_offset).astype(np.int64)
        if scene_ids is not None:
            self.scene_id = np.asarray(scene_ids, dtype=np.int64).reshape(S)
        self.obs_dim = tables.base_obs_dim + (1 if cfg.append_lcf else 0)

    # ------------------------------------------------------------------------------------------
    def _draw(self, mask):
        """One u32 per scene where mask; advances the per-scene counter there."""
        u = rng_u32(self.cfg.seed, self.scene_id, self.episode, self.rng_ctr)
        self.rng_ctr = np.where(mask, self.rng_ctr + 1, self.rng_ctr).astype(np.int32)
        return u

    def _seg(self, seg_id):
        return self.m.seg[seg_id]            # [..., 12]

    def _localize(self, seg_id, x, y):
        sg = self._seg(seg_id)
        x0, y0, slen, kappa, c0, s0, cx, cy, r = (sg[..., k] for k in (0, 1, 3, 4, 7, 8, 9, 10, 11))
        # straight
        dx = x - x0
        dy = y - y0
        s_st = dx * c0 + dy * s0
        l_st = dy * c0 - dx * s0
        # arc
        ex = x - cx
        ey = y - cy
        rho = np.sqrt(ex * ex + ey * ey)
        sgn = np.where(kappa > ZERO, ONE, -ONE).astype(f32)
        px = sgn * (ex * s0 - ey * c0)
        py = ex * c0 + ey * s0
        dl = det_atan2(py, px)
        thr = (slen * np.abs(kappa) - TWO_PI) * HALF
        dl = np.where(dl < thr, dl + TWO_PI, dl)
        s_arc = r * dl
        l_arc = sgn * (r - rho)
        is_arc = kappa != ZERO
        return np.where(is_arc, s_arc, s_st).astype(f32), np.where(is_arc, l_arc, l_st).astype(f32)

    # ------------------------------------------------------------------------------------------
    def reset(self, new_episode=False):
        S, A = self.S, self.A
        if new_episode:
            self.episode = (self.episode + 1).astype(np.int32)
        self.status[:] = DISABLED
        self.status[:, :self.cfg.num_agents] = EMPTY
        self.linger[:] = 0
        self.ep_step[:] = 0
        self.next_id[:] = 0
        self.rng_ctr[:] = 0
        zero_r = np.zeros((S, A), f32)
        zero_f = np.zeros((S, A), np.int32)
        return self._finish_step(zero_r, zero_f, np.zeros((S, A), bool), force_spawn=True)

    def step(self, actions):
        """actions [S, A, 2] float32.  Returns dict of per-step outputs."""
        S, A, m, cfg = self.S, self.A, self.m, self.cfg
        act = np.asarray(actions, dtype=f32)
        active = self.status == ACTIVE
        self.ep_step = (self.ep_step + 1).astype(np.int32)
        self.agent_steps = (self.agent_steps + active.sum(axis=1)).astype(np.int32)

        # ---- 1. kinematic bicycle, NSUB sub-steps ------------------------------------------------
        a0 = np.where(act[..., 0] < -ONE, -ONE, np.where(act[..., 0] > ONE, ONE, act[..., 0])).astype(f32)
        # MetaDrive's steering sign: a positive action turns the vehicle towards DEcreasing heading in this frame.  Identified
        # from the reference's shipped MetaDrive-trained policies (profiles/r02_e_metadrive_crosscheck.md): with this sign
        # and obs[0] / obs[1] in the order below they drive this simulator's roads (single agent: success 0.76 - 0.98),
        # with the opposite sign or order none of them ever arrives.  The stored steering (obs[4], obs[5]) keeps this sign.
        a0 = (-a0).astype(f32)
        a1 = np.where(act[..., 1] < -ONE, -ONE, np.where(act[..., 1] > ONE, ONE, act[..., 1])).astype(f32)
        st = a0 * MAX_STEER
        ss, cs = det_sincos(st)
        tan_st = ss / cs
        acc = np.where(a1 >= ZERO, a1 * ACC, a1 * BRAKE).astype(f32)
        x, y, h, v = self.x.copy(), self.y.copy(), self.h.copy(), self.v.copy()
        yaw = np.zeros_like(v)
        for _ in range(NSUB):
            v = v + acc * DT
            v = np.where(v < ZERO, ZERO, v)
            v = np.where(v > VMAX, VMAX, v).astype(f32)
            yaw = (v * tan_st) / WHEELBASE
            h = h + yaw * DT
            sh, ch = det_sincos(h)
            x = x + (v * ch) * DT
            y = y + (v * sh) * DT
        h = wrap_pi(h)
        long_last = self.done_len + self.seg_s
        self.x = np.where(active, x, self.x).astype(f32)
        self.y = np.where(active, y, self.y).astype(f32)
        self.h = np.where(active, h, self.h).astype(f32)
        self.v = np.where(active, v, self.v).astype(f32)
        self.yaw = np.where(active, yaw, self.yaw).astype(f32)
        self.steer = np.where(active, a0, self.steer).astype(f32)
        self.thr = np.where(active, a1, self.thr).astype(f32)

        # ---- 2. localisation on the route --------------------------------------------------------
        k = self.seg_k.copy()
        done_len = self.done_len.copy()
        nseg = m.route_nseg[self.route]
        seg_id = m.route_seg[self.route, k]
        s, l = self._localize(seg_id, self.x, self.y)
        for _ in range(2):
            slen = m.seg[seg_id, 3]
            adv = active & (s > slen) & (k < nseg - 1)
            done_len = np.where(adv, done_len + slen, done_len).astype(f32)
            k = np.where(adv, k + 1, k).astype(np.int32)
            seg_id = m.route_seg[self.route, k]
            s2, l2 = self._localize(seg_id, self.x, self.y)
            s = np.where(adv, s2, s)
            l = np.where(adv, l2, l)
        self.seg_k = np.where(active, k, self.seg_k).astype(np.int32)
        self.done_len = np.where(active, done_len, self.done_len).astype(f32)
        self.seg_s = np.where(active, s, self.seg_s).astype(f32)
        sg = m.seg[seg_id]
        slen, wl, wr = sg[..., 3], sg[..., 5], sg[..., 6]
        last = k == nseg - 1
        out = (l > wl) | (l < -wr) | (s < -BACK_MARGIN) | (last & (s > slen + END_MARGIN))
        arrive = last & (s > slen - ARRIVE_DIST) & ~out
        long_now = done_len + s

        # ---- 3. vehicle-vehicle overlap (SAT on oriented boxes) ----------------------------------
        present = (self.status == ACTIVE) | (self.status == LINGER)
        sn, cn = det_sincos(self.h)
        crash = self._sat_crash(present, active, self.x, self.y, cn, sn)

        # ---- 4. reward / termination -------------------------------------------------------------
        q = self.v / VMAX
        q = SPEED_W * q
        rew = (long_now - long_last) + q
        rew = np.where(arrive, SUCCESS_REWARD, np.where(out, -OUT_PENALTY, np.where(crash, -CRASH_PENALTY, rew)))
        rew = np.where(active, rew, ZERO).astype(f32)
        ep_len = self.ep_len + 1
        done = arrive | out | crash
        maxstep = ~done & ((ep_len >= cfg.agent_horizon) | (self.ep_step >= cfg.horizon)[:, None])
        done = (done | maxstep) & active
        self.ep_len = np.where(active, ep_len, self.ep_len).astype(np.int32)
        self.ep_rew = np.where(active, self.ep_rew + rew, self.ep_rew).astype(f32)
        flags = np.zeros((S, A), np.int32)
        flags |= np.where(active, F_VALID, 0)
        flags |= np.where(done, F_DONE, 0)
        flags |= np.where(active & arrive, F_ARRIVE, 0)
        flags |= np.where(active & crash, F_CRASH, 0)
        flags |= np.where(active & out, F_OUT, 0)
        flags |= np.where(active & maxstep, F_MAXSTEP, 0)

        # ---- 5. lingering wrecks count down; just-finished agents start lingering ----------------
        ling = self.status == LINGER
        self.linger = np.where(ling, self.linger - 1, self.linger).astype(np.int32)
        self.status = np.where(ling & (self.linger <= 0), EMPTY, self.status).astype(np.int32)
        if cfg.delay_done > 0:
            self.status = np.where(done, LINGER, self.status).astype(np.int32)
            self.linger = np.where(done, cfg.delay_done, self.linger).astype(np.int32)
            self.v = np.where(done, ZERO, self.v).astype(f32)
        else:
            self.status = np.where(done, EMPTY, self.status).astype(np.int32)
        return self._finish_step(rew, flags, active)

    # ------------------------------------------------------------------------------------------
    def _sat_crash(self, present, active, x, y, c, s):
        xi, xj = x[:, :, None], x[:, None, :]
        yi, yj = y[:, :, None], y[:, None, :]
        ci, cj = c[:, :, None], c[:, None, :]
        si, sj = s[:, :, None], s[:, None, :]
        dx = xj - xi
        dy = yj - yi
        abs_c = np.abs(ci * cj + si * sj)
        abs_s = np.abs(ci * sj - si * cj)
        ext_l = HALF_L + (HALF_L * abs_c + HALF_W * abs_s)
        ext_w = HALF_W + (HALF_L * abs_s + HALF_W * abs_c)
        sep = np.abs(dx * ci + dy * si) > ext_l
        sep |= np.abs(dy * ci - dx * si) > ext_w
        sep |= np.abs(dx * cj + dy * sj) > ext_l
        sep |= np.abs(dy * cj - dx * sj) > ext_w
        eye = np.eye(self.A, dtype=bool)[None]
        hit = ~sep & present[:, :, None] & present[:, None, :] & ~eye
        hit &= active[:, :, None] | active[:, None, :]
        return hit.any(axis=2)

    # ------------------------------------------------------------------------------------------
    def _finish_step(self, rew, flags, acted, force_spawn=False):
        S, A, m, cfg = self.S, self.A, self.m, self.cfg
        # ---- 6. scene horizon: everything is cleared and the scene restarts ----------------------
        scene_done = (self.ep_step >= cfg.horizon) & (not force_spawn)
        if cfg.auto_reset:
            rs = scene_done
            self.status = np.where(rs[:, None] & (self.status != DISABLED), EMPTY, self.status).astype(np.int32)
            self.episode = np.where(rs, self.episode + 1, self.episode).astype(np.int32)
            self.ep_step = np.where(rs, 0, self.ep_step).astype(np.int32)
            self.next_id = np.where(rs, 0, self.next_id).astype(np.int32)
            self.rng_ctr = np.where(rs, 0, self.rng_ctr).astype(np.int32)
            self.linger = np.where(rs[:, None], 0, self.linger).astype(np.int32)
            can_spawn = np.ones(S, bool)
        else:
            can_spawn = ~scene_done
        if not (cfg.allow_respawn or force_spawn):
            can_spawn = can_spawn & (self.ep_step == 0) & (self.next_id == 0)

        # ---- 7. respawn into EMPTY slots, slot order ---------------------------------------------
        spawned = np.zeros((S, A), bool)
        n_sp = m.n_spawn
        for i in range(A):
            need = (self.status[:, i] == EMPTY) & can_spawn
            if not need.any():
                continue
            u = self._draw(need)
            start = (u % np.uint32(n_sp)).astype(np.int64)
            found = np.zeros(S, bool)
            place = np.zeros(S, np.int64)
            present = (self.status == ACTIVE) | (self.status == LINGER)
            for qq in range(n_sp):
                p = (start + qq) % n_sp
                px, py, pc, ps = (m.spawn_f[p, kk] for kk in (0, 1, 3, 4))
                dx = self.x - px[:, None]
                dy = self.y - py[:, None]
                lon = dx * pc[:, None] + dy * ps[:, None]
                lat = dy * pc[:, None] - dx * ps[:, None]
                block = present & (np.abs(lon) < SPAWN_LONG) & (np.abs(lat) < SPAWN_LAT)
                free = ~block.any(axis=1)
                take = need & ~found & free
                place = np.where(take, p, place)
                found |= take
            if not found.any():
                continue
            u_route = self._draw(found)
            nr = m.spawn_nroute[place].astype(np.uint32)
            ridx = (u_route % np.maximum(nr, 1)).astype(np.int64)
            rid = m.spawn_route[place, ridx]
            # LCF draw: Irwin-Hall(12) - 6 standard normal surrogate, 12 sequential float32 adds
            z = np.zeros(S, f32)
            for _ in range(12):
                z = z + u32_to_unit(self._draw(found))
            z = z - f32(6.0)
            if cfg.lcf_uniform:
                uni = u32_to_unit(self._draw(found)) * f32(2.0) - ONE
            forced = cfg.force_lcf != f32(-100.0)
            if cfg.lcf_uniform:
                lcf = np.full(S, cfg.force_lcf, f32) if forced else uni
            else:
                lcf = (cfg.force_lcf if forced else cfg.lcf_mean) + cfg.lcf_std * z
            lcf = np.where(lcf < -ONE, -ONE, np.where(lcf > ONE, ONE, lcf)).astype(f32)
            if not cfg.append_lcf:
                lcf = np.zeros(S, f32)
            sf = m.spawn_f[place]
            upd = lambda arr, val: np.where(found, val, arr[:, i]).astype(arr.dtype)
            self.x[:, i] = upd(self.x, sf[:, 0])
            self.y[:, i] = upd(self.y, sf[:, 1])
            self.h[:, i] = upd(self.h, sf[:, 2])
            self.v[:, i] = upd(self.v, ZERO)
            self.steer[:, i] = upd(self.steer, ZERO)
            self.thr[:, i] = upd(self.thr, ZERO)
            self.yaw[:, i] = upd(self.yaw, ZERO)
            self.seg_s[:, i] = upd(self.seg_s, sf[:, 5])
            self.done_len[:, i] = upd(self.done_len, ZERO)
            self.route[:, i] = upd(self.route, rid)
            self.seg_k[:, i] = upd(self.seg_k, 0)
            self.ep_len[:, i] = upd(self.ep_len, 0)
            self.ep_rew[:, i] = upd(self.ep_rew, ZERO)
            self.lcf[:, i] = upd(self.lcf, lcf)
            self.agent_id[:, i] = upd(self.agent_id, self.next_id)
            self.status[:, i] = upd(self.status, ACTIVE)
            self.linger[:, i] = upd(self.linger, 0)
            self.next_id = np.where(found, self.next_id + 1, self.next_id).astype(np.int32)
            spawned[:, i] = found
        flags = flags | np.where(spawned, F_SPAWNED, 0)
        flags = flags | np.where(self.status == ACTIVE, F_ALIVE, 0)

        # ---- 8. neighbours and shared rewards (env_wrappers.py:125-158, 313-326) -----------------
        part = acted | spawned
        nb = self._neighbours(part, rew)
        # ---- 9. observations ---------------------------------------------------------------------
        lcf_now = self._step_lcf()
        obs = self._observe(part, lcf_now)
        out = dict(obs=obs, reward=rew.astype(f32), flags=flags.astype(np.uint8), scene_done=scene_done,
                   agent_id=self.agent_id.copy(), lcf=lcf_now.copy(), **nb)
        return out

    def _step_lcf(self):
        """The LCF the wrapper hands out this step: the episode value (lcf_map), except under a forced mean with the
        normal distribution, where the reference redraws it at every call of _add_lcf, i.e. every step of every agent,
        and leaves lcf_map alone (env_wrappers.py:337-342, 398-403).  Counter-based: (scene, episode, episode step,
        slot), counters with bit 31 set (the sequential scene stream never gets there)."""
        cfg = self.cfg
        if (not cfg.append_lcf) or cfg.lcf_uniform or cfg.force_lcf == f32(-100.0):
            return self.lcf
        S, A = self.S, self.A
        slot = np.arange(A, dtype=np.uint64)[None, :]
        ctr0 = np.uint64(0x80000000) | (self.ep_step.astype(np.uint64)[:, None] << np.uint64(10)) | (slot << np.uint64(4))
        z = np.zeros((S, A), f32)
        for t in range(12):
            z = z + u32_to_unit(rng_u32(cfg.seed, self.scene_id[:, None], self.episode[:, None], ctr0 + np.uint64(t)))
        z = z - f32(6.0)
        lcf = cfg.force_lcf + cfg.lcf_std * z
        return np.where(lcf < -ONE, -ONE, np.where(lcf > ONE, ONE, lcf)).astype(f32)

    # ------------------------------------------------------------------------------------------
    def _neighbours(self, part, rew):
        S, A, cfg = self.S, self.A, self.cfg
        dx = self.x[:, :, None] - self.x[:, None, :]
        dy = self.y[:, :, None] - self.y[:, None, :]
        # Distances are compared SQUARED (float32 dx*dx + dy*dy against the float32 square of the radius): the
        # reference compares float64 Euclidean norms (env_wrappers.py:133, 156; algo_ccppo.py:283), which is the same
        # test in exact arithmetic, and the squared form has one rounding less than a float32 square root (fewer
        # artificial ties in the nearest-neighbour order).
        dist = (dx * dx + dy * dy).astype(f32)                     # squared
        eye = np.eye(A, dtype=bool)[None]
        pair = part[:, :, None] & part[:, None, :] & ~eye
        inr = pair & (dist < f32(cfg.neighbours_distance * cfg.neighbours_distance))
        mf = inr & ~(dist > f32(cfg.mf_nei_distance * cfg.mf_nei_distance))
        bits = (np.uint64(1) << np.arange(A, dtype=np.uint64))[None, None, :]
        nei_mask = np.where(inr, bits, np.uint64(0)).sum(axis=2, dtype=np.uint64)
        mf_mask = np.where(mf, bits, np.uint64(0)).sum(axis=2, dtype=np.uint64)
        cnt = inr.sum(axis=2).astype(np.int32)
        # sequential float32 sums in slot order
        nsum = np.zeros((S, A), f32)
        gsum = np.zeros(S, f32)
        for j in range(A):
            nsum = nsum + np.where(inr[:, :, j], rew[:, j][:, None], ZERO)
            gsum = gsum + np.where(part[:, j], rew[:, j], ZERO)
        with np.errstate(divide="ignore", invalid="ignore"):
            nei_rew = np.where(cnt > 0, nsum / np.maximum(cnt, 1).astype(f32), ZERO).astype(f32)
            pc = part.sum(axis=1).astype(np.int32)
            glob = np.where(pc > 0, gsum / np.maximum(pc, 1).astype(f32), ZERO).astype(f32)
        # K nearest in (distance, slot) order: rank by counting
        dkey = np.where(inr, dist, f32(np.inf))
        lt = dkey[:, :, None, :] < dkey[:, :, :, None]            # [s,i,j,k]: d_ik < d_ij
        eq = (dkey[:, :, None, :] == dkey[:, :, :, None]) & (np.arange(A)[None, None, None, :] <
                                                             np.arange(A)[None, None, :, None])
        rank = (lt | eq).sum(axis=3)
        rank = np.where(inr, rank, A + NEI_K)          # sentinel beyond every k (also when A <= NEI_K)
        nei_list = np.full((S, A, NEI_K), -1, np.int8)
        for kk in range(NEI_K):
            sel = rank == kk
            has = sel.any(axis=2)
            nei_list[:, :, kk] = np.where(has, sel.argmax(axis=2), -1)
        nei_mask = np.where(part, nei_mask, np.uint64(0))
        return dict(nei_mask=nei_mask, mf_mask=np.where(part, mf_mask, np.uint64(0)),
                    nei_count=np.where(part, cnt, 0).astype(np.int32),
                    nei_reward=np.where(part, nei_rew, ZERO).astype(f32),
                    global_reward=glob, nei_list=np.where(part[:, :, None], nei_list, -1).astype(np.int8),
                    dist=dist)

    # ------------------------------------------------------------------------------------------
    def _observe(self, part, lcf_now=None):
        S, A, m, cfg = self.S, self.A, self.m, self.cfg
        D = self.obs_dim
        obs = np.zeros((S, A, D), f32)
        nseg = m.route_nseg[self.route]
        seg_id = m.route_seg[self.route, self.seg_k]
        sg = m.seg[seg_id]
        h0, slen, kappa, wl, wr = sg[..., 2], sg[..., 3], sg[..., 4], sg[..., 5], sg[..., 6]
        s, l = self._localize(seg_id, self.x, self.y)
        sn, cn = det_sincos(self.h)
        tw = wl + wr
        obs[..., 0] = _clip01((l + wr) / tw)              # MetaDrive's order of the two lateral distances (see step())
        obs[..., 1] = _clip01((wl - l) / tw)
        lane_h = h0 + kappa * s
        hd = wrap_pi(self.h - lane_h)
        obs[..., 2] = _clip01(hd * INV_PI + HALF)
        obs[..., 3] = _clip01(self.v / VMAX)
        obs[..., 4] = _clip01(self.steer * HALF + HALF)
        obs[..., 5] = _clip01(self.steer * HALF + HALF)
        obs[..., 6] = _clip01(self.thr * HALF + HALF)
        obs[..., 7] = _clip01(self.yaw * YAW_SCALE + HALF)
        obs[..., 8] = _clip01(l / tw + HALF)
        # navigation: end points of the current and the next route segment, in the ego frame
        k2 = np.where(self.seg_k + 1 < nseg, self.seg_k + 1, self.seg_k)
        for c, kk in enumerate((self.seg_k, k2)):
            sid = m.route_seg[self.route, kk]
            g = m.seg[sid]
            ex, ey = self._seg_end(g)
            rx = ex - self.x
            ry = ey - self.y
            ahead = rx * cn + ry * sn
            side = ry * cn - rx * sn
            b = EGO_DIM + 5 * c
            obs[..., b + 0] = _clip01(ahead * NAVI_SCALE + HALF)
            obs[..., b + 1] = _clip01(side * NAVI_SCALE + HALF)
            kap = g[..., 4]
            obs[..., b + 2] = _clip01(np.abs(kap) * KAPPA_SCALE)
            obs[..., b + 3] = np.where(kap > ZERO, ONE, np.where(kap < ZERO, ZERO, HALF))
            obs[..., b + 4] = _clip01((g[..., 3] * np.abs(kap)) * INV_PI)
        # lidar
        lid = self._lidar(cn, sn)
        b = EGO_DIM + NAVI_DIM
        obs[..., b:b + m.n_ray] = lid
        b += m.n_ray
        # side detector (tollgate / bottleneck widths): distance to the drivable edge along fixed bearings
        if m.n_side > 0:
            obs[..., b:b + m.n_side] = self._side(l, wl, wr, hd, m.n_side)
            b += m.n_side
        if cfg.append_lcf:
            obs[..., b] = ((self.lcf if lcf_now is None else lcf_now) + ONE) * HALF
        return np.where(part[:, :, None], obs, ZERO).astype(f32)

    @staticmethod
    def _seg_end(g):
        x0, y0, h0, slen, kappa, c0, s0, cx, cy, r = (g[..., k] for k in (0, 1, 2, 3, 4, 7, 8, 9, 10, 11))
        xs = x0 + slen * c0
        ys = y0 + slen * s0
        h1 = h0 + kappa * slen
        s1, c1 = det_sincos(h1)
        sgn = np.where(kappa > ZERO, ONE, -ONE).astype(f32)
        xa = cx + (sgn * r) * s1
        ya = cy - (sgn * r) * c1
        arc = kappa != ZERO
        return np.where(arc, xa, xs).astype(f32), np.where(arc, ya, ys).astype(f32)

    def _side(self, l, wl, wr, hd, n_side):
        """Distance to the lane edge along n_side bearings spread over the half plane ahead (own spec)."""
        out = np.zeros(l.shape + (n_side,), f32)
        for k in range(n_side):
            ang = f32(-1.5) + f32(3.0) * f32(k) / f32(max(n_side - 1, 1))
            sa, ca = det_sincos(hd + ang)
            with np.errstate(divide="ignore", invalid="ignore"):
                dl = np.where(sa > ZERO, (wl - l) / sa, np.where(sa < ZERO, (-wr - l) / sa, LIDAR_RANGE))
            dl = np.where(dl < ZERO, ZERO, dl)
            out[..., k] = _clip01(dl * INV_LIDAR_RANGE)
        return out

    def _lidar(self, cn, sn):
        """Brute force: every laser of every slot against every other present box (slab test in the box frame).

        Per ordered pair (i observes j): ego position (ox, oy) in j's frame and the relative rotation
        C = cos(th_i - th_j), SN = sin(th_i - th_j); per laser k with ego-frame direction (rx, ry):
        direction in j's frame (rx*C - ry*SN, ry*C + rx*SN), two correctly rounded reciprocals, four products.
        """
        S, A, m = self.S, self.A, self.m
        R = m.n_ray
        present = (self.status == ACTIVE) | (self.status == LINGER)
        rx = m.ray[:, 0][None, None, :]
        ry = m.ray[:, 1][None, None, :]
        best = np.full((S, A, R), LIDAR_RANGE * INV_LIDAR_RANGE, f32)      # == 1.0
        with np.errstate(divide="ignore", invalid="ignore"):
            for j in range(A):
                relx = self.x - self.x[:, j][:, None]            # [S,A]
                rely = self.y - self.y[:, j][:, None]
                cj = cn[:, j][:, None]
                sj = sn[:, j][:, None]
                ox = relx * cj + rely * sj
                oy = rely * cj - relx * sj
                cc = (cn * cj + sn * sj)[:, :, None]
                ss = (sn * cj - cn * sj)[:, :, None]
                nx1 = (-HALF_L - ox)[:, :, None]
                nx2 = (HALF_L - ox)[:, :, None]
                ny1 = (-HALF_W - oy)[:, :, None]
                ny2 = (HALF_W - oy)[:, :, None]
                ddx = rx * cc - ry * ss                          # [S,A,R]
                ddy = ry * cc + rx * ss
                ix = ONE / ddx
                iy = ONE / ddy
                t1 = nx1 * ix
                t2 = nx2 * ix
                tnx = np.where(t1 < t2, t1, t2)
                tfx = np.where(t1 < t2, t2, t1)
                t3 = ny1 * iy
                t4 = ny2 * iy
                tny = np.where(t3 < t4, t3, t4)
                tfy = np.where(t3 < t4, t4, t3)
                tn = np.where(tnx > tny, tnx, tny)
                tf = np.where(tfx < tfy, tfx, tfy)
                hit = (tn <= tf) & (tf >= ZERO)
                t = np.where(tn > ZERO, tn, ZERO)
                ts = (t * INV_LIDAR_RANGE).astype(f32)
                ok = hit & present[:, j][:, None, None] & (np.arange(A) != j)[None, :, None]
                best = np.where(ok & (ts < best), ts, best)
        return best.astype(f32)

    # convenience for tests -------------------------------------------------------------------------
    def state_dict(self):
        keys = ("x", "y", "h", "v", "steer", "thr", "seg_s", "done_len", "route", "seg_k", "ep_len", "ep_rew", "lcf",
                "status", "linger", "agent_id", "yaw", "ep_step", "next_id", "episode", "rng_ctr", "agent_steps")
        return {k: getattr(self, k).copy() for k in keys}
