"""Shared helpers for comparing an implementation of the scene step with oracle/sim.py (tests only)."""
import ctypes
import os
import subprocess

import numpy as np

from oracle import sim as osim

FIELDS = ("x", "y", "h", "v", "steer", "thr", "seg_s", "done_len", "route", "seg_k", "ep_len", "ep_rew", "lcf",
          "status", "agent_id", "yaw")
INT_FIELDS = {"route", "seg_k", "ep_len", "status", "agent_id"}
HDR = ("ep_step", "next_id", "episode", "rng_ctr", "agent_steps")
OUT_KEYS = ("obs", "reward", "flags", "nei_mask", "mf_mask", "nei_reward", "global_reward", "nei_list", "agent_id",
            "lcf", "scene_done")


class EnvConfigC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("S", "A", "AP", "D", "num_agents", "delay_done", "horizon",
                                            "agent_horizon", "allow_respawn", "auto_reset", "append_lcf",
                                            "lcf_uniform", "do_reset", "new_episode", "scene_offset")] + \
               [("seed", ctypes.c_uint32)] + \
               [(n, ctypes.c_float) for n in ("nei_dist", "mf_dist", "lcf_mean", "lcf_std", "force_lcf")]


class HostIOC(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in OUT_KEYS]


def unpack_tiles(tiles, S, A):
    """tiles: uint32 [S, 16*AP+8] -> dict of [S, A] arrays in oracle naming."""
    AP = (A + 3) // 4 * 4
    body = tiles[:, :16 * AP].reshape(S, 16, AP)[:, :, :A]
    out = {}
    for k, name in enumerate(FIELDS):
        w = np.ascontiguousarray(body[:, k, :])
        if name == "status":
            wi = w.view(np.int32)
            out["status"] = wi & 0xff
            out["linger"] = wi >> 8
        elif name in INT_FIELDS:
            out[name] = w.view(np.int32).copy()
        else:
            out[name] = w.view(np.float32).copy()
    hdr = tiles[:, 16 * AP:16 * AP + 8].view(np.int32)
    for k, name in enumerate(HDR):
        out[name] = hdr[:, k].copy()
    return out


def alloc_outputs(S, A, D):
    return dict(obs=np.zeros((S, A, D), np.float32), reward=np.zeros((S, A), np.float32),
                flags=np.zeros((S, A), np.uint8), nei_mask=np.zeros((S, A), np.uint64),
                mf_mask=np.zeros((S, A), np.uint64), nei_reward=np.zeros((S, A), np.float32),
                global_reward=np.zeros(S, np.float32), nei_list=np.zeros((S, A, 4), np.int8),
                agent_id=np.zeros((S, A), np.int32), lcf=np.zeros((S, A), np.float32),
                scene_done=np.zeros(S, np.uint8))


def compare_outputs(ref, got, tag=""):
    """Bit-exact comparison of a step's outputs (floats compared as raw bits)."""
    for k in OUT_KEYS:
        a = np.asarray(ref[k])
        b = np.asarray(got[k])
        if k == "scene_done":
            a = a.astype(np.uint8)
        if a.dtype == np.float32:
            same = a.view(np.uint32) == b.view(np.uint32)
            same |= (a == b)            # +0 / -0
        else:
            same = a == b
        if not same.all():
            idx = np.argwhere(~same)[0]
            raise AssertionError("%s output %s differs at %s: ref %r got %r (%d mismatches)" %
                                 (tag, k, tuple(idx), a[tuple(idx)], b[tuple(idx)], (~same).sum()))


def compare_state(sim, st, tag=""):
    live = sim.status != osim.DISABLED
    present = (sim.status == osim.ACTIVE) | (sim.status == osim.LINGER)
    for name in FIELDS + ("linger",) + HDR:
        a = getattr(sim, name)
        b = st[name]
        if name in HDR:
            ok = a == b
        elif name == "status":
            ok = a == b
        elif name == "linger":
            ok = (a == b) | (sim.status != osim.LINGER)
        else:
            if a.dtype == np.float32:
                ok = (a.view(np.uint32) == b.view(np.uint32)) | (a == b)
            else:
                ok = a == b
            ok = ok | ~present
        if not np.all(ok):
            idx = np.argwhere(~ok)[0]
            raise AssertionError("%s state %s differs at %s: ref %r got %r" % (tag, name, tuple(idx), a[tuple(idx)],
                                                                               b[tuple(idx)]))


def compare_tiles(a, b, tag=""):
    """Two unpacked state dicts (unpack_tiles) of the same batch: headers, status and - for present vehicles - every
    field, bit for bit."""
    present = (a["status"] == osim.ACTIVE) | (a["status"] == osim.LINGER)
    for name in a:
        x, y = a[name], b[name]
        if x.dtype == np.float32:
            ok = (x.view(np.uint32) == y.view(np.uint32)) | (x == y)
        else:
            ok = x == y
        if name == "linger":
            ok = ok | (a["status"] != osim.LINGER)
        elif name not in HDR and name != "status":
            ok = ok | ~present
        if not np.all(ok):
            idx = np.argwhere(~ok)[0]
            raise AssertionError("%s state %s differs at %s: %r vs %r" % (tag, name, tuple(idx), x[tuple(idx)], y[tuple(idx)]))


def make_cfg_c(cfg, S, A, D, do_reset=0, new_episode=0, scene_offset=0):
    c = EnvConfigC()
    c.S, c.A, c.AP, c.D = S, A, (A + 3) // 4 * 4, D
    c.num_agents = cfg.num_agents or A
    c.delay_done, c.horizon, c.agent_horizon = cfg.delay_done, cfg.horizon, cfg.agent_horizon
    c.allow_respawn, c.auto_reset, c.append_lcf = int(cfg.allow_respawn), int(cfg.auto_reset), int(cfg.append_lcf)
    c.lcf_uniform, c.do_reset, c.new_episode, c.scene_offset = int(cfg.lcf_uniform), do_reset, new_episode, scene_offset
    c.seed = cfg.seed
    c.nei_dist, c.mf_dist = float(cfg.neighbours_distance), float(cfg.mf_nei_distance)
    c.lcf_mean, c.lcf_std, c.force_lcf = float(cfg.lcf_mean), float(cfg.lcf_std), float(cfg.force_lcf)
    return c


_HOSTSIM = None


def load_hostsim():
    """Builds (once) and loads the host build of the device phases."""
    global _HOSTSIM
    if _HOSTSIM is None:
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
        so = os.path.join(here, "_hostsim.so")
        src = os.path.join(here, "hostsim.cpp")
        core = os.path.join(here, "..", "..", "copo_b200", "csrc", "sim_core.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so])
        _HOSTSIM = ctypes.CDLL(so)
    return _HOSTSIM


class HostSim:
    """Drives tests/hostsim/_hostsim.so with the same interface as the oracle (reset/step -> outputs dict)."""

    def __init__(self, tables, S, A, cfg, scene_offset=0):
        self.lib = load_hostsim()
        self.m, self.S, self.A, self.cfg = tables, S, A, cfg
        self.D = tables.base_obs_dim + (1 if cfg.append_lcf else 0)
        self.AP = (A + 3) // 4 * 4
        self.tiles = np.zeros((S, 16 * self.AP + 8), np.uint32)
        self.scene_offset = scene_offset

    def _run(self, actions, do_reset, new_episode=0):
        out = alloc_outputs(self.S, self.A, self.D)
        io = HostIOC(*[out[k].ctypes.data for k in OUT_KEYS])
        c = make_cfg_c(self.cfg, self.S, self.A, self.D, do_reset, new_episode, self.scene_offset)
        act = np.ascontiguousarray(actions, np.float32) if actions is not None else np.zeros(1, np.float32)
        rc = self.lib.hostsim_step(ctypes.byref(c), self.m.blob.ctypes.data_as(ctypes.c_void_p), int(self.m.blob.size),
                                   self.tiles.ctypes.data_as(ctypes.c_void_p), act.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.byref(io))
        assert rc == 0
        return out

    def reset(self, new_episode=False):
        return self._run(None, 1, int(new_episode))

    def step(self, actions):
        return self._run(actions, 0)

    def state(self):
        return unpack_tiles(self.tiles, self.S, self.A)
