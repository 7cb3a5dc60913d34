"""Device-resident batch of S driving scenes x A agent slots behind the C ABI (b2c_env_*).

This is the native tensor API of the environment; the RLlib-style dict API of the reference
(`MultiAgent*Env.reset/step` wrapped by `get_lcf_env` / `get_ccenv`, utils/env_wrappers.py:30-471) is a
thin host view over it in copo_b200/envs.py.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .maps import build_map

MAP_OF_ENV = {
    "MultiAgentIntersectionEnv": "intersection",
    "MultiAgentRoundaboutEnv": "roundabout",
    "MultiAgentTollgateEnv": "tollgate",
    "MultiAgentBottleneckEnv": "bottleneck",
    "MultiAgentParkingLotEnv": "parking_lot",
    "MultiAgentMetaDrive": "pg",
}
# the reference's eval script passes these explicitly (eval/evaluate_population.py:106-132); "pg" (MultiAgentMetaDrive's
# own default) is recalled, not verifiable here
DEFAULT_NUM_AGENTS = {"intersection": 30, "roundabout": 40, "tollgate": 40, "bottleneck": 20, "parking_lot": 10,
                      "pg": 15}

FLAG_VALID, FLAG_DONE, FLAG_ARRIVE, FLAG_CRASH, FLAG_OUT, FLAG_MAXSTEP, FLAG_SPAWNED, FLAG_ALIVE = (
    1 << k for k in range(8))


class BatchedDrivingEnv:
    def __init__(self, map_name="intersection", num_scenes=1, num_slots=None, num_agents=None, delay_done=25,
                 horizon=1000, agent_horizon=1000, neighbours_distance=40.0, mf_nei_distance=10.0,
                 allow_respawn=True, auto_reset=True, append_lcf=True, lcf_uniform=False, seed=0, scene_offset=0,
                 lcf_mean=0.0, lcf_std=0.1, force_lcf=-100.0, device=None, map_kwargs=None):
        self.lib = _lib.require_device()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.tables = build_map(map_name, **(map_kwargs or {}))
        self.map_name = map_name
        num_agents = num_agents or DEFAULT_NUM_AGENTS[map_name]
        num_slots = num_slots or num_agents
        self.S, self.A = int(num_scenes), int(num_slots)
        self.cfg = _lib.EnvConfig(
            num_scenes=self.S, num_slots=self.A, num_agents=int(num_agents), delay_done=int(delay_done),
            horizon=int(horizon), agent_horizon=int(agent_horizon), allow_respawn=int(allow_respawn),
            auto_reset=int(auto_reset), append_lcf=int(append_lcf), lcf_uniform=int(lcf_uniform),
            scene_offset=int(scene_offset), seed=int(seed) & 0xFFFFFFFF,
            neighbours_distance=float(neighbours_distance), mf_nei_distance=float(mf_nei_distance),
            lcf_mean=float(lcf_mean), lcf_std=float(lcf_std), force_lcf=float(force_lcf))
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            blob = self.tables.blob
            _lib.check(self.lib.b2c_env_create(ctypes.byref(self.cfg), blob.ctypes.data_as(ctypes.c_void_p),
                                               int(blob.size), ctypes.byref(self._h)))
        self.D = int(self.lib.b2c_env_obs_dim(self._h))
        self.tile_words = int(self.lib.b2c_env_state_words(self._h))
        self.split_width = int(self.lib.b2c_env_obs_split_width(self._h))
        self.kernels_per_step = int(self.lib.b2c_env_kernels_per_step(self._h))
        self.num_agents = int(num_agents)
        self.lcf_mean, self.lcf_std, self.force_lcf = float(lcf_mean), float(lcf_std), float(force_lcf)
        self.out = self.alloc_outputs()

    # -- buffers -------------------------------------------------------------------------------------------
    def alloc_outputs(self):
        S, A, D, dev = self.S, self.A, self.D, self.device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        return dict(obs=z((S, A, D), torch.float32), reward=z((S, A), torch.float32), flags=z((S, A), torch.uint8),
                    nei_mask=z((S, A), torch.int64), mf_mask=z((S, A), torch.int64),
                    nei_reward=z((S, A), torch.float32), global_reward=z((S,), torch.float32),
                    nei_list=z((S, A, 4), torch.int8), agent_id=z((S, A), torch.int32), lcf=z((S, A), torch.float32),
                    scene_done=z((S,), torch.uint8))

    def alloc_obs_split(self):
        """[S, A, 2*Kp] bf16 buffer for the optional `obs_split` output (the policy's tensor-core operand)."""
        return torch.zeros((self.S, self.A, self.split_width), dtype=torch.bfloat16, device=self.device)

    def _io(self, out):
        return _lib.EnvIO(*[_lib.ptr(out.get(k)) for k in _lib.ENV_IO_FIELDS])

    # -- stepping ------------------------------------------------------------------------------------------
    def reset(self, out=None, new_episode=False):
        out = out if out is not None else self.out
        io = self._io(out)
        _lib.check(self.lib.b2c_env_reset(self._h, ctypes.byref(io), int(new_episode), _lib.stream_ptr()))
        return out

    def step(self, actions, out=None, scenes=None):
        """actions: float32 device tensor [S, A, 2]; returns the dict of output tensors (views, not copies).
        scenes=(first, count) steps only that range of scenes: `actions` and every tensor of `out` then hold just
        those scenes ([count, A, ...]); stepping the ranges of a partition equals one full step."""
        out = out if out is not None else self.out
        assert actions.dtype == torch.float32 and actions.is_cuda and actions.is_contiguous()
        if scenes is None:
            assert tuple(actions.shape) == (self.S, self.A, 2), actions.shape
            io = self._io(out)
            _lib.check(self.lib.b2c_env_step(self._h, _lib.ptr(actions), ctypes.byref(io), _lib.stream_ptr()))
        else:
            first, count = int(scenes[0]), int(scenes[1])
            assert tuple(actions.shape) == (count, self.A, 2), actions.shape
            assert all(v is None or v.shape[0] == count for v in out.values()), "outputs must hold the scene range"
            io = self._io(out)
            _lib.check(self.lib.b2c_env_step_scenes(self._h, _lib.ptr(actions), ctypes.byref(io), first, count,
                                                    _lib.stream_ptr()))
        _lib.LAUNCHES += self.kernels_per_step - 1          # the launch counter counts kernels, not calls
        return out

    def relaunch_lidar(self, out=None):
        """Two-kernel mode: runs the lidar kernel again on the last step's poses / pairs (same result; for timing)."""
        io = self._io(out if out is not None else self.out)
        _lib.check(self.lib.b2c_env_relaunch_lidar(self._h, ctypes.byref(io), _lib.stream_ptr()))

    # -- host-buffer stepping (what a CPU-side caller of the reference's env.step sees) ---------------------
    HOST_KEYS = ("obs", "reward", "flags", "nei_mask", "nei_reward", "global_reward", "nei_list", "agent_id", "lcf",
                 "scene_done")
    _host = None

    def _host_buffers(self):
        """One device arena and one pinned host arena hold every output of a host-facing step (the kernels write
        straight into the device arena; observations first, everything else behind them).  The scenes are stepped in
        `host_chunks` ranges: a range's observations start crossing PCIe on a dedicated copy stream while the next
        range is still being computed; the small outputs follow as one last copy."""
        if self._host is None:
            offs, size = {}, 0
            for k in self.HOST_KEYS:
                t = self.out[k]
                offs[k] = size
                size += (t.numel() * t.element_size() + 255) & ~255
            self._arena_dev = torch.zeros(size, dtype=torch.uint8, device=self.device)
            self._arena_host = torch.zeros(size, dtype=torch.uint8).pin_memory()
            self._rest_off = offs[self.HOST_KEYS[1]]

            def views(arena):
                d = {}
                for k in self.HOST_KEYS:
                    t = self.out[k]
                    n = t.numel() * t.element_size()
                    d[k] = arena[offs[k]:offs[k] + n].view(t.dtype).view(t.shape)
                return d
            self._host = views(self._arena_host)
            self.host_step_out = dict(self.out)          # device side of the same step (arena views + mf_mask)
            self.host_step_out.update(views(self._arena_dev))
            self.host_step_out["obs_split"] = None
            want = getattr(self, "host_chunks", None) or int(os.environ.get("B2C_HOST_CHUNKS", "4"))
            K = max(1, min(int(want), self.S // 64 if self.S >= 128 else 1))
            per = (self.S + K - 1) // K
            self._chunks = [(f, min(per, self.S - f)) for f in range(0, self.S, per)]
            self._chunk_out = [{k: (v[f:f + n] if v is not None else None) for k, v in self.host_step_out.items()}
                               for f, n in self._chunks]
            self._dev_act = torch.empty((self.S, self.A, 2), dtype=torch.float32, device=self.device)
            self._pin_act = torch.empty((self.S, self.A, 2), dtype=torch.float32).pin_memory()
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._ev_chunk = [torch.cuda.Event() for _ in self._chunks]
            self._ev_copy = torch.cuda.Event()
            self._ev_act = torch.cuda.Event()
            self._act_staged = False
            self._copy_pending = False
            self.h2d_bytes_per_step = self._dev_act.numel() * 4
            self.d2h_bytes_per_step = sum(v.numel() * v.element_size() for v in self._host.values())
        return self._host

    def step_host(self, actions_host, obs_split=None, wait=True, outputs=None):
        """actions_host: float32 [S, A, 2] numpy array or CPU tensor (pinned tensors are read in place).  Copies it
        to the device, steps every scene and returns pinned host tensors of every output (valid until the next call).
        With wait=False the call returns as soon as the work is enqueued - the outputs travel on the copy stream
        while the caller queues more device work on `host_step_out` - and are valid after `wait_host()`.
        `outputs`: names (of HOST_KEYS) the caller reads on the host; the others stay on the device (their host tensors
        keep stale values) - a host policy that needs `("obs", "reward", "flags")` does not pay PCIe time for the masks,
        lists and ids.  Default: everything.  `last_d2h_bytes` is what this call moved."""
        host = self._host_buffers()
        sel = tuple(self.HOST_KEYS) if outputs is None else tuple(outputs)
        assert all(k in self.HOST_KEYS for k in sel), sel
        a = torch.as_tensor(actions_host, dtype=torch.float32)
        assert tuple(a.shape) == (self.S, self.A, 2), a.shape
        cur = torch.cuda.current_stream(self.device)
        if not (a.is_pinned() and a.is_contiguous()):
            if self._act_staged:                             # the previous call's H2D copy may still be reading the
                self._ev_act.synchronize()                   # staging buffer: wait before overwriting it
            a = self._pin_act.copy_(a)
            self._act_staged = True
        if self._copy_pending:                               # a forgotten wait_host() must not race the arena
            cur.wait_event(self._ev_copy)
        self._dev_act.copy_(a, non_blocking=True)
        self._ev_act.record(cur)
        self.host_step_out["obs_split"] = obs_split
        dev_obs, host_obs = self.host_step_out["obs"], host["obs"]
        for c, (f, n) in enumerate(self._chunks):
            out = self._chunk_out[c]
            out["obs_split"] = obs_split[f:f + n] if obs_split is not None else None
            self.step(self._dev_act[f:f + n], out=out, scenes=(f, n))
            self._ev_chunk[c].record(cur)
            self._copy_stream.wait_event(self._ev_chunk[c])
            if "obs" in sel:
                with torch.cuda.stream(self._copy_stream):
                    host_obs[f:f + n].copy_(dev_obs[f:f + n], non_blocking=True)
        moved = host_obs.numel() * 4 if "obs" in sel else 0
        with torch.cuda.stream(self._copy_stream):
            if outputs is None:                              # everything behind the observations: one copy of the arena tail
                self._arena_host[self._rest_off:].copy_(self._arena_dev[self._rest_off:], non_blocking=True)
                moved = self.d2h_bytes_per_step
            else:
                for k in sel:
                    if k != "obs":
                        host[k].copy_(self.host_step_out[k], non_blocking=True)
                        moved += host[k].numel() * host[k].element_size()
            self._ev_copy.record(self._copy_stream)
        self.last_d2h_bytes = moved
        self._copy_pending = True
        if wait:
            self.wait_host()
        return host

    def wait_host(self):
        """Blocks until the outputs of the last `step_host(..., wait=False)` are in the pinned host tensors."""
        if self._host is not None and self._copy_pending:
            self._ev_copy.synchronize()
            self._copy_pending = False

    def agent_steps(self):
        """Running count of agent-env-steps (agents that received an action), summed over scenes."""
        st = self.get_state()
        ap = (self.A + 3) // 4 * 4
        return int(st[:, 16 * ap + 4].astype(np.int64).sum())

    # -- trainer-driven controls (utils/env_wrappers.py:420-430, 450-454) ----------------------------------
    def set_lcf_dist(self, mean, std):
        _lib.check(self.lib.b2c_env_set_lcf_dist(self._h, ctypes.c_float(mean), ctypes.c_float(std)))
        self.lcf_mean, self.lcf_std = float(mean), float(std)

    def set_force_lcf(self, v):
        _lib.check(self.lib.b2c_env_set_force_lcf(self._h, ctypes.c_float(v)))
        self.force_lcf = float(v)

    def set_num_agents(self, n):
        _lib.check(self.lib.b2c_env_set_num_agents(self._h, int(n)))
        self.num_agents = int(n)

    # -- state access (info dicts, tests) -------------------------------------------------------------------
    def get_state(self):
        buf = np.zeros((self.S, self.tile_words), np.uint32)
        _lib.check(self.lib.b2c_env_get_state(self._h, buf.ctypes.data_as(ctypes.c_void_p), _lib.stream_ptr()))
        return buf

    def set_state(self, tiles):
        tiles = np.ascontiguousarray(tiles, np.uint32)
        assert tiles.shape == (self.S, self.tile_words)
        _lib.check(self.lib.b2c_env_set_state(self._h, tiles.ctypes.data_as(ctypes.c_void_p), _lib.stream_ptr()))

    def close(self):
        if self._h:
            self.lib.b2c_env_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
