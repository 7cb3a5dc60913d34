"""TEST-ONLY: a host stand-in for copo_b200.batched_env.BatchedDrivingEnv built on tests/hostsim (the host build of the
device phases), so the dict-API environments and wrappers of copo_b200/envs.py - host logic - run in the CPU suite.
Never imported by the package."""
import numpy as np
import torch

import simcheck as sc
from copo_b200.maps import build_map
from oracle import sim as osim


class HostBatchedEnv:
    def __init__(self, map_name="intersection", num_scenes=1, num_slots=None, num_agents=None, delay_done=25,
                 horizon=1000, agent_horizon=1000, neighbours_distance=40.0, mf_nei_distance=10.0, allow_respawn=True,
                 auto_reset=True, append_lcf=True, lcf_uniform=False, seed=0, scene_offset=0, lcf_mean=0.0, lcf_std=0.1,
                 force_lcf=-100.0, device=None, map_kwargs=None):
        self.tables = build_map(map_name, **(map_kwargs or {}))
        self.S, self.A = int(num_scenes), int(num_slots or num_agents)
        self.cfg = osim.SimConfig(num_agents=num_agents or self.A, delay_done=delay_done, horizon=horizon,
                                  agent_horizon=agent_horizon, neighbours_distance=neighbours_distance,
                                  mf_nei_distance=mf_nei_distance, allow_respawn=allow_respawn, auto_reset=auto_reset,
                                  append_lcf=append_lcf, seed=seed, lcf_mean=lcf_mean, lcf_std=lcf_std,
                                  force_lcf=force_lcf, lcf_uniform=lcf_uniform)
        self.sim = sc.HostSim(self.tables, self.S, self.A, self.cfg, scene_offset)
        self.D = self.sim.D
        self.device = torch.device("cpu")

    @staticmethod
    def _wrap(out):
        d = {}
        for k, v in out.items():
            v = np.ascontiguousarray(v)
            d[k] = torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v)
        return d

    def reset(self, out=None, new_episode=False):
        return self._wrap(self.sim.reset(new_episode))

    def step(self, actions, out=None):
        return self._wrap(self.sim.step(actions.numpy()))

    def get_state(self):
        return self.sim.tiles.copy()

    def set_lcf_dist(self, mean, std):
        self.cfg.lcf_mean, self.cfg.lcf_std = osim.f32(mean), osim.f32(std)

    def set_force_lcf(self, v):
        self.cfg.force_lcf = osim.f32(v)

    def set_num_agents(self, n):
        self.cfg.num_agents = int(n)

    def close(self):
        pass
