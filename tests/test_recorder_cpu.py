"""`copo_b200.recorder.RecorderEnv` (the reference's evaluation wrapper, copo/eval/recoder.py) against the dict-keyed
restatement in oracle/recorder.py, on episodes of the dict-API environment (host simulator behind it)."""
import numpy as np
import pytest

import hostenv
from copo_b200 import envs
from copo_b200.recorder import RecorderEnv
from oracle import recorder as orec

REFERENCE_COLUMNS = [       # the columns eval.py writes (shipped eval/demo_results/evaluate_results/*.csv, minus the old coll_*)
    "velocity_step_mean_episode_min", "velocity_step_mean_episode_mean", "velocity_step_mean_episode_max",
    "energy_step_mean_episode_min", "energy_step_mean_episode_mean", "energy_step_mean_episode_max",
    "num_neighbours_mean_episode_mean", "num_neighbours_mean_episode_max", "num_agents_total",
    "num_agents_total_per_300_steps", "success_rate", "num_agents_success", "num_agents_success_per_300_steps",
    "num_agents_failed_per_300_steps", "episode_reward_mean", "episode_reward_min", "episode_reward_max",
    "episode_cost_mean", "episode_cost_min", "episode_cost_max", "episode_cost_sum", "crash_rate", "num_agents_crash",
    "out_rate", "num_agents_out", "episode_length_mean", "success_episode_length_mean", "svo_estimate_deg_mean",
    "svo_estimate_deg_min", "svo_estimate_deg_max", "svo_reward"]


@pytest.fixture(autouse=True)
def host_simulator(monkeypatch):
    monkeypatch.setattr(envs.MultiAgentDrivingEnv, "SIM_FACTORY", staticmethod(hostenv.HostBatchedEnv))


def _drive(obs):
    """lane keeping + speed control from the observation (heading error, lateral offset, speed)."""
    return np.array([-np.clip(-1.5 * (obs[2] - 0.5) * 3.14 - 2.0 * (obs[8] - 0.5), -1, 1),    # a positive action turns right
                     0.6 if obs[3] < 0.35 else 0.0], np.float32)


@pytest.mark.parametrize("cls,agents,seed", [(envs.MultiAgentIntersectionEnv, 16, 1), (envs.MultiAgentRoundaboutEnv, 12, 2)])
def test_episode_report_matches_the_reference_recorder(cls, agents, seed):
    env = RecorderEnv(envs.get_lcf_env(cls)({"num_agents": agents, "start_seed": seed, "horizon": 260}))
    assert env.eval_config == {"neighbours_distance": 20} and env.unwrapped is env.env
    ref = orec.Recorder(20)
    rng = np.random.default_rng(seed)
    for episode in range(2):
        o = env.reset()
        ref.episode_step = 0
        d = {"__all__": False}
        steps = 0
        while not d["__all__"]:
            acts = {k: _drive(v) + rng.normal(0, 0.05, 2).astype(np.float32) for k, v in o.items() if k in env.vehicles}
            o, r, d, i = env.step(acts)
            ref.step({k: v.position for k, v in env.vehicles.items()}, r, d, i)
            steps += 1
            if steps % 50 == 0 and r:
                a, b = env.get_step_result(), ref.step_result()
                assert set(a) == set(b), (set(a) ^ set(b))
                for k in a:
                    assert np.isclose(a[k], b[k], rtol=1e-9, atol=1e-12), (k, a[k], b[k])
        got, want = env.get_episode_result(), ref.episode_result()
        assert list(got.keys()) == list(want.keys()) == REFERENCE_COLUMNS
        for k in want:
            assert np.isclose(got[k], want[k], rtol=1e-9, atol=1e-12), (episode, k, got[k], want[k])
        assert got["num_agents_total"] > agents / 2 and 0 < got["velocity_step_mean_episode_mean"] < 80
        assert got["success_rate"] + got["crash_rate"] + got["out_rate"] <= 1 + 1e-9          # the rest: max_step
        assert got["num_agents_success"] > 0 and got["num_agents_crash"] + got["num_agents_out"] > 0


class _Replay:
    """Feeds a recorded step stream (tests/golden/recorder_golden.json.gz) to a recorder wrapper."""

    def __init__(self, episode):
        self.episode, self.t, self.vehicles = episode, 0, {}

    def reset(self):
        self.t = 0
        return {}

    def step(self, actions=None):
        import types
        s = self.episode["steps"][self.t]
        self.t += 1
        self.vehicles = {k: types.SimpleNamespace(position=np.asarray(p, np.float64)) for k, p in s["vehicles"].items()}
        return {}, dict(s["reward"]), dict(s["done"]), {k: dict(v) for k, v in s["info"].items()}


@pytest.mark.parametrize("index", [0, 1])
def test_episode_report_equals_what_the_reference_class_returned(index):
    """The reference's own `RecorderEnv` (copo/eval/recoder.py, executed in the build container by
    tests/golden/make_recorder_golden.py over recorded episodes of the dict-API env) against this repo's wrapper and the
    oracle restatement, replaying the same stream: step reports every 25 steps and the 31-column episode report."""
    import gzip
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "recorder_golden.json.gz")
    ep = json.load(gzip.open(path, "rt"))[index]
    env = RecorderEnv(_Replay(ep))
    ref = orec.Recorder(20)
    env.reset()
    ref.episode_step = 0
    for t in range(len(ep["steps"])):
        _, r, d, i = env.step({})
        ref.step({k: v.position for k, v in env.vehicles.items()}, r, d, i)
        want = ep["step_results"].get(str(t))
        if want is not None:
            for got in (env.get_step_result(), ref.step_result()):
                assert set(got) == set(want), set(got) ^ set(want)
                for k, v in want.items():
                    assert np.isclose(got[k], v, rtol=1e-9, atol=1e-12), (t, k, got[k], v)
    want = ep["episode_result"]
    for got in (env.get_episode_result(), ref.episode_result()):
        assert [k for k in ep["episode_result_keys"] if k in got] == list(got.keys())
        assert set(REFERENCE_COLUMNS) <= set(want)
        for k in got:
            assert np.isclose(got[k], want[k], rtol=1e-9, atol=1e-12), (k, got[k], want[k])
    assert want["num_agents_total"] >= 3 and len(ep["step_results"]) >= 4
