"""Generates tests/golden/recorder_golden.json.gz by running the REFERENCE's own evaluation wrapper
(copo/eval/recoder.py: `norm`, `DistanceMap`, `RecorderEnv`, extracted with `ast` from /root/reference and exec'd against
stand-ins for gym.Wrapper / deep_update) over episodes of this repo's dict-API environment (host simulator, CPU):

    python tests/golden/make_recorder_golden.py          (build container only: needs /root/reference)

Stored: the recorded stream of every step (rewards, dones, the info entries the recorder reads, vehicle positions) and the
reference's `get_step_result()` every 25 steps and `get_episode_result()` at the end of each episode.  The test replays the
stream through copo_b200.recorder.RecorderEnv and oracle/recorder.py and compares with what the reference returned."""
import ast
import json
import math
import os
import sys
import textwrap
import time
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostenv  # noqa: E402
from copo_b200 import envs  # noqa: E402

REF = "/root/reference/copo_code/copo/eval/recoder.py"
INFO_KEYS = ("velocity", "steering", "step_reward", "acceleration", "cost", "episode_length", "episode_reward", "step_energy",
             "raw_action", "arrive_dest", "crash", "out_of_road", "max_step", "episode_energy")


class Wrapper:                                        # gym.Wrapper, as far as RecorderEnv uses it
    def __init__(self, env):
        self.env = env

    def reset(self, *a, **k):
        return self.env.reset(*a, **k)

    def step(self, *a, **k):
        return self.env.step(*a, **k)

    @property
    def unwrapped(self):
        return self.env


def deep_update(a, b):
    out = dict(a)
    out.update(b)
    return out


def load_reference():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = dict(math=math, time=time, np=np, defaultdict=defaultdict, Wrapper=Wrapper, deep_update=deep_update,
              pretty_print=lambda *a, **k: "")
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in ("norm", "DistanceMap", "RecorderEnv"):
            exec(compile(textwrap.dedent(ast.get_source_segment(src, node)), REF + ":" + node.name, "exec"), ns)
    return ns["RecorderEnv"]


class Replay:
    """Hands the recorded stream to the wrapper: reset() / step() / vehicles, nothing else."""

    def __init__(self, episode):
        self.episode, self.t, self.vehicles = episode, 0, {}

    def reset(self):
        self.t = 0
        return {}

    def step(self, actions=None):
        s = self.episode["steps"][self.t]
        self.t += 1
        self.vehicles = {k: type("V", (), {"position": np.asarray(p, np.float64)})() for k, p in s["vehicles"].items()}
        return {}, dict(s["reward"]), dict(s["done"]), {k: dict(v) for k, v in s["info"].items()}


def _drive(obs):
    return np.array([-np.clip(-1.5 * (obs[2] - 0.5) * 3.14 - 2.0 * (obs[8] - 0.5), -1, 1), 0.6 if obs[3] < 0.35 else 0.0],
                    np.float32)


def record(cls, agents, seed, horizon):
    envs.MultiAgentDrivingEnv.SIM_FACTORY = staticmethod(hostenv.HostBatchedEnv)
    env = envs.get_lcf_env(cls)({"num_agents": agents, "start_seed": seed, "horizon": horizon, "delay_done": 5})
    rng = np.random.default_rng(seed)
    o = env.reset()
    d = {"__all__": False}
    steps = []
    while not d["__all__"]:
        acts = {k: _drive(v) + rng.normal(0, 0.08, 2).astype(np.float32) for k, v in o.items() if k in env.vehicles}
        o, r, d, i = env.step(acts)
        jn = lambda v: [float(x) for x in v] if isinstance(v, (tuple, list, np.ndarray)) else (
            bool(v) if isinstance(v, (bool, np.bool_)) else float(v))
        steps.append(dict(reward={k: float(v) for k, v in r.items()}, done={k: bool(v) for k, v in d.items()},
                          info={k: {q: jn(v[q]) for q in INFO_KEYS if q in v} for k, v in i.items()},
                          vehicles={k: [float(v.position[0]), float(v.position[1])] for k, v in env.vehicles.items()}))
    env.close()
    return dict(map=cls.__name__, agents=agents, seed=seed, steps=steps)


def main():
    Recorder = load_reference()
    out = []
    for cls, agents, seed, horizon in ((envs.MultiAgentIntersectionEnv, 10, 1, 140), (envs.MultiAgentRoundaboutEnv, 8, 2, 120)):
        ep = record(cls, agents, seed, horizon)
        rec = Recorder(Replay(ep), None)
        rec.reset()
        step_results = {}
        for t in range(len(ep["steps"])):
            _, r, d, i = rec.step({})
            if (t + 1) % 25 == 0 and r:
                step_results[str(t)] = {k: float(v) for k, v in rec.get_step_result().items()}
        res = rec.get_episode_result()
        ep["step_results"] = step_results
        ep["episode_result"] = {k: (float(v) if v is not None else None) for k, v in res.items()}
        ep["episode_result_keys"] = list(res.keys())
        print(ep["map"], len(ep["steps"]), "steps;", {k: round(float(v), 4) for k, v in list(res.items())[:6]})
        out.append(ep)
    import gzip
    path = os.path.join(HERE, "recorder_golden.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
