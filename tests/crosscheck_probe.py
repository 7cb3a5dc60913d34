"""Probe behind profiles/r02_e_metadrive_crosscheck.md (test infrastructure: runs the numpy oracle, CPU only).

Drives oracle/sim.py with the reference's shipped MetaDrive-trained Intersection policies (weights in
tests/golden/metadrive_crosscheck.npz, forward = eval/get_policy_function.py:54-80) under switchable convention variants
of the observation / action vector, and prints arrivals / crashes / out-of-road per finished agent.  The two conventions
it identified (steering sign, order of the two lateral distances) are part of the specification now; `undo_steer` and
`undo_lat` switch them back, the other switches are the hypotheses that were tested and rejected.

    python tests/crosscheck_probe.py copo|ippo [agents] [scenes] [steps] [variant,variant,...]
    python tests/crosscheck_probe.py copo 1 12 400 search      # the 64-way search, single agent per scene
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from copo_b200.maps import build_map  # noqa: E402
from oracle import sim as osim  # noqa: E402

FX = np.load(os.path.join(HERE, "golden", "metadrive_crosscheck.npz"))
SWITCHES = ("undo_steer", "undo_lat", "flip_hd", "flip_latpos", "flip_yaw", "flip_navside")


def policy(name):
    sfx = "_1" if name.startswith("copo") else ""
    W = [FX["%s/default/fc_%s%s/kernel" % (name, layer, sfx)] for layer in ("1", "2", "out")]
    b = [FX["%s/default/fc_%s%s/bias" % (name, layer, sfx)] for layer in ("1", "2", "out")]

    def f(obs, rng):
        x = np.tanh(obs @ W[0] + b[0])
        x = np.tanh(x @ W[1] + b[1])
        x = x @ W[2] + b[2]
        return x[:, :2] + np.exp(x[:, 2:]) * rng.standard_normal((len(x), 2))
    return f


def run(name, variant=(), S=4, A=30, T=600, seed=3):
    copo = name.startswith("copo")
    cfg = osim.SimConfig(seed=seed, append_lcf=copo, lcf_mean=float(FX["copo_inter/lcf"][0]) if copo else 0.0,
                         lcf_std=1e-3, horizon=1000)
    cfg.num_agents = A
    sim = osim.OracleSim(build_map("intersection"), S, A, cfg)
    r, rng, pol = sim.reset(), np.random.default_rng(0), policy(name)
    done = succ = crash = out = 0
    vel = []
    for _ in range(T):
        obs = r["obs"].reshape(S * A, -1).astype(np.float64).copy()
        if "undo_lat" in variant:
            obs[:, [0, 1]] = obs[:, [1, 0]]
        if "flip_hd" in variant:
            obs[:, 2] = 1 - obs[:, 2]
        if "flip_latpos" in variant:
            obs[:, 8] = 1 - obs[:, 8]
        if "flip_yaw" in variant:
            obs[:, 7] = 1 - obs[:, 7]
        if "flip_navside" in variant:
            obs[:, [10, 15]] = 1 - obs[:, [10, 15]]
        if "steer_obs" in variant:
            obs[:, [4, 5]] = 1 - obs[:, [4, 5]]
        for v in variant:
            if v.startswith("lidar_scale="):
                obs[:, 19:91] = np.clip(obs[:, 19:91] * float(v.split("=")[1]), 0, 1)
            if v.startswith("lidar_roll="):
                obs[:, 19:91] = np.roll(obs[:, 19:91], int(v.split("=")[1]), axis=1)
        if "flip_lidar" in variant:
            obs[:, 19:91] = obs[:, 19:91][:, ::-1]
        a = pol(obs, rng).reshape(S, A, 2).astype(np.float32)
        if "undo_steer" in variant:
            a[..., 0] = -a[..., 0]
        r = sim.step(np.clip(a, -1, 1))
        f = r["flags"]
        done += int(((f & osim.F_DONE) > 0).sum())
        succ += int(((f & osim.F_ARRIVE) > 0).sum())
        crash += int(((f & osim.F_CRASH) > 0).sum())
        out += int(((f & osim.F_OUT) > 0).sum())
        valid = (f & osim.F_VALID) > 0
        if valid.any():
            vel.append(float((r["obs"][..., 3] * valid).sum() / valid.sum()) * 80.0)
    d = max(done, 1)
    return dict(variant="+".join(variant) or "spec", finished=done, success=round(succ / d, 3), crash=round(crash / d, 3),
                out=round(out / d, 3), kmh=round(float(np.mean(vel)), 1))


if __name__ == "__main__":
    name = "copo_inter" if (len(sys.argv) < 2 or sys.argv[1] == "copo") else "ippo_inter"
    A, S, T = (int(sys.argv[k]) if len(sys.argv) > k else d for k, d in ((2, 30), (3, 4), (4, 600)))
    what = sys.argv[5] if len(sys.argv) > 5 else ""
    if what == "search":
        rows = [run(name, [n for k, n in enumerate(SWITCHES) if (m >> k) & 1], S, A, T) for m in range(64)]
        for row in sorted(rows, key=lambda q: -q["success"]):
            print(row)
    else:
        print(run(name, [v for v in what.split(",") if v], S, A, T))
