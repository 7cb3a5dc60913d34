"""TEST INFRASTRUCTURE (oracle) - CPU restatement of the rollout bookkeeping the reference runs per trajectory.

  discount_cumsum / compute_advantages   ray 2.2.0 rllib/evaluation/postprocessing.py (not vendored; used at
                                         torch_copo/algo_ccppo.py:14,366 and algo_copo.py:17,192,201); restated
                                         from its published form: A = lfilter([1],[1,-g*l], delta[::-1])[::-1]
  compute_nei_advantage                  torch_copo/algo_copo.py:189-195
  compute_global_advantage               torch_copo/algo_copo.py:198-204   (called with gamma = 1.0, :498-500)
  bootstrap rule                         torch_copo/algo_ccppo.py:362-365, algo_copo.py:492-496: last_r = 0 when the
                                         trajectory's last row is done, else the value prediction OF THAT LAST ROW
  lcf_mix_standardize                    torch_copo/algo_copo.py:539-551 (+ rllib.utils.sgd.standardized:
                                         (x - mean) / max(1e-4, std), population std)

PINNING: `discount_cumsum` is pinned by scipy.signal.lfilter itself (the very call rllib makes) and by the closed form
checked in tests/test_bookkeeping_cpu.py; compute_nei_advantage / compute_global_advantage by the reference's own
functions executed in the build container (tests/golden/ref_golden.npz, bit-equal in tests/test_ref_golden_cpu.py).

`rollout_gae3` is the driver that cuts the batched [T, N] rollout columns (N = scenes x slots) into the
per-agent trajectories RLlib would hand to postprocess_trajectory and applies the functions above to each.
Only tests/, __graft_entry__.smoke() and bench.py may import this.
"""
import numpy as np
import scipy.signal

FLAG_VALID, FLAG_DONE = 1, 2


def discount_cumsum(x, gamma):
    return scipy.signal.lfilter([1], [1, float(-gamma)], x[::-1], axis=0)[::-1]


def compute_advantages(rewards, vf_preds, last_r, gamma=0.99, lambda_=0.95):
    """use_gae=True, use_critic=True branch.  Returns (advantages f32, value_targets f32)."""
    vpred_t = np.concatenate([vf_preds, np.array([last_r])])
    delta_t = rewards + gamma * vpred_t[1:] - vpred_t[:-1]
    adv = discount_cumsum(delta_t, gamma * lambda_)
    targets = (adv + vf_preds).astype(np.float32)
    return adv.astype(np.float32), targets


def compute_nei_advantage(nei_rewards, nei_values, last_r, gamma=0.99, lambda_=0.95):      # algo_copo.py:189-195
    return compute_advantages(nei_rewards, nei_values, last_r, gamma, lambda_)


def compute_global_advantage(global_rewards, global_values, last_r, gamma=1.0, lambda_=0.95):   # :198-204
    return compute_advantages(global_rewards, global_values, last_r, gamma, lambda_)


def trajectories(flags_col):
    """Row index lists of the per-agent trajectories inside one (scene, slot) column of a rollout fragment:
    consecutive VALID rows, cut after each DONE row (a new agent may take the slot later)."""
    out, cur = [], []
    for t, f in enumerate(flags_col):
        if not (f & FLAG_VALID):
            continue
        cur.append(t)
        if f & FLAG_DONE:
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def rollout_gae3(flags, rewards, values, nei_rewards, nei_values, glob_rewards, glob_values, gamma=0.99,
                 lambda_=0.95, heads=3, next_values=None):
    """All inputs [T, N] (flags uint8, the rest float32).  Returns dict of [T, N] float32 arrays, zero where the
    row is not valid.  heads=1 computes only the native advantage (IPPO / CCPPO).
    next_values ([N], heads=1 only): the value of the observation after the fragment's last row - stock rllib PPO
    (`compute_gae_for_sample_batch`, which IPPOPolicy inherits: algo_ippo.py:79 subclasses PPOTorchPolicy without
    overriding postprocess_trajectory) bootstraps a trajectory that is not done with V(NEXT_OBS of its last row)."""
    T, N = flags.shape
    names = ["advantages", "value_targets", "nei_advantage", "nei_target", "global_advantages", "global_target"]
    out = {k: np.zeros((T, N), np.float32) for k in names[:2 * heads]}
    for n in range(N):
        for rows in trajectories(flags[:, n]):
            rows = np.asarray(rows)
            done = bool(flags[rows[-1], n] & FLAG_DONE)
            v = values[rows, n]
            last_r = 0.0 if done else v[-1]                                  # algo_ccppo.py:362-365
            if next_values is not None and not done and rows[-1] == T - 1:
                last_r = next_values[n]                                      # stock rllib: V(NEXT_OBS)
            a, tg = compute_advantages(rewards[rows, n], v, last_r, gamma, lambda_)
            out["advantages"][rows, n], out["value_targets"][rows, n] = a, tg
            if heads == 3:
                nv, gv = nei_values[rows, n], glob_values[rows, n]
                last_nei = 0.0 if done else nv[-1]                           # algo_copo.py:492-496
                last_glob = 0.0 if done else gv[-1]
                a, tg = compute_nei_advantage(nei_rewards[rows, n], nv, last_nei, gamma, lambda_)
                out["nei_advantage"][rows, n], out["nei_target"][rows, n] = a, tg
                a, tg = compute_global_advantage(glob_rewards[rows, n], gv, last_glob, 1.0, lambda_)
                out["global_advantages"][rows, n], out["global_target"][rows, n] = a, tg
    return out


def standardized(x):
    """rllib.utils.sgd.standardized (ray 2.2.0): (x - x.mean()) / max(1e-4, x.std())."""
    return (x - x.mean()) / max(1e-4, x.std())


def lcf_mix_standardize(advantages, nei_advantage, step_lcf, global_advantages):
    """algo_copo.py:539-551 on the flat train batch (1-D float32 arrays of the valid rows).
    Returns (normalized_advantages, raw_mean, raw_std, standardized global advantages)."""
    used_lcf = step_lcf * np.pi / 2
    norm = np.cos(used_lcf) * advantages + np.sin(used_lcf) * nei_advantage
    raw_mean = norm.mean()
    raw_std = max(1e-4, norm.std())
    return standardized(norm), raw_mean, raw_std, standardized(global_advantages)
