"""TEST INFRASTRUCTURE (oracle) - float64, dict-keyed restatement of the CoPO-owned env wrappers.

Follows copo_code/copo/torch_copo/utils/env_wrappers.py line by line for the arithmetic the
reference itself owns (everything below MetaDrive's `super().step`):

  update_distance_map   env_wrappers.py:141-158   (all pairs, np.linalg.norm in float64, symmetric fill,
                                                   insertion order of the inner dicts preserved)
  find_in_range         env_wrappers.py:125-139   (stable sort by distance, strict `<`, distance<=0 -> empty)
  lcf_step              env_wrappers.py:307-361   (global reward, nei reward = mean or 0.0, coordinated reward)
  add_lcf               env_wrappers.py:393-418   (normal / uniform / forced LCF, clip, obs gets (lcf+1)/2)
  set_lcf_dist          env_wrappers.py:420-426

It works on the per-scene python dicts the reference works on, so tests can feed it the positions and
native rewards of one scene of oracle/sim.py (or of the CUDA kernel) and compare neighbour lists, masks,
nei/global rewards.

PINNING: the reference's own CCEnv.step / _update_distance_map / _find_in_range / LCFEnv.step / _add_lcf, extracted from
/root/reference and executed over a scripted base env (tests/golden/make_ref_golden.py -> tests/golden/ref_golden.npz),
give the neighbour lists, distances, rewards and LCF values tests/test_ref_golden_cpu.py holds this file to.
Only tests/, __graft_entry__.smoke() and bench.py may import this.
"""
from collections import defaultdict
from math import cos, sin

import numpy as np


def update_distance_map(vehicles):
    """vehicles: ordered dict name -> position (2-vector) or None.  env_wrappers.py:141-158."""
    distance_map = defaultdict(lambda: defaultdict(lambda: float("inf")))
    keys = [k for k, v in vehicles.items() if v is not None]            # :149
    for c1 in range(0, len(keys) - 1):                                  # :150
        for c2 in range(c1 + 1, len(keys)):
            k1, k2 = keys[c1], keys[c2]
            p1 = np.asarray(vehicles[k1], dtype=np.float64)
            p2 = np.asarray(vehicles[k2], dtype=np.float64)
            distance = np.linalg.norm(p1 - p2)                          # :156
            distance_map[k1][k2] = distance
            distance_map[k2][k1] = distance
    return distance_map


def find_in_range(distance_map, v_id, distance):
    """env_wrappers.py:125-139: (names, distances) sorted ascending, ties keep insertion order."""
    if distance <= 0:
        return [], []
    dist_to_others = distance_map[v_id]
    order = sorted(dist_to_others, key=lambda k: dist_to_others[k])      # stable
    names = [k for k in order if dist_to_others[k] < distance]
    dists = [dist_to_others[k] for k in order if dist_to_others[k] < distance]
    return names, dists


def add_lcf(agent_obs, lcf, rng, enable_copo=True, force_lcf=-100, lcf_dist="normal", mean=0.0, std=0.1):
    """env_wrappers.py:393-418.  `rng` stands in for metadrive.utils.get_np_random() (unseeded upstream)."""
    if not enable_copo:
        return 0.0, agent_obs
    if force_lcf != -100:
        if lcf_dist == "normal":
            assert -1.0 <= force_lcf <= 1.0
            lcf = float(np.clip(rng.normal(loc=force_lcf, scale=std), -1, 1))
        else:
            lcf = force_lcf
    elif lcf is not None:
        pass
    else:
        if lcf_dist == "normal":
            assert -1.0 <= mean <= 1.0
            lcf = float(np.clip(rng.normal(loc=mean, scale=std), -1, 1))
        else:
            lcf = rng.uniform(-1, 1)
    assert -1.0 <= lcf <= 1.0
    output_lcf = (lcf + 1) / 2
    return lcf, np.float32(np.concatenate([agent_obs, [output_lcf]]))


def lcf_step(rewards, infos, lcf_map, lcf_mode="angle", return_native_reward=True):
    """env_wrappers.py:313-361 on one scene.  rewards: name -> float, infos: name -> {"neighbours": [...]}.

    Fills nei_rewards / global_rewards / lcf / lcf_deg / coordinated_rewards / native_rewards into infos
    and returns the reward dict the env would return.
    """
    new_rewards = {}
    global_reward = sum(rewards.values()) / len(rewards.values())        # :313
    for name, info in infos.items():
        nei = [rewards[n] for n in info["neighbours"]]                   # :321
        info["nei_rewards"] = sum(nei) / len(nei) if nei else 0.0        # :322-325
        info["global_rewards"] = global_reward
        agent_lcf = lcf_map[name]
        info["lcf"] = agent_lcf
        info["lcf_deg"] = agent_lcf * 90
        if lcf_mode == "linear":
            new_r = agent_lcf * rewards[name] + (1 - agent_lcf) * info["nei_rewards"]
        else:
            lcf_rad = agent_lcf * np.pi / 2
            new_r = cos(lcf_rad) * rewards[name] + sin(lcf_rad) * info["nei_rewards"]   # :352-353
        info["coordinated_rewards"] = new_r
        info["native_rewards"] = rewards[name]
        new_rewards[name] = rewards[name] if return_native_reward else new_r
    return new_rewards


def cc_step(positions, rewards, neighbours_distance=40):
    """CCEnv.step info fill (env_wrappers.py:96-102) for one scene: positions and rewards keyed by name, in
    vehicle order.  Returns infos with all_agents / neighbours / neighbours_distance."""
    dm = update_distance_map(positions)
    infos = {}
    for k in rewards.keys():
        names, dists = find_in_range(dm, k, neighbours_distance)
        infos[k] = dict(all_agents=list(rewards.keys()), neighbours=names, neighbours_distance=dists)
    return infos


# ---- disabled-by-default branches of CCEnv / LCFEnv ---------------------------------------
def traffic_light_msg(counter, fix_interval=30):
    """env_wrappers.py:259-266: a saw-tooth in [0.9, 1] / [0, 0.1] that flips every `fix_interval` steps."""
    increment = (counter % fix_interval) / fix_interval * 0.1
    if ((counter // fix_interval) % 2) == 1:
        return 0 + increment
    else:
        return 1 - increment


def agent_traffic_light_msg(msg, pos, b_box):
    """env_wrappers.py:268-272: message + the agent's position normalised by the map's bounding box
    (x_min, x_max, y_min, y_max), clipped to [0, 1], float32."""
    pos0 = (pos[0] - b_box[0]) / (b_box[1] - b_box[0])
    pos1 = (pos[1] - b_box[2]) / (b_box[3] - b_box[2])
    return np.clip(np.array([msg, pos0, pos1]), 0, 1).astype(np.float32)


def comm_current_obs(neighbours, comm_actions, comm_size=4, comm_neighbours=4):
    """env_wrappers.py:104-121 without `add_pos_in_comm`: what agent k hears = the message part of the action of
    each of its first `comm_neighbours` neighbours (zeros for a neighbour that did not act)."""
    out = []
    for n in neighbours[:comm_neighbours]:
        out.append(comm_actions[n] if n in comm_actions else np.zeros((comm_size,)))
    return out


def lcf_comm_obs(old_obs, comm_obs, comm_size=4, comm_neighbours=4):
    """env_wrappers.py:363-372: the heard messages, zero-padded to `comm_neighbours`, appended to the observation."""
    comm_obs = list(comm_obs)
    if len(comm_obs) < comm_neighbours:
        comm_obs.extend([np.zeros((comm_size,))] * (comm_neighbours - len(comm_obs)))
    return np.concatenate([old_obs] + comm_obs).astype(np.float32)
