"""Trains CoPO on the batched Intersection simulator for a few dozen iterations and logs the learning curve
(success rate, reward, LCF) - an end-to-end sanity check of rollout -> GAE x3 -> LCF mix -> PPO -> meta-gradient."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from copo_b200.trainer import CoPOTrainer

def main(iters=30, scenes=256, out="gpurun_out/learn_curve.json"):
    tr = CoPOTrainer(dict(env="MultiAgentIntersectionEnv", num_scenes=scenes, rollout_fragment_length=200,
                          sgd_minibatch_size=8192, num_sgd_iter=5, lcf_num_iters=5, env_config={"num_agents": 30},
                          seed=0))
    log = []
    t0 = time.time()
    for it in range(iters):
        res = tr.train()
        cm = res["custom_metrics"]
        st = res["info"]["learner"]["default"]["learner_stats"]
        row = dict(iter=it, env_steps=res["timesteps_total"], agent_steps=res["agent_timesteps_total"],
                   success_rate=cm["success_rate"], crash_rate=cm["crash_rate"], out_rate=cm["out_of_road_rate"],
                   step_reward=cm["step_reward_mean"], episodes=cm["episodes"], lcf=cm["meta_update"]["lcf"],
                   lcf_std=cm["meta_update"]["lcf_std"], kl=st["kl"], vf_loss=st["vf_loss"], entropy=st["entropy"],
                   sample_ms=res["timers"]["sample_time_ms"], learn_ms=res["timers"]["learn_time_ms"],
                   wall_s=time.time() - t0)
        log.append(row)
        print(json.dumps(row), flush=True)
    json.dump(log, open(out, "w"), indent=1)
    tr.stop()

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 30)
