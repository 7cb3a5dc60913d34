"""torch_copo/utils/callbacks.py of the reference.  Its per-episode hooks (on_episode_start / step / end, :15-111) walk
RLlib's episode objects agent by agent; on the batched simulator the same quantities are device-side reductions over the
rollout columns (copo_b200.trainer.IPPOTrainer._episode_metrics -> `custom_metrics["*_mean"]`), so only
`on_train_result` (:118-147) has work left: it lifts them to the top-level result keys Tune reports."""
import numpy as np


class MultiAgentDrivingCallbacks:
    def on_episode_start(self, **kwargs):
        pass

    def on_episode_step(self, **kwargs):
        pass

    def on_episode_end(self, **kwargs):
        pass

    def on_train_result(self, *, algorithm=None, result=None, **kwargs):
        cm = result.get("custom_metrics", {})
        result["success"] = cm.get("success_rate_mean", np.nan)
        result["crash"] = cm.get("crash_rate_mean", np.nan)
        result["out"] = cm.get("out_of_road_rate_mean", np.nan)
        result["max_step"] = cm.get("max_step_rate_mean", np.nan)
        result["length"] = result.get("episode_len_mean", np.nan)
        result["rc"] = cm.get("route_completion_mean", np.nan)
        result["cost"] = cm.get("episode_cost_mean", np.nan)
        result["raw_episode_reward_mean"] = result.get("episode_reward_mean", np.nan)
        policy_reward_mean = list(result.get("policy_reward_mean", {}).values())
        if len(policy_reward_mean) == 0:
            if "episode_reward_mean" in cm:
                result["episode_reward_mean"] = cm["episode_reward_mean"]
        else:
            result["episode_reward_mean"] = float(np.mean(policy_reward_mean))
