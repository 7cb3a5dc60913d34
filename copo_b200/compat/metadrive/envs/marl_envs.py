"""`from metadrive.envs.marl_envs import MultiAgent*Env` (torch_copo/train_copo.py:1-2) -> copo_b200.envs."""
from copo_b200.envs import (MultiAgentBottleneckEnv, MultiAgentIntersectionEnv, MultiAgentMetaDrive,  # noqa: F401
                            MultiAgentParkingLotEnv, MultiAgentRoundaboutEnv, MultiAgentTollgateEnv)

__all__ = ["MultiAgentParkingLotEnv", "MultiAgentRoundaboutEnv", "MultiAgentBottleneckEnv", "MultiAgentMetaDrive",
           "MultiAgentTollgateEnv", "MultiAgentIntersectionEnv"]
