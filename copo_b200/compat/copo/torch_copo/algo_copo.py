"""torch_copo/algo_copo.py of the reference -> copo_b200 (column names :48-60, CoPOConfig :63-92, CoPOModel :96-182,
CoPOPolicy :207-502, CoPOTrainer :506-661)."""
from copo_b200.models import CoPOModel  # noqa: F401
from copo_b200.policy import CoPOPolicy, copo_config  # noqa: F401
from copo_b200.trainer import CoPOTrainer  # noqa: F401

NEI_REWARDS = "nei_rewards"
NEI_VALUES = "nei_values"
NEI_ADVANTAGE = "nei_advantage"
NEI_TARGET = "nei_target"
LCF_LR = "lcf_lr"
GLOBAL_VALUES = "global_values"
GLOBAL_REWARDS = "global_rewards"
GLOBAL_ADVANTAGES = "global_advantages"
GLOBAL_TARGET = "global_target"
USE_CENTRALIZED_CRITIC = "use_centralized_critic"
CENTRALIZED_CRITIC_OBS = "centralized_critic_obs"
COUNTERFACTUAL = "counterfactual"
USE_DISTRIBUTIONAL_LCF = "use_distributional_lcf"


def CoPOConfig(algo_class=None):
    return copo_config()
