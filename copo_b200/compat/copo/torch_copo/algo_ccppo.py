"""torch_copo/algo_ccppo.py of the reference -> copo_b200 (CCPPOConfig :37-52, get_centralized_critic_obs_dim :55-71,
CCModel :74-219, get_ccppo_env :314-315, CCPPOPolicy :318-472, CCPPOTrainer :475-482)."""
from copo_b200.models import CCModel  # noqa: F401
from copo_b200.models import centralized_critic_obs_dim as _cc_dim
from copo_b200.policy import CCPPOPolicy, ccppo_config  # noqa: F401
from copo_b200.trainer import CCPPOTrainer  # noqa: F401
from copo.torch_copo.utils.env_wrappers import get_ccenv, get_rllib_compatible_env

CENTRALIZED_CRITIC_OBS = "centralized_critic_obs"
COUNTERFACTUAL = "counterfactual"


def CCPPOConfig(algo_class=None):
    return ccppo_config()


def get_centralized_critic_obs_dim(observation_space_shape, action_space_shape, counterfactual, num_neighbours,
                                   fuse_mode):
    odim = observation_space_shape[0] if hasattr(observation_space_shape, "__len__") else observation_space_shape
    adim = action_space_shape[0] if hasattr(action_space_shape, "__len__") else action_space_shape
    return _cc_dim(odim, adim, counterfactual, num_neighbours, fuse_mode)


def get_ccppo_env(env_class):
    return get_rllib_compatible_env(get_ccenv(env_class))
