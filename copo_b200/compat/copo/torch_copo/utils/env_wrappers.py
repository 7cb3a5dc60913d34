"""torch_copo/utils/env_wrappers.py of the reference -> copo_b200.envs (same factory names and return conventions)."""
from copo_b200.envs import (COMM_CURRENT_OBS, COMM_METHOD, NEI_OBS, get_ccenv, get_change_n_env,  # noqa: F401
                            get_lcf_env)
from copo_b200 import envs as _envs


def get_rllib_compatible_env(env_class, return_class=False):
    """env_wrappers.py:559-597: registers the class under its name; returns the name, or (name, class)."""
    name = _envs.get_rllib_compatible_env(env_class)
    return (name, env_class) if return_class else name
