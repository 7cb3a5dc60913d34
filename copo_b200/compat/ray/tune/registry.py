"""`ray.tune.registry.register_env` (torch_copo/utils/env_wrappers.py:590): the creator is kept next to the classes
copo_b200.envs.get_rllib_compatible_env registers, so trainers find either by name."""
from copo_b200 import envs as _envs

_CREATORS = {}


def register_env(name, env_creator):
    _CREATORS[name] = env_creator


def make(name, config=None):
    if name in _CREATORS:
        return _CREATORS[name](config or {})
    return _envs.make_env(name, config)
