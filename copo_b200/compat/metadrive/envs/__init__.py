from .marl_envs import *  # noqa: F401,F403
