"""The two helpers the reference's wrappers take from metadrive.utils (torch_copo/utils/env_wrappers.py:7)."""
import numpy as np


def get_np_random(seed=None, return_seed=False):
    rng = np.random.RandomState(seed)
    return (rng, seed) if return_seed else rng


def clip(a, low, high):
    return min(max(a, low), high)
