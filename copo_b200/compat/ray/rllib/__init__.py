"""Placeholder so `from ray import rllib` resolves; the RLlib pieces of the hot path live in copo_b200."""
