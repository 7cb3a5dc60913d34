"""torch_copo/algo_ippo.py of the reference -> copo_b200 (IPPOConfig :17-75, IPPOPolicy :78-172, IPPOTrainer :175-183)."""
from copo_b200.policy import IPPOPolicy, ippo_config  # noqa: F401
from copo_b200.trainer import IPPOTrainer  # noqa: F401
from metadrive.constants import DEFAULT_AGENT  # noqa: F401


def IPPOConfig(algo_class=None):
    return ippo_config()
