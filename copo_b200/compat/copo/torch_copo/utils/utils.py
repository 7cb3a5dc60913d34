"""torch_copo/utils/utils.py of the reference: what the training scripts take from it (`get_train_parser`,
`setup_logger`; :206-223)."""
from copo.train.utils import get_train_parser, initialize_ray, setup_logger  # noqa: F401
