"""copo/train/utils.py of the reference: `initialize_ray`, `get_train_parser`, `setup_logger`."""
import argparse
import logging
import os

import ray


def initialize_ray(local_mode=False, num_gpus=None, test_mode=False, **kwargs):
    os.environ["OMP_NUM_THREADS"] = "1"
    kwargs.pop("redis_password", None)
    ray.init(logging_level=logging.ERROR if not test_mode else logging.DEBUG, log_to_driver=test_mode,
             local_mode=local_mode, num_gpus=num_gpus, ignore_reinit_error=True, **kwargs)
    print("Successfully initialize Ray! (copo_b200 shim: no actor runtime, trials run in this process)")
    print("Available resources: ", ray.available_resources())


def get_train_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--exp-name", type=str, default="")
    parser.add_argument("--num-gpus", type=int, default=0)
    parser.add_argument("--num-seeds", type=int, default=3)
    parser.add_argument("--num-cpus-per-worker", type=float, default=0.5)
    parser.add_argument("--num-gpus-per-trial", type=float, default=0.25)
    parser.add_argument("--test", action="store_true")
    return parser


def setup_logger(debug=False):
    logging.basicConfig(level=logging.DEBUG if debug else logging.WARNING,
                        format="%(asctime)s - %(filename)s[line:%(lineno)d] - %(levelname)s: %(message)s")
