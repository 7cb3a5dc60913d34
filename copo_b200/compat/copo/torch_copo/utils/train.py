"""torch_copo/utils/train.py of the reference: `train(trainer, config, stop, exp_name, ...)` (:27-199), same
arguments.  Seeds are a grid over `num_seeds`, `stop` a scalar number of env steps or a dict, trials go through
`tune.run`; the training progress is pickled at the end like the reference does.  wandb arguments are accepted and
ignored (no network)."""
import copy
import os
import pickle

import numpy as np
from ray import tune
from ray.tune import CLIReporter

from copo.train.utils import initialize_ray


def train(trainer, config, stop, exp_name, num_seeds=1, num_gpus=0, test_mode=False, suffix="", checkpoint_freq=10,
          keep_checkpoints_num=None, start_seed=0, local_mode=False, save_pkl=True, custom_callback=None,
          max_failures=1, wandb_key_file=None, wandb_project=None, wandb_team="copo", wandb_log_config=True,
          init_kws=None, **kwargs):
    initialize_ray(test_mode=test_mode, local_mode=local_mode, num_gpus=num_gpus, **(init_kws or {}))
    used_config = {
        "seed": tune.grid_search([i * 100 + start_seed for i in range(num_seeds)]) if num_seeds is not None else None,
        "log_level": "DEBUG" if test_mode else "INFO",
        "callbacks": custom_callback if custom_callback else False,
    }
    if custom_callback is False:
        used_config.pop("callbacks")
    if config:
        used_config.update(config)
    config = copy.deepcopy(used_config)
    if isinstance(trainer, str):
        trainer_name = trainer
    elif hasattr(trainer, "_name"):
        trainer_name = trainer._name
    else:
        trainer_name = trainer.__name__
    if not isinstance(stop, dict) and stop is not None:
        assert np.isscalar(stop)
        stop = {"timesteps_total": int(stop)}
    if test_mode and not os.environ.get("B2C_COMPAT_MAX_ITERS"):
        os.environ["B2C_COMPAT_MAX_ITERS"] = "1"         # --test: one iteration per trial is the smoke run
    if (keep_checkpoints_num is not None) and (not test_mode) and (keep_checkpoints_num != 0):
        assert isinstance(keep_checkpoints_num, int)
        kwargs["keep_checkpoints_num"] = keep_checkpoints_num
        kwargs["checkpoint_score_attr"] = "episode_reward_mean"
    if "verbose" not in kwargs:
        kwargs["verbose"] = 1 if not test_mode else 2
    progress_reporter = CLIReporter(metric_columns=CLIReporter.DEFAULT_COLUMNS.copy())
    for col in ("success", "crash", "out", "max_step", "length", "cost", "takeover", "rc"):
        progress_reporter.add_metric_column(col)
    kwargs["progress_reporter"] = progress_reporter
    analysis = tune.run(trainer, name=exp_name, checkpoint_freq=checkpoint_freq, checkpoint_at_end=True, stop=stop,
                        config=config, max_failures=max_failures if not test_mode else 0, reuse_actors=False,
                        local_dir=".", **kwargs)
    if save_pkl and int(os.environ.get("RANK", "0")) == 0 and os.environ.get("B2C_COMPAT_DRY_RUN") != "1":
        pkl_path = "{}-{}{}.pkl".format(exp_name, trainer_name, "" if not suffix else "-" + suffix)
        with open(pkl_path, "wb") as f:
            pickle.dump(analysis.fetch_trial_dataframes(), f)
            print("Result is saved at: <{}>".format(pkl_path))
    return analysis
