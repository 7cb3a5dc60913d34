"""Import shim: lets the reference's own training scripts drive this package UNCHANGED.

`copo_code/copo/torch_copo/train_copo.py` (and train_ippo.py / train_ccppo.py) import
`metadrive.envs.marl_envs`, `ray.tune`, `copo.torch_copo.{algo_copo,algo_ippo,algo_ccppo}` and
`copo.torch_copo.utils.{env_wrappers,train,utils,callbacks}` (train_copo.py:1-9).  None of MetaDrive, Ray or gym is
installed (or needed) here: the directories beside this file are packages of those names that re-export this repo's
CUDA-backed implementations, so

    PYTHONPATH=/root/repo:/root/repo/copo_b200/compat python /path/to/copo/torch_copo/train_copo.py --exp-name t

runs the reference script as written on the B200 path.  `install()` does the same from Python.  The shim is NOT
imported by `copo_b200` itself and shadows `ray` / `metadrive` / `copo` only for processes that put it on their path.

What maps to what (reference file -> here):
  metadrive/envs/marl_envs.py  MultiAgent*Env            -> copo_b200.envs (dict API over the batched CUDA scene step)
  ray.tune                     grid_search, run, CLIReporter -> compat/ray/tune: sequential trials in this process
  torch_copo/algo_*.py         *Trainer, *Config, *Policy, models, column names -> copo_b200.{trainer,policy,models}
  torch_copo/utils/train.py    train(trainer, config, stop, exp_name, ...)     -> compat/copo/torch_copo/utils/train.py
  torch_copo/utils/callbacks.py MultiAgentDrivingCallbacks.on_train_result     -> compat/.../callbacks.py

Environment variables: B2C_NUM_SCENES (scenes per GPU; default train_batch_size / rollout_fragment_length, i.e. the
reference's number of concurrent envs), B2C_COMPAT_MAX_ITERS (stop every trial after this many iterations),
B2C_COMPAT_DRY_RUN=1 (expand the trials and return without touching the GPU).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def path():
    return HERE


def install():
    """Puts the shim (and the repo root) at the front of sys.path; returns the shim directory."""
    root = os.path.dirname(os.path.dirname(HERE))
    for p in (HERE, root):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    return HERE
