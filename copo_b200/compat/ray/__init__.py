"""Stand-in for the `ray` package name (see copo_b200/compat/__init__.py).  There is no actor runtime: one process drives
one GPU (torchrun gives data parallelism), so `init` / `shutdown` only keep the bookkeeping the scripts look at."""
from . import tune  # noqa: F401

__version__ = "2.2.0+copo_b200"
_state = {"initialized": False, "kwargs": {}}


def init(*args, **kwargs):
    _state["initialized"] = True
    _state["kwargs"] = dict(kwargs)
    return dict(_state)


def is_initialized():
    return _state["initialized"]


def shutdown():
    _state["initialized"] = False


def available_resources():
    import os
    out = {"CPU": float(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)}
    try:
        import torch
        if torch.cuda.is_available():
            out["GPU"] = float(torch.cuda.device_count())
    except Exception:
        pass
    return out
