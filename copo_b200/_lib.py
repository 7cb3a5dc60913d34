"""ctypes binding of libcopo_b200.so (the C ABI declared in include/copo_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is current when a compute
entry point is called, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcopo_b200.so")
_lib = None

c_void_p, c_int, c_float, c_u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint32


class B2CError(RuntimeError):
    pass


class EnvConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("num_scenes", "num_slots", "num_agents", "delay_done", "horizon",
                                              "agent_horizon", "allow_respawn", "auto_reset", "append_lcf",
                                              "lcf_uniform", "scene_offset")] + \
               [("seed", ctypes.c_uint32)] + \
               [(n, ctypes.c_float) for n in ("neighbours_distance", "mf_nei_distance", "lcf_mean", "lcf_std",
                                              "force_lcf")]


ENV_IO_FIELDS = ("obs", "reward", "flags", "nei_mask", "mf_mask", "nei_reward", "global_reward", "nei_list",
                 "agent_id", "lcf", "scene_done", "obs_split")


class EnvIO(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ENV_IO_FIELDS]


def load():
    """Loads the library (building is __graft_entry__.build()'s job); raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2CError("libcopo_b200.so is not built: run `python -m copo_b200.build` "
                           "(there is no CPU fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b2c_last_error.restype = ctypes.c_char_p
    return _lib


LAUNCHES = 0      # C-ABI compute calls issued by this process (each enqueues at least one of the library's kernels)


def check(rc):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        raise B2CError("libcopo_b200 error %d: %s" % (rc, load().b2c_last_error().decode()))


def require_device():
    lib = load()
    if not lib.b2c_device_ok():
        raise B2CError("no sm_100 (B200) CUDA device is current; the hot path has no CPU fallback")
    return lib


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous()
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
