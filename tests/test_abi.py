"""The C-ABI library loads without a GPU and exports every entry point include/copo_b200.h declares; argument
errors come back as codes + messages, never as exceptions across the ABI.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "copo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from copo_b200 import build
    build.build()
    from copo_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_torch_in_the_boundary():
    src = open(os.path.join(ROOT, "include", "copo_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "#include <cuda" not in code


def test_errors_are_codes_not_exceptions(lib):
    lib.b2c_last_error.restype = ctypes.c_char_p
    assert lib.b2c_version() >= 100
    rc = lib.b2c_env_create(None, None, 0, None)
    assert rc == -1 and b"null" in lib.b2c_last_error()
    rc = lib.b2c_env_set_lcf_dist(None, ctypes.c_float(0.0), ctypes.c_float(0.1))
    assert rc == -1


def test_product_refuses_to_run_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from copo_b200 import _lib
    from copo_b200.batched_env import BatchedDrivingEnv
    with pytest.raises(_lib.B2CError):
        BatchedDrivingEnv("intersection", num_scenes=1)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "copo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "_hostsim.so" not in txt, f


def test_ctypes_structs_match_the_header(tmp_path):
    """Sizes and field offsets of every struct crossing the ABI: the C header compiled with gcc vs the ctypes mirrors."""
    import subprocess
    from copo_b200 import _lib, ops
    pairs = {"b2c_env_config": _lib.EnvConfig, "b2c_env_io": _lib.EnvIO, "b2c_ppo_head_args": ops.PpoHeadArgs,
             "b2c_gae_args": ops.GaeArgs, "b2c_tc_head": ops.TcHead, "b2c_lcf_meta_finish_args": ops.LcfMetaFinishArgs}
    fields = {"b2c_env_config": ["num_scenes", "seed", "neighbours_distance", "force_lcf"],
              "b2c_env_io": ["obs", "scene_done", "obs_split"],
              "b2c_ppo_head_args": ["logits", "v_cur", "dlogits", "dv", "stats", "rows", "mode", "clip_param", "kl_coeff"],
              "b2c_gae_args": ["flags", "rewards", "targets", "T", "global_reward_per_scene", "gamma", "lambda_",
                               "bootstrap"],
              "b2c_tc_head": ["weight", "out", "n", "actions", "logp", "seed", "step"],
              "b2c_lcf_meta_finish_args": ["grad_value", "sums", "rows", "raw_std", "lcf_parameters", "lcf_grad", "stats", "lr",
                                           "eps", "step"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "copo_b200.h"', 'int main(void) {']
    for s, fs in fields.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (s, s))
        for f in fs:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (s, f, s, f))
    lines += ["return 0; }"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    got = dict(line.rsplit(" ", 1) for line in out.strip().splitlines())
    for s, cls in pairs.items():
        assert int(got["%s size" % s]) == ctypes.sizeof(cls), s
        for f in fields[s]:
            assert int(got["%s.%s" % (s, f)]) == getattr(cls, f).offset, (s, f)
