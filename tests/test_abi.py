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
