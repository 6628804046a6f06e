"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mapf_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "mapf_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mapf_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from mapf_rl_b200 import _native
    lib = _native.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mapf_b200.h but not exported"
        assert n in _native.SIGNATURES, f"{n} has no ctypes signature in _native.SIGNATURES"
    assert set(_native.SIGNATURES) == set(names)
    assert lib.mapf_abi_version() == _native.ABI_VERSION == 2


def test_library_is_sm100a_only():
    import shutil
    import subprocess
    from mapf_rl_b200 import build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mapf_rl_b200 import _native
    lib = _native.lib()
    cfg = _native.EnvConfig(4, 2, 8, 4, 0, (C.c_float * 5)(-0.075, 0, -0.075, -0.5, 3))
    h = C.c_void_p()
    rc = lib.mapf_env_create(C.byref(cfg), C.byref(h))
    assert rc == _native.MAPF_ECUDA and not h
    assert b"no CUDA device" in lib.mapf_last_error()
    t = C.c_void_p()
    assert lib.mapf_per_create(1024, 0, C.byref(t)) == _native.MAPF_ECUDA
    from mapf_rl_b200 import BatchedEnvironment, SumTree
    with pytest.raises(RuntimeError):
        BatchedEnvironment(4, 2, 8)
    with pytest.raises(RuntimeError):
        SumTree(1024)


def test_argument_validation_without_gpu():
    from mapf_rl_b200 import _native
    lib = _native.lib()
    h = C.c_void_p()
    bad = _native.EnvConfig(4, 2, 8, 3, 0, (C.c_float * 5)(0, 0, 0, 0, 0))  # obs_radius != 4
    assert lib.mapf_env_create(C.byref(bad), C.byref(h)) == _native.MAPF_EINVAL
    bad = _native.EnvConfig(4, 500, 8, 4, 0, (C.c_float * 5)(0, 0, 0, 0, 0))
    assert lib.mapf_env_create(C.byref(bad), C.byref(h)) == _native.MAPF_EINVAL
    t = C.c_void_p()
    assert lib.mapf_per_create(1000, 0, C.byref(t)) == _native.MAPF_EINVAL  # not a power of two, buffer.py:23
    assert lib.mapf_env_step_observe(None, None, None, None, None, None, None) == _native.MAPF_EINVAL


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under mapf_rl_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "mapf_rl_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|libmapf_oracle|mapf_oracle)")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    assert not pat.search(f.read()), f"{fn} references the oracle"
