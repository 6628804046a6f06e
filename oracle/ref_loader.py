"""TEST INFRASTRUCTURE — loads the *real* reference at run time (dev container only).

The reference (`/root/reference`, read-only, never copied) does not import on
numpy >= 1.24 / without matplotlib.  This loader applies the three shims
described in SURVEY.md Appendix B and returns live reference modules, so that
`tests/golden/make_golden.py` can generate golden vectors from the reference
itself and so CPU tests running in the dev container can pin `oracle/` against
it.  `/root/reference` does not exist on the GPU box: nothing under `-m gpu`,
`smoke()` or `bench.py` may call this module (they use the committed fixtures).

Shims (each is the exact old-numpy meaning, nothing else is altered):
  1. stub `matplotlib`, `matplotlib.pyplot` (`ion`), `matplotlib.colors`,
     `matplotlib.animation`                         (environment.py:3-5)
  2. `numpy.int = int`, `numpy.bool = bool`         (environment.py:12,26,102 / buffer.py:64)
  3. `if target_agent_id:` -> empty -> False, one element -> its truthiness
                                                     (environment.py:343)
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_DIR = os.environ.get("MAPF_REFERENCE_DIR", "/root/reference")

_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "environment.py"))


def _install_shims() -> None:
    import numpy as np

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        plt.ion = lambda: None
        colors = types.ModuleType("matplotlib.colors")
        anim = types.ModuleType("matplotlib.animation")
        mpl.pyplot, mpl.colors, mpl.animation = plt, colors, anim
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.colors"] = colors
        sys.modules["matplotlib.animation"] = anim
    for name, typ in (("int", int), ("bool", bool), ("float", float)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)


def load_environment():
    """Return the reference `environment` module (shimmed)."""
    if "environment" in _cache:
        return _cache["environment"]
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_DIR}")
    _install_shims()
    path = os.path.join(REFERENCE_DIR, "environment.py")
    with open(path, "r") as f:
        src = f.read()
    needle = "if target_agent_id:"
    assert src.count(needle) == 1, "reference environment.py changed; shim 3 no longer applies"
    src = src.replace(needle, "if target_agent_id.size > 0 and bool(target_agent_id[0]):")
    mod = types.ModuleType("ref_environment")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    _cache["environment"] = mod
    return mod


def load_module(name: str):
    """Import `buffer`, `search` or `config` from the reference unmodified."""
    key = "mod:" + name
    if key in _cache:
        return _cache[key]
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_DIR}")
    _install_shims()
    saved = sys.modules.pop(name, None)
    try:
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REFERENCE_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules[name] = saved
    _cache[key] = mod
    return mod


def load_worker():
    """Import the reference's worker.py (GlobalBuffer / Learner / Actor) with a no-op `ray` stub: `ray.remote`
    returns the class unchanged, `ray.put` / `ray.get` are identities.  Only GlobalBuffer's storage / sampling
    methods are exercised (they do not touch Ray beyond `ray.put` in __init__)."""
    if "worker" in _cache:
        return _cache["worker"]
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_DIR}")
    _install_shims()
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        def remote(*a, **k):
            if a and callable(a[0]) and not k:
                return a[0]
            return lambda cls: cls
        ray.remote, ray.put, ray.get = remote, (lambda x: x), (lambda x: x)
        sys.modules["ray"] = ray
    saved = {k: sys.modules.get(k) for k in ("environment", "model", "buffer", "config")}
    sys.modules["environment"] = load_environment()
    try:
        spec = importlib.util.spec_from_file_location("ref_worker", os.path.join(REFERENCE_DIR, "worker.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["worker"] = mod
    return mod


def load_pkl(num_agents: int):
    """Load test{N}_40_0.3.pkl -> (maps, agents, goals) lists."""
    import pickle

    with open(os.path.join(REFERENCE_DIR, f"test{num_agents}_40_0.3.pkl"), "rb") as f:
        d = pickle.load(f)
    return d["maps"], d["agents"], d["goals"]
