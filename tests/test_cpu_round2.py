"""CPU: round-2 additions to the test infrastructure and the bench's CPU arm.

* the oracle's n-step actor TD against numpy's own `np.convolve` (the arithmetic LocalBuffer.finish runs, buffer.py:170-177)
  for config.forward_steps = 1, 2, 3, 5;
* the batched navi helper against the per-environment one;
* `bench.py --impl reference` (the C port on the host cores) prints the contract's JSON line with the same config keys as
  the GPU arm;
* the Python-reference timing script returns a sane number where the reference tree is mounted."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle, ref_loader

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("n,gamma", [(1, 0.99), (2, 0.99), (3, 0.99), (5, 0.9)])
def test_oracle_actor_td_n_equals_numpy_convolve(n, gamma):
    rng = np.random.default_rng(n)
    for size in (1, 2, 7, 64):
        rew = rng.choice([-0.075, -0.5, 0.0, 3.0], size=size).astype(np.float16)
        q = rng.normal(size=(size, 5)).astype(np.float32)
        act = rng.integers(0, 5, size=size).astype(np.uint8)
        ret = rew.tolist() + [0 for _ in range(n - 1)]                                           # buffer.py:174
        reward = np.convolve(ret, [gamma ** (n - 1 - i) for i in range(n)], 'valid') + np.max(q, axis=1)   # :175
        exp = np.zeros(64)
        exp[:size] = np.abs(reward - q[np.arange(size), act])                                    # :176-177
        got = oracle.actor_td_n(rew.astype(np.float64), q, act, 64, n, gamma)
        assert np.array_equal(got, exp), (n, size)
        if n == 2 and gamma == 0.99:
            assert np.array_equal(got, oracle.actor_td(rew, q, act, 64))


def test_navi_batch_equals_per_env():
    rng = np.random.default_rng(0)
    B, L, N = 9, 17, 5
    maps = (rng.random((B, L, L)) < 0.3).astype(np.uint8)
    goals = np.zeros((B, N, 2), np.int32)
    for b in range(B):
        free = np.argwhere(maps[b] == 0)
        goals[b] = free[rng.choice(len(free), N, replace=False)]
    nb = oracle.navi_batch(maps, goals)
    for b in range(B):
        assert np.array_equal(oracle.navi(maps[b], goals[b])[1], nb[b])


def test_bench_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--num-envs", "48", "--steps", "5",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "env_agent_steps_per_sec_step_observe"
    assert line["steps"] == 5 and line["value"] > 0 and line["higher_is_better"] is True
    assert line["config"]["num_envs"] == 48 and line["config"]["num_agents"] == 32 and line["config"]["map_length"] == 40
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert set(line["config"]) >= {"workload", "num_envs", "num_agents", "map_length", "obstacle_density", "actions", "config", "l2"}
    # `config` is computed from the arguments alone: the GPU arm prints the same object (what differs between the arms is `method`)
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    a = argparse.Namespace(config="c2", num_envs=48, num_agents=32, map_length=40, density=0.3, max_steps=256, scaling="weak", gpus=1)
    assert bench.bench_config(a) == line["config"]
    assert "timing" in line["method"]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_python_reference_timing_script():
    from oracle import time_python_reference
    r = time_python_reference.measure(num_agents=16, seconds=0.5, procs=2, steps_per_instance=8)
    assert r["kind"] == "reference" and r["cores"] == 2
    assert 2e3 < r["per_core_mean"] < 1e6     # ~20-30 k agent-steps/s per core in the dev container


def test_committed_python_reference_fixture():
    with open(os.path.join(ROOT, "profiles", "r2_python_reference_cpu.json")) as f:
        r = json.load(f)
    assert r["kind"] == "reference" and r["unit"] == "agent-steps/s" and r["cores"] >= 1 and r["value"] > 1e4
