"""the instance generators against the reference's own Environment.__init__ / reset().

The reference's RNG stream is not reproduced (parity of step/observe goes through load()); its DISTRIBUTION is.
tests/golden/generator_stats.npz holds histograms over 3000 instances per geometry drawn by the live reference
(make_golden.py generator); the host generator (mapf_rl_b200.instances, used by the drop-in Environment) and the
device generator (mapf_env_reset) must produce the same histograms up to sampling noise: two-sample chi-square,
rejected below p = 1e-4 (a wrong component weighting, a biased start or goal draw or a wrong density law shift these
histograms by many standard deviations; see test_statistic_detects_a_biased_generator)."""
import numpy as np
import pytest

from helpers import GEN_CONFIGS, generator_stats, golden, histograms_agree

P_MIN = 1e-4
SAMPLES = 3000


def reference_hist(name, L, N):
    return golden("generator_stats.npz")[f"{name}_{L}_{N}"]


def check_against_reference(st, L, N):
    for name, h in st.items():
        ref = reference_hist(name, L, N)
        assert h.sum() > 0 and ref.sum() > 0
        p = histograms_agree(h, ref)
        assert p > P_MIN, f"{name} histogram differs from the reference generator at {L}x{L}/{N} agents (p = {p:.2e})"


@pytest.mark.parametrize("L,N", GEN_CONFIGS)
def test_host_generator_matches_reference_distribution(L, N):
    from mapf_rl_b200.instances import generate_instance
    rng = np.random.default_rng(31 + L)
    inst = [generate_instance(rng, L, N, density=None) for _ in range(SAMPLES)]
    st = generator_stats(np.stack([i[0] for i in inst]), np.stack([i[1] for i in inst]), np.stack([i[2] for i in inst]))
    check_against_reference(st, L, N)


def test_statistic_detects_a_biased_generator():
    """Power check of the statistic itself: a uniform(0, 0.5) density law instead of triangular(0, 0.33, 0.5), and a
    goal draw biased towards the start (the far tail of the distance histogram folded into the short distances),
    are both rejected at the same sample size."""
    from mapf_rl_b200.instances import generate_instance
    L, N = GEN_CONFIGS[0]
    rng = np.random.default_rng(5)
    inst = [generate_instance(rng, L, N, density=rng.uniform(0.0, 0.5)) for _ in range(SAMPLES)]
    st = generator_stats(np.stack([i[0] for i in inst]), np.stack([i[1] for i in inst]), np.stack([i[2] for i in inst]))
    assert histograms_agree(st["density"], reference_hist("density", L, N)) < P_MIN
    near = reference_hist("distance", L, N).copy()
    near[1:6] += near[20:].sum() // 5
    near[20:] = 0
    assert histograms_agree(near, reference_hist("distance", L, N)) < P_MIN


@pytest.mark.gpu
@pytest.mark.parametrize("L,N", GEN_CONFIGS)
def test_device_generator_matches_reference_distribution(L, N):
    from mapf_rl_b200 import BatchedEnvironment
    env = BatchedEnvironment(SAMPLES, N, L)
    env.reset(seed=77 + L, density=None)
    env.check()
    st = generator_stats(env.map.cpu().numpy(), env.agents_pos.cpu().numpy(), env.goals_pos.cpu().numpy())
    check_against_reference(st, L, N)
