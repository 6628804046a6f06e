"""Communication mask of the actor-side inference glue (model.py:196-208).

CPU: the C oracle against a torch restatement of the reference lines (and, in the dev container, against the
live `model.Network.step`).  torch.topk leaves the order of equal distances unspecified, so rows whose
k-th and (k+1)-th nearest agents are equidistant are compared as sets of distances, all other rows exactly.
GPU: the CUDA kernel against the oracle, bit for bit (both give ties to the lower agent id)."""
import numpy as np
import pytest
import torch

from helpers import instances, random_instance
from oracle import oracle, ref_loader


def torch_comm_mask(pos, max_comm_agents=3, obs_radius=4):
    """model.py:196-208 verbatim in meaning; pos float tensor [N,2]."""
    num_agents = pos.shape[0]
    pos_mat = (pos.unsqueeze(1) - pos.unsqueeze(0)).abs()
    dis_mat = (pos_mat[:, :, 0] ** 2 + pos_mat[:, :, 1] ** 2).sqrt()
    in_obs_mask = (pos_mat <= obs_radius).all(2)
    _, ranking = dis_mat.topk(min(max_comm_agents, num_agents), dim=1, largest=False)
    dis_mask = torch.zeros((num_agents, num_agents), dtype=torch.bool)
    dis_mask.scatter_(1, ranking, True)
    return torch.bitwise_and(in_obs_mask, dis_mask).numpy(), dis_mat.numpy()


def assert_equal_modulo_ties(ref_mask, dist, got, k):
    N = dist.shape[0]
    k = min(k, N)
    for i in range(N):
        order = np.sort(dist[i])
        tie = k < N and order[k - 1] == order[k]
        if not tie:
            assert np.array_equal(ref_mask[i], got[i].astype(bool)), i
        else:
            # same number of partners at every distance below the tied one, and never more than the reference allows
            below = dist[i] < order[k - 1]
            assert np.array_equal(ref_mask[i][below], got[i].astype(bool)[below]), i
            assert got[i].sum() <= k


@pytest.mark.parametrize("N", [1, 2, 3, 6, 16, 32, 64])
def test_oracle_vs_torch_restatement(N):
    rng = np.random.default_rng(N)
    for trial in range(20):
        L = int(rng.integers(max(3, int(np.ceil(np.sqrt(N))) + 1), 24))
        _, a, _ = random_instance(rng, L, N, 0.0)
        ref, dist = torch_comm_mask(torch.from_numpy(a.astype(np.float32)))
        got = oracle.comm_mask(a)
        assert_equal_modulo_ties(ref, dist, got, 3)
        assert (np.diag(got) == 1).all()          # an agent always communicates with itself (distance 0)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_oracle_vs_live_network_step():
    model = ref_loader.load_module("model")
    torch.manual_seed(0)
    net = model.Network()
    net.eval()
    maps, agents, goals = instances(16)
    for k in (0, 5):
        o = oracle.OracleEnv()
        o.load(maps[k], agents[k], goals[k])
        obs, pos = o.observe()
        net.reset()
        _, _, _, comm = net.step(torch.from_numpy(obs.astype(np.float32)), torch.from_numpy(pos.astype(np.float32)))
        _, dist = torch_comm_mask(torch.from_numpy(pos.astype(np.float32)))
        assert_equal_modulo_ties(comm, dist, oracle.comm_mask(pos), 3)


@pytest.mark.gpu
@pytest.mark.parametrize("L,N,B", [(40, 32, 200), (40, 64, 64), (12, 100, 16), (6, 30, 32), (10, 1, 4), (80, 128, 8)])
def test_gpu_comm_mask_vs_oracle(L, N, B):
    from mapf_rl_b200 import BatchedEnvironment
    rng = np.random.default_rng(L * N)
    ms, as_, gs = [], [], []
    for _ in range(B):
        m, a, g = random_instance(rng, L, N, 0.1 if N < L * L // 2 else 0.0)
        ms.append(m), as_.append(a), gs.append(g)
    env = BatchedEnvironment(B, N, L, device="cuda:0")
    env.load(np.stack(ms), np.stack(as_), np.stack(gs))
    for kk in (3, 2, 1):
        got = env.comm_mask(kk).cpu().numpy()
        for b in range(B):
            assert np.array_equal(got[b], oracle.comm_mask(as_[b], kk)), (b, kk)
    # follows the agents after a step
    acts = rng.integers(0, 5, size=(B, N)).astype(np.uint8)
    env.step(acts)
    pos = env.agents_pos.cpu().numpy()
    got = env.comm_mask().cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], oracle.comm_mask(pos[b]))
