"""GPU tier: fused brute-force kNN against the oracle -- the step after the distance
path (pairwise_distance + raft::matrix::select_k, cpp/include/raft/matrix/select_k.cuh:73-106)."""
import numpy as np
import pytest
import torch

import oracle
from raft_b200 import LogicError
from raft_b200.distance import fused_l2_nn, pairwise_distance
from raft_b200.neighbors import brute_force

pytestmark = pytest.mark.gpu


def check_knn(x, y, kk, gd, gi, sqrt=False, rel=1e-4):
    """Indices valid and unique per row (the property cpp/tests/matrix/select_k_edgecases.cu checks),
    distances ascending, equal to the true distance of the returned index, and the returned set is a
    valid k-nearest set up to fp32-resolution ties at the k-th position."""
    ri, rv = oracle.knn_l2(x, y, kk, sqrt=sqrt)
    m, n = x.shape[0], y.shape[0]
    assert gi.shape == (m, kk) and gd.shape == (m, kk)
    assert gi.min() >= 0 and gi.max() < n
    assert all(len(set(r.tolist())) == kk for r in gi)
    assert np.all(np.diff(gd, axis=1) >= 0)
    true_d = ((x[:, None, :].astype(np.float64) - y[gi].astype(np.float64)) ** 2).sum(-1)
    if sqrt:
        true_d = np.sqrt(true_d)
    scale = np.maximum(rv[:, -1:], 1e-3)
    assert np.all(np.abs(gd - true_d) <= rel * np.maximum(true_d, 1e-3) + 1e-6), np.abs(gd - true_d).max()
    # nothing returned may be farther than the true k-th neighbour (beyond rounding)
    assert np.all(true_d <= rv[:, -1:] + 2e-5 * scale)
    ok, msg = oracle.match_approx(gd, rv, rel)
    assert ok, msg
    mism = (gi != ri).mean()
    assert mism < 0.01, f"{mism:.4f} of the indices differ from the oracle"


@pytest.mark.parametrize("shape", [(1000, 5000, 32, 10), (513, 33001, 96, 64), (200, 300, 200, 5), (64, 129, 16, 1),
                                   (300, 4000, 128, 33), (7, 130, 3, 64)])
@pytest.mark.parametrize("metric", ["sqeuclidean", "euclidean"])
def test_knn_vs_oracle(shape, metric):
    m, n, k, kk = shape
    rng = np.random.default_rng(m + n + k)
    x = (rng.standard_normal((m, k)) * 2).astype(np.float32)
    y = (rng.standard_normal((n, k)) * 2 + 0.3).astype(np.float32)
    gd, gi = brute_force.knn(torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda(), k=kk, metric=metric)
    check_knn(x, y, kk, gd.cpu().numpy(), gi.cpu().numpy(), sqrt=metric == "euclidean")


def test_knn_k1_matches_fused_l2_nn_and_topk_of_pairwise():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((700, 64)).astype(np.float32)
    y = rng.standard_normal((9000, 64)).astype(np.float32)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    gd, gi = brute_force.knn(yt, xt, k=1)
    ni, nv = fused_l2_nn(xt, yt, sqrt=False)
    assert (gi[:, 0] == ni.long()).all() and torch.allclose(gd[:, 0], nv, rtol=1e-5, atol=1e-5)
    # the composition the reference's callers used: full matrix, then select_k
    d = torch.as_tensor(pairwise_distance(xt, yt, metric="sqeuclidean"), device="cuda")
    td, ti = torch.topk(d, 16, dim=1, largest=False, sorted=True)
    gd, gi = brute_force.knn(yt, xt, k=16)
    assert torch.allclose(gd, td, rtol=1e-4, atol=1e-4)
    assert (gi == ti).float().mean() > 0.999


def test_knn_adversarial_order_and_ties():
    """Database sorted by decreasing distance from every query: each pass finds only better candidates,
    the per-row lists overflow and the pass is repeated in halves -- still exact.  Duplicated rows tie:
    ascending index order among equal distances (raft::argmin_op's law applied to k results)."""
    rng = np.random.default_rng(9)
    k, n = 24, 6000
    direction = rng.standard_normal(k).astype(np.float32)
    direction /= np.linalg.norm(direction)
    y = (np.linspace(50.0, 1.0, n, dtype=np.float32)[:, None] * direction[None, :]).astype(np.float32)
    x = (rng.standard_normal((130, k)) * 0.01).astype(np.float32)
    gd, gi = brute_force.knn(torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda(), k=32)
    check_knn(x, y, 32, gd.cpu().numpy(), gi.cpu().numpy())
    assert (gi.cpu().numpy()[:, 0] == n - 1).all()
    # exact duplicates of the nearest row at several indices
    y2 = (rng.standard_normal((3000, 40)) * 3).astype(np.float32)
    x2 = y2[[17, 400]] + 1e-3
    y2[[900, 2500]] = y2[17]
    y2[[401, 2999]] = y2[400]
    gd, gi = brute_force.knn(torch.from_numpy(y2).cuda(), torch.from_numpy(x2).cuda(), k=3)
    assert gi.cpu().tolist() == [[17, 900, 2500], [400, 401, 2999]]


def test_knn_arguments_and_outputs():
    x = torch.randn(50, 8, device="cuda")
    y = torch.randn(90, 8, device="cuda")
    idx = torch.empty(50, 4, dtype=torch.int64, device="cuda")
    dist = torch.empty(50, 4, dtype=torch.float32, device="cuda")
    d2, i2 = brute_force.knn(y, x, indices=idx, distances=dist, global_id_offset=1000)
    assert d2 is dist and i2 is idx and idx.min() >= 1000 and idx.max() < 1090
    with pytest.raises(ValueError):
        brute_force.knn(y, x)                       # k unknown
    with pytest.raises(ValueError):
        brute_force.knn(y, x, k=3, metric="cityblock")
    with pytest.raises(LogicError):
        brute_force.knn(y, x, k=65)                 # above the supported k
    with pytest.raises(LogicError):
        brute_force.knn(y[:3], x, k=4)              # k > n
    from pylibraft.neighbors import brute_force as bf2
    assert bf2.knn is brute_force.knn


@pytest.mark.parametrize("metric", ["cosine", "correlation"])
@pytest.mark.parametrize("shape", [(400, 6000, 48, 12), (129, 20000, 96, 40)])
def test_knn_cosine_family(metric, shape):
    m, n, k, kk = shape
    rng = np.random.default_rng(m + kk)
    x = (rng.standard_normal((m, k)) + 0.4).astype(np.float32)
    y = (rng.standard_normal((n, k)) + 0.4).astype(np.float32)
    mt = oracle.DistanceType.CosineExpanded if metric == "cosine" else oracle.DistanceType.CorrelationExpanded
    ri, rv = oracle.knn(x, y, kk, mt)
    gd, gi = brute_force.knn(torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda(), k=kk, metric=metric)
    gd, gi = gd.cpu().numpy(), gi.cpu().numpy()
    assert np.all(np.diff(gd, axis=1) >= 0) and all(len(set(r.tolist())) == kk for r in gi)
    assert np.all(np.abs(gd - rv) <= 1e-4 * np.abs(rv) + 4e-7)
    assert (gi != ri).mean() < 0.01
    true_d = np.take_along_axis(oracle.pairwise_distance(x, y, mt), gi, axis=1)
    assert np.all(np.abs(gd - true_d) <= 1e-4 * np.abs(true_d) + 4e-7)
