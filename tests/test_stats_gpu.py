"""GPU tier: silhouette_score on the distance engine against the oracle -- the caller
whose pairwise_distance call is dangling in the reference (stats/detail/silhouette_score.cuh:205-206)."""
import numpy as np
import pytest
import torch

import oracle
from raft_b200 import LogicError
from raft_b200.stats import silhouette_score, trustworthiness_score

pytestmark = pytest.mark.gpu


def data(n, k, n_labels, seed):
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, n_labels, n).astype(np.int32)
    x = (rng.standard_normal((n, k)) + labels[:, None] * 1.5).astype(np.float32)
    return x, labels


@pytest.mark.parametrize("metric,dt", [("sqeuclidean_unexpanded", oracle.DistanceType.L2Unexpanded),
                                       ("euclidean", oracle.DistanceType.L2SqrtExpanded),
                                       ("cityblock", oracle.DistanceType.L1), ("cosine", oracle.DistanceType.CosineExpanded)])
@pytest.mark.parametrize("shape", [(1500, 24, 6), (777, 130, 2), (300, 5, 37)])
def test_silhouette_vs_oracle(metric, dt, shape):
    n, k, nl = shape
    x, labels = data(n, k, nl, seed=n + k)
    labels[: nl] = np.arange(nl)                       # every label present
    ref, ref_s = oracle.silhouette_score(x, labels, nl, metric=dt, return_samples=True)
    got, per = silhouette_score(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), nl, metric=metric,
                                return_samples=True)
    assert abs(got - ref) < 2e-4, (got, ref)
    assert np.allclose(per.cpu().numpy(), ref_s, atol=5e-4)


def test_silhouette_chunked_singletons_and_empty_labels():
    """Chunked slabs (the reference's batched pattern, detail/batched/silhouette_score.cuh:213-243) give
    the same score; singleton clusters score 0; an unused label id is ignored."""
    x, labels = data(1000, 16, 4, seed=3)
    labels[labels == 2] = 1                            # label 2 unused
    labels[5] = 4                                      # singleton
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    ref, ref_s = oracle.silhouette_score(x, labels, 5, return_samples=True)
    whole, per = silhouette_score(xt, lt, 5, return_samples=True)
    assert abs(whole - ref) < 2e-4 and per[5].item() == 0.0
    for chunk in (128, 256, 896):
        got = silhouette_score(xt, lt, 5, chunk=chunk)
        assert abs(got - whole) < 1e-6
    assert abs(silhouette_score(xt, lt) - whole) < 1e-6          # n_labels inferred


def test_silhouette_argument_errors():
    x = torch.randn(40, 4, device="cuda")
    lab = torch.zeros(40, dtype=torch.int32, device="cuda")
    with pytest.raises(LogicError):
        silhouette_score(x, lab, 1)                     # fewer than 2 labels
    lab[3] = 7
    with pytest.raises(LogicError):
        silhouette_score(x, lab, 3)                     # label outside [0, n_labels)
    with pytest.raises(ValueError):
        silhouette_score(x, lab[:10], 8)
    from pylibraft.stats import silhouette_score as s2
    assert s2 is silhouette_score


@pytest.mark.parametrize("shape", [(1200, 20, 2, 5), (700, 64, 8, 12), (300, 7, 3, 63)])
@pytest.mark.parametrize("quality", ["projection", "random"])
def test_trustworthiness_vs_oracle(shape, quality):
    n, m, d, k = shape
    rng = np.random.default_rng(n + k)
    x = rng.standard_normal((n, m)).astype(np.float32)
    if quality == "projection":
        e = (x @ rng.standard_normal((m, d)).astype(np.float32)).astype(np.float32)
    else:
        e = rng.standard_normal((n, d)).astype(np.float32)
    ref = oracle.trustworthiness_score(x, e, k)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(e).cuda()
    got = trustworthiness_score(xt, et, n_neighbors=k)
    assert abs(got - ref) < 1e-4, (got, ref)
    assert abs(trustworthiness_score(xt, et, n_neighbors=k, metric="sqeuclidean_unexpanded", batch_size=256) - ref) < 1e-4
    assert trustworthiness_score(xt, xt, n_neighbors=k) == 1.0


def test_trustworthiness_uses_the_callers_metric_in_both_spaces():
    """Cosine neighbourhoods differ from L2 ones when the rows have different lengths: the embedded-space kNN must use
    the metric the caller names (round 1 hard-wired L2 there)."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((600, 24)).astype(np.float32) * rng.uniform(0.1, 10, (600, 1)).astype(np.float32)
    e = (x @ rng.standard_normal((24, 4)).astype(np.float32)).astype(np.float32)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(e).cuda()
    ref_cos = oracle.trustworthiness_score(x, e, 7, metric=oracle.DistanceType.CosineExpanded)
    ref_l2 = oracle.trustworthiness_score(x, e, 7)
    assert abs(ref_cos - ref_l2) > 1e-3                      # the two definitions are told apart by this data
    assert abs(trustworthiness_score(xt, et, n_neighbors=7, metric="cosine") - ref_cos) < 1e-4
    assert abs(trustworthiness_score(xt, et, n_neighbors=7) - ref_l2) < 1e-4


def test_trustworthiness_argument_errors():
    x = torch.randn(50, 6, device="cuda")
    with pytest.raises(LogicError):
        trustworthiness_score(x, x[:, :2].contiguous(), n_neighbors=64)
    with pytest.raises(LogicError):
        trustworthiness_score(x, x[:, :2].contiguous(), n_neighbors=40)      # >= n / 2
    with pytest.raises(ValueError):
        trustworthiness_score(x, x[:40, :2].contiguous())
