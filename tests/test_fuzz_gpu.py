"""GPU tier: randomised shapes / strides / metrics against the oracle (tile edges, ragged k, single
rows and columns, several work items per CTA)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import DistanceType as DT
from raft_b200.distance import fused_l2_nn, pairwise_distance

pytestmark = pytest.mark.gpu


def test_random_shapes_all_paths():
    rng = np.random.default_rng(2024)
    metrics = [DT.L2Expanded, DT.L2SqrtExpanded, DT.CosineExpanded, DT.CorrelationExpanded, DT.L1, DT.Linf,
               DT.L2Unexpanded, DT.Canberra]
    for it in range(48):
        m = int(rng.choice([1, 2, 31, 127, 128, 129, 255, 257, 400, 640]))
        n = int(rng.choice([1, 3, 64, 255, 256, 257, 511, 513, 700]))
        k = int(rng.choice([1, 2, 7, 31, 32, 33, 64, 96, 97, 128, 129, 200, 330]))
        metric = metrics[it % len(metrics)]
        if metric == DT.CorrelationExpanded and k < 3:
            k = 5
        x, _, c = oracle.make_blobs(m, k, seed=it)
        y, _, _ = oracle.make_blobs(n, k, seed=1000 + it, centers=c)
        # every third case: padded leading dimensions (ld != k), odd pitch for the output
        if it % 3 == 0:
            xp = torch.zeros(m, k + 3, device="cuda"); xp[:, :k] = torch.from_numpy(x).cuda()
            yp = torch.zeros(n, k + 5, device="cuda"); yp[:, :k] = torch.from_numpy(y).cuda()
            outp = torch.zeros(m, n + 1, device="cuda")
            pairwise_distance(xp[:, :k], yp[:, :k], out=outp[:, :n], metric=metric)
            got = outp[:, :n].cpu().numpy()
        else:
            got = pairwise_distance(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), metric=metric).copy_to_host()
        ref = oracle.pairwise_distance(x, y, metric)
        ok = oracle.compare_approx(got, ref, 1e-4)
        if metric in (DT.L2Expanded, DT.L2SqrtExpanded):
            # expanded form in fp32: xn + yn - 2xy cannot resolve d^2 below a few ulp of (xn + yn); pairs
            # that close (random k = 1 collisions) are held to that floor instead of 1e-4 relative --
            # the same limit the reference's expanded fp32 kernels have
            sq = metric == DT.L2SqrtExpanded
            g2 = got.astype(np.float64) ** 2 if sq else got.astype(np.float64)
            r2 = ref ** 2 if sq else ref
            floor = 8 * 2.0 ** -24 * (oracle.row_norm_sq(x)[:, None] + oracle.row_norm_sq(y)[None, :])
            ok |= np.abs(g2 - r2) <= floor
        assert ok.all(), (it, m, n, k, metric, int((~ok).sum()), np.argwhere(~ok)[0])


def test_random_shapes_fused_nn():
    rng = np.random.default_rng(7)
    for it in range(16):
        m = int(rng.choice([1, 33, 128, 129, 300, 1000]))
        n = int(rng.choice([1, 2, 255, 256, 257, 1000, 3000]))
        k = int(rng.choice([1, 5, 32, 64, 96, 128, 160]))
        x, _, c = oracle.make_blobs(m, k, seed=50 + it)
        y, _, _ = oracle.make_blobs(n, k, seed=90 + it, centers=c)
        ri, rv = oracle.fused_l2_nn(x, y)
        gi, gv = fused_l2_nn(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), sqrt=False)
        gi, gv = gi.cpu().numpy(), gv.cpu().numpy()
        bad = np.nonzero(gi != ri)[0]
        for i in bad:   # only provable near-ties may differ
            d = ((x[i].astype(np.float64) - y[gi[i]].astype(np.float64)) ** 2).sum()
            assert abs(d - rv[i]) <= 1e-5 * max(rv[i], 1.0), (it, m, n, k, i)
        assert len(bad) <= 1
        ok, msg = oracle.match_approx(gv, rv, 1e-4)
        assert ok, (it, m, n, k, msg)


def test_many_work_items_per_cta_nn():
    # 274 y blocks x 32 x tiles: several items per CTA, keys updated from many CTAs concurrently
    x, _, c = oracle.make_blobs(4000, 96, seed=3)
    y, _, _ = oracle.make_blobs(70000, 96, seed=4, centers=c)
    ri, rv = oracle.fused_l2_nn(x, y)
    gi, gv = fused_l2_nn(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), sqrt=False)
    gi, gv = gi.cpu().numpy(), gv.cpu().numpy()
    bad = np.nonzero(gi != ri)[0]
    assert len(bad) <= 4
    for i in bad:
        d = ((x[i].astype(np.float64) - y[gi[i]].astype(np.float64)) ** 2).sum()
        assert abs(d - rv[i]) <= 1e-5 * max(rv[i], 1.0)
    ok, msg = oracle.match_approx(gv, rv, 1e-4)
    assert ok, msg
