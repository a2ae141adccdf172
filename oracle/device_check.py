"""TEST INFRASTRUCTURE -- fp64 re-evaluation ON THE DEVICE of sampled results at the full BASELINE.json sizes.

The CPU oracle (oracle.py) finishes in seconds only for small shapes; at 100000 x 100000 (10^10 pairs) or
1M x 8M the same definitions are re-evaluated here in fp64 with plain torch ops on the GPU, for a random
sample of pairs (plus the corners of the kernels' 128 x 256 tiles and the matrix edges), or -- for fusedL2NN --
for a sample of query rows against the WHOLE database.  Same semantics as oracle.pairwise_distance /
oracle.fused_l2_nn (metric definitions SURVEY.md 8(a3),(a4); tolerance rule raft::CompareApprox,
cpp/tests/test_utils.h:31-45; tie law raft::argmin_op, cpp/include/raft/core/operators.hpp:187-194).
Only tests/ and bench.py's `parity` leg import this; the product never does.
"""
from __future__ import annotations

import torch

from .oracle import DistanceType as DT


def sample_pairs(m: int, n: int, count: int, seed: int, device, tile_m: int = 128, tile_n: int = 256):
    """`count` random (i, j) plus every combination of a few tile-corner / edge rows and columns."""
    g = torch.Generator(device=device).manual_seed(seed)
    i = torch.randint(0, m, (count,), device=device, generator=g)
    j = torch.randint(0, n, (count,), device=device, generator=g)

    def edges(size, tile):
        e = {0, size - 1}
        for t in (1, max(1, size // tile // 2), max(1, size // tile)):
            for d in (-1, 0, 1):
                v = t * tile + d
                if 0 <= v < size:
                    e.add(v)
        return torch.tensor(sorted(e), device=device)

    ei, ej = edges(m, tile_m), edges(n, tile_n)
    gi, gj = torch.meshgrid(ei, ej, indexing="ij")
    return torch.cat([i, gi.reshape(-1)]), torch.cat([j, gj.reshape(-1)])


def pair_metric_fp64(x, y, i, j, metric, p: float = 2.0):
    """metric(x[i], y[j]) in fp64 for index vectors i, j (definitions as oracle._block)."""
    metric = DT(int(metric))
    a = x[i].to(torch.float64)
    b = y[j].to(torch.float64)
    if metric in (DT.L2Expanded, DT.L2Unexpanded):
        return ((a - b) ** 2).sum(1)
    if metric in (DT.L2SqrtExpanded, DT.L2SqrtUnexpanded):
        return ((a - b) ** 2).sum(1).sqrt()
    if metric == DT.InnerProduct:
        return (a * b).sum(1)
    if metric in (DT.CosineExpanded, DT.CorrelationExpanded):
        if metric == DT.CorrelationExpanded:
            a = a - a.mean(1, keepdim=True)
            b = b - b.mean(1, keepdim=True)
        return 1.0 - (a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1))
    if metric == DT.L1:
        return (a - b).abs().sum(1)
    if metric == DT.Linf:
        return (a - b).abs().amax(1)
    if metric == DT.Canberra:
        den = a.abs() + b.abs()
        return torch.where(den > 0, (a - b).abs() / den.clamp_min(1e-300), torch.zeros_like(den)).sum(1)
    if metric == DT.LpUnexpanded:
        return ((a - b).abs() ** p).sum(1) ** (1.0 / p)
    raise ValueError(f"device_check: metric {metric!r} not covered")


def compare_approx(got, ref, eps: float):
    """raft::CompareApprox (cpp/tests/test_utils.h:31-45): |a-b| <= eps * max(|a|,|b|), or both below eps.
    Returns (n_bad, max relative error over the entries that are not both below eps)."""
    got = got.to(torch.float64)
    diff = (got - ref).abs()
    big = torch.maximum(got.abs(), ref.abs())
    small = big < eps
    rel = torch.where(small, torch.zeros_like(diff), diff / big.clamp_min(1e-300))
    bad = (~small) & ~(rel <= eps)          # NaN in got -> bad
    return int(bad.sum().item()), float(rel.max().item()) if rel.numel() else 0.0


def check_pairwise_sampled(out, x, y, metric, count: int = 200_000, eps: float = 1e-4, p: float = 2.0, seed: int = 7,
                           row_offset: int = 0):
    """out [rows, n] holds metric(x[row_offset + r], y[c]).  Returns {"checked", "n_bad", "max_rel_err"}."""
    rows, n = out.shape
    i, j = sample_pairs(rows, n, count, seed, out.device)
    ref = pair_metric_fp64(x, y, i + row_offset, j, metric, p)
    n_bad, max_rel = compare_approx(out[i, j], ref, eps)
    return {"checked": int(i.numel()), "n_bad": n_bad, "max_rel_err": max_rel}


def nn_exact_fp64(xq, y, chunk_rows: int = 65536, idx_offset: int = 0):
    """Exact fp64 (min squared L2 distance, smallest arg-min index + idx_offset) of every row of xq over ALL rows of y.
    fp64 expanded form: its cancellation error (~1e-13 of the norms) is far below any fp32 near-tie."""
    q = xq.to(torch.float64)
    qn = (q * q).sum(1)
    best = torch.full((q.shape[0],), float("inf"), dtype=torch.float64, device=q.device)
    arg = torch.full((q.shape[0],), -1, dtype=torch.int64, device=q.device)
    for j0 in range(0, y.shape[0], chunk_rows):
        yb = y[j0:j0 + chunk_rows].to(torch.float64)
        d = torch.addmm((yb * yb).sum(1)[None, :].expand(q.shape[0], -1), q, yb.t(), alpha=-2.0)   # yn - 2 q.y
        v, loc = d.min(dim=1)          # first occurrence == smallest index (ties -> smaller j)
        upd = v < best                 # strict: an equal value in a later chunk has a larger index
        best = torch.where(upd, v, best)
        arg = torch.where(upd, loc + j0 + idx_offset, arg)
        del d, yb
    return (best + qn).clamp_min(0.0), arg


def check_nn(got_idx, got_val, ref_val, ref_idx, x, y_lookup, eps: float = 1e-4, tie_rel: float = 1e-6):
    """Compares an engine result (for the sampled queries x) with the exact fp64 answer.
    y_lookup(idx) -> rows of the database for GLOBAL indices idx (fp32).  strict = same index; tie-aware = the
    chosen row's fp64 distance is within tie_rel (relative) of the true minimum."""
    same = got_idx.to(torch.int64) == ref_idx
    a = x.to(torch.float64)
    b = y_lookup(got_idx.to(torch.int64)).to(torch.float64)
    d_got = ((a - b) ** 2).sum(1)
    gap = (d_got - ref_val).abs() / ref_val.clamp_min(1e-300)
    tie_ok = same | (gap <= tie_rel)
    n_bad_val, max_rel = compare_approx(got_val, ref_val, eps)
    n = int(same.numel())
    return {"checked": n, "idx_strict_match": float(same.sum().item()) / n, "idx_tie_aware_match": float(tie_ok.sum().item()) / n,
            "max_gap_of_mismatches": float(gap[~same].max().item()) if int((~same).sum().item()) else 0.0,
            "val_n_bad": n_bad_val, "val_max_rel_err": max_rel}
