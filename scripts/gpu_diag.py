"""GPU diagnostics: error statistics of every kernel family against the CPU oracle, plus quick
timings.  Usage: python scripts/gpu_diag.py <group>   (groups: ux, tc_small, tc_more, nn, time)
Each group runs in its own process so that a device-side trap in one does not poison the rest."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from raft_b200.distance import pairwise_distance, fused_l2_nn, DistanceType as DT
from raft_b200.common import DeviceResources


def data(m, n, k, seed=0, kind="blobs"):
    if kind == "blobs":
        x, _, c = oracle.make_blobs(m, k, seed=1234 + seed)
        y, _, _ = oracle.make_blobs(n, k, seed=4321 + seed, centers=c)
    else:
        rng = np.random.default_rng(seed)
        x = rng.uniform(-1, 1, (m, k)).astype(np.float32)
        y = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    return x, y


def err_stats(got, ref, eps=1e-4):
    got = got.astype(np.float64)
    diff = np.abs(got - ref)
    den = np.maximum(np.abs(got), np.abs(ref))
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(diff > eps, diff / den, diff)
    bad = (~oracle.compare_approx(got, ref, eps)).sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        true_rel = np.where(den > 0, diff / den, 0.0)
    return (f"cmp_max={np.nanmax(rel):.3e} true_rel_max={np.nanmax(true_rel):.3e} "
            f"true_rel_p99.9={np.nanquantile(true_rel, 0.999):.3e} max_abs={np.nanmax(diff):.3e} bad={bad}/{got.size}")


def run_pw(metric, m, n, k, kind="blobs", p=2.0):
    x, y = data(m, n, k, kind=kind)
    ref = oracle.pairwise_distance(x, y, metric, p)
    out = pairwise_distance(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), metric=metric, p=p)
    got = out.copy_to_host()
    print(f"  {DT(metric).name:20s} {m}x{n}x{k} {kind}: {err_stats(got, ref)}", flush=True)


def group_ux():
    for metric in (DT.L1, DT.L2Unexpanded, DT.L2SqrtUnexpanded, DT.Linf, DT.Canberra, DT.LpUnexpanded):
        for shape in ((1024, 1024, 32), (333, 257, 45)):
            run_pw(metric, *shape, p=3.0)


def group_tc_small():
    run_pw(DT.L2Expanded, 128, 256, 32)
    run_pw(DT.L2Expanded, 128, 256, 32, kind="uniform")
    run_pw(DT.L2Expanded, 1024, 1024, 32)
    run_pw(DT.L2Expanded, 1000, 777, 100)
    run_pw(DT.L2Expanded, 2048, 4096, 128)


def group_tc_more():
    run_pw(DT.L2Expanded, 513, 300, 160)      # streaming variant
    run_pw(DT.L2Expanded, 1024, 32, 1024)
    run_pw(DT.L2SqrtExpanded, 1024, 1024, 32)
    run_pw(DT.CosineExpanded, 1000, 777, 100)
    run_pw(DT.CorrelationExpanded, 1000, 777, 100)
    run_pw(DT.InnerProduct, 1000, 777, 100)
    run_pw(DT.L2Expanded, 1000, 777, 100, kind="uniform")
    run_pw(DT.CosineExpanded, 1024, 1024, 32, kind="uniform")


def group_nn():
    for (m, n, k) in ((1024, 1024, 32), (5000, 3000, 96), (777, 10000, 128), (300, 500, 200)):
        x, y = data(m, n, k)
        ri, rv = oracle.fused_l2_nn(x, y, sqrt=False)
        gi, gv = fused_l2_nn(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), sqrt=False)
        gi, gv = gi.cpu().numpy(), gv.cpu().numpy()
        print(f"  fusedL2NN {m}x{n}x{k}: idx_mismatch={(gi != ri).sum()} val {err_stats(gv, rv)}", flush=True)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]


def group_time():
    h = DeviceResources()
    g = torch.Generator(device="cuda").manual_seed(0)
    for (m, n, k) in ((16384, 16384, 128), (50000, 50000, 128), (100000, 100000, 128)):
        x = torch.randn(m, k, device="cuda", generator=g) * 3
        y = torch.randn(n, k, device="cuda", generator=g) * 3
        out = torch.empty(m, n, device="cuda")
        ms = timeit(lambda: pairwise_distance(x, y, out=out, metric="sqeuclidean", handle=h))
        print(f"  L2Expanded {m}x{n}x{k}: {ms:.3f} ms  {m*n/ms/1e9:.1f} Gpairs/s  store {m*n*4/ms/1e6:.0f} GB/s", flush=True)
        del out
    for (m, n, k) in ((100000, 1000000, 96), (1000000, 1000000, 96)):
        x = torch.randn(m, k, device="cuda", generator=g) * 3
        y = torch.randn(n, k, device="cuda", generator=g) * 3
        ms = timeit(lambda: fused_l2_nn(x, y, sqrt=False, handle=h), iters=3, warm=1)
        print(f"  fusedL2NN {m}x{n}x{k}: {ms:.3f} ms  {m*n/ms/1e9:.1f} Gpairs/s  {3*2*m*n*k/ms/1e12:.1f} TF/s(3x)", flush=True)
    for metric in ("cityblock", "sqeuclidean_unexpanded", "chebyshev"):
        m = n = 20000; k = 256
        x = torch.randn(m, k, device="cuda", generator=g)
        y = torch.randn(n, k, device="cuda", generator=g)
        out = torch.empty(m, n, device="cuda")
        ms = timeit(lambda: pairwise_distance(x, y, out=out, metric=metric, handle=h), iters=3, warm=1)
        print(f"  {metric} {m}x{n}x{k}: {ms:.3f} ms  {m*n/ms/1e9:.2f} Gpairs/s  {2*m*n*k/ms/1e9:.0f} Glane-ops/s", flush=True)


if __name__ == "__main__":
    grp = sys.argv[1]
    print(f"== {grp} ==", flush=True)
    t0 = time.time()
    globals()["group_" + grp]()
    torch.cuda.synchronize()
    print(f"== {grp} done in {time.time()-t0:.1f}s ==", flush=True)
