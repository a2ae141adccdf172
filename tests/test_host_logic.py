"""CPU tier: host-side mirror of the pylibraft conventions and the multi-GPU exchange logic
(world_size-2 gloo run of the packed min-loc all-reduce)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import oracle
from raft_b200.common.cai_wrapper import cai_wrapper
from raft_b200.distance.distance_type import DISTANCE_TYPES, DistanceType, resolve_metric
from raft_b200.distance.fused_l2_nn import shard_bounds
from raft_b200.distance.pairwise_distance import _layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeCai:
    def __init__(self, arr, ptr=0x1000):
        self.__cuda_array_interface__ = {"shape": arr.shape, "typestr": arr.dtype.str, "version": 3,
                                         "data": (ptr, False),
                                         "strides": None if arr.flags.c_contiguous else arr.strides}


def test_distance_type_values_match_reference_enum():
    assert int(DistanceType.L2Expanded) == 0 and int(DistanceType.L2SqrtExpanded) == 1
    assert int(DistanceType.CosineExpanded) == 2 and int(DistanceType.L1) == 3
    assert int(DistanceType.LpUnexpanded) == 9 and int(DistanceType.CorrelationExpanded) == 10
    assert int(DistanceType.Precomputed) == 100
    assert {int(v) for v in DistanceType} == {int(v) for v in oracle.DistanceType}


def test_metric_strings():
    assert resolve_metric("euclidean") == DistanceType.L2SqrtExpanded
    assert resolve_metric("sqeuclidean") == DistanceType.L2Expanded
    assert resolve_metric("cityblock") == DistanceType.L1
    assert resolve_metric("chebyshev") == DistanceType.Linf
    assert resolve_metric("minkowski") == DistanceType.LpUnexpanded
    assert resolve_metric("L2Unexpanded") == DistanceType.L2Unexpanded
    assert resolve_metric(8) == DistanceType.Canberra
    assert resolve_metric("hamming") == DistanceType.HammingUnexpanded
    assert resolve_metric("kl_divergence") == DistanceType.KLDivergence
    with pytest.raises(ValueError):
        resolve_metric("haversine")
    assert set(DISTANCE_TYPES) >= {"l2", "l1", "cosine", "correlation", "canberra", "inner_product", "lp"}


def test_cai_wrapper_and_layout():
    a = np.zeros((5, 7), np.float32)
    w = cai_wrapper(FakeCai(a))
    assert w.shape == (5, 7) and w.c_contiguous and not w.f_contiguous and w.data == 0x1000
    assert _layout(w) == (True, 7)
    f = np.asfortranarray(a)
    wf = cai_wrapper(FakeCai(f))
    assert wf.f_contiguous and _layout(wf) == (False, 5)
    padded = np.zeros((5, 16), np.float32)[:, :7]
    assert _layout(cai_wrapper(FakeCai(padded))) == (True, 16)
    with pytest.raises(TypeError):
        cai_wrapper(FakeCai(a)).validate_shape_dtype(expected_dtype=np.float64)
    with pytest.raises(ValueError):
        cai_wrapper(FakeCai(a)).validate_shape_dtype(expected_dims=1)
    with pytest.raises(TypeError):
        cai_wrapper(a)


def test_shard_bounds_cover_database_exactly():
    for n, w in ((8_000_000, 8), (1001, 4), (7, 8), (5, 1)):
        cuts = [shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.environ["B2D_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    import oracle
    from raft_b200.distance.fused_l2_nn import shard_bounds
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["B2D_PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    x, _, c = oracle.make_blobs(257, 24, seed=11)
    y, _, _ = oracle.make_blobs(1003, 24, seed=12, centers=c)
    y[500] = y[3]                      # a tie across shards: the smaller global index must win
    x[0] = y[3]
    lo, hi = shard_bounds(y.shape[0], world, rank)
    # what the GPU kernel leaves in `keys` for this shard (host model: oracle.pack_minloc)
    d = oracle.pairwise_distance(x, y[lo:hi], oracle.DistanceType.L2Expanded)
    loc = np.argmin(d, axis=1)
    v = d[np.arange(len(x)), loc].astype(np.float32)            # ||x - y||^2, as the kernel packs it
    keys = torch.from_numpy(oracle.pack_minloc(v, loc + lo))
    dist.all_reduce(keys, op=dist.ReduceOp.MIN)                 # the ONE exchange step
    val, idx = oracle.unpack_minloc(keys.numpy())
    ref_idx, ref_val = oracle.fused_l2_nn(x, y)
    assert (idx == ref_idx).all(), (rank, np.nonzero(idx != ref_idx))
    assert idx[0] == 3
    assert np.allclose(np.maximum(val.astype(np.float64), 0), ref_val, rtol=1e-4, atol=1e-3)
    # the two-exchange form of fused_l2_nn_sharded: head of every shard, all-reduce (global bounds),
    # the rest of the shard continuing from the reduced keys (init_keys = 0 == elementwise MIN), all-reduce
    head = 100
    def shard_keys(a, b):
        dd = oracle.pairwise_distance(x, y[a:b], oracle.DistanceType.L2Expanded)
        ll = np.argmin(dd, axis=1)
        return torch.from_numpy(oracle.pack_minloc(dd[np.arange(len(x)), ll].astype(np.float32), ll + a))
    k2 = shard_keys(lo, lo + head)
    dist.all_reduce(k2, op=dist.ReduceOp.MIN)
    k2 = torch.minimum(k2, shard_keys(lo + head, hi))
    dist.all_reduce(k2, op=dist.ReduceOp.MIN)
    assert torch.equal(k2, keys)
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


def test_minloc_allreduce_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29000 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", B2D_ROOT=ROOT, B2D_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert "ok" in out


def test_pylibraft_alias_namespace():
    """`from pylibraft.distance import pairwise_distance` keeps working (drop-in alias package)."""
    import pylibraft.common
    import pylibraft.config
    import pylibraft.distance as d
    import raft_b200.distance as r
    assert d.pairwise_distance is r.pairwise_distance and d.fused_l2_nn_argmin is r.fused_l2_nn_argmin
    assert "euclidean" in d.DISTANCE_TYPES and callable(pylibraft.config.set_output_as)
    assert pylibraft.common.DeviceResources is not None


def test_exchange_plan_is_rank_invariant_and_covers_every_shard():
    """ranks whose shards differ by one row must issue the SAME number of all_reduces.  The
    plan depends on (n_total // world, world) only; a rank's last sub-chunk takes what its shard has left."""
    from raft_b200.distance import shard_bounds
    from raft_b200.distance.fused_l2_nn import SHARD_HEAD_ROWS, plan_exchanges
    for n_total, world in [(8_000_000, 8), (8_000_001, 8), (1_048_577, 8), (4 * SHARD_HEAD_ROWS * 8 + 3, 8),
                           (4 * SHARD_HEAD_ROWS * 2 - 1, 2), (1000, 4), (5, 8), (8_000_000, 2), (8_000_000, 1)]:
        plan = plan_exchanges(n_total // world, world)
        assert all(p >= 0 for p in plan) and sum(plan) == n_total // world
        assert len(plan) <= 4
        for r in range(world):
            lo, hi = shard_bounds(n_total, world, r)
            n_local, done = hi - lo, 0
            for c, rows in enumerate(plan):          # the loop of fused_l2_nn_sharded
                rows = n_local - done if c == len(plan) - 1 else max(0, min(rows, n_local - done))
                assert rows >= 0
                done += rows
            assert done == n_local
    assert plan_exchanges(1_000_000, 8)[0] == SHARD_HEAD_ROWS and len(plan_exchanges(1_000_000, 8)) >= 2
