"""fused_l2_nn_argmin / fused_l2_nn -- mirror of ``pylibraft.distance.fused_l2_nn_argmin``
(C++: raft::distance::fusedL2NNMinReduce, runtime ABI
raft::runtime::distance::fused_l2_nn_min_arg) plus the multi-GPU, db-row-sharded variant of
SURVEY.md 8(e): per-GPU K3 with an index offset, then ONE exchange step -- an all-reduce(MIN) of
packed 64-bit (distance, index) keys -- then unpack."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..common import auto_convert_output, auto_sync_handle, cai_wrapper, device_ndarray
from .distance_type import resolve_metric

KVP_DTYPE = np.dtype([("key", np.int32), ("value", np.float32)])  # raft::KeyValuePair<int,float>
SHARD_HEAD_ROWS = 1 << 15  # rows of every shard searched before the first exchange of bounds (fused_l2_nn_sharded)


def _check_xy(X, Y):
    x_cai, y_cai = cai_wrapper(X), cai_wrapper(Y)
    x_cai.validate_shape_dtype(expected_dims=2, expected_dtype=np.float32)
    y_cai.validate_shape_dtype(expected_dims=2, expected_dtype=np.float32)
    if x_cai.shape[1] != y_cai.shape[1]:
        raise ValueError("Inputs must have same number of columns. a=%s, b=%s" % (x_cai.shape[1], y_cai.shape[1]))
    if not (x_cai.c_contiguous and y_cai.c_contiguous):
        raise ValueError("Inputs must be C contiguous")
    return x_cai, y_cai


@auto_sync_handle
def fused_l2_nn(X, Y, sqrt=True, xn=None, yn=None, handle=None):
    """Returns (indices int32 [m], distances float32 [m]) as torch tensors: the
    raft::KeyValuePair<int,float> output of fusedL2NN split into its two fields."""
    x_cai, y_cai = _check_xy(X, Y)
    m, k = x_cai.shape
    n = y_cai.shape[0]
    L = _lib.lib()
    need = L.b2d_fused_l2_nn_workspace_bytes(m, n, k)
    ws = handle.workspace(need)
    with torch.cuda.stream(handle.torch_stream):
        kvp = torch.empty((m, 2), dtype=torch.int32, device=handle.device)
    xn_p = cai_wrapper(xn).data if xn is not None else None
    yn_p = cai_wrapper(yn).data if yn is not None else None
    _lib.check(L.b2d_fused_l2_nn(handle.stream_ptr, kvp.data_ptr(), x_cai.data, k, y_cai.data, k, xn_p, yn_p,
                                 m, n, k, 1 if sqrt else 0, 1, ws.data_ptr(), ws.numel()))
    return kvp[:, 0], kvp[:, 1].view(torch.float32)


@auto_sync_handle
def fused_distance_nn(X, Y, metric="euclidean", handle=None):
    """raft::distance::fusedDistanceNN: (indices int32 [m], distances float32 [m]) of the nearest row
    of Y for every row of X under metric in {sqeuclidean, euclidean, cosine, correlation}."""
    x_cai, y_cai = _check_xy(X, Y)
    m, k = x_cai.shape
    n = y_cai.shape[0]
    L = _lib.lib()
    ws = handle.workspace(L.b2d_fused_l2_nn_workspace_bytes(m, n, k))
    with torch.cuda.stream(handle.torch_stream):
        kvp = torch.empty((m, 2), dtype=torch.int32, device=handle.device)
    _lib.check(L.b2d_fused_distance_nn(handle.stream_ptr, kvp.data_ptr(), int(resolve_metric(metric)), x_cai.data, k,
                                       y_cai.data, k, None, None, m, n, k, 1, ws.data_ptr(), ws.numel()))
    return kvp[:, 0], kvp[:, 1].view(torch.float32)


@auto_sync_handle
@auto_convert_output
def fused_l2_nn_argmin(X, Y, out=None, sqrt=True, handle=None):
    """argmin_j ||X_i - Y_j|| for every row of X: int32 [m] (ties -> smaller j)."""
    x_cai, _ = _check_xy(X, Y)
    m = x_cai.shape[0]
    idx, _ = fused_l2_nn(X, Y, sqrt=sqrt, handle=handle)
    if out is None:
        return device_ndarray(idx.contiguous())
    o_cai = cai_wrapper(out)
    o_cai.validate_shape_dtype(expected_dims=1, expected_dtype=np.int32)
    if o_cai.shape[0] != m:
        raise ValueError("out must have shape (%d,)" % m)
    with torch.cuda.stream(handle.torch_stream):
        torch.as_tensor(out, device=handle.device).copy_(idx)
    return out


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous row-block [lo, hi) of the database owned by `rank`."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


def plan_exchanges(shard_rows: int, world: int):
    """Sub-chunks (rows) a shard is searched in, with one exchange of the packed keys after each: a small head
    first (every rank then knows the best distance found ANYWHERE in world x head rows), growing geometrically --
    the screened search (screen_tc.cuh) keeps far fewer candidates once its bounds are global.  Depends only on
    (shard_rows, world): pass the SAME shard_rows on every rank (n_total // world), the last sub-chunk takes
    whatever a rank's shard has left."""
    if world <= 1 or shard_rows <= 0:
        return [max(shard_rows, 0)]
    sizes, left, step = [], shard_rows, SHARD_HEAD_ROWS
    while left > 4 * step and len(sizes) < 3:
        sizes.append(step)
        left -= step
        step *= 4
    sizes.append(left)
    return sizes


def fused_l2_nn_sharded(X, Y_shard, idx_offset, sqrt=True, handle=None, group=None, keys=None, head_rows=None,
                        n_total=None):
    """Multi-GPU fusedL2NN: every rank holds all queries X [m,k] and its own row-block Y_shard of
    the database (global row index of its first row = idx_offset).  One process per GPU;
    `group` is a torch.distributed process group (NCCL over NVLink on the GPU box).

    Returns (indices int32 [m] -- GLOBAL database rows, distances float32 [m]) on every rank.
    The only collective is all_reduce(MIN) over m packed int64 keys: NCCL has no MINLOC, and
    signed 64-bit MIN over (ordered distance bits << 32 | index) is exactly raft::argmin_op
    (smaller value first, then smaller index; cpp/include/raft/core/operators.hpp:187-194).

    Exchange plan (the SAME number of collectives on every rank, whatever the local shard size -- shards from
    shard_bounds differ by a row): with n_total (rows of the whole database) the shard is searched in the
    sub-chunks of plan_exchanges(n_total // world, world), one all_reduce after each; without it, in a head of
    min(SHARD_HEAD_ROWS, n) rows and the rest (two all_reduces, always).  No row is visited twice: every call
    continues behind the previous one with the reduced keys, which also serve as the search's starting bounds.
    head_rows (tests): an explicit head size, single exchange plan [head, rest]."""
    import torch.distributed as dist

    own = handle is None
    if handle is None:
        from ..common import DeviceResources
        handle = DeviceResources()
    x_cai, y_cai = _check_xy(X, Y_shard)
    m, k = x_cai.shape
    n = y_cai.shape[0]
    L = _lib.lib()
    ws = handle.workspace(L.b2d_fused_l2_nn_workspace_bytes(m, n, k))
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if multi else 1
    if head_rows is not None:
        head = max(0, min(int(head_rows), n))
        plan = [head, n - head] if head else [n]
    elif not multi:
        plan = [n]
    elif n_total is not None:
        plan = plan_exchanges(int(n_total) // world, world)
    else:
        plan = [SHARD_HEAD_ROWS, -1]          # -1: the rest (rank-invariant count: always two exchanges)
    with torch.cuda.stream(handle.torch_stream):
        if keys is None:
            keys = torch.empty(m, dtype=torch.int64, device=handle.device)
        kvp = torch.empty((m, 2), dtype=torch.int32, device=handle.device)
        done = 0
        for c, rows in enumerate(plan):
            last = c == len(plan) - 1
            rows = n - done if (last or rows < 0) else max(0, min(rows, n - done))
            # (a call with 0 rows only initialises the keys; the collective below still takes place)
            _lib.check(L.b2d_fused_l2_nn_keys(handle.stream_ptr, keys.data_ptr(), x_cai.data, k,
                                              y_cai.data + done * k * 4, k, None, None, m, rows, k,
                                              int(idx_offset) + done, 1 if c == 0 else 0, ws.data_ptr(), ws.numel()))
            done += rows
            if multi:
                dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
        _lib.check(L.b2d_fused_l2_nn_finalize(handle.stream_ptr, kvp.data_ptr(), keys.data_ptr(), m,
                                              1 if sqrt else 0, ws.data_ptr(), ws.numel()))
    if own:
        handle.sync()
    return kvp[:, 0], kvp[:, 1].view(torch.float32)
