"""ctypes binding of include/raft_b200.h.  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os

from . import _build

_lib = None

# every symbol include/raft_b200.h declares
SYMBOLS = [
    "b2d_version", "b2d_last_error", "b2d_set_option", "b2d_debug_nn_stats", "b2d_pairwise_workspace_bytes", "b2d_pairwise_distance",
    "b2d_fused_l2_nn_workspace_bytes", "b2d_fused_l2_nn", "b2d_fused_distance_nn", "b2d_fused_l2_nn_keys",
    "b2d_fused_l2_nn_finalize", "b2d_row_norm", "b2d_knn_l2_workspace_bytes", "b2d_knn_l2", "b2d_knn",
    "b2d_silhouette_score_workspace_bytes", "b2d_silhouette_score",
    "b2d_trustworthiness_score_workspace_bytes", "b2d_trustworthiness_score", "b2d_profile_begin", "b2d_profile_end",
    "b2d_row_argmin", "b2d_fused_l2_nn_multi",
]

B2D_OK, B2D_ERR_INVALID_ARG, B2D_ERR_CUDA, B2D_ERR_UNSUPPORTED, B2D_ERR_WORKSPACE = range(5)
B2D_F32, B2D_F16, B2D_F64 = 0, 1, 2


class RaftB200Error(RuntimeError):
    """Mirrors raft::exception (cpp/include/raft/core/error.hpp:37-58)."""


class LogicError(RaftB200Error, ValueError):
    """raft::logic_error (RAFT_EXPECTS failures), error.hpp:218-239."""


class CudaError(RaftB200Error):
    """raft::cuda_error, cpp/include/raft/util/cuda_rt_essentials.hpp:23-52."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RAFT_B200_LIB", _build.SO_PATH)   # developer hook: an experimental build of the same ABI
    if not os.path.exists(path):
        if os.environ.get("RAFT_B200_NO_AUTOBUILD"):
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _build.build()
    L = ctypes.CDLL(path)
    i64, vp, sz, ci, cf = ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
    L.b2d_version.restype = ci
    L.b2d_last_error.restype = ctypes.c_char_p
    L.b2d_set_option.restype = ci
    L.b2d_set_option.argtypes = [ctypes.c_char_p, ctypes.c_double]
    L.b2d_debug_nn_stats.restype = ci
    L.b2d_debug_nn_stats.argtypes = [vp, vp, i64, i64, i64, vp]
    L.b2d_pairwise_workspace_bytes.restype = sz
    L.b2d_pairwise_workspace_bytes.argtypes = [ci, ci, i64, i64, i64]
    L.b2d_pairwise_distance.restype = ci
    L.b2d_pairwise_distance.argtypes = [vp, ci, ci, vp, i64, vp, i64, vp, i64, i64, i64, i64, ci, cf, vp, sz]
    L.b2d_fused_l2_nn_workspace_bytes.restype = sz
    L.b2d_fused_l2_nn_workspace_bytes.argtypes = [i64, i64, i64]
    L.b2d_fused_l2_nn.restype = ci
    L.b2d_fused_l2_nn.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp, i64, i64, i64, ci, ci, vp, sz]
    L.b2d_fused_distance_nn.restype = ci
    L.b2d_fused_distance_nn.argtypes = [vp, vp, ci, vp, i64, vp, i64, vp, vp, i64, i64, i64, ci, vp, sz]
    L.b2d_fused_l2_nn_keys.restype = ci
    L.b2d_fused_l2_nn_keys.argtypes = [vp, vp, vp, i64, vp, i64, vp, vp, i64, i64, i64, i64, ci, vp, sz]
    L.b2d_fused_l2_nn_finalize.restype = ci
    L.b2d_fused_l2_nn_finalize.argtypes = [vp, vp, vp, i64, ci, vp, sz]
    L.b2d_knn_l2_workspace_bytes.restype = sz
    L.b2d_knn_l2_workspace_bytes.argtypes = [i64, i64, i64, i64]
    L.b2d_knn.restype = ci
    L.b2d_knn.argtypes = [vp, vp, vp, ci, vp, i64, vp, i64, i64, i64, i64, i64, vp, sz]
    L.b2d_knn_l2.restype = ci
    L.b2d_knn_l2.argtypes = [vp, vp, vp, vp, i64, vp, i64, i64, i64, i64, i64, ci, vp, sz]
    L.b2d_silhouette_score_workspace_bytes.restype = sz
    L.b2d_silhouette_score_workspace_bytes.argtypes = [i64, i64, ci, ci, i64]
    L.b2d_silhouette_score.restype = ci
    L.b2d_silhouette_score.argtypes = [vp, vp, vp, vp, i64, vp, i64, i64, ci, ci, cf, i64, vp, sz]
    L.b2d_trustworthiness_score_workspace_bytes.restype = sz
    L.b2d_trustworthiness_score_workspace_bytes.argtypes = [i64, i64, i64, ci, ci, i64]
    L.b2d_trustworthiness_score.restype = ci
    L.b2d_trustworthiness_score.argtypes = [vp, vp, vp, i64, vp, i64, i64, i64, i64, ci, ci, i64, vp, sz]
    L.b2d_profile_begin.restype = ci
    L.b2d_profile_begin.argtypes = [ci]
    L.b2d_profile_end.restype = ci
    L.b2d_profile_end.argtypes = [vp, ci, vp]
    L.b2d_row_argmin.restype = ci
    L.b2d_row_argmin.argtypes = [vp, vp, vp, i64, i64, i64]
    L.b2d_row_norm.restype = ci
    L.b2d_row_norm.argtypes = [vp, vp, vp, i64, i64, i64, ci, ci]
    _lib = L
    return L


def check(status: int) -> None:
    if status == B2D_OK:
        return
    msg = lib().b2d_last_error().decode("utf-8", "replace")
    if status in (B2D_ERR_INVALID_ARG, B2D_ERR_UNSUPPORTED, B2D_ERR_WORKSPACE):
        raise LogicError(msg)
    raise CudaError(msg)
