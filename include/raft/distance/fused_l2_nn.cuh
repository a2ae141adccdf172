// raft::distance::fusedL2NNMinReduce -- shim over b2d_fused_l2_nn (include/raft_b200.h).
// Signature of the removed raft/distance/fused_l2_nn.cuh; OutT is
// raft::KeyValuePair<IdxT, DataT> (cpp/include/raft/core/kvp.hpp:20-62) with IdxT = int,
// DataT = float; xn / yn are the SQUARED row norms; `workspace` (the reference's per-row mutex
// array) is accepted and ignored -- this implementation needs no mutex.
#pragma once
#include "distance.cuh"

namespace raft {
namespace distance {

template <typename DataT, typename OutT, typename IdxT>
void fusedL2NNMinReduce(OutT* min, const DataT* x, const DataT* y, const DataT* xn, const DataT* yn, IdxT m, IdxT n,
                        IdxT k, void* /*workspace*/, bool sqrt, bool initOutBuffer, raft::resources const& handle)
{
  static_assert(std::is_same<DataT, float>::value && sizeof(OutT) == sizeof(b2d_kvp_if),
                "raft_b200: fusedL2NN is provided for <float, KeyValuePair<int,float>, int>");
  const size_t need = b2d_fused_l2_nn_workspace_bytes(m, n, k);
  void* ws          = handle.workspace(need);
  detail::b2d_check(b2d_fused_l2_nn(raft::resource::get_cuda_stream(handle), reinterpret_cast<b2d_kvp_if*>(min), x, k, y,
                                    k, xn, yn, m, n, k, sqrt ? 1 : 0, initOutBuffer ? 1 : 0, ws, need));
}

// Legacy stream-only signature (no handle): fusedL2NNMinReduce(min, x, y, xn, yn, m, n, k, workspace,
// sqrt, initOutBuffer, stream) as listed in SURVEY.md 8(a5).  The reference's `workspace` holds m
// ints for its per-row mutexes -- too small for the packed operands this engine stages -- so the
// scratch comes from the stream-ordered allocator and is returned to it behind the kernels.
template <typename DataT, typename OutT, typename IdxT>
void fusedL2NNMinReduce(OutT* min, const DataT* x, const DataT* y, const DataT* xn, const DataT* yn, IdxT m, IdxT n,
                        IdxT k, void* /*workspace*/, bool sqrt, bool initOutBuffer, cudaStream_t stream)
{
  static_assert(std::is_same<DataT, float>::value && sizeof(OutT) == sizeof(b2d_kvp_if),
                "raft_b200: fusedL2NN is provided for <float, KeyValuePair<int,float>, int>");
  const size_t need = b2d_fused_l2_nn_workspace_bytes(m, n, k);
  void* ws          = nullptr;
  if (cudaMallocAsync(&ws, need, stream) != cudaSuccess) throw raft::cuda_error("fusedL2NN: scratch allocation failed");
  const int st = b2d_fused_l2_nn(stream, reinterpret_cast<b2d_kvp_if*>(min), x, k, y, k, xn, yn, m, n, k, sqrt ? 1 : 0,
                                 initOutBuffer ? 1 : 0, ws, need);
  cudaFreeAsync(ws, stream);
  detail::b2d_check(st);
}

// raft::distance::fusedDistanceNNMinReduce: L2 (sqrt flag) or cosine, SURVEY.md 8(f1)
template <typename DataT, typename OutT, typename IdxT>
void fusedDistanceNNMinReduce(OutT* min, const DataT* x, const DataT* y, const DataT* xn, const DataT* yn, IdxT m, IdxT n,
                              IdxT k, void* /*workspace*/, bool sqrt, bool initOutBuffer, bool /*isRowMajor*/,
                              raft::distance::DistanceType metric, float /*metric_arg*/, raft::resources const& handle)
{
  static_assert(std::is_same<DataT, float>::value && sizeof(OutT) == sizeof(b2d_kvp_if), "float / KeyValuePair<int,float>");
  int mt = static_cast<int>(metric);
  if (metric == DistanceType::L2Expanded && sqrt) mt = static_cast<int>(DistanceType::L2SqrtExpanded);
  const size_t need = b2d_fused_l2_nn_workspace_bytes(m, n, k);
  void* ws          = handle.workspace(need);
  detail::b2d_check(b2d_fused_distance_nn(raft::resource::get_cuda_stream(handle), reinterpret_cast<b2d_kvp_if*>(min), mt,
                                          x, k, y, k, xn, yn, m, n, k, initOutBuffer ? 1 : 0, ws, need));
}

}  // namespace distance
}  // namespace raft
