// raft::distance::pairwise_distance -- header-only shim with the signatures of the removed
// raft/distance/distance.cuh, dispatching through the C ABI (include/raft_b200.h) into the sm_100a
// kernels.  Call shapes kept (verbatim from the surviving callers):
//   pairwise_distance(handle, x, y, dist, m, n, k, metric)             cpp/include/raft/stats/detail/silhouette_score.cuh:205-206
//   pairwise_distance(handle, x, y, dist, m, n, k, metric, isRowMajor, metric_arg)
//                                                                      cpp/include/raft/stats/detail/trustworthiness_score.cuh:152-153
//   pairwise_distance(handle, x_view, y_view, dist_view, metric, metric_arg)   (mdspan form, README.md:115-127 analogue)
// Semantics kept: asynchronous on raft::resource::get_cuda_stream(handle), never synchronises,
// temporaries from the handle's workspace, errors as raft::logic_error / raft::cuda_error.
#pragma once
#include <type_traits>

#include "../../raft_b200.h"
#include "../core/b200_compat.hpp"
#include "distance_types.hpp"

namespace raft {
namespace distance {
namespace detail {
inline void b2d_check(int status)
{
  if (status == B2D_OK) return;
  std::string msg = b2d_last_error();
  if (status == B2D_ERR_CUDA) throw raft::cuda_error(msg);
  throw raft::logic_error(msg);
}
template <typename T>
constexpr int b2d_dtype()
{
  static_assert(std::is_same<T, float>::value, "raft_b200: fp32 inputs (fp16 through the C ABI)");
  return B2D_F32;
}
}  // namespace detail

template <typename DataT, typename IdxT = int>
void pairwise_distance(raft::resources const& handle, const DataT* x, const DataT* y, DataT* dist, IdxT m, IdxT n,
                       IdxT k, raft::distance::DistanceType metric, bool isRowMajor = true, DataT metric_arg = 2.0f)
{
  const int dt      = detail::b2d_dtype<DataT>();
  const size_t need = b2d_pairwise_workspace_bytes(static_cast<int>(metric), dt, m, n, k);
  RAFT_EXPECTS(need != static_cast<size_t>(-1), "pairwise_distance: metric not supported by raft_b200");
  void* ws = need ? handle.workspace(need) : nullptr;
  const int64_t ldx = isRowMajor ? k : m, ldy = isRowMajor ? k : n, ldd = isRowMajor ? n : m;
  detail::b2d_check(b2d_pairwise_distance(raft::resource::get_cuda_stream(handle), static_cast<int>(metric), dt, x, ldx,
                                          y, ldy, dist, ldd, m, n, k, isRowMajor ? 1 : 0,
                                          static_cast<float>(metric_arg), ws, need));
}

// overload with a caller-supplied workspace (the reference took an rmm::device_uvector<char>&;
// here: any device buffer of at least b2d_pairwise_workspace_bytes() bytes)
template <typename DataT, typename IdxT = int>
void pairwise_distance(raft::resources const& handle, const DataT* x, const DataT* y, DataT* dist, IdxT m, IdxT n,
                       IdxT k, void* workspace, size_t workspace_bytes, raft::distance::DistanceType metric,
                       bool isRowMajor = true, DataT metric_arg = 2.0f)
{
  const int64_t ldx = isRowMajor ? k : m, ldy = isRowMajor ? k : n, ldd = isRowMajor ? n : m;
  detail::b2d_check(b2d_pairwise_distance(raft::resource::get_cuda_stream(handle), static_cast<int>(metric),
                                          detail::b2d_dtype<DataT>(), x, ldx, y, ldy, dist, ldd, m, n, k,
                                          isRowMajor ? 1 : 0, static_cast<float>(metric_arg), workspace,
                                          workspace_bytes));
}

template <typename DataT, typename IdxT, typename Layout>
void pairwise_distance(raft::resources const& handle, raft::device_matrix_view<const DataT, IdxT, Layout> x,
                       raft::device_matrix_view<const DataT, IdxT, Layout> y,
                       raft::device_matrix_view<DataT, IdxT, Layout> dist, raft::distance::DistanceType metric,
                       DataT metric_arg = 2.0f)
{
  RAFT_EXPECTS(x.extent(1) == y.extent(1), "Number of columns must be equal.");
  RAFT_EXPECTS(dist.extent(0) == x.extent(0), "Number of rows in output must be equal to number of rows in X");
  RAFT_EXPECTS(dist.extent(1) == y.extent(0), "Number of columns in output must be equal to number of rows in Y");
  constexpr bool rm = std::is_same<Layout, raft::row_major>::value;
  pairwise_distance<DataT, IdxT>(handle, x.data_handle(), y.data_handle(), dist.data_handle(), x.extent(0), y.extent(0),
                                 x.extent(1), metric, rm, metric_arg);
}

}  // namespace distance
}  // namespace raft
