// raft::neighbors::brute_force::fused_l2_knn / knn -- shim over b2d_knn_l2 (include/raft_b200.h).
// The reference dropped these together with the distance package (CHANGELOG.md:59-60); what is left
// of the step is raft::matrix::select_k (cpp/include/raft/matrix/select_k.cuh:73-106), which callers
// ran over a materialised distance matrix.  Here the selection is fused behind the distance tiles
// (SURVEY.md 8(f2)): index [n, d], query [m, d] row-major fp32, out_inds int64 [m, k], out_dists [m, k]
// in ascending (distance, index) order.  Metrics: L2Expanded / L2Unexpanded (squared) and the Sqrt forms.
#pragma once
#include "../distance/distance.cuh"

namespace raft {
namespace neighbors {
namespace brute_force {

template <typename value_t, typename idx_t, typename idx_layout, typename query_layout>
void fused_l2_knn(raft::resources const& handle, raft::device_matrix_view<const value_t, idx_t, idx_layout> index,
                  raft::device_matrix_view<const value_t, idx_t, query_layout> query,
                  raft::device_matrix_view<idx_t, idx_t, raft::row_major> out_inds,
                  raft::device_matrix_view<value_t, idx_t, raft::row_major> out_dists,
                  raft::distance::DistanceType metric)
{
  static_assert(std::is_same<value_t, float>::value && sizeof(idx_t) == 8, "raft_b200: fused_l2_knn is provided for <float, int64_t>");
  using raft::distance::DistanceType;
  const bool sq = metric == DistanceType::L2Expanded || metric == DistanceType::L2Unexpanded;
  const bool rt = metric == DistanceType::L2SqrtExpanded || metric == DistanceType::L2SqrtUnexpanded;
  if (!sq && !rt) throw raft::logic_error("fused_l2_knn: L2 metrics only");
  if (index.extent(1) != query.extent(1)) throw raft::logic_error("fused_l2_knn: index and query must have the same number of columns");
  if (out_inds.extent(0) != query.extent(0) || out_dists.extent(0) != query.extent(0) ||
      out_inds.extent(1) != out_dists.extent(1))
    throw raft::logic_error("fused_l2_knn: output shapes must be [n_queries, k]");
  const int64_t m = query.extent(0), n = index.extent(0), d = index.extent(1), k = out_inds.extent(1);
  const size_t need = b2d_knn_l2_workspace_bytes(m, n, d, k);
  if (need == static_cast<size_t>(-1)) throw raft::logic_error("fused_l2_knn: k must be in [1, 64]");
  raft::b200::scoped_workspace ws(handle, need);
  raft::distance::detail::b2d_check(b2d_knn_l2(raft::resource::get_cuda_stream(handle),
                                               reinterpret_cast<int64_t*>(out_inds.data_handle()), out_dists.data_handle(),
                                               query.data_handle(), d, index.data_handle(), d, m, n, d, k, rt ? 1 : 0,
                                               ws.data(), need));
}

// single-partition form of brute_force::knn (the reference's vector-of-partitions overload merged
// per-partition results with knn_merge_parts; one partition needs no merge)
template <typename value_t, typename idx_t>
void knn(raft::resources const& handle, raft::device_matrix_view<const value_t, idx_t, raft::row_major> index,
         raft::device_matrix_view<const value_t, idx_t, raft::row_major> search,
         raft::device_matrix_view<idx_t, idx_t, raft::row_major> indices,
         raft::device_matrix_view<value_t, idx_t, raft::row_major> distances,
         raft::distance::DistanceType metric = raft::distance::DistanceType::L2Unexpanded)
{
  fused_l2_knn(handle, index, search, indices, distances, metric);
}

}  // namespace brute_force
}  // namespace neighbors
}  // namespace raft
