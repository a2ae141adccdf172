// raft::stats::silhouette_score -- shim over b2d_silhouette_score (include/raft_b200.h).
// Pointer signature of cpp/include/raft/stats/silhouette_score.cuh (detail: stats/detail/
// silhouette_score.cuh:186-196), whose pairwise_distance call dangles in the reference snapshot
// (SURVEY.md 8(f3)).  DataT = float, LabelT = int.  Returns the mean score (synchronises the stream).
#pragma once
#include "../distance/distance.cuh"

namespace raft {
namespace stats {

template <typename DataT, typename LabelT>
DataT silhouette_score(raft::resources const& handle, const DataT* X_in, int nRows, int nCols, const LabelT* labels,
                       int nLabels, DataT* silhouette_scorePerSample, cudaStream_t stream,
                       raft::distance::DistanceType metric = raft::distance::DistanceType::L2Unexpanded,
                       int chunk = 0)
{
  static_assert(std::is_same<DataT, float>::value && sizeof(LabelT) == 4, "raft_b200: silhouette_score is provided for <float, int>");
  const size_t need = b2d_silhouette_score_workspace_bytes(nRows, nCols, nLabels, static_cast<int>(metric), chunk);
  if (need == static_cast<size_t>(-1)) throw raft::logic_error("silhouette_score: metric not supported");
  raft::b200::scoped_workspace scratch(handle, need + 256);
  char* ws       = static_cast<char*>(scratch.data());
  float* d_score = reinterpret_cast<float*>(ws);
  raft::distance::detail::b2d_check(b2d_silhouette_score(stream, d_score, silhouette_scorePerSample, X_in, nCols,
                                                         reinterpret_cast<const int*>(labels), nRows, nCols, nLabels,
                                                         static_cast<int>(metric), 2.0f, chunk, ws + 256, need));
  float h = 0.f;
  if (cudaMemcpyAsync(&h, d_score, sizeof(float), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess)
    throw raft::cuda_error("silhouette_score: result copy failed");
  return h;
}

}  // namespace stats
}  // namespace raft
