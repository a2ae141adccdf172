// raft::stats::trustworthiness_score -- shim over b2d_trustworthiness_score (include/raft_b200.h).
// Pointer signature of cpp/include/raft/stats/trustworthiness_score.cuh:29-41 (its pairwise_distance
// call, detail/trustworthiness_score.cuh:152-153, dangles in the reference snapshot).
#pragma once
#include "../distance/distance.cuh"

namespace raft {
namespace stats {

template <typename math_t, raft::distance::DistanceType distance_type>
double trustworthiness_score(const raft::resources& h, const math_t* X, math_t* X_embedded, int n, int m, int d,
                             int n_neighbors, int batchSize = 512)
{
  static_assert(std::is_same<math_t, float>::value, "raft_b200: trustworthiness_score is provided for float");
  const size_t need = b2d_trustworthiness_score_workspace_bytes(n, m, d, n_neighbors, static_cast<int>(distance_type), batchSize);
  if (need == static_cast<size_t>(-1)) throw raft::logic_error("trustworthiness_score: unsupported n_neighbors / metric");
  raft::b200::scoped_workspace scratch(h, need + 256);
  char* ws        = static_cast<char*>(scratch.data());
  double* d_score = reinterpret_cast<double*>(ws);   // the C ABI writes the score to device memory and never synchronises;
  cudaStream_t s  = raft::resource::get_cuda_stream(h);  // the reference returns it by value, so THIS function waits for it
  raft::distance::detail::b2d_check(b2d_trustworthiness_score(s, d_score, X, m, X_embedded, d, n, m, d, n_neighbors,
                                                              static_cast<int>(distance_type), batchSize, ws + 256, need));
  double score = 0.0;
  if (cudaMemcpyAsync(&score, d_score, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    throw raft::cuda_error("trustworthiness_score: result copy failed");
  return score;
}

}  // namespace stats
}  // namespace raft
