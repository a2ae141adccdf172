// raft::runtime::distance -- the compiled-ABI layer pylibraft bound to (pattern:
// cpp/include/raft_runtime/random/rmat_rectangular_generator.hpp:18-34; shape SURVEY.md 8(b)).
#pragma once
#include "../../raft/distance/fused_l2_nn.cuh"

namespace raft::runtime::distance {
inline void pairwise_distance(raft::resources const& handle, float* x, float* y, float* dists, int m, int n, int k,
                              raft::distance::DistanceType metric, bool isRowMajor, float metric_arg)
{
  raft::distance::pairwise_distance<float, int>(handle, x, y, dists, m, n, k, metric, isRowMajor, metric_arg);
}
// KeyValuePair output (the reference's fused_l2_nn_min_arg returned only the keys; the Python
// mirror raft_b200.distance.fused_l2_nn_argmin slices them out of this result)
inline void fused_l2_nn_min(raft::resources const& handle, raft::KeyValuePair<int, float>* min, const float* x,
                            const float* y, int m, int n, int k, bool sqrt)
{
  raft::distance::fusedL2NNMinReduce<float, raft::KeyValuePair<int, float>, int>(min, x, y, nullptr, nullptr, m, n, k,
                                                                                  nullptr, sqrt, true, handle);
}
}  // namespace raft::runtime::distance
