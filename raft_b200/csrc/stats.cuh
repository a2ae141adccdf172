// Callers of the distance path: silhouette score on top of the pairwise engine.
//
// Semantics follow raft::stats::silhouette_score (cpp/include/raft/stats/detail/silhouette_score.cuh:
// 186-328): a(i) = mean distance to the other members of i's cluster (-1 marks a singleton cluster ->
// score 0), b(i) = min over the other non-empty clusters of the mean distance to that cluster,
// s(i) = 0 if a == b, (b - a) / max(a, b) otherwise; result = mean of s.  The reference materialises
// the n x n matrix (its batched variant, detail/batched/silhouette_score.cuh:213-243, works in chunks);
// here the database side is a copy of X with rows grouped by label, so the per-cluster sums of a row
// are sums over contiguous column segments of a [chunk x n] slab that is produced by the engine and
// consumed straight away.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace b2d {

__global__ void sil_count_kernel(const int* labels, int* counts, int64_t n, int n_labels, unsigned* bad)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = labels[i];
  if (l < 0 || l >= n_labels) { *bad = 1u; return; }
  atomicAdd(&counts[l], 1);
}

// exclusive scan of the label counts (n_labels is small next to n: one block)
__global__ void sil_scan_kernel(const int* counts, int* offsets, int* cursor, int n_labels)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int l = 0; l < n_labels; ++l) { offsets[l] = run; cursor[l] = run; run += counts[l]; }
    offsets[n_labels] = run;
  }
}

// y_sorted[pos] = x[i] with pos inside the segment of labels[i] (order inside a segment is irrelevant)
__global__ void sil_gather_kernel(const float* x, int64_t ldx, const int* labels, int* cursor, float* ys, int* where,
                                  int64_t n, int k, int n_labels)
{
  const int lane    = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int l = labels[row];
  if (l < 0 || l >= n_labels) return;
  int pos = 0;
  if (lane == 0) pos = atomicAdd(&cursor[l], 1);
  pos = __shfl_sync(0xffffffffu, pos, 0);
  if (lane == 0) where[row] = pos;
  for (int t = lane; t < k; t += 32) ys[static_cast<int64_t>(pos) * k + t] = x[row * ldx + t];
}

// one block per row of the slab: per-cluster sums over the column segments, then a, b, s
__global__ void __launch_bounds__(256) sil_row_kernel(const float* slab, int64_t ld, int64_t row0, int64_t rows,
                                                      const int* labels, const int* counts, const int* offsets,
                                                      const int* where, int n_labels, float* per_sample, double* total)
{
  extern __shared__ float sums[];  // [n_labels]
  const int64_t r = blockIdx.x;
  if (r >= rows) return;
  const float* d = slab + r * ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int l = threadIdx.x; l < n_labels; l += blockDim.x) sums[l] = 0.f;
  __syncthreads();
  for (int l = 0; l < n_labels; ++l) {
    // every warp takes an equal slice of every segment: balanced for 2 clusters as for 2000
    const int64_t a = offsets[l], b = offsets[l + 1], len = b - a;
    const int64_t lo = a + len * warp / nw, hi = a + len * (warp + 1) / nw;
    float acc = 0.f;
    for (int64_t j = lo + lane; j < hi; j += 32) acc += d[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && hi > lo) atomicAdd(&sums[l], acc);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int own = labels[row0 + r];
    float s = 0.f;
    if (counts[own] > 1) {
      // the sample's distance to itself is 0 by definition (the reference gets that from x == y aliasing;
      // here the column side is a permuted copy, so whatever rounding left in d(i, i) is taken out)
      const float av = (sums[own] - d[where[row0 + r]]) / static_cast<float>(counts[own] - 1);
      float bv       = 3.402823466e38f;
      for (int l = 0; l < n_labels; ++l)
        if (l != own && counts[l] > 0) bv = fminf(bv, sums[l] / static_cast<float>(counts[l]));
      if (!((av == 0.f && bv == 0.f) || av == bv)) s = av > bv ? (bv - av) / av : (bv - av) / bv;
    }
    if (per_sample) per_sample[row0 + r] = s;
    atomicAdd(total, static_cast<double>(s));
  }
}

__global__ void sil_finish_kernel(const double* total, float* score, int64_t n)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) *score = static_cast<float>(*total / static_cast<double>(n));
}

}  // namespace b2d
