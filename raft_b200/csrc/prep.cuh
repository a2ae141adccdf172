// Operand preparation for the tensor-core (expanded-metric) path.
//
// One pass per input matrix, O(rows*k) work (<1% of the pair loop).  For every row it
//   * optionally centres the row (CorrelationExpanded == cosine of centred rows),
//   * computes the squared L2 norm in fp64 (the numerical spec of the reference's row-norm
//     prologue is a compensated fp32 sum: cpp/include/raft/linalg/detail/
//     coalesced_reduction-inl.cuh:34-44; an fp64 accumulator rounds to the same fp32 value),
//   * picks a power-of-two row scale s so that max|x|*s is in [2^14, 2^15) and splits the scaled
//     row into fp16 hi + fp16 lo (x*s = hi + lo + O(2^-22 |x*s|)), written in the packed layout
//     the MMA kernel's TMA descriptor expects:
//         row r, k-block b (32 source columns):  [ hi[0..32) | lo[0..32) ]   = 128 bytes
//     so that one 128-byte SWIZZLE_128B shared-memory row carries both halves of a k-block and
//     hi*hi, hi*lo and lo*hi are three K=16 x2 tcgen05.mma steps over the same tile.
//   * emits the per-row epilogue pair used by   d = (acc * a.x) * b.x + (a.y + b.y):
//         L2       x:(1/s, |x|^2)        y:(-2/s, |y|^2)
//         cosine   x:(1/(s|x|), 1)       y:(-1/(s|y|), 0)      (also correlation, on centred rows)
//         inner    x:(1/s, 0)            y:(1/s, 0)
// k need not be a multiple of 32 and the source may have any row/column stride (row- or
// column-major, ld != k): the packed copy is what makes the TMA legal for every input, which is
// the job the reference's tile loader did with zero-fill and Veclen fallback
// (cpp/include/raft/linalg/detail/contractions.cuh:186-193).
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace b2d {

enum PrepMode : int { PREP_L2 = 0, PREP_COSINE = 1, PREP_INNER = 2 };

struct PrepParams {
  const void* src;     // float or half
  int64_t rs, cs;      // element strides of src: element (r,t) at src[r*rs + t*cs]
  int64_t rows;
  int k;
  int nkb;             // ceil(k/32)
  __half* op;          // [rows][nkb][64]
  float2* vec;         // [rows]
  const float* ext_norm_sq;  // optional caller-provided squared norms (fusedL2NN xn/yn)
  int mode;            // PrepMode
  int side;            // 0 = x (rows of the output), 1 = y (columns)
  int center;          // subtract the row mean first (correlation)
};

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(__ldg(p)); }

template <typename T>
__global__ void __launch_bounds__(256) prep_rows_kernel(PrepParams p)
{
  const int lane   = threadIdx.x & 31;
  const int64_t r  = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  const T* row = static_cast<const T*>(p.src) + r * p.rs;

  float mean = 0.f;
  if (p.center) {
    double s = 0.0;
    for (int t = lane; t < p.k; t += 32) s += static_cast<double>(ld_as_float(row + t * p.cs));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = static_cast<float>(s / static_cast<double>(p.k));
  }
  float amax = 0.f;
  double ss  = 0.0;
  for (int t = lane; t < p.k; t += 32) {
    float v = ld_as_float(row + t * p.cs) - mean;
    amax    = fmaxf(amax, fabsf(v));
    ss += static_cast<double>(v) * static_cast<double>(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  int e = 0;
  if (amax > 0.f && amax < 3.0e38f) {
    int q;
    frexpf(amax, &q);  // amax = f * 2^q, f in [0.5,1)
    e = 15 - q;        // amax * 2^e in [2^14, 2^15)
    e = max(-100, min(100, e));
  }
  const float scale = ldexpf(1.f, e);
  const float inv   = ldexpf(1.f, -e);

  __half* orow = p.op + r * static_cast<int64_t>(p.nkb) * 64;
  const int kpad = p.nkb * 32;
  for (int t = lane; t < kpad; t += 32) {
    float xs = 0.f;
    if (t < p.k) xs = (ld_as_float(row + t * p.cs) - mean) * scale;
    __half h = __float2half_rn(xs);
    __half l = __float2half_rn(xs - __half2float(h));
    const int b = t >> 5, j = t & 31;
    orow[b * 64 + j]      = h;
    orow[b * 64 + 32 + j] = l;
  }
  if (lane == 0) {
    float nsq = p.ext_norm_sq ? p.ext_norm_sq[r] : static_cast<float>(ss);
    float2 v;
    if (p.mode == PREP_L2) {
      v.x = p.side == 0 ? inv : -2.f * inv;
      v.y = nsq;
    } else if (p.mode == PREP_COSINE) {
      float nrm = p.ext_norm_sq ? sqrtf(nsq) : static_cast<float>(sqrt(ss));
      v.x       = p.side == 0 ? inv / nrm : -inv / nrm;
      v.y       = p.side == 0 ? 1.f : 0.f;
    } else {
      v.x = inv;
      v.y = 0.f;
    }
    p.vec[r] = v;
  }
}

}  // namespace b2d
