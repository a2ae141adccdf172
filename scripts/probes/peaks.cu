// Builder-measured pipe peaks for the roofline denominators that MEASURED_PEAKS.json does not hold (SURVEY.md 8(d):
// "tensor peak via a synthetic back-to-back tcgen05.mma loop per kind; FP32 pipe = SMs x 128 lanes x sustained clock"):
//   1. tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, issued back to back by one thread per SM on operands that sit
//      in shared memory (SWIZZLE_128B descriptors), accumulating into TMEM -- no loads, no epilogue: the tensor pipe alone;
//   2. the same with N = 128 and with kind::tf32 (K = 8);
//   3. FP32 pipe: 8 independent FFMA chains per thread, 1024 threads per SM, and the packed FFMA2 form.
// Each test runs ~0.5 s so that the power cap acts (sustained figure) and prints the first-50-ms burst figure too.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../raft_b200/csrc -o peaks peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b2d;

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
               "l"(da), "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
// kind::tf32 instruction descriptor: D = f32 (1 << 4), A = B = tf32 (format 2 at bits [7,10) and [10,13))
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

template <int kKind, int kN, int kData>   // kKind 0: f16, 1: tf32; kData 1: random operands, 0: constant
__global__ void __launch_bounds__(128, 1) mma_peak(int iters)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // operands: pseudo-random finite values (constant operands barely toggle the datapath: the first version of this probe
  // ran at 2235 TF/s WITHOUT ever reaching the power cap); kData = 0 keeps the constant fill for comparison
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // two fp16 in [-2, 2): sign + exponent 0x3c00 +- 1 + random mantissa  (as tf32 bits: a finite value of modest size as well)
    const uint32_t lo = (h & 0x83ffu) | 0x3c00u, hi = ((h >> 16) & 0x83ffu) | 0x3800u;
    reinterpret_cast<uint32_t*>(smem)[i] = kData ? (lo | (hi << 16)) : 0x3c003c00u;
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(smem));
    const uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(smem + 16384));
    constexpr uint32_t idf = ptx::umma_idesc_f16(128, kN), idt = idesc_tf32(128, kN);
    for (int it = 0; it < iters; ++it) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // 8 MMAs alternate between two accumulators (512 columns in use for N = 256)
          if (kKind == 0) ptx::mma_f16_ss(tmem + (u & 1) * 256, da + 2 * (u & 3), db + 2 * (u & 3), idf, 1u);
          else mma_tf32_ss(tmem + (u & 1) * 256, da + 2 * (u & 3), db + 2 * (u & 3), idt, 1u);
        }
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::mma_commit(&bar);
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem); }
}

template <bool kPacked>
__global__ void __launch_bounds__(1024, 1) ffma_peak(float* out, int iters, float a, float b)
{
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-6f + i;
  if (!kPacked) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  } else {
    uint64_t v[4], a2, b2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(a2), "l"(b2));
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static void run(const char* name, double work_per_launch, const char* unit, F launch)
{
  cudaEvent_t e[4];
  for (auto& x : e) cudaEventCreate(&x);
  launch();   // warm-up
  cudaDeviceSynchronize();
  // burst: one launch from idle; sustained: the last of a ~0.6 s train
  cudaEventRecord(e[0]); launch(); cudaEventRecord(e[1]);
  float ms1 = 0; cudaEventSynchronize(e[1]); cudaEventElapsedTime(&ms1, e[0], e[1]);
  const int reps = (int)(600.0 / ms1) + 1;
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e[2]); launch(); cudaEventRecord(e[3]);
  float ms2 = 0; cudaEventSynchronize(e[3]); cudaEventElapsedTime(&ms2, e[2], e[3]);
  printf("%-46s burst %9.1f %s   sustained %9.1f %s   (%.2f / %.2f ms per launch, %s)\n", name, work_per_launch / ms1 / 1e9, unit,
         work_per_launch / ms2 / 1e9, unit, ms1, ms2, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, max SM clock %d MHz\n", prop.name, sms, prop.clockRate / 1000);
  const int iters = 20000;
  const size_t smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(mma_peak<0, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mma_peak<0, 256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mma_peak<0, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mma_peak<1, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  run("tcgen05.mma kind::f16 M128 N256 K16, random data", 1.0 * sms * (double)iters * 8 * 2.0 * 128 * 256 * 16, "TF/s",
      [&] { mma_peak<0, 256, 1><<<sms, 128, smem>>>(iters); });
  run("tcgen05.mma kind::f16 M128 N256 K16, constant data", 1.0 * sms * (double)iters * 8 * 2.0 * 128 * 256 * 16, "TF/s",
      [&] { mma_peak<0, 256, 0><<<sms, 128, smem>>>(iters); });
  run("tcgen05.mma kind::f16 M128 N128 K16, random data", 1.0 * sms * (double)iters * 8 * 2.0 * 128 * 128 * 16, "TF/s",
      [&] { mma_peak<0, 128, 1><<<sms, 128, smem>>>(iters); });
  run("tcgen05.mma kind::tf32 M128 N256 K8, random data", 1.0 * sms * (double)iters * 8 * 2.0 * 128 * 256 * 8, "TF/s",
      [&] { mma_peak<1, 256, 1><<<sms, 128, smem>>>(iters); });
  float* out; cudaMalloc(&out, 4);
  const int fi = 200000;
  run("FP32 pipe FFMA (lane-ops/s, 1 FFMA = 1 lane-op)", 1.0 * sms * 1024.0 * fi * 8, "Tlop/s",
      [&] { ffma_peak<false><<<sms, 1024>>>(out, fi, 1.000001f, 0.5f); });
  run("FP32 pipe FFMA2 (lane-ops/s, 1 FFMA2 = 2)", 1.0 * sms * 1024.0 * fi * 8, "Tlop/s",
      [&] { ffma_peak<true><<<sms, 1024>>>(out, fi, 1.000001f, 0.5f); });
  return 0;
}
