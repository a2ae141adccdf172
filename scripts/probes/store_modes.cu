// Probe: ceiling of different store mechanisms for the 128x256 tile pattern (no compute).
// modes: 0 TMA tensor store 32x32 box from smem; 1 LSU st.v4 4 rows x 128 B per instr;
//        2 LSU st.v4 thread==row (32 rows x 16 B per instr); 3 LSU st.v2 sectors (8 rows x 32 B per instr)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm128, float* out, long n, long m, int mode)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const long tiles_n = n / 256, tiles_m = m / 128, total = tiles_n * tiles_m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, g = warp >> 2;
  float* stg = reinterpret_cast<float*>(smem) + warp * 1024;
  for (int i = lane; i < 1024; i += 32) stg[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    const long tn = t % tiles_n, tm_ = t / tiles_n;
    const long row0 = tm_ * 128 + q * 32, col0 = tn * 256 + g * 128;
    for (int chunk = 0; chunk < 4; ++chunk) {
      const long c0 = col0 + chunk * 32;
      if (mode == 0) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm),
                       "r"(smem_u32(stg)), "r"((int)c0), "r"((int)row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      } else if (mode == 4) {   // 4 warps, 64-row boxes (8 KB per store)
        if (warp < 4 && lane == 0) {
          const long r64 = tm_ * 128 + (warp & 1) * 64, c64 = tn * 256 + (warp >> 1) * 128 + chunk * 32;
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm64),
                       "r"(smem_u32(smem + warp * 8192)), "r"((int)c64), "r"((int)r64) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      } else if (mode == 5) {   // 2 warps, 128-row boxes (16 KB per store)
        if (warp < 2 && lane == 0) {
          const long r128 = tm_ * 128, c128 = tn * 256 + warp * 128 + chunk * 32;
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm128),
                       "r"(smem_u32(smem + warp * 16384)), "r"((int)c128), "r"((int)r128) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      } else if (mode == 6) {   // 8 warps, two 4 KB stores in flight per warp (double buffer)
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm),
                       "r"(smem_u32(stg)), "r"((int)c0), "r"((int)row0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      } else if (mode == 1) {
        for (int it = 0; it < 8; ++it) {
          float* p = out + (row0 + it * 4 + (lane >> 3)) * n + c0 + (lane & 7) * 4;
          asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"((float)t) : "memory");
        }
      } else if (mode == 2) {
        for (int it = 0; it < 8; ++it) {
          float* p = out + (row0 + lane) * n + c0 + it * 4;
          asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"((float)t) : "memory");
        }
      } else {
        for (int rh = 0; rh < 2; ++rh)
          for (int it = 0; it < 4; ++it) {
            float* p0 = out + (row0 + rh * 16 + (lane >> 2)) * n + c0 + it * 8 + (lane & 3) * 2;
            asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p0), "f"(1.f), "f"((float)t) : "memory");
            asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p0 + 8 * n), "f"(1.f), "f"((float)t) : "memory");
          }
      }
    }
  }
  if ((mode == 0 || mode >= 4) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
  const long m = 100352, n = 100352;
  float* out; cudaMalloc(&out, m * n * 4);
  void* fp; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m}, str[1] = {(cuuint64_t)n * 4}; cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
  ((Enc)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUtensorMap tm64, tm128; cuuint32_t box64[2] = {32, 64}, box128[2] = {32, 128};
  ((Enc)fp)(&tm64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, str, box64, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ((Enc)fp)(&tm128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, str, box128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int mode = 0; mode < 7; ++mode) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148, 256, 32768>>>(tm, tm64, tm128, out, n, m, mode);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) k<<<148, 256, 32768>>>(tm, tm64, tm128, out, n, m, mode);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("mode %d: %.3f ms  %.0f GB/s  (%s)\n", mode, ms, m * n * 4.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
