// Probe: how fast does ONE SM's TMA unit bring operand tiles from L2 into shared memory, as a function of the box shape?
// The coarse fusedL2NN pass loads 128-row x-tiles whose rows are 64-byte pieces (the fp16 hi halves) of 128-byte packed
// k-blocks: nkb boxes of {32 fp16, 128 rows}, SWIZZLE_64B, per tile.  Candidates: 128-byte rows (SWIZZLE_128B).
// Every SM loops over an L2-resident matrix (`rows` x `pitch` bytes) with `depth` boxes in flight; reports cycles per box
// and bytes per clock per SM, all SMs running (L2 -> SM fabric shared) and one SM alone.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../raft_b200/csrc -o tma_rate tma_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b2d;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows, int kboxes, int rows,
                                            int iters, int depth, unsigned long long* cyc)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(&full[i], 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const uint64_t pol = ptx::policy_evict_last();
  const int tiles = rows / box_rows;
  long long t0 = 0;
  // one "tile" = kboxes boxes (k-blocks) on one barrier, like the kernel's x stage
  for (int it = 0; it < iters + depth; ++it) {
    if (it == depth) t0 = clock64();
    const int s = it % depth;
    if (it >= depth) ptx::mbar_wait(&full[s], ((it / depth) - 1) & 1);
    if (it < iters) {
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(&full[s], (uint32_t)(box_bytes * kboxes));
        const int tile = (int)((blockIdx.x * 7919u + it) % tiles);
        for (int kb = 0; kb < kboxes; ++kb)
          ptx::tma_load_2d(smem + (size_t)s * box_bytes * kboxes + (size_t)kb * box_bytes, &map, &full[s], kb * 64, tile * box_rows, pol);
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  EncodeFn enc = (EncodeFn)fn;
  const int rows = 1 << 17;                 // 131072 rows x 512 B = 64 MB: L2-resident
  const int pitch = 512;                    // bytes per row: 4 packed k-blocks of 128 B ([hi32|lo32] fp16)
  uint8_t* buf; cudaMalloc(&buf, (size_t)rows * pitch); cudaMemset(buf, 1, (size_t)rows * pitch);
  unsigned long long* cyc; cudaMalloc(&cyc, sms * 8);
  struct Cfg { const char* name; int inner_elems; int box_rows; CUtensorMapSwizzle sw; int kboxes; int kstride_elems; };
  // inner_elems fp16 per box row; kstride: element offset between successive k boxes (64 = next packed k-block)
  Cfg cfgs[] = {
    {"64 B rows (hi half), 128 rows, SW64, 3 boxes/tile", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 3, 64},
    {"64 B rows (hi half), 128 rows, SW64, 4 boxes/tile", 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, 4, 64},
    {"128 B rows (whole k-block), 128 rows, SW128, 3 boxes/tile", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 3, 64},
    {"128 B rows, 128 rows, SW128, 2 boxes/tile (= 4 hi k-blocks, hi-only layout)", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 2, 64},
    {"64 B rows, 256 rows, SW64, 3 boxes/tile", 32, 256, CU_TENSOR_MAP_SWIZZLE_64B, 3, 64},
  };
  for (auto& c : cfgs) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)(pitch / 2), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)c.inner_elems, (cuuint32_t)c.box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int box_bytes = c.inner_elems * 2 * c.box_rows;
    for (int grid : {sms, 1}) {
      for (int depth : {2, 4, 6}) {
        const size_t smem = (size_t)depth * box_bytes * c.kboxes;
        if (smem > 200 * 1024) continue;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int iters = 20000;
        k<<<grid, 128, smem>>>(map, box_bytes, c.box_rows, c.kboxes, rows, 200, depth, cyc);   // warm-up (L2 fill)
        k<<<grid, 128, smem>>>(map, box_bytes, c.box_rows, c.kboxes, rows, iters, depth, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[256]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
        printf("%-78s grid %3d depth %d: %7.1f cyc/tile  %6.1f B/clk/SM  (%s)\n", c.name, grid, depth, mean / iters,
               (double)box_bytes * c.kboxes * iters / mean, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
