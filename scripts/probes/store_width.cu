// Probe: does the width of the contiguous row segment per store request decide the HBM write
// bandwidth of the 128x256-tile output pattern?  (r01: every mechanism tried wrote 128-byte row
// segments and topped out at ~20 B/clk/SM; a linear fill reaches 25+.)
//   family L  LSU st.global.cs.v4, one warp instruction = R rows x SEG bytes (SEG = 128, 256, 512)
//   family B  1-D bulk copies shared -> global (cp.async.bulk.global.shared::cta), one op = one row
//             segment of SEG bytes (128 .. 1024), `depth` commit groups in flight per warp
//   family T  TMA tensor stores, box {C cols x R rows} = 4 KB .. 16 KB, SWIZZLE_128B for C = 32,
//             no swizzle for wider boxes, `depth` groups in flight per warp
// All walk the output like expanded_tc_kernel: item = (256-column block, run of 32 row tiles),
// consecutive CTAs on consecutive column blocks.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_width store_width.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Maps { CUtensorMap t[8]; };

__device__ __forceinline__ void wait_read(int depth)
{
  if (depth <= 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else if (depth == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
}

// family: 0 = L, 1 = B, 2 = T.  a = SEG bytes (L, B) or map index (T); bc/br = box cols/rows (T)
__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ Maps maps, float* out, long n, long m, int family,
                                            int a, int bc, int br, int depth, int order, int evict_first)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tw = (family == 5 && a == 768) ? 192 : 256;
  const int tiles_n = (int)(n / tw), tiles_m = (int)(m / 128);
  const int chunk = 32, chunks_m = (tiles_m + chunk - 1) / chunk;
  const long items = (long)tiles_n * chunks_m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384; i += 256) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (long item = blockIdx.x; item < items; item += gridDim.x) {
    int tn, ch;
    if (order == 0) { tn = (int)(item % tiles_n); ch = (int)(item / tiles_n); }
    else { ch = (int)(item % chunks_m); tn = (int)(item / chunks_m); }
    const int mt1 = min((ch + 1) * chunk, tiles_m);
    for (int mt = ch * chunk; mt < mt1; ++mt) {
      const long trow = (long)mt * 128, tcol = (long)tn * 256;
      if (family == 0) {
        // warp w: rows [16w, 16w+16) of the tile, all 256 columns; one instruction = R rows x SEG bytes
        const int lanes_per_row = a / 16, rows_per_instr = 32 / lanes_per_row;
        for (int r0 = 0; r0 < 16; r0 += rows_per_instr)
          for (int c0 = 0; c0 < 1024; c0 += a) {
            float* p = out + (trow + warp * 16 + r0 + lane / lanes_per_row) * n + tcol + (c0 + (lane % lanes_per_row) * 16) / 4;
            asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"((float)mt) : "memory");
          }
      } else if (family == 1) {
        // warp w: rows [16w, 16w+16); per step the 32 lanes issue 32 row segments of SEG bytes
        const int segs_per_row = 1024 / a;              // 8, 4, 2, 1
        const int steps = 16 * segs_per_row / 32;       // 4, 2, 1, (1 with half the lanes)
        for (int s = 0; s < (steps ? steps : 1); ++s) {
          const int idx = s * 32 + lane;
          const bool on = idx < 16 * segs_per_row;
          const int r = idx / segs_per_row, sg = idx % segs_per_row;
          wait_read(depth);
          if (on) {
            float* p = out + (trow + warp * 16 + r) * n + tcol + sg * (a / 4);
            const uint32_t src = smem_u32(smem + warp * 8192 + ((idx * a) & 8191));
            if (evict_first)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p), "r"(src), "r"(a), "l"(pol) : "memory");
            else
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p), "r"(src), "r"(a) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (family == 5) {
        // pair pattern: warps (2p, 2p+1) own rows [32p, 32p+32) of the tile; per round the EVEN warp of the pair sends 16 rows x
        // (a / 256) segments of 256 B from ONE elected lane (warp-uniform loop: no per-lane waterfall); a = row bytes
        // (1024: 128x256 tile, 768: 128x192 tile -- then tcol/tiles use 192 columns, see tw)
        if ((warp & 1) == 0) {
          const int p = warp >> 1, segs = a / 256;
          for (int round = 0; round < 2; ++round) {
            wait_read(depth);
            if (lane == 0) {
              for (int r = 0; r < 16; ++r)
                for (int sg = 0; sg < segs; ++sg) {
                  float* pp = out + (trow + p * 32 + round * 16 + r) * n + (long)tn * (a / 4) + sg * 64;
                  const uint32_t src = smem_u32(smem + p * 16384 + r * (a + 32) - (r ? 0 : 0) + sg * 256);
                  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(pp), "r"(src), "r"(256), "l"(pol) : "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if (family == 3) {
        // D1: warp w owns rows [16w, 16w+16); 4 steps of 64 columns; a = 0: lanes < 16 issue one 256-byte row segment each;
        // a = 1: 8 rows x 2 adjacent 256-byte ops per step (rows [8(s&1).., cols 128(s>>1)..)); a = 2: all 32 lanes, 128-byte halves
        for (int s = 0; s < 4; ++s) {
          wait_read(depth);
          if (a == 0) {
            if (lane < 16) {
              float* p = out + (trow + warp * 16 + lane) * n + tcol + s * 64;
              const uint32_t src = smem_u32(smem + warp * 4352 + lane * 272);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p), "r"(src), "r"(256), "l"(pol) : "memory");
            }
          } else if (a == 1) {
            if (lane < 16) {
              float* p = out + (trow + warp * 16 + (s & 1) * 8 + (lane >> 1)) * n + tcol + (s >> 1) * 128 + (lane & 1) * 64;
              const uint32_t src = smem_u32(smem + warp * 4352 + lane * 272);
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p), "r"(src), "r"(256), "l"(pol) : "memory");
            }
          } else {
            float* p = out + (trow + warp * 16 + (lane >> 1)) * n + tcol + s * 64 + (lane & 1) * 32;
            const uint32_t src = smem_u32(smem + warp * 4352 + (lane >> 1) * 272 + (lane & 1) * 128);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(p), "r"(src), "r"(128), "l"(pol) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (family == 4) {
        // T box 64x16 arranged like D1 (warp w rows [16w,+16), step s columns [64s,+64))
        for (int s = 0; s < 4; ++s) {
          if (lane == 0) {
            wait_read(depth);
            const uint32_t src = smem_u32(smem + warp * 4096);
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"((uint64_t)&maps.t[1]),
                         "r"(src), "r"((int)(tcol + s * 64)), "r"((int)(trow + warp * 16)), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        }
      } else {
        // boxes of bc x br floats; the tile holds (256/bc) x (128/br) of them, dealt round-robin to the warps
        const int bpr = 256 / bc, nbox = bpr * (128 / br);
        for (int b = warp; b < nbox; b += 8) {
          const int bx = b % bpr, by = b / bpr;
          if (lane == 0) {
            wait_read(depth);
            const uint32_t src = smem_u32(smem + (warp & 3) * 16384);
            if (evict_first)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"((uint64_t)&maps.t[a]),
                           "r"(src), "r"((int)(tcol + bx * bc)), "r"((int)(trow + by * br)), "l"(pol) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&maps.t[a]),
                           "r"(src), "r"((int)(tcol + bx * bc)), "r"((int)(trow + by * br)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
        }
      }
    }
  }
  if (family != 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
  const long m = 100352, n = 100352;
  float* out;
  if (cudaMalloc(&out, m * n * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  void* fp; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Maps maps;
  const int boxes[8][2] = {{32, 32}, {64, 16}, {128, 8}, {256, 4}, {64, 32}, {128, 16}, {256, 8}, {256, 16}};
  for (int i = 0; i < 8; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m}, str[1] = {(cuuint64_t)n * 4};
    cuuint32_t box[2] = {(cuuint32_t)boxes[i][0], (cuuint32_t)boxes[i][1]}, es[2] = {1, 1};
    CUresult r = ((Enc)fp)(&maps.t[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           i == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode %d failed: %d\n", i, (int)r);
  }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920);
  struct Cfg { int family, a, bc, br, depth, order, ef; const char* name; };
  const Cfg cfgs[] = {
    {5, 1024, 0, 0, 1, 0, 1, "P pair 16 rows x 4x256 B, one lane issues, depth 1"},
    {5, 1024, 0, 0, 2, 0, 1, "P pair 16 rows x 4x256 B, one lane issues, depth 2"},
    {5, 768, 0, 0, 1, 0, 1, "P pair 16 rows x 3x256 B (128x192 tile), depth 1"},
    {5, 768, 0, 0, 2, 0, 1, "P pair 16 rows x 3x256 B (128x192 tile), depth 2"},
    {3, 0, 0, 0, 1, 0, 1, "D1 16 rows x 256 B per step (16 lanes), depth 1"},
    {3, 0, 0, 0, 2, 0, 1, "D1 16 rows x 256 B per step (16 lanes), depth 2"},
    {3, 1, 0, 0, 1, 0, 1, "D1c 8 rows x 2x256 B per step, depth 1"},
    {3, 2, 0, 0, 1, 0, 1, "D1h 16 rows x 2x128 B per step (32 lanes), depth 1"},
    {4, 0, 0, 0, 1, 0, 1, "T box 64x16 laid out like D1, depth 1"},
    {1, 256, 0, 0, 1, 0, 1, "B bulk 256 B segs depth 1 (again)"},
    {3, 0, 0, 0, 1, 0, 1, "D1 again"},
    {0, 128, 0, 0, 0, 0, 0, "L st.v4 4 rows x 128 B"},
    {0, 256, 0, 0, 0, 0, 0, "L st.v4 2 rows x 256 B"},
    {0, 512, 0, 0, 0, 0, 0, "L st.v4 1 row  x 512 B"},
    {0, 512, 0, 0, 0, 1, 0, "L st.v4 1 row  x 512 B, row-chunk-major items"},
    {1, 128, 0, 0, 1, 0, 1, "B bulk 128 B segs depth 1"},
    {1, 256, 0, 0, 1, 0, 1, "B bulk 256 B segs depth 1"},
    {1, 512, 0, 0, 1, 0, 1, "B bulk 512 B segs depth 1"},
    {1, 1024, 0, 0, 1, 0, 1, "B bulk 1024 B segs depth 1"},
    {1, 128, 0, 0, 2, 0, 1, "B bulk 128 B segs depth 2"},
    {1, 512, 0, 0, 2, 0, 1, "B bulk 512 B segs depth 2"},
    {1, 1024, 0, 0, 2, 0, 1, "B bulk 1024 B segs depth 2"},
    {1, 1024, 0, 0, 4, 0, 1, "B bulk 1024 B segs depth 4"},
    {1, 1024, 0, 0, 2, 0, 0, "B bulk 1024 B segs depth 2, no L2 hint"},
    {2, 0, 32, 32, 1, 0, 1, "T box 32x32 sw128 depth 1 (r01 kernel)"},
    {2, 0, 32, 32, 2, 0, 1, "T box 32x32 sw128 depth 2"},
    {2, 0, 32, 32, 4, 0, 1, "T box 32x32 sw128 depth 4"},
    {2, 1, 64, 16, 1, 0, 1, "T box 64x16 depth 1"},
    {2, 2, 128, 8, 1, 0, 1, "T box 128x8 depth 1"},
    {2, 3, 256, 4, 1, 0, 1, "T box 256x4 depth 1"},
    {2, 3, 256, 4, 2, 0, 1, "T box 256x4 depth 2"},
    {2, 3, 256, 4, 4, 0, 1, "T box 256x4 depth 4"},
    {2, 4, 64, 32, 1, 0, 1, "T box 64x32 (8 KB) depth 1"},
    {2, 5, 128, 16, 1, 0, 1, "T box 128x16 (8 KB) depth 1"},
    {2, 6, 256, 8, 1, 0, 1, "T box 256x8 (8 KB) depth 1"},
    {2, 6, 256, 8, 2, 0, 1, "T box 256x8 (8 KB) depth 2"},
    {2, 7, 256, 16, 1, 0, 1, "T box 256x16 (16 KB) depth 1"},
    {2, 3, 256, 4, 2, 0, 0, "T box 256x4 depth 2, no L2 hint"},
    {2, 3, 256, 4, 2, 1, 1, "T box 256x4 depth 2, row-chunk-major items"},
  };
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
  for (const Cfg& c : cfgs) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 256, 81920>>>(maps, out, n, m, c.family, c.a, c.bc, c.br, c.depth, c.order, c.ef);
    cudaEventRecord(e0);
    for (int i = 0; i < 4; ++i) k<<<148, 256, 81920>>>(maps, out, n, m, c.family, c.a, c.bc, c.br, c.depth, c.order, c.ef);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 4;
    printf("%-52s %7.3f ms  %6.0f GB/s  %s\n", c.name, ms, m * n * 4.0 / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
    if (err != cudaSuccess) return 1;
  }
  // reference: linear fill
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaMemsetAsync(out, 0, m * n * 4);
  cudaEventRecord(e0);
  for (int i = 0; i < 4; ++i) cudaMemsetAsync(out, 1, m * n * 4);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 4;
  printf("%-52s %7.3f ms  %6.0f GB/s\n", "cudaMemset (linear)", ms, m * n * 4.0 / ms / 1e6);
  return 0;
}
