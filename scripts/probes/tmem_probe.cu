// Probe: (A) TMEM -> register read throughput per SM for the tcgen05.ld shapes the epilogues use,
//        (B) layout and rounding of an fp16 accumulator (kind::f16 with D = f16).
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I raft_b200/csrc -o scripts/probes/tmem_probe scripts/probes/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace b2d::ptx;

__device__ __forceinline__ void ld_pack16_x32(uint32_t taddr, uint32_t (&r)[32])
{
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
      "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
      "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
    : "r"(taddr)
    : "memory");
}

// mode 0: 32x32b.x32 (4 KB/instr)  1: 16x256b.x8 (4 KB/instr)  2: 32x32b.pack::16b.x32 (64 columns -> 32 regs)
__global__ void __launch_bounds__(512, 1) bw_kernel(int mode, int iters, int cols_per_iter, long long* cycles, uint32_t* sink)
{
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tbase;
  const int q = warp & 3, g = warp >> 2, G = nwarps >> 2;
  const int cols_per_group = cols_per_iter / G;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int stage = (it & 1) * 256;
    if (mode == 0) {
      for (int c = 0; c < cols_per_group; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(base + (uint32_t(q * 32) << 16) + stage + g * cols_per_group + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    } else if (mode == 1) {
      for (int c = 0; c < cols_per_group; c += 32) {   // 32 lanes x 32 columns as two 16-lane x 64-col? keep bytes equal: 16 lanes x 64 cols
        uint32_t r[32];
        const int cc = (c / 64) * 64, rh = (c / 32) & 1;
        tmem_ld_16x256_x8(base + (uint32_t(q * 32 + rh * 16) << 16) + stage + g * cols_per_group + cc, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    } else {
      for (int c = 0; c < cols_per_group; c += 64) {
        uint32_t r[32];
        ld_pack16_x32(base + (uint32_t(q * 32) << 16) + stage + g * cols_per_group + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(base);
}

// (B) one 128x256xK16 MMA with an fp16 accumulator, then dump TMEM raw.
__global__ void __launch_bounds__(128, 1) f16acc_kernel(uint32_t* dump, uint32_t* dump_pack, int test)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tbase;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* sA = smem;            // 128 rows x 128 B
  uint8_t* sB = smem + 16384;    // 256 rows x 128 B
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  auto put = [&](uint8_t* base, int r, int kk, float v) {
    const int c = kk >> 3;
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) * 16) + (kk & 7) * 2;
    *reinterpret_cast<__half*>(base + off) = __float2half_rn(v);
  };
  // pass 0 operands
  if (test == 0) {
    put(sA, tid, 0, float(tid % 64 + 1) * 0.25f);
    put(sB, tid, 0, float(tid % 32 + 1));
    put(sB, tid + 128, 0, float((tid + 128) % 32 + 1));
  } else {  // rounding: D = 2048 (k=0) ; second MMA adds 3 (test 1) ; or single MMA 2048+3 (test 2); test 3: 2048 + 1 + 1 + 1 within one MMA
    put(sA, tid, 0, 1.f);
    put(sB, tid, 0, 2048.f);
    put(sB, tid + 128, 0, 2048.f);
    if (test == 2) { put(sA, tid, 1, 1.f); put(sB, tid, 1, 3.f); put(sB, tid + 128, 1, 3.f); }
    if (test == 3) {
      for (int kk = 1; kk <= 3; ++kk) { put(sA, tid, kk, 1.f); put(sB, tid, kk, 1.f); put(sB, tid + 128, kk, 1.f); }
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tbase;
  const uint32_t idesc = (0u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);  // D = f16
  if (warp == 0) {
    if (elect_one()) {
      mma_f16_ss(base, umma_desc_sw128(smem_u32(sA)), umma_desc_sw128(smem_u32(sB)), idesc, 0);
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (test == 1) {
    __syncthreads();
    // second MMA: A = 1, B = 3 -> +3
    put(sB, tid, 0, 3.f);
    put(sB, tid + 128, 0, 3.f);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        mma_f16_ss(base, umma_desc_sw128(smem_u32(sA)), umma_desc_sw128(smem_u32(sB)), idesc, 1);
        mma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, 1);
    tc_fence_after();
  }
  for (int c = 0; c < 256; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(base + (uint32_t(warp * 32) << 16) + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) dump[tid * 256 + c + i] = r[i];
  }
  for (int c = 0; c < 256; c += 64) {
    uint32_t r[32];
    ld_pack16_x32(base + (uint32_t(warp * 32) << 16) + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) dump_pack[tid * 128 + c / 2 + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(base);
}

static float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }

int main()
{
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      bw_kernel<<<148, warps * 32>>>(mode, iters, 256, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("bw mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
      printf("bw mode %d warps %2d: %.1f cycles per 256-column pass (32-bit cols; mode 2 = packed)\n", mode, warps, double(mx) / iters);
    }
  uint32_t *dump, *dump_pack;
  cudaMalloc(&dump, 128 * 256 * 4); cudaMalloc(&dump_pack, 128 * 128 * 4);
  cudaFuncSetAttribute(f16acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
  for (int test = 0; test < 4; ++test) {
    cudaMemset(dump, 0xff, 128 * 256 * 4);
    f16acc_kernel<<<1, 128, 16384 + 32768 + 1024>>>(dump, dump_pack, test);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("f16acc test %d: %s\n", test, cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> h(128 * 256), hp(128 * 128);
    cudaMemcpy(h.data(), dump, h.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hp.data(), dump_pack, hp.size() * 4, cudaMemcpyDeviceToHost);
    for (int row : {0, 5, 77}) {
      printf("test %d row %3d raw:", test, row);
      for (int c = 0; c < 8; ++c) printf(" %08x", h[row * 256 + c]);
      printf(" | c128: %08x c255: %08x\n", h[row * 256 + 128], h[row * 256 + 255]);
      printf("            as f16 lo/hi:");
      for (int c = 0; c < 6; ++c) printf(" (%g,%g)", h2f(h[row * 256 + c] & 0xffff), h2f(h[row * 256 + c] >> 16));
      printf("\n            packed ld  :");
      for (int c = 0; c < 6; ++c) printf(" (%g,%g)", h2f(hp[row * 128 + c] & 0xffff), h2f(hp[row * 128 + c] >> 16));
      printf("\n");
    }
  }
  return 0;
}
