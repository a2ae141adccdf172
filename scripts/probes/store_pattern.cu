// Probe: achievable HBM write bandwidth for 2-D tiled store patterns (no compute).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern store_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) tiled_store(float* out, long n, long m, int TM, int TN, int order)
{
  const long tiles_n = n / TN, tiles_m = m / TM, total = tiles_n * tiles_m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long t = blockIdx.x; t < total; t += gridDim.x) {
    long tm, tn;
    if (order == 0) { tn = t % tiles_n; tm = t / tiles_n; } else { tm = t % tiles_m; tn = t / tiles_m; }
    // 8 warps: each takes rows [warp*TM/8, +TM/8); lanes cover 128 B row segments (8 lanes x 16 B), 4 rows/instr
    const int rows_per_warp = TM / 8;
    for (int c0 = 0; c0 < TN; c0 += 32) {
      for (int r = lane >> 3; r < rows_per_warp; r += 4) {
        float* p = out + (tm * TM + warp * rows_per_warp + r) * n + tn * TN + c0 + (lane & 7) * 4;
        float4 v = make_float4(1.f, 2.f, 3.f, (float)t);
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
    }
  }
}
int main()
{
  const long m = 100352, n = 100352;  // multiples of 1024
  float* out; cudaMalloc(&out, m * n * 4);
  int cfg[][3] = {{128, 256, 0}, {128, 256, 1}, {64, 512, 0}, {32, 1024, 0}, {16, 2048, 0}, {8, 4096, 0}, {256, 128, 0}, {128, 1024, 0}, {128, 512, 0}};
  for (auto& c : cfg) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    tiled_store<<<148, 256>>>(out, n, m, c[0], c[1], c[2]);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) tiled_store<<<148, 256>>>(out, n, m, c[0], c[1], c[2]);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("tile %4dx%4d order %d: %.3f ms  %.0f GB/s\n", c[0], c[1], c[2], ms, m * n * 4.0 / ms / 1e6);
  }
  // multiple CTAs per SM
  for (int g : {296, 592}) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) tiled_store<<<g, 256>>>(out, n, m, 128, 256, 0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("tile 128x256 grid %d: %.3f ms  %.0f GB/s\n", g, ms, m * n * 4.0 / ms / 1e6);
  }
  return 0;
}
