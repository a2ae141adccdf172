// Dependency-free test of b2d_fused_l2_nn_multi (single process, N devices, one ncclComm_t per device from
// ncclCommInitAll -- the reference's SNMG pattern, cpp/include/raft/core/resource/nccl_comm.hpp:43-62).
// usage: multi_test [ngpu]   (default: all visible devices, at most 8).  Prints PASS / SKIP / FAIL.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <nccl.h>
#include <raft_b200.h>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { std::printf("FAIL %s: %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

int main(int argc, char** argv)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::printf("SKIP no gpu\n"); return 0; }
  int G = argc > 1 ? std::atoi(argv[1]) : (ndev > 8 ? 8 : ndev);
  if (G > ndev) { std::printf("SKIP %d devices wanted, %d visible\n", G, ndev); return 0; }
  const int64_t m = 20000, n = 600000 + 5, k = 96;   // shards of ~75k..300k rows: head + sub-chunk exchanges, screened search
  std::vector<float> hx(m * k), hy(n * k);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  std::vector<float> cen(5 * k);
  for (auto& v : cen) v = rnd() * 20.f;
  for (int64_t i = 0; i < m; ++i) { const int c = (int)(((unsigned)i * 2654435761u) >> 29) % 5; for (int t = 0; t < k; ++t) hx[i * k + t] = cen[c * k + t] + 2.f * rnd(); }
  for (int64_t j = 0; j < n; ++j) { const int c = (int)(((unsigned)j * 40503u + 7u) >> 7) % 5; for (int t = 0; t < k; ++t) hy[j * k + t] = cen[c * k + t] + 2.f * rnd(); }

  // reference: the single-device entry point on device 0
  std::vector<b2d_kvp_if> ref(m);
  {
    CK(cudaSetDevice(0));
    float *x, *y; b2d_kvp_if* o; void* ws;
    const size_t need = b2d_fused_l2_nn_workspace_bytes(m, n, k);
    CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&y, hy.size() * 4)); CK(cudaMalloc(&o, m * sizeof(b2d_kvp_if))); CK(cudaMalloc(&ws, need));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(y, hy.data(), hy.size() * 4, cudaMemcpyHostToDevice));
    if (b2d_fused_l2_nn(nullptr, o, x, k, y, k, nullptr, nullptr, m, n, k, 0, 1, ws, need)) { std::printf("FAIL single: %s\n", b2d_last_error()); return 1; }
    CK(cudaMemcpy(ref.data(), o, m * sizeof(b2d_kvp_if), cudaMemcpyDeviceToHost));
    cudaFree(x); cudaFree(y); cudaFree(o); cudaFree(ws);
  }
  std::vector<int> devs(G);
  for (int g = 0; g < G; ++g) devs[g] = g;
  std::vector<ncclComm_t> comms(G);
  if (G > 1 && ncclCommInitAll(comms.data(), G, devs.data()) != ncclSuccess) { std::printf("FAIL ncclCommInitAll\n"); return 1; }
  std::vector<void*> streams(G), cm(G), ws(G);
  std::vector<b2d_kvp_if*> out(G);
  std::vector<const float*> xs(G), ys(G);
  std::vector<int64_t> nsh(G), off(G);
  std::vector<int64_t*> keys(G);
  std::vector<size_t> wsb(G);
  for (int g = 0; g < G; ++g) {
    CK(cudaSetDevice(g));
    const int64_t lo = n * g / G, hi = n * (g + 1) / G;
    nsh[g] = hi - lo; off[g] = lo;
    cudaStream_t st; CK(cudaStreamCreate(&st)); streams[g] = st;
    cm[g] = G > 1 ? (void*)comms[g] : nullptr;
    float *x, *y;
    CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&y, (size_t)nsh[g] * k * 4));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(y, hy.data() + lo * k, (size_t)nsh[g] * k * 4, cudaMemcpyHostToDevice));
    xs[g] = x; ys[g] = y;
    CK(cudaMalloc(&out[g], m * sizeof(b2d_kvp_if))); CK(cudaMalloc(&keys[g], m * 8));
    wsb[g] = b2d_fused_l2_nn_workspace_bytes(m, nsh[g], k);
    CK(cudaMalloc(&ws[g], wsb[g]));
  }
  if (b2d_fused_l2_nn_multi(G, devs.data(), streams.data(), cm.data(), out.data(), xs.data(), k, ys.data(), k, nsh.data(), off.data(),
                            m, k, 0, keys.data(), ws.data(), wsb.data())) { std::printf("FAIL multi: %s\n", b2d_last_error()); return 1; }
  int bad = 0; double worst = 0;
  for (int g = 0; g < G; ++g) {
    CK(cudaSetDevice(g));
    CK(cudaStreamSynchronize((cudaStream_t)streams[g]));
    std::vector<b2d_kvp_if> got(m);
    CK(cudaMemcpy(got.data(), out[g], m * sizeof(b2d_kvp_if), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < m; ++i) {
      if (got[i].key != ref[i].key) {
        // a different index is only acceptable for an fp32-inseparable tie
        double da = 0, db = 0;
        for (int t = 0; t < k; ++t) { double a = (double)hx[i * k + t] - hy[(int64_t)got[i].key * k + t]; double b = (double)hx[i * k + t] - hy[(int64_t)ref[i].key * k + t]; da += a * a; db += b * b; }
        if (std::fabs(da - db) > 1e-4 * db + 4e-6 * (da + db)) ++bad;   // (the engine's 1e-4 bar: tensor-path vs direct fp32 finalists)
      }
      // the single-device call may finish through the tensor-path kernel (3-term split: ~5e-5 of d, and an absolute
      // error ~1e-6 |x|^2 on near-duplicate pairs) while a sub-chunked shard search re-measures its finalists directly
      // in fp32: both are inside the engine's 1e-4 bar, so that is the tolerance here
      double xn = 0;
      for (int t = 0; t < k; ++t) xn += (double)hx[i * k + t] * hx[i * k + t];
      const double diff = std::fabs((double)got[i].value - ref[i].value);
      const double rel  = diff / std::fmax(ref[i].value, 1e-6);
      if (diff <= 4e-6 * xn) continue;
      if (rel > worst) worst = rel;
      if (rel > 1e-4) {
        if (bad < 6) std::printf("  dev %d row %lld: got (%d, %g) ref (%d, %g)\n", g, (long long)i, got[i].key, got[i].value, ref[i].key, ref[i].value);
        ++bad;
      }
    }
  }
  std::printf("devices %d  max rel diff of the values vs the single-device call %.3g\n", G, worst);
  std::printf(bad ? "FAIL %d\n" : "PASS\n", bad);
  return bad ? 1 : 0;
}
