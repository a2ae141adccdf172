// C-ABI entry points (include/raft_b200.h) and host-side dispatch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/raft_b200.h"
#include "expanded_tc.cuh"
#include "expanded_tc2.cuh"
#include "screen_tc.cuh"
#include "prep.cuh"
#include "unexpanded_simt.cuh"
#include "stats.cuh"
#include "fp64_simt.cuh"

namespace b2d {

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}
#define B2D_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(B2D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));        \
  } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// process-wide tuning / test hooks (b2d_set_option): nothing on the hot path reads the environment
static std::atomic<float> g_nn_tau{6.0f};   // screened fusedL2NN: candidates per row above which the trial calls screening off
static std::atomic<int> g_nn_screen{1};     // 0: always the exact arg-min kernel

// ------------------------------------------------------------------------------------------
// TMA descriptor encode through the driver entry point (no link-time libcuda dependency, so the
// library still loads -- and reports a clean error -- on a machine without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// packed operand [rows][nkb*64] fp16, box = 64 x box_rows, SWIZZLE_128B
// (hi_only: boxes of the 32 hi halves of a k-block, 64 B rows, SWIZZLE_64B -- the coarse screening pass)
static int make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int nkb, int box_rows,
                            int kb0 = 0, int nkb_total = 0, bool hi_only = false)
{
  if (nkb_total == 0) nkb_total = nkb;
  base = static_cast<const char*>(base) + static_cast<size_t>(kb0) * 128;  // K chunk [kb0, kb0 + nkb) of every row
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2]    = {static_cast<cuuint64_t>(nkb) * 64, static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(nkb_total) * 128};
  cuuint32_t box[2]     = {hi_only ? 32u : 64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2]    = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, hi_only ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return B2D_OK;
}

static int device_sms(int* sms, int* cc_major)
{
  int dev = 0;
  B2D_CUDA(cudaGetDevice(&dev));
  B2D_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  B2D_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------
// Optional timing of the dominant kernel (bench.py's roofline): CUDA events on the call's stream around
// the main kernel of every b2d_pairwise_distance call between b2d_profile_begin and b2d_profile_end.
// Nothing is synchronised until b2d_profile_end, so the timed region of the caller is not perturbed.
namespace {
struct ProfileRing {
  std::vector<cudaEvent_t> ev;  // pairs
  int used = 0;
  bool on  = false;
} g_prof;
std::mutex g_prof_mu;
std::atomic<bool> g_prof_armed{false};   // fast path: a call outside a profiling region touches no lock

struct ProfileScope {
  cudaStream_t s;
  int slot = -1;
  explicit ProfileScope(cudaStream_t st) : s(st)
  {
    if (!g_prof_armed.load(std::memory_order_acquire)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof.on && 2 * (g_prof.used + 1) <= static_cast<int>(g_prof.ev.size())) {
      slot = g_prof.used++;
      cudaEventRecord(g_prof.ev[2 * slot], s);
    }
  }
  ~ProfileScope()
  {
    if (slot >= 0) cudaEventRecord(g_prof.ev[2 * slot + 1], s);
  }
};
}  // namespace

// ------------------------------------------------------------------------------------------
// workspace layout of the tensor-core path
struct TcWorkspace {
  __half* xop;
  __half* yop;
  float* xt;
  float* yt;
  float* xsc;        // [m] per-row scale 2^(E - e_row) of x (1 unless the row took its own exponent, prep.cuh)
  float* ysc;        // [n] the same for y
  unsigned* nonuni;  // [2] some row of x / y took its own exponent
  long long* keys;
  unsigned* gmax;  // [2]
  float* coef;     // [1]
  unsigned* has_lo;  // [1]
  float4* aux;       // [m]    screened NN: per-row (bound - |x|^2, |x|, |x - hi part|, -)
  float* xlo;        // [m]    screened NN: |x_i - hi part| (prep.cuh lvec)
  float* ylo;        // [n]
  int2* cand;        // [cap]  screened NN: candidate list
  unsigned* cand_cnt;  // [1] (+ overflow flag right behind it)
  unsigned cand_cap;
  // screened NN: the database chunk is worked through in the order of its squared row norms (screen_tc.cuh)
  unsigned* sort_key[2];  // [min(n, 2^20)] each
  int* sort_val[2];       // [min(n, 2^20)] each: position -> source row
  void* sort_tmp;
  size_t bytes;
};
constexpr int64_t kNnChunkRows  = 1 << 20;   // database rows per screened chunk
constexpr size_t kSortTmpBytes  = 4u << 20;  // cub::DeviceRadixSort scratch for <= 2^20 pairs (checked at run time)

static TcWorkspace tc_layout(void* base, int64_t m, int64_t n, int64_t k, bool with_keys)
{
  const int nkb = static_cast<int>((k + 31) / 32);
  size_t off    = 0;
  auto take     = [&](size_t b) {
    size_t o = off;
    off += align_up(b, 1024);
    return o;
  };
  char* c = static_cast<char*>(base);
  TcWorkspace w;
  // scalars, xt and keys come first so that their offsets depend on m only
  // (b2d_fused_l2_nn_finalize finds |x_i|^2 again without knowing n or k)
  w.gmax = reinterpret_cast<unsigned*>(c + take(32));
  w.coef = reinterpret_cast<float*>(w.gmax + 2);
  w.has_lo = w.gmax + 3;
  w.nonuni = w.gmax + 4;
  w.xt   = reinterpret_cast<float*>(c + take(static_cast<size_t>(m) * 4));
  w.xsc  = reinterpret_cast<float*>(c + take(static_cast<size_t>(m) * 4));
  w.keys = reinterpret_cast<long long*>(c + take(with_keys ? static_cast<size_t>(m) * 8 : 0));
  // screened fusedL2NN scratch: thresholds, counters, candidate list (128 per row and 1M-row chunk of y;
  // ~54 measured on far-from-origin clusters; overflow falls back to the exact pass on the device,
  // never to a wrong answer)
  // (+2^19: the coarse kernel's warps take list slots in blocks of 64 -- 148 x 16 warps x 2 launches leave up to
  // ~300k slots unused, which must not look like an overflow when m is small)
  w.cand_cap = with_keys ? static_cast<unsigned>(std::min<int64_t>(128 * m + (1 << 19), 0x7fffffff)) : 0u;
  w.aux      = reinterpret_cast<float4*>(c + take(with_keys ? static_cast<size_t>(m) * 16 : 0));
  w.cand_cnt = reinterpret_cast<unsigned*>(c + take(with_keys ? 32 : 0));
  w.cand     = reinterpret_cast<int2*>(c + take(static_cast<size_t>(w.cand_cap) * 8));
  w.yt   = reinterpret_cast<float*>(c + take(static_cast<size_t>(n) * 4));
  w.ysc  = reinterpret_cast<float*>(c + take(static_cast<size_t>(n) * 4));
  w.xop  = reinterpret_cast<__half*>(c + take(static_cast<size_t>(m) * nkb * 128));
  w.yop  = reinterpret_cast<__half*>(c + take(static_cast<size_t>(n) * nkb * 128));
  const size_t ns = with_keys ? static_cast<size_t>(std::min<int64_t>(n, kNnChunkRows)) : 0;
  for (int i = 0; i < 2; ++i) {
    w.sort_key[i] = reinterpret_cast<unsigned*>(c + take(ns * 4));
    w.sort_val[i] = reinterpret_cast<int*>(c + take(ns * 4));
  }
  w.sort_tmp = c + take(with_keys ? kSortTmpBytes : 0);
  w.xlo = with_keys ? reinterpret_cast<float*>(c + take(static_cast<size_t>(m) * 4)) : nullptr;
  w.ylo = with_keys ? reinterpret_cast<float*>(c + take(static_cast<size_t>(n) * 4)) : nullptr;
  w.bytes = off;
  return w;
}

static bool is_expanded(int metric)
{
  return metric == B2D_L2Expanded || metric == B2D_L2SqrtExpanded || metric == B2D_CosineExpanded ||
         metric == B2D_CorrelationExpanded || metric == B2D_InnerProduct || metric == B2D_HellingerExpanded ||
         metric == B2D_RusselRaoExpanded || metric == B2D_JaccardExpanded || metric == B2D_DiceExpanded;
}
static bool is_unexpanded(int metric)
{
  return metric == B2D_L1 || metric == B2D_L2Unexpanded || metric == B2D_L2SqrtUnexpanded ||
         metric == B2D_Linf || metric == B2D_Canberra || metric == B2D_LpUnexpanded ||
         metric == B2D_HammingUnexpanded || metric == B2D_KLDivergence || metric == B2D_JensenShannon ||
         metric == B2D_BrayCurtis;
}

template <typename T>
static int launch_prep(cudaStream_t s, const TcWorkspace& w, const void* x, int64_t xrs, int64_t xcs, int64_t m,
                       const void* y, int64_t yrs, int64_t ycs, int64_t n, int64_t k, const float* xn, const float* yn,
                       int mode, int center, int xform = 0, float coef_mul = 1.f, float tx_const = 0.f,
                       const int* y_gather = nullptr)
{
  PrepParams p;
  p.side[0] = PrepSide{x, xrs, xcs, m, w.xop, w.xt, w.xsc, xn, w.xlo, nullptr};
  p.side[1] = PrepSide{y, yrs, ycs, n, w.yop, w.yt, w.ysc, yn, w.ylo, y_gather};
  p.k = static_cast<int>(k); p.nkb = static_cast<int>((k + 31) / 32); p.mode = mode; p.center = center;
  p.gmax = w.gmax; p.coef = w.coef; p.has_lo = w.has_lo; p.nonuni = w.nonuni;
  p.xform = xform; p.coef_mul = coef_mul; p.tx_const = tx_const;
  B2D_CUDA(cudaMemsetAsync(w.gmax, 0, 32, s));
  const int64_t blocks = (m + n + 7) / 8;
  prep_max_kernel<T><<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
  B2D_CUDA(cudaGetLastError());
  prep_split_kernel<T><<<static_cast<unsigned>(blocks), 256, 0, s>>>(p);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// dist [m][n] fp32 (row pitch ldd) seen as [m][chunks of 128 columns][128]: box {132, 2, 8} = 8 rows x 2 chunks, the
// 4 trailing floats of every chunk fall outside dimension 0 and are never written (padded smem rows, see
// expanded_tc.cuh).  `first_col`/`cols`: the column range the map covers (the ragged last chunk has its own map).
static int make_dist_map3(CUtensorMap* map, const float* base, int64_t m, int64_t first_col, int64_t cols, int64_t ldd)
{
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  const cuuint64_t inner = cols >= 128 ? 128 : static_cast<cuuint64_t>(cols);
  cuuint64_t dims[3]    = {inner, static_cast<cuuint64_t>(cols >= 128 ? cols / 128 : 1), static_cast<cuuint64_t>(m)};
  cuuint64_t strides[2] = {512, static_cast<cuuint64_t>(ldd) * 4};
  cuuint32_t box[3]     = {132, 2, 8};
  cuuint32_t estr[3]    = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base + first_col), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled(dist) failed: " + std::to_string((int)r));
  return B2D_OK;
}

template <bool kRes, int kEpi, int kPost, bool kTma>
static int launch_tc_inst(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                          const CUtensorMap& mr, const TcParams& p, int grid)
{
  // the attribute is per device and per function: cheap, set on every launch
  B2D_CUDA(cudaFuncSetAttribute(expanded_tc_kernel<kRes, kEpi, kPost, kTma>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(TC_SMEM_BYTES)));
  expanded_tc_kernel<kRes, kEpi, kPost, kTma><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(ma, mb, md, mr, p);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

template <bool kRes, bool kTma>
static int launch_tc_store(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                           const CUtensorMap& mr, const TcParams& p, int grid, int post)
{
  if (post == POST_NONE) return launch_tc_inst<kRes, EPI_STORE, POST_NONE, kTma>(s, ma, mb, md, mr, p, grid);
  if (post == POST_CLAMP) return launch_tc_inst<kRes, EPI_STORE, POST_CLAMP, kTma>(s, ma, mb, md, mr, p, grid);
  if (post == POST_JACCARD) return launch_tc_inst<kRes, EPI_STORE, POST_JACCARD, false>(s, ma, mb, md, mr, p, grid);
  if (post == POST_DICE) return launch_tc_inst<kRes, EPI_STORE, POST_DICE, false>(s, ma, mb, md, mr, p, grid);
  return launch_tc_inst<kRes, EPI_STORE, POST_CLAMP_SQRT, kTma>(s, ma, mb, md, mr, p, grid);
}

static std::atomic<int> g_pair_kernel{0};  // b2d_set_option("pairwise_2cta", 1): the CTA-pair kernel (expanded_tc2.cuh) for k <= 128

// K1, 2-CTA form (expanded_tc2.cuh): EPI_STORE, k <= 128, aligned output
template <int kPost>
static int launch_tc2_inst(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                           const TcParams& p, int grid)
{
  B2D_CUDA(cudaFuncSetAttribute(expanded_tc2_kernel<kPost>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(T2_SMEM_BYTES)));
  expanded_tc2_kernel<kPost><<<grid, T2_THREADS, T2_SMEM_BYTES, s>>>(ma, mb, md, p);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

static int launch_tc2(cudaStream_t s, const TcWorkspace& w, TcParams p, int64_t k, int post, int sms)
{
  p.nkb     = static_cast<int>((k + 31) / 32);
  p.tiles_m = static_cast<int>((p.m + TC_BM - 1) / TC_BM);
  p.tiles_n = static_cast<int>((p.n + TC_BN - 1) / TC_BN);
  const int clusters  = std::max(1, sms / 2);
  const int pairs_m   = (p.tiles_m + 1) / 2;
  const int64_t total = static_cast<int64_t>(pairs_m) * p.tiles_n;
  int64_t chunk       = total / (static_cast<int64_t>(clusters) * 6);
  chunk               = std::max<int64_t>(1, std::min<int64_t>(chunk, 16));
  chunk               = std::min<int64_t>(chunk, pairs_m);
  p.chunk    = static_cast<int>(chunk);        // PAIRS of x tiles per work item
  p.chunks_m = (pairs_m + p.chunk - 1) / p.chunk;
  p.n_items  = static_cast<int64_t>(p.tiles_n) * p.chunks_m;
  p.xt = w.xt; p.yt = w.yt; p.coef = w.coef; p.has_lo = w.has_lo; p.xsc = w.xsc; p.ysc = w.ysc; p.nonuni = w.nonuni;
  if (p.n_items == 0) return B2D_OK;
  CUtensorMap ma, mb, md;
  memset(&md, 0, sizeof(md));
  int rc = make_operand_map(&ma, w.xop, p.m, p.nkb, TC_BM);
  if (rc) return rc;
  rc = make_operand_map(&mb, w.yop, p.n, p.nkb, 128);   // each CTA of a pair loads ITS 128 rows of the y block
  if (rc) return rc;
  if (p.n >= 128) { rc = make_dist_map3(&md, p.dist, p.m, 0, p.n / 128 * 128, p.ldd); if (rc) return rc; }
  const int grid = 2 * static_cast<int>(std::min<int64_t>(p.n_items, clusters));
  if (post == POST_NONE) return launch_tc2_inst<POST_NONE>(s, ma, mb, md, p, grid);
  if (post == POST_CLAMP) return launch_tc2_inst<POST_CLAMP>(s, ma, mb, md, p, grid);
  return launch_tc2_inst<POST_CLAMP_SQRT>(s, ma, mb, md, p, grid);
}

static int launch_tc(cudaStream_t s, const TcWorkspace& w, TcParams p, int64_t k, int epi, int post, int kb0 = 0,
                     int nkb_chunk = 0)
{
  int sms = 0, cc = 0;
  int rc  = device_sms(&sms, &cc);
  if (rc) return rc;
  if (cc != 10) return fail(B2D_ERR_CUDA, "raft_b200 requires an sm_100 (B200) device; found cc major " + std::to_string(cc));
  const int nkb_total = static_cast<int>((k + 31) / 32);
  p.nkb     = nkb_chunk ? nkb_chunk : nkb_total;
  p.tiles_m = static_cast<int>((p.m + TC_BM - 1) / TC_BM);
  p.tiles_n = static_cast<int>((p.n + TC_BN - 1) / TC_BN);
  p.tiles_sel = sel_count(p.tiles_n, p.sel_s, p.sel_lo, p.sel_hi);
  int64_t total = static_cast<int64_t>(p.tiles_m) * p.tiles_sel;
  int64_t chunk = total / (static_cast<int64_t>(sms) * 6);
  if (chunk < 1) chunk = 1;
  if (chunk > 32) chunk = 32;
  if (chunk > p.tiles_m) chunk = p.tiles_m;
  p.chunk    = static_cast<int>(chunk);
  p.chunks_m = (p.tiles_m + p.chunk - 1) / p.chunk;
  p.n_items  = static_cast<int64_t>(p.tiles_sel) * p.chunks_m;
  p.xt       = w.xt;
  p.yt       = w.yt;
  p.coef     = w.coef;
  p.has_lo   = w.has_lo;
  p.xsc      = w.xsc;
  p.ysc      = w.ysc;
  p.nonuni   = w.nonuni;
  if (p.n_items == 0) return B2D_OK;
  CUtensorMap ma, mb, md, mr;
  memset(&md, 0, sizeof(md));
  memset(&mr, 0, sizeof(mr));
  rc = make_operand_map(&ma, w.xop, p.m, p.nkb, TC_BM, kb0, nkb_total);
  if (rc) return rc;
  rc = make_operand_map(&mb, w.yop, p.n, p.nkb, TC_BN, kb0, nkb_total);
  if (rc) return rc;
  const int grid      = static_cast<int>(p.n_items < sms ? p.n_items : sms);
  const bool resident = p.nkb <= TC_MAX_RES_KB;
  if (epi == EPI_MINLOC) {
    return resident ? launch_tc_inst<true, EPI_MINLOC, POST_NONE, false>(s, ma, mb, md, mr, p, grid)
                    : launch_tc_inst<false, EPI_MINLOC, POST_NONE, false>(s, ma, mb, md, mr, p, grid);
  }
  if (epi == EPI_TOPK) {
    return resident ? launch_tc_inst<true, EPI_TOPK, POST_NONE, false>(s, ma, mb, md, mr, p, grid)
                    : launch_tc_inst<false, EPI_TOPK, POST_NONE, false>(s, ma, mb, md, mr, p, grid);
  }
  // bulk row stores (cp.async.bulk shared -> global) need 16-byte aligned addresses and sizes, i.e. a
  // 16-byte aligned base and row pitch and n % 4 == 0; ragged or unaligned outputs take the direct
  // register->global path
  bool tma = (reinterpret_cast<uintptr_t>(p.dist) % 16 == 0) && (p.ldd % 4 == 0) && (p.n % 4 == 0) &&
             p.acc_mode == 0 && post < POST_JACCARD;  // the K-chunked read-modify-write epilogue lives in the direct path
  // k <= 128 with an aligned output: the CTA-pair kernel (full-width output rows, expanded_tc2.cuh)
  if (tma && resident && p.sel_s <= 1 && p.run_flag == nullptr && sms >= 2 && g_pair_kernel.load(std::memory_order_relaxed))
    return launch_tc2(s, w, p, k, post, sms);
  if (tma) {
    // dist [m][n] fp32 (row pitch ldd), box = 32 x 32, SWIZZLE_128B (inner box = 128 bytes); edge clipping works in
    // 16-byte units, hence n % 4 == 0 above
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t dims[2]    = {static_cast<cuuint64_t>(p.n), static_cast<cuuint64_t>(p.m)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(p.ldd) * 4};
    cuuint32_t box[2]     = {32, 32};
    cuuint32_t estr[2]    = {1, 1};
    CUresult r = enc(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.dist, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled(dist) failed: " + std::to_string((int)r));
    mr = md;
    // k <= 64: the full-width store path of the same kernel (3-D box {132, 2, 8} over the 128-column chunks)
    if (p.n >= 128) { rc = make_dist_map3(&mr, p.dist, p.m, 0, p.n / 128 * 128, p.ldd); if (rc) return rc; }
    return resident ? launch_tc_store<true, true>(s, ma, mb, md, mr, p, grid, post)
                    : launch_tc_store<false, true>(s, ma, mb, md, mr, p, grid, post);
  }
  return resident ? launch_tc_store<true, false>(s, ma, mb, md, mr, p, grid, post)
                  : launch_tc_store<false, false>(s, ma, mb, md, mr, p, grid, post);
}

// coarse screening pass of the screened fusedL2NN (screen_tc.cuh) over the y blocks with
// sel_lo <= index % sel_s < sel_hi
static int launch_screen(cudaStream_t s, const TcWorkspace& w, int64_t m, int64_t n, int64_t k, int sel_s, int sel_lo,
                         int sel_hi, unsigned* overflow, const unsigned* run_flag, int unit_norm, const int* col_map)
{
  int sms = 0, cc = 0;
  int rc  = device_sms(&sms, &cc);
  if (rc) return rc;
  if (cc != 10) return fail(B2D_ERR_CUDA, "raft_b200 requires an sm_100 (B200) device; found cc major " + std::to_string(cc));
  ScreenParams p;
  memset(&p, 0, sizeof(p));
  p.m = m; p.n = n;
  p.nkb     = static_cast<int>((k + 31) / 32);
  if (p.nkb > SC_MAX_KB) return fail(B2D_ERR_UNSUPPORTED, "internal: screening needs k <= 128");
  p.n_stages = static_cast<int>(std::min<size_t>(SC_MAX_STAGES, SC_A_RING / (static_cast<size_t>(p.nkb) * SC_A_KB_BYTES)));
  p.tiles_m = static_cast<int>((m + TC_BM - 1) / TC_BM);
  const int tiles_n = static_cast<int>((n + TC_BN - 1) / TC_BN);
  p.sel_s = sel_s; p.sel_lo = sel_lo; p.sel_hi = sel_hi;
  p.tiles_sel = sel_count(tiles_n, sel_s, sel_lo, sel_hi);
  // long runs of x tiles per y block: the y block load is not overlapped with the previous item
  int64_t total = static_cast<int64_t>(p.tiles_m) * p.tiles_sel;
  int64_t chunk = total / (static_cast<int64_t>(sms) * 4);
  chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, 128));
  chunk = std::min<int64_t>(chunk, p.tiles_m);
  p.chunk    = static_cast<int>(chunk);
  p.chunks_m = (p.tiles_m + p.chunk - 1) / p.chunk;
  p.n_items  = static_cast<int64_t>(p.tiles_sel) * p.chunks_m;
  p.xsc = w.xsc; p.nonuni = w.nonuni;
  p.yt = w.yt; p.ylo = w.ylo; p.coef = w.coef; p.aux = w.aux; p.cand = w.cand; p.cand_cnt = w.cand_cnt; p.cand_cap = w.cand_cap;
  p.overflow = overflow; p.run_flag = run_flag; p.unit_norm = unit_norm; p.col_map = col_map;
  if (p.n_items == 0) return B2D_OK;
  CUtensorMap ma, mb;
  rc = make_operand_map(&ma, w.xop, m, p.nkb, TC_BM, 0, 0, true);
  if (rc) return rc;
  rc = make_operand_map(&mb, w.yop, n, p.nkb, TC_BN, 0, 0, true);
  if (rc) return rc;
  const int grid = static_cast<int>(p.n_items < sms ? p.n_items : sms);
  B2D_CUDA(cudaFuncSetAttribute(screen_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(SC_SMEM_BYTES)));
  screen_tc_kernel<<<grid, SC_THREADS, SC_SMEM_BYTES, s>>>(ma, mb, p);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// fp32 row-major matrix [rows][k] with row pitch ld: box = 32 floats x 128 rows, SWIZZLE_128B
static int make_f32_map(CUtensorMap* map, const float* base, int64_t rows, int64_t k, int64_t ld)
{
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2]    = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2]     = {UX_KB, UX_BM};
  cuuint32_t estr[2]    = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2D_ERR_CUDA, "cuTensorMapEncodeTiled(f32 operand) failed: " + std::to_string((int)r));
  return B2D_OK;
}

template <int kMetric>
static int launch_ux_inst(cudaStream_t s, const UxParams& p, int64_t tiles)
{
  if (p.vec_x && p.vec_y) {  // TMA-legal operands: hardware-staged tiles
    CUtensorMap mx, my;
    int rc = make_f32_map(&mx, p.x, p.m, p.k, p.xrs);
    if (rc) return rc;
    rc = make_f32_map(&my, p.y, p.n, p.k, p.yrs);
    if (rc) return rc;
    B2D_CUDA(cudaFuncSetAttribute(unexpanded_tma_kernel<kMetric>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  UX_TMA_SMEM_BYTES));
    unexpanded_tma_kernel<kMetric><<<static_cast<unsigned>(tiles), UX_THREADS, UX_TMA_SMEM_BYTES, s>>>(mx, my, p);
    B2D_CUDA(cudaGetLastError());
    return B2D_OK;
  }
  cudaError_t e = cudaFuncSetAttribute(unexpanded_simt_kernel<kMetric>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       UX_SMEM_BYTES);
  B2D_CUDA(e);
  unexpanded_simt_kernel<kMetric><<<static_cast<unsigned>(tiles), UX_THREADS, UX_SMEM_BYTES, s>>>(p);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------
__global__ void row_norm_kernel(float* out, const float* x, int64_t ldx, int64_t rows, int k, int type, int do_sqrt)
{
  const int lane  = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* row = x + r * ldx;
  double acc = 0.0;
  float mx   = 0.f;
  for (int t = lane; t < k; t += 32) {
    const float v = __ldg(row + t);
    if (type == B2D_L0PseudoNorm) acc += (v != 0.f) ? 1.0 : 0.0;
    else if (type == B2D_L1Norm) acc += fabs(static_cast<double>(v));
    else if (type == B2D_L2Norm) acc += static_cast<double>(v) * v;
    else mx = fmaxf(mx, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    double v = (type == B2D_LinfNorm) ? static_cast<double>(mx) : acc;
    if (do_sqrt) v = sqrt(v);
    out[r] = static_cast<float>(v);
  }
}

// raft::matrix::argmin (cpp/include/raft/matrix/argmin.cuh:25-37, detail/math.cuh:290-343): column index of the
// minimum of every row, ties -> the smaller index (cub::ArgMin), start value (0, +inf): a row of NaN / +inf gives 0.
// One pass over the matrix: HBM-read bound.  kWarp: one warp per row (short rows), else one 256-thread block per row.
template <bool kWarp>
__global__ void __launch_bounds__(256) row_argmin_kernel(int* out, const float* in, int64_t ld, int64_t rows, int64_t n)
{
  const int lane    = threadIdx.x & 31;
  const int64_t r   = kWarp ? static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5) : blockIdx.x;
  const int t0      = kWarp ? lane : threadIdx.x;
  const int stride  = kWarp ? 32 : 256;
  float v = __int_as_float(0x7f800000);
  int ix  = 0;
  if (r < rows) {
    const float* row = in + r * ld;
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0 && (ld & 3) == 0) {
      const int64_t n4 = n >> 2;
      for (int64_t c = t0; c < n4; c += stride) {
        const float4 q4 = __ldcs(reinterpret_cast<const float4*>(row) + c);
        const int j     = static_cast<int>(c << 2);
        if (q4.x < v) { v = q4.x; ix = j; }
        if (q4.y < v) { v = q4.y; ix = j + 1; }
        if (q4.z < v) { v = q4.z; ix = j + 2; }
        if (q4.w < v) { v = q4.w; ix = j + 3; }
      }
      for (int64_t c = (n4 << 2) + t0; c < n; c += stride) {
        const float q1 = __ldcs(row + c);
        if (q1 < v) { v = q1; ix = static_cast<int>(c); }
      }
    } else {
      for (int64_t c = t0; c < n; c += stride) {
        const float q1 = __ldcs(row + c);
        if (q1 < v) { v = q1; ix = static_cast<int>(c); }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi   = __shfl_xor_sync(0xffffffffu, ix, o);
    if (ov < v || (ov == v && oi < ix)) { v = ov; ix = oi; }
  }
  if (kWarp) {
    if (lane == 0 && r < rows) out[r] = ix;
    return;
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  if (lane == 0) { sv[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = ix; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] < v || (sv[w] == v && si[w] < ix)) { v = sv[w]; ix = si[w]; }
    out[r] = ix;
  }
}

}  // namespace b2d

using namespace b2d;

// fp64 path: one SIMT kernel per metric (fp64_simt.cuh)
template <int kMetric>
static int launch_f64_inst(cudaStream_t s, const F64Params& p, int64_t tiles)
{
  f64_pairwise_kernel<kMetric><<<static_cast<unsigned>(tiles), 256, 0, s>>>(p);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

static int pairwise_f64(cudaStream_t s, int metric, const double* x, int64_t xrs, int64_t xcs, const double* y, int64_t yrs,
                        int64_t ycs, double* dist, int64_t ldd, int64_t m, int64_t n, int64_t k, int swapped, double metric_arg,
                        void* workspace, size_t workspace_bytes)
{
  const size_t need = align_up(static_cast<size_t>(m + n) * 16, 1024);
  if (!workspace || workspace_bytes < need) return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  double* xs = static_cast<double*>(workspace);
  double* ys = xs + 2 * m;
  f64_row_stats_kernel<<<static_cast<unsigned>((m + 7) / 8), 256, 0, s>>>(xs, x, xrs, xcs, m, static_cast<int>(k));
  f64_row_stats_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(ys, y, yrs, ycs, n, static_cast<int>(k));
  B2D_CUDA(cudaGetLastError());
  F64Params p;
  p.x = x; p.y = y; p.dist = dist; p.xrs = xrs; p.xcs = xcs; p.yrs = yrs; p.ycs = ycs; p.ldd = ldd; p.m = m; p.n = n;
  p.k = static_cast<int>(k); p.metric = metric; p.swapped = swapped; p.p = metric_arg; p.inv_p = 1.0 / metric_arg;
  p.xs = xs; p.ys = ys;
  p.tiles_n = static_cast<int>((n + F64_T - 1) / F64_T);
  const int64_t tiles = ((m + F64_T - 1) / F64_T) * p.tiles_n;
  switch (metric) {
#define B2D_F64_CASE(M) case M: return launch_f64_inst<M>(s, p, tiles);
    B2D_F64_CASE(B2D_L2Expanded) B2D_F64_CASE(B2D_L2SqrtExpanded) B2D_F64_CASE(B2D_CosineExpanded) B2D_F64_CASE(B2D_L1)
    B2D_F64_CASE(B2D_L2Unexpanded) B2D_F64_CASE(B2D_L2SqrtUnexpanded) B2D_F64_CASE(B2D_InnerProduct) B2D_F64_CASE(B2D_Linf)
    B2D_F64_CASE(B2D_Canberra) B2D_F64_CASE(B2D_LpUnexpanded) B2D_F64_CASE(B2D_CorrelationExpanded)
    B2D_F64_CASE(B2D_JaccardExpanded) B2D_F64_CASE(B2D_HellingerExpanded) B2D_F64_CASE(B2D_BrayCurtis)
    B2D_F64_CASE(B2D_JensenShannon) B2D_F64_CASE(B2D_HammingUnexpanded) B2D_F64_CASE(B2D_KLDivergence)
    B2D_F64_CASE(B2D_RusselRaoExpanded) B2D_F64_CASE(B2D_DiceExpanded)
#undef B2D_F64_CASE
    default: return fail(B2D_ERR_UNSUPPORTED, "metric");
  }
}

extern "C" {

int b2d_profile_begin(int capacity)
{
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (capacity < 1 || capacity > 4096) return fail(B2D_ERR_INVALID_ARG, "capacity must be in [1, 4096]");
  while (static_cast<int>(g_prof.ev.size()) < 2 * capacity) {
    cudaEvent_t e;
    B2D_CUDA(cudaEventCreate(&e));
    g_prof.ev.push_back(e);
  }
  g_prof.used = 0;
  g_prof.on   = true;
  g_prof_armed.store(true, std::memory_order_release);
  return B2D_OK;
}

int b2d_profile_end(float* ms, int max_count, int* count)
{
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.on = false;
  g_prof_armed.store(false, std::memory_order_release);
  if (!ms || !count || max_count < 0) return fail(B2D_ERR_INVALID_ARG, "null ms / count");
  const int c = std::min(g_prof.used, max_count);
  for (int i = 0; i < c; ++i) {
    B2D_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
    B2D_CUDA(cudaEventElapsedTime(&ms[i], g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
  }
  *count = c;
  return B2D_OK;
}


int b2d_version(void) { return 200; }

int b2d_set_option(const char* name, double value)
{
  if (!name) return fail(B2D_ERR_INVALID_ARG, "null option name");
  const std::string n(name);
  if (n == "nn_tau") {
    if (!(value >= 0.0)) return fail(B2D_ERR_INVALID_ARG, "nn_tau must be >= 0");
    g_nn_tau.store(static_cast<float>(value));
  } else if (n == "nn_screen") {
    g_nn_screen.store(value != 0.0 ? 1 : 0);
  } else if (n == "pairwise_2cta") {
    g_pair_kernel.store(value != 0.0 ? 1 : 0);
  } else {
    return fail(B2D_ERR_INVALID_ARG, "unknown option '" + n + "' (nn_tau, nn_screen, pairwise_2cta)");
  }
  return B2D_OK;
}

// Diagnostic (synchronises `stream`): the control words the last screened fusedL2NN chunk left in `workspace`
// [0] list slots in use (one incumbent per row + blocks of 64 reserved by the screen) [1] list overflow [2] go_screen [3] go_exact [4] redo_trial [5] candidates after the trial [6] candidates found by the screen (out7: 7 words)
int b2d_debug_nn_stats(void* stream, const void* workspace, int64_t m, int64_t n, int64_t k, unsigned* out6 /* 7 words */)
{
  if (!workspace || !out6 || m < 0 || n < 0 || k < 0) return fail(B2D_ERR_INVALID_ARG, "null workspace / out");
  TcWorkspace w = tc_layout(const_cast<void*>(workspace), m, n, k, true);
  B2D_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  B2D_CUDA(cudaMemcpy(out6, w.cand_cnt, 7 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  return B2D_OK;
}
const char* b2d_last_error(void) { return g_err.c_str(); }
#ifdef SC_TRACE
int b2d_debug_screen_trace(long long* out_host)
{
  B2D_CUDA(cudaDeviceSynchronize());
  B2D_CUDA(cudaMemcpyFromSymbol(out_host, g_sc_trace, sizeof(long long) * 8 * 4096));
  B2D_CUDA(cudaMemcpyFromSymbol(out_host + 8 * 4096, g_sc_wtrace, sizeof(long long) * 3 * 16 * 2048));
  return B2D_OK;
}
#endif

size_t b2d_pairwise_workspace_bytes(int metric, int dtype, int64_t m, int64_t n, int64_t k)
{
  if (dtype != B2D_F32 && dtype != B2D_F16 && dtype != B2D_F64) return static_cast<size_t>(-1);
  if (m < 0 || n < 0 || k < 0) return static_cast<size_t>(-1);
  if (dtype == B2D_F64)   // per-row (sum, sum of squares) of x and y
    return (is_expanded(metric) || is_unexpanded(metric)) ? align_up(static_cast<size_t>(m + n) * 16, 1024) : static_cast<size_t>(-1);
  if (is_unexpanded(metric)) return dtype == B2D_F32 ? 0 : static_cast<size_t>(-1);
  if (!is_expanded(metric)) return static_cast<size_t>(-1);
  return tc_layout(nullptr, m, n, k, false).bytes;
}

int b2d_pairwise_distance(void* stream, int metric, int dtype, const void* x, int64_t ldx, const void* y,
                          int64_t ldy, void* dist_v, int64_t ldd, int64_t m, int64_t n, int64_t k,
                          int row_major, float metric_arg, void* workspace, size_t workspace_bytes)
{
  float* dist = static_cast<float*>(dist_v);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m < 0 || n < 0 || k < 0) return fail(B2D_ERR_INVALID_ARG, "negative extent");
  if (m == 0 || n == 0) return B2D_OK;
  if (k == 0) return fail(B2D_ERR_INVALID_ARG, "k must be positive");
  if (!x || !y || !dist) return fail(B2D_ERR_INVALID_ARG, "null x / y / dist");
  if (m > 0x7fffffffll * 128 || n > 0x7fffffffll || k > (1 << 24)) return fail(B2D_ERR_INVALID_ARG, "extent too large");
  if (row_major) {
    if (ldx < k || ldy < k || ldd < n) return fail(B2D_ERR_INVALID_ARG, "leading dimension smaller than the row length");
  } else {
    if (ldx < m || ldy < n || ldd < m) return fail(B2D_ERR_INVALID_ARG, "leading dimension smaller than the column length");
  }
  if (!is_expanded(metric) && !is_unexpanded(metric))
    return fail(B2D_ERR_UNSUPPORTED, "metric " + std::to_string(metric) + " is not on the B200 distance path");
  if (dtype != B2D_F32 && dtype != B2D_F16 && dtype != B2D_F64) return fail(B2D_ERR_UNSUPPORTED, "dtype");
  if (metric == B2D_LpUnexpanded && !(metric_arg > 0.f)) return fail(B2D_ERR_INVALID_ARG, "LpUnexpanded needs p > 0");
  if (dtype == B2D_F64) {   // double in / double out: the SIMT fp64 path, every metric
    if (row_major)
      return pairwise_f64(s, metric, static_cast<const double*>(x), ldx, 1, static_cast<const double*>(y), ldy, 1,
                          static_cast<double*>(dist_v), ldd, m, n, k, 0, metric_arg, workspace, workspace_bytes);
    return pairwise_f64(s, metric, static_cast<const double*>(y), 1, ldy, static_cast<const double*>(x), 1, ldx,
                        static_cast<double*>(dist_v), ldd, n, m, k, 1, metric_arg, workspace, workspace_bytes);
  }

  // Fortran order: D^T (row-major [n,m]) = metric(y_j, x_i).  Every metric here is symmetric except
  // KLDivergence, which switches to the kernel with the operand roles exchanged (UX_KL_REV).
  const void *xa = x, *ya = y;
  int64_t ma = m, na = n;
  int64_t xrs, xcs, yrs, ycs;
  if (row_major) { xrs = ldx; xcs = 1; yrs = ldy; ycs = 1; }
  else { xa = y; ya = x; ma = n; na = m; xrs = 1; xcs = ldy; yrs = 1; ycs = ldx; }

  if (is_expanded(metric)) {
    const size_t need = tc_layout(nullptr, ma, na, k, false).bytes;
    if (!workspace || workspace_bytes < need)
      return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
    if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(B2D_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
    TcWorkspace w = tc_layout(workspace, ma, na, k, false);
    int mode = PREP_L2, center = 0, post = POST_NONE, xform = 0;
    float coef_mul = 1.f, tx_const = 0.f;
    if (metric == B2D_L2Expanded) { post = POST_CLAMP; }
    else if (metric == B2D_L2SqrtExpanded) { post = POST_CLAMP_SQRT; }
    else if (metric == B2D_CosineExpanded) { mode = PREP_COSINE; }
    else if (metric == B2D_CorrelationExpanded) { mode = PREP_COSINE; center = 1; }
    else if (metric == B2D_HellingerExpanded) { mode = PREP_INNER; post = POST_CLAMP_SQRT; xform = 1; coef_mul = -1.f; tx_const = 1.f; }
    else if (metric == B2D_RusselRaoExpanded) { mode = PREP_INNER; coef_mul = -1.f / static_cast<float>(k); tx_const = 1.f; }
    else if (metric == B2D_JaccardExpanded) { mode = PREP_INNER_NORM; post = POST_JACCARD; }
    else if (metric == B2D_DiceExpanded) { mode = PREP_INNER_NORM; post = POST_DICE; }
    else { mode = PREP_INNER; }
    int rc = dtype == B2D_F32
               ? launch_prep<float>(s, w, xa, xrs, xcs, ma, ya, yrs, ycs, na, k, nullptr, nullptr, mode, center, xform,
                                    coef_mul, tx_const)
               : launch_prep<__half>(s, w, xa, xrs, xcs, ma, ya, yrs, ycs, na, k, nullptr, nullptr, mode, center, xform,
                                     coef_mul, tx_const);
    if (rc) return rc;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.m = ma; p.n = na; p.dist = dist; p.ldd = ldd;
    p.diag_zero = (post != POST_NONE && x == y && m == n && ldx == ldy && mode == PREP_L2) ? 1 : 0;
    p.pair_ok   = (reinterpret_cast<uintptr_t>(dist) % 8 == 0 && ldd % 2 == 0) ? 1 : 0;
    // k > 320: accumulate K in chunks of 256 columns, each added to dist with a round-to-nearest fp32
    // add -- bounds the truncation bias of long MMA chains (numerics)
    const int nkb_total = static_cast<int>((k + 31) / 32);
    constexpr int kChunk = 8;
    ProfileScope prof(s);  // the main kernel(s) only: the operand preparation above is outside
    if (nkb_total > 10 && post < POST_JACCARD) {  // (the ratio metrics need the whole inner product at once)
      for (int kb0 = 0; kb0 < nkb_total; kb0 += kChunk) {
        const int nk = nkb_total - kb0 < kChunk ? nkb_total - kb0 : kChunk;
        p.acc_mode   = kb0 == 0 ? 1 : (kb0 + nk >= nkb_total ? 3 : 2);
        rc           = launch_tc(s, w, p, k, EPI_STORE, post, kb0, nk);
        if (rc) return rc;
      }
      return B2D_OK;
    }
    return launch_tc(s, w, p, k, EPI_STORE, post);
  }

  if (dtype != B2D_F32) return fail(B2D_ERR_UNSUPPORTED, "unexpanded metrics take fp32 inputs");
  UxParams p;
  memset(&p, 0, sizeof(p));
  p.x = static_cast<const float*>(xa); p.y = static_cast<const float*>(ya); p.dist = dist;
  p.xrs = xrs; p.xcs = xcs; p.yrs = yrs; p.ycs = ycs; p.ldd = ldd; p.m = ma; p.n = na; p.k = static_cast<int>(k);
  p.vec_x = (xcs == 1 && xrs % 4 == 0 && k % 4 == 0 && reinterpret_cast<uintptr_t>(xa) % 16 == 0) ? 1 : 0;
  p.vec_y = (ycs == 1 && yrs % 4 == 0 && k % 4 == 0 && reinterpret_cast<uintptr_t>(ya) % 16 == 0) ? 1 : 0;
  p.p = metric_arg; p.inv_p = metric == B2D_HammingUnexpanded ? 1.f / static_cast<float>(k) : 1.f / metric_arg;
  p.tiles_n = static_cast<int>((na + UX_BN - 1) / UX_BN);
  const int64_t tiles = ((ma + UX_BM - 1) / UX_BM) * p.tiles_n;
  int cc = 0, sms = 0;
  int rc = device_sms(&sms, &cc);
  if (rc) return rc;
  if (cc != 10) return fail(B2D_ERR_CUDA, "raft_b200 requires an sm_100 (B200) device");
  switch (metric) {
    case B2D_L1: return launch_ux_inst<UX_L1>(s, p, tiles);
    case B2D_L2Unexpanded: return launch_ux_inst<UX_L2>(s, p, tiles);
    case B2D_L2SqrtUnexpanded: return launch_ux_inst<UX_L2SQRT>(s, p, tiles);
    case B2D_Linf: return launch_ux_inst<UX_LINF>(s, p, tiles);
    case B2D_Canberra: return launch_ux_inst<UX_CANBERRA>(s, p, tiles);
    case B2D_HammingUnexpanded: return launch_ux_inst<UX_HAMMING>(s, p, tiles);
    case B2D_KLDivergence:
      return row_major ? launch_ux_inst<UX_KL>(s, p, tiles) : launch_ux_inst<UX_KL_REV>(s, p, tiles);
    case B2D_JensenShannon: return launch_ux_inst<UX_JS>(s, p, tiles);
    case B2D_BrayCurtis: return launch_ux_inst<UX_BRAYCURTIS>(s, p, tiles);
    default: return launch_ux_inst<UX_LP>(s, p, tiles);
  }
}

size_t b2d_fused_l2_nn_workspace_bytes(int64_t m, int64_t n, int64_t k)
{
  if (m < 0 || n < 0 || k < 0) return static_cast<size_t>(-1);
  return tc_layout(nullptr, m, n, k, true).bytes;
}

// one chunk of y (prep + search); keys carry the result so far (and act as the rows' bounds)
static int fused_nn_keys_chunk(cudaStream_t s, int64_t* keys, const float* x, int64_t ldx, const float* y, int64_t ldy,
                               const float* xn, const float* yn, int64_t m, int64_t n, int64_t k, int64_t idx_offset,
                               void* workspace, int mode, int center, bool allow_screen, bool have_bounds)
{
  TcWorkspace w = tc_layout(workspace, m, n, k, true);
  const int nkb = static_cast<int>((k + 31) / 32);
  // (k <= 64: the exact kernel is already epilogue-bound, screening would not pay)
  const bool screen = (mode == PREP_L2 || mode == PREP_COSINE) && nkb >= 3 && nkb <= TC_MAX_RES_KB && n >= 16384 && allow_screen;
  const int unit_norm = mode == PREP_COSINE ? 1 : 0;
  // The coarse pass bounds a whole 32-column group at once: c * max_j(acc_ij) + min_j |y_j|^2 (screen_tc.cuh).  That is
  // only tight when the rows of a group have nearly equal norms, so the L2 search packs the chunk in the order of
  // its squared row norms (stable radix sort on the top 24 bits of the fp32 norm: deterministic); the kernels
  // translate a packed position back to the source row where an index leaves them (col_map).  The cosine family
  // needs none of this: its column term is the constant 0.
  const int* col_map = nullptr;
  if (screen && !unit_norm && n > 0) {
    if (n > kNnChunkRows) return fail(B2D_ERR_INVALID_ARG, "internal: screened chunk larger than 2^20 rows");
    nn_sortkey_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(y, ldy, n, static_cast<int>(k), yn, w.sort_key[0], w.sort_val[0]);
    B2D_CUDA(cudaGetLastError());
    cub::DoubleBuffer<unsigned> dk(w.sort_key[0], w.sort_key[1]);
    cub::DoubleBuffer<int> dv(w.sort_val[0], w.sort_val[1]);
    size_t tmp_need = 0;
    B2D_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_need, dk, dv, static_cast<int>(n), 8, 32, s));
    if (tmp_need > kSortTmpBytes) return fail(B2D_ERR_WORKSPACE, "internal: radix sort scratch " + std::to_string(tmp_need) + " bytes");
    B2D_CUDA(cub::DeviceRadixSort::SortPairs(w.sort_tmp, tmp_need, dk, dv, static_cast<int>(n), 8, 32, s));
    col_map = dv.Current();
  }
  int rc = launch_prep<float>(s, w, x, ldx, 1, m, y, ldy, 1, n, k, xn, yn, mode, center, 0, 1.f, 0.f, col_map);
  if (rc) return rc;
  if (n == 0) return B2D_OK;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.m = m; p.n = n; p.keys = reinterpret_cast<long long*>(keys); p.idx_offset = idx_offset;
  p.col_map = col_map;
  if (!screen) return launch_tc(s, w, p, k, EPI_MINLOC, POST_NONE);

  // Screened search (screen_tc.cuh).  With S = 32 and r = index of a 256-row y block modulo S:
  //   A  exact arg-min kernel on the blocks r == 0            -> every row has an upper bound U_i
  //   T  coarse 1-product screen on the SAME blocks (trial)   -> the incumbent again and everything within the margin
  //      of it.  Two jobs: (1) A chooses among near-ties (1e-5 relative) with tensor arithmetic and keeps one column;
  //      the others must reach the exact re-evaluation like any candidate of the other blocks, or the result is only
  //      as good as A's arithmetic; (2) the number of candidates per row is the measure of whether screening pays:
  //      many (data whose norms dwarf the nearest-neighbour distances: the margin is relative to |x||y|) -> the
  //      remaining blocks take the exact kernel
  //   R  coarse screen on the blocks r >= 1                   (if screening pays)
  //   E  exact re-evaluation of every candidate, straight from the fp32 inputs
  //   X  exact kernel on r >= 1                               (only where candidates were dropped -- list overflow --
  //                                                            or where screening was called off)
  // Every launch after T is conditional on a device flag: no host round trip, no wrong answer.
  constexpr int kSel = 32;
  const float tau = g_nn_tau.load(std::memory_order_relaxed);
  unsigned* flags = w.cand_cnt;  // [0] slots [1] overflow [2] go_screen [3] go_exact [4] - [5] candidates after T [6] candidates
  // `have_bounds`: the keys already carry every row's best over earlier chunks / shards (a later chunk of this call, or
  // a call that continues from exchanged keys).  A sample of this chunk would add next to nothing to such a bound, so A
  // is skipped -- T and R then cover every block, and so does X if it has to run.  (Correct for ANY keys: a row
  // without a finite bound makes everything a candidate, the list overflows and X takes over; only slower.)
  p.sel_s = kSel; p.sel_lo = 0; p.sel_hi = 1;
  if (!have_bounds) {
    rc = launch_tc(s, w, p, k, EPI_MINLOC, POST_NONE);
    if (rc) return rc;
  }
  nn_seed_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(reinterpret_cast<long long*>(keys), w.aux, w.xt, w.xlo,
                                                                         w.cand, flags, m, n, idx_offset);
  B2D_CUDA(cudaGetLastError());
  rc = launch_screen(s, w, m, n, k, kSel, 0, 1, flags + 1, nullptr, unit_norm, col_map);
  if (rc) return rc;
  // (tau + 1: the trial finds every row's incumbent again)
  nn_decide_kernel<<<1, 1, 0, s>>>(flags, static_cast<unsigned>(m), tau + 1.f, 1, w.nonuni);
  B2D_CUDA(cudaGetLastError());
  rc = launch_screen(s, w, m, n, k, kSel, 1, kSel, flags + 1, flags + 2, unit_norm, col_map);
  if (rc) return rc;
  nn_decide_kernel<<<1, 1, 0, s>>>(flags, static_cast<unsigned>(m), tau, 2, w.nonuni);
  B2D_CUDA(cudaGetLastError());
  int sms = 0, cc = 0;
  rc = device_sms(&sms, &cc);
  if (rc) return rc;
  nn_exact_kernel<<<sms * 8, 256, 0, s>>>(reinterpret_cast<long long*>(keys), w.cand, w.cand_cnt, w.cand_cap, x, ldx, y,
                                          ldy, static_cast<int>(k), idx_offset, unit_norm, center, static_cast<unsigned>(m), col_map);
  B2D_CUDA(cudaGetLastError());
  p.sel_lo = have_bounds ? 0 : 1; p.sel_hi = kSel; p.run_flag = flags + 3;
  rc = launch_tc(s, w, p, k, EPI_MINLOC, POST_NONE);
  if (rc) return rc;
  return B2D_OK;
}

static int fused_nn_keys(void* stream, int64_t* keys, const float* x, int64_t ldx, const float* y, int64_t ldy,
                         const float* xn, const float* yn, int64_t m, int64_t n, int64_t k, int64_t idx_offset,
                         int init_keys, void* workspace, size_t workspace_bytes, int mode, int center)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m < 0 || n < 0 || k < 0) return fail(B2D_ERR_INVALID_ARG, "negative extent");
  if (m == 0) return B2D_OK;
  if (k == 0) return fail(B2D_ERR_INVALID_ARG, "k must be positive");
  if (!keys || !x || (n > 0 && !y)) return fail(B2D_ERR_INVALID_ARG, "null keys / x / y");
  if (ldx < k || ldy < k) return fail(B2D_ERR_INVALID_ARG, "leading dimension smaller than k");
  if (n + idx_offset > 0xFFFFFFFFll || idx_offset < 0) return fail(B2D_ERR_INVALID_ARG, "index range exceeds 32 bits");
  const size_t need = tc_layout(nullptr, m, n, k, true).bytes;
  if (!workspace || workspace_bytes < need)
    return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(B2D_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  const bool screen_off = g_nn_screen.load(std::memory_order_relaxed) == 0;
  if (init_keys) {
    minloc_init_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(reinterpret_cast<long long*>(keys), m);
    B2D_CUDA(cudaGetLastError());
  }
  // The screened search (below) works through y in chunks of 2^20 rows: the candidate list is sized per
  // chunk, and every chunk starts from the bounds the earlier ones left in the keys.
  const int nkb_all    = static_cast<int>((k + 31) / 32);
  const bool screen_ok = (mode == PREP_L2 || mode == PREP_COSINE) && nkb_all >= 3 && nkb_all <= TC_MAX_RES_KB && !screen_off;
  constexpr int64_t kChunkRows = kNnChunkRows;
  if (!screen_ok || n <= kChunkRows)
    return fused_nn_keys_chunk(s, keys, x, ldx, y, ldy, xn, yn, m, n, k, idx_offset, workspace, mode, center, screen_ok,
                               init_keys == 0);
  for (int64_t off = 0; off < n; off += kChunkRows) {
    const int64_t nc = std::min<int64_t>(kChunkRows, n - off);
    int rc = fused_nn_keys_chunk(s, keys, x, ldx, y + off * ldy, ldy, xn, yn ? yn + off : nullptr, m, nc, k,
                                 idx_offset + off, workspace, mode, center, true, init_keys == 0 || off > 0);
    if (rc) return rc;
  }
  return B2D_OK;
}

int b2d_fused_l2_nn_keys(void* stream, int64_t* keys, const float* x, int64_t ldx, const float* y, int64_t ldy,
                         const float* xn, const float* yn, int64_t m, int64_t n, int64_t k, int64_t idx_offset,
                         int init_keys, void* workspace, size_t workspace_bytes)
{
  return fused_nn_keys(stream, keys, x, ldx, y, ldy, xn, yn, m, n, k, idx_offset, init_keys, workspace,
                       workspace_bytes, PREP_L2, 0);
}

int b2d_fused_l2_nn_finalize(void* stream, b2d_kvp_if* out, const int64_t* keys, int64_t m, int do_sqrt,
                             const void* workspace, size_t workspace_bytes)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m < 0) return fail(B2D_ERR_INVALID_ARG, "negative extent");
  if (m == 0) return B2D_OK;
  if (!out || !keys || !workspace) return fail(B2D_ERR_INVALID_ARG, "null out / keys / workspace");
  if (workspace_bytes < 1024 + static_cast<size_t>(m) * 4) return fail(B2D_ERR_WORKSPACE, "workspace too small");
  // (the keys hold the full distance, so nothing is read back from the workspace any more; the argument stays
  // for ABI stability)
  minloc_finalize_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(
    reinterpret_cast<KvpIF*>(out), reinterpret_cast<const long long*>(keys), m, do_sqrt, 0);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

int b2d_fused_l2_nn(void* stream, b2d_kvp_if* out, const float* x, int64_t ldx, const float* y, int64_t ldy,
                    const float* xn, const float* yn, int64_t m, int64_t n, int64_t k, int do_sqrt, int init_out,
                    void* workspace, size_t workspace_bytes)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m < 0 || n < 0 || k < 0) return fail(B2D_ERR_INVALID_ARG, "negative extent");
  if (m == 0) return B2D_OK;
  if (!out) return fail(B2D_ERR_INVALID_ARG, "null out");
  const size_t need = tc_layout(nullptr, m, n, k, true).bytes;
  if (!workspace || workspace_bytes < need)
    return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  TcWorkspace w = tc_layout(workspace, m, n, k, true);
  int rc = b2d_fused_l2_nn_keys(stream, reinterpret_cast<int64_t*>(w.keys), x, ldx, y, ldy, xn, yn, m, n, k, 0, 1,
                                workspace, workspace_bytes);
  if (rc) return rc;
  minloc_finalize_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(
    reinterpret_cast<KvpIF*>(out), w.keys, m, do_sqrt, init_out ? 0 : 1);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------
// Single-process multi-GPU fusedL2NN: the reference's SNMG pattern -- one ncclComm_t per device,
// created by the caller with a grouped ncclCommInitRank / ncclCommInitAll
// (cpp/include/raft/core/resource/nccl_comm.hpp:43-62, core/device_resources_snmg.hpp:35-154).  NCCL is resolved at
// run time (dlopen of the libnccl.so.2 the process already uses), so the library has no link-time NCCL dependency.
namespace {
struct NcclApi {
  int (*all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*group_start)()                                                          = nullptr;
  int (*group_end)()                                                            = nullptr;
  const char* (*err_string)(int)                                                = nullptr;
  bool ok                                                                       = false;
};
const NcclApi& nccl_api()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.all_reduce  = reinterpret_cast<decltype(api.all_reduce)>(dlsym(h, "ncclAllReduce"));
    api.group_start = reinterpret_cast<decltype(api.group_start)>(dlsym(h, "ncclGroupStart"));
    api.group_end   = reinterpret_cast<decltype(api.group_end)>(dlsym(h, "ncclGroupEnd"));
    api.err_string  = reinterpret_cast<decltype(api.err_string)>(dlsym(h, "ncclGetErrorString"));
    api.ok          = api.all_reduce && api.group_start && api.group_end;
  });
  return api;
}
constexpr int kNcclInt64 = 4, kNcclMin = 3;  // ncclInt64 / ncclMin of nccl.h (stable ABI values since NCCL 2.0)
constexpr int64_t kShardHeadRows = 1 << 15;
}  // namespace

int b2d_fused_l2_nn_multi(int ngpu, const int* devices, void* const* streams, void* const* comms, b2d_kvp_if* const* out,
                          const float* const* x, int64_t ldx, const float* const* y, int64_t ldy, const int64_t* n_shard,
                          const int64_t* idx_offset, int64_t m, int64_t k, int do_sqrt, int64_t* const* keys,
                          void* const* workspace, const size_t* workspace_bytes)
{
  if (ngpu < 1 || ngpu > 64) return fail(B2D_ERR_INVALID_ARG, "ngpu must be in [1, 64]");
  if (!devices || !streams || !out || !x || !y || !n_shard || !idx_offset || !keys || !workspace || !workspace_bytes)
    return fail(B2D_ERR_INVALID_ARG, "null argument array");
  if (ngpu > 1 && !comms) return fail(B2D_ERR_INVALID_ARG, "null comms");
  const NcclApi& nccl = nccl_api();
  if (ngpu > 1 && !nccl.ok) return fail(B2D_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
  int prev = 0;
  B2D_CUDA(cudaGetDevice(&prev));
  // exchange plan: identical on every device (it depends on the SMALLEST shard): a 32768-row head, then sub-chunks
  // growing x4, one all-reduce of the packed keys after each (raft_b200/distance/fused_l2_nn.py: plan_exchanges)
  int64_t smin = n_shard[0];
  for (int g = 1; g < ngpu; ++g) smin = std::min(smin, n_shard[g]);
  std::vector<int64_t> plan;
  if (ngpu > 1 && smin > 0) {
    int64_t left = smin, step = kShardHeadRows;
    while (left > 4 * step && plan.size() < 3) { plan.push_back(step); left -= step; step *= 4; }
    plan.push_back(left);
  } else {
    plan.push_back(smin);
  }
  std::vector<int64_t> done(ngpu, 0);
  int rc = B2D_OK;
  for (size_t c = 0; c < plan.size() && rc == B2D_OK; ++c) {
    const bool last = c + 1 == plan.size();
    for (int g = 0; g < ngpu && rc == B2D_OK; ++g) {
      if (cudaSetDevice(devices[g]) != cudaSuccess) { rc = fail(B2D_ERR_CUDA, "cudaSetDevice failed"); break; }
      const int64_t rows = last ? n_shard[g] - done[g] : std::min(plan[c], n_shard[g] - done[g]);
      rc = b2d_fused_l2_nn_keys(streams[g], keys[g], x[g], ldx, y[g] + done[g] * ldy, ldy, nullptr, nullptr, m, rows, k,
                                idx_offset[g] + done[g], c == 0 ? 1 : 0, workspace[g], workspace_bytes[g]);
      done[g] += rows;
    }
    if (rc == B2D_OK && ngpu > 1) {
      int st = nccl.group_start();
      for (int g = 0; g < ngpu && st == 0; ++g)
        st = nccl.all_reduce(keys[g], keys[g], static_cast<size_t>(m), kNcclInt64, kNcclMin, comms[g],
                             static_cast<cudaStream_t>(streams[g]));
      const int st2 = nccl.group_end();
      if (st == 0) st = st2;
      if (st != 0) rc = fail(B2D_ERR_CUDA, std::string("ncclAllReduce(int64, min): ") + (nccl.err_string ? nccl.err_string(st) : "error"));
    }
  }
  for (int g = 0; g < ngpu && rc == B2D_OK; ++g) {
    if (cudaSetDevice(devices[g]) != cudaSuccess) { rc = fail(B2D_ERR_CUDA, "cudaSetDevice failed"); break; }
    rc = b2d_fused_l2_nn_finalize(streams[g], out[g], keys[g], m, do_sqrt, workspace[g], workspace_bytes[g]);
  }
  cudaSetDevice(prev);
  return rc;
}

int b2d_fused_distance_nn(void* stream, b2d_kvp_if* out, int metric, const float* x, int64_t ldx, const float* y,
                          int64_t ldy, const float* xn, const float* yn, int64_t m, int64_t n, int64_t k, int init_out,
                          void* workspace, size_t workspace_bytes)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (metric == B2D_L2Expanded || metric == B2D_L2SqrtExpanded)
    return b2d_fused_l2_nn(stream, out, x, ldx, y, ldy, xn, yn, m, n, k, metric == B2D_L2SqrtExpanded ? 1 : 0, init_out,
                           workspace, workspace_bytes);
  if (metric != B2D_CosineExpanded && metric != B2D_CorrelationExpanded)
    return fail(B2D_ERR_UNSUPPORTED, "fusedDistanceNN supports L2Expanded, L2SqrtExpanded, CosineExpanded, CorrelationExpanded");
  if (m < 0 || n < 0 || k < 0) return fail(B2D_ERR_INVALID_ARG, "negative extent");
  if (m == 0) return B2D_OK;
  if (!out) return fail(B2D_ERR_INVALID_ARG, "null out");
  const size_t need = tc_layout(nullptr, m, n, k, true).bytes;
  if (!workspace || workspace_bytes < need)
    return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  TcWorkspace w = tc_layout(workspace, m, n, k, true);
  // cosine family: rows are normalised in prep, keys hold -cos, the row term t_x = 1 turns it into 1 - cos
  int rc = fused_nn_keys(stream, reinterpret_cast<int64_t*>(w.keys), x, ldx, y, ldy, nullptr, nullptr, m, n, k, 0, 1,
                         workspace, workspace_bytes, PREP_COSINE, metric == B2D_CorrelationExpanded ? 1 : 0);
  if (rc) return rc;
  minloc_finalize_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(
    reinterpret_cast<KvpIF*>(out), w.keys, m, 0, init_out ? 0 : 1);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------
// Fused brute-force kNN: raft::neighbors::brute_force::knn / fused_l2_knn for the
// L2 metrics without ever writing the m x n matrix (expanded_tc.cuh, EPI_TOPK).

static size_t knn_topk_offset(int64_t m, int64_t n, int64_t k) { return align_up(tc_layout(nullptr, m, n, k, true).bytes, 1024); }

size_t b2d_knn_l2_workspace_bytes(int64_t m, int64_t n, int64_t k, int64_t n_neighbors)
{
  if (m < 0 || n < 0 || k < 0 || n_neighbors < 1 || n_neighbors > KNN_MAX_K) return static_cast<size_t>(-1);
  return knn_topk_offset(m, n, k) + align_up(static_cast<size_t>(m) * n_neighbors * 8, 1024) +
         align_up(static_cast<size_t>(m) * 8, 1024) + align_up(static_cast<size_t>(m) * 4, 1024) + 1024;  // + dropmin, dirty list, counters
}

namespace {
struct KnnCtx {
  cudaStream_t s;
  TcWorkspace w;
  TcParams p;       // m, knn_* set; n / idx_offset per pass
  int64_t k;
  int nkb;
  long long* topk;
  float* thr;
  unsigned* cnt;
  long long* dropmin;   // [m] best key that did not fit into a row's list this pass
  int* dirty;           // [m] rows to redo for the current pass
  unsigned* dirty_cnt;  // [64] one counter per pass
  int pass;
  const float* x; int64_t ldx;
  const float* y; int64_t ldy;
  int family, center;
  int kk;
};
}  // namespace

// one pass over y rows [off, off + width): append, fold into the top-k, redo the rows whose lists overflowed in a way
// that matters (knn_merge_kernel / knn_fix_kernel) -- all on the device: the call only enqueues work
static int knn_pass(KnnCtx& c, int64_t off, int64_t width)
{
  TcWorkspace w2 = c.w;
  w2.yop         = c.w.yop + static_cast<size_t>(off) * c.nkb * 64;
  w2.yt          = c.w.yt + off;
  w2.ysc         = c.w.ysc + off;
  TcParams p     = c.p;
  p.n            = width;
  p.idx_offset   = off;
  int rc         = launch_tc(c.s, w2, p, c.k, EPI_TOPK, POST_NONE);
  if (rc) return rc;
  unsigned* dc = c.dirty_cnt + (c.pass & 63);
  knn_merge_kernel<<<static_cast<unsigned>((c.p.m + 7) / 8), 256, 0, c.s>>>(c.topk, c.p.knn_cand, c.cnt, c.thr, c.dropmin, c.dirty,
                                                                            dc, c.p.m, c.kk);
  B2D_CUDA(cudaGetLastError());
  if (width > KNN_CAP) {  // (a pass of at most KNN_CAP columns cannot overflow)
    int sms = 0, cc = 0;
    rc = device_sms(&sms, &cc);
    if (rc) return rc;
    knn_fix_kernel<<<sms * 2, 256, 0, c.s>>>(c.topk, c.thr, c.dirty, dc, c.x, c.ldx, c.y, c.ldy, static_cast<int>(c.k), c.kk, off,
                                             width, c.family, c.center);
    B2D_CUDA(cudaGetLastError());
  }
  ++c.pass;
  return B2D_OK;
}

int b2d_knn_l2(void* stream, int64_t* out_idx, float* out_dist, const float* x, int64_t ldx, const float* y,
               int64_t ldy, int64_t m, int64_t n, int64_t k, int64_t n_neighbors, int do_sqrt, void* workspace,
               size_t workspace_bytes)
{
  return b2d_knn(stream, out_idx, out_dist, do_sqrt ? B2D_L2SqrtExpanded : B2D_L2Expanded, x, ldx, y, ldy, m, n, k,
                 n_neighbors, workspace, workspace_bytes);
}

int b2d_knn(void* stream, int64_t* out_idx, float* out_dist, int metric, const float* x, int64_t ldx, const float* y,
            int64_t ldy, int64_t m, int64_t n, int64_t k, int64_t n_neighbors, void* workspace, size_t workspace_bytes)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int mode = PREP_L2, center = 0, do_sqrt = 0;
  if (metric == B2D_L2Expanded || metric == B2D_L2Unexpanded) {}
  else if (metric == B2D_L2SqrtExpanded || metric == B2D_L2SqrtUnexpanded) do_sqrt = 1;
  else if (metric == B2D_CosineExpanded) mode = PREP_COSINE;
  else if (metric == B2D_CorrelationExpanded) { mode = PREP_COSINE; center = 1; }
  else return fail(B2D_ERR_UNSUPPORTED, "kNN supports the L2 metrics, CosineExpanded and CorrelationExpanded");
  if (m < 0 || n < 0 || k < 0) return fail(B2D_ERR_INVALID_ARG, "negative extent");
  if (n_neighbors < 1 || n_neighbors > n) return fail(B2D_ERR_INVALID_ARG, "n_neighbors must be in [1, n]");
  if (n_neighbors > KNN_MAX_K) return fail(B2D_ERR_UNSUPPORTED, "n_neighbors > 64 is not supported");
  if (m == 0) return B2D_OK;
  if (k == 0) return fail(B2D_ERR_INVALID_ARG, "k must be positive");
  if (k > 320) return fail(B2D_ERR_UNSUPPORTED, "kNN supports k <= 320 (see DESIGN.md section 3)");
  if (!out_idx || !out_dist || !x || !y) return fail(B2D_ERR_INVALID_ARG, "null out / x / y");
  if (ldx < k || ldy < k) return fail(B2D_ERR_INVALID_ARG, "leading dimension smaller than k");
  if (n > 0xFFFFFFFFll) return fail(B2D_ERR_INVALID_ARG, "index range exceeds 32 bits");
  if (m * static_cast<int64_t>(KNN_CAP) > tc_layout(nullptr, m, n, k, true).cand_cap)
    return fail(B2D_ERR_UNSUPPORTED, "kNN: too many query rows for one call (m * 128 must stay below 2^31): split the queries");
  const size_t need = b2d_knn_l2_workspace_bytes(m, n, k, n_neighbors);
  if (!workspace || workspace_bytes < need)
    return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(B2D_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  KnnCtx c;
  c.s   = s;
  c.w   = tc_layout(workspace, m, n, k, true);
  c.k   = k;
  c.nkb = static_cast<int>((k + 31) / 32);
  c.kk  = static_cast<int>(n_neighbors);
  c.topk     = reinterpret_cast<long long*>(static_cast<char*>(workspace) + knn_topk_offset(m, n, k));
  c.thr      = reinterpret_cast<float*>(c.w.aux);          // [m] floats ...
  c.cnt      = reinterpret_cast<unsigned*>(c.w.aux) + m;   // ... and [m] counters share the aux block (8 B per row)
  {
    char* q     = reinterpret_cast<char*>(c.topk) + align_up(static_cast<size_t>(m) * n_neighbors * 8, 1024);
    c.dropmin   = reinterpret_cast<long long*>(q);
    q += align_up(static_cast<size_t>(m) * 8, 1024);
    c.dirty     = reinterpret_cast<int*>(q);
    q += align_up(static_cast<size_t>(m) * 4, 1024);
    c.dirty_cnt = reinterpret_cast<unsigned*>(q);
  }
  c.pass = 0; c.x = x; c.ldx = ldx; c.y = y; c.ldy = ldy;
  c.family = mode == PREP_COSINE ? 1 : 0; c.center = center;
  int rc = launch_prep<float>(s, c.w, x, ldx, 1, m, y, ldy, 1, n, k, nullptr, nullptr, mode, center);
  if (rc) return rc;
  const int64_t total = m * n_neighbors;
  knn_init_kernel<<<static_cast<unsigned>((std::max<int64_t>(std::max<int64_t>(total, m), 64) + 255) / 256), 256, 0, s>>>(c.topk, c.thr, c.cnt, c.dropmin, c.dirty_cnt, m, c.kk);
  B2D_CUDA(cudaGetLastError());
  memset(&c.p, 0, sizeof(c.p));
  c.p.m = m;
  c.p.knn_thr = c.thr; c.p.knn_cnt = c.cnt; c.p.knn_cand = reinterpret_cast<long long*>(c.w.cand);
  c.p.knn_cap = KNN_CAP; c.p.knn_dropmin = c.dropmin;
  // pass widths KNN_CAP, KNN_CAP, 2 KNN_CAP, 4 KNN_CAP, ...: with the k-th best of the s columns seen so far as
  // threshold, a pass over the next s columns appends ~n_neighbors entries per row on unordered data
  int64_t off = 0;
  while (off < n) {
    const int64_t width = std::min<int64_t>(off == 0 ? KNN_CAP : off, n - off);
    rc = knn_pass(c, off, width);
    if (rc) return rc;
    off += width;
  }
  knn_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(out_idx, out_dist, c.topk, total, do_sqrt);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// ------------------------------------------------------------------------------------------
// raft::stats::silhouette_score on the distance engine (stats.cuh)

namespace {
struct SilLayout {
  size_t counts, offsets, cursor, bad, total, where, ys, slab, pw, bytes;
  int64_t chunk, ld;
  size_t pw_bytes;
};
SilLayout sil_layout(int64_t n, int64_t k, int n_labels, int metric, int64_t chunk_rows)
{
  SilLayout L;
  L.ld = (n + 3) / 4 * 4;
  int64_t c = chunk_rows > 0 ? chunk_rows : (int64_t(1) << 30) / (4 * std::max<int64_t>(L.ld, 1));
  c         = std::max<int64_t>(128, c / 128 * 128);
  L.chunk   = std::min<int64_t>(c, std::max<int64_t>(n, 1));
  size_t off = 0;
  auto take  = [&](size_t b) { size_t o = off; off += align_up(b, 1024); return o; };
  L.counts  = take(static_cast<size_t>(n_labels) * 4);
  L.offsets = take(static_cast<size_t>(n_labels + 1) * 4);
  L.cursor  = take(static_cast<size_t>(n_labels) * 4);
  L.bad     = take(4);
  L.total   = take(8);
  L.where   = take(static_cast<size_t>(n) * 4);
  L.ys      = take(static_cast<size_t>(n) * k * 4);
  L.slab    = take(static_cast<size_t>(L.chunk) * L.ld * 4);
  L.pw_bytes = b2d_pairwise_workspace_bytes(metric, B2D_F32, L.chunk, n, k);
  L.pw      = take(L.pw_bytes == static_cast<size_t>(-1) ? 0 : L.pw_bytes);
  L.bytes   = off;
  return L;
}
}  // namespace

size_t b2d_silhouette_score_workspace_bytes(int64_t n, int64_t k, int n_labels, int metric, int64_t chunk_rows)
{
  if (n < 0 || k < 0 || n_labels < 1 || chunk_rows < 0) return static_cast<size_t>(-1);
  if (b2d_pairwise_workspace_bytes(metric, B2D_F32, 1, 1, k) == static_cast<size_t>(-1)) return static_cast<size_t>(-1);
  return sil_layout(n, k, n_labels, metric, chunk_rows).bytes;
}

int b2d_silhouette_score(void* stream, float* score, float* per_sample, const float* x, int64_t ldx, const int* labels,
                         int64_t n, int64_t k, int n_labels, int metric, float metric_arg, int64_t chunk_rows,
                         void* workspace, size_t workspace_bytes)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n < 0 || k <= 0 || chunk_rows < 0) return fail(B2D_ERR_INVALID_ARG, "bad extent");
  if (!score || !x || !labels) return fail(B2D_ERR_INVALID_ARG, "null score / x / labels");
  if (n_labels < 2 || n_labels > n - 1)
    return fail(B2D_ERR_INVALID_ARG, "silhouette score is not defined for this number of labels (need 2 <= n_labels <= n - 1)");
  if (n_labels > 12000) return fail(B2D_ERR_UNSUPPORTED, "n_labels > 12000");
  if (ldx < k) return fail(B2D_ERR_INVALID_ARG, "leading dimension smaller than k");
  const size_t need = b2d_silhouette_score_workspace_bytes(n, k, n_labels, metric, chunk_rows);
  if (need == static_cast<size_t>(-1)) return fail(B2D_ERR_UNSUPPORTED, "metric " + std::to_string(metric) + " is not on the B200 distance path");
  if (!workspace || workspace_bytes < need)
    return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(B2D_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  const SilLayout L = sil_layout(n, k, n_labels, metric, chunk_rows);
  char* base    = static_cast<char*>(workspace);
  int* counts   = reinterpret_cast<int*>(base + L.counts);
  int* offsets  = reinterpret_cast<int*>(base + L.offsets);
  int* cursor   = reinterpret_cast<int*>(base + L.cursor);
  unsigned* bad = reinterpret_cast<unsigned*>(base + L.bad);
  double* total = reinterpret_cast<double*>(base + L.total);
  int* where    = reinterpret_cast<int*>(base + L.where);
  float* ys     = reinterpret_cast<float*>(base + L.ys);
  float* slab   = reinterpret_cast<float*>(base + L.slab);
  B2D_CUDA(cudaMemsetAsync(base, 0, L.where, s));  // counts, offsets, cursor, bad, total
  const unsigned nb = static_cast<unsigned>((n + 255) / 256);
  sil_count_kernel<<<nb, 256, 0, s>>>(labels, counts, n, n_labels, bad);
  sil_scan_kernel<<<1, 32, 0, s>>>(counts, offsets, cursor, n_labels);
  sil_gather_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(x, ldx, labels, cursor, ys, where, n, static_cast<int>(k), n_labels);
  B2D_CUDA(cudaGetLastError());
  // labels outside [0, n_labels) are detected on the device (no host round trip: the call stays asynchronous);
  // they poison the result: *score and the per-sample values of the offending rows become NaN
  for (int64_t r0 = 0; r0 < n; r0 += L.chunk) {
    const int64_t rows = std::min<int64_t>(L.chunk, n - r0);
    int rc = b2d_pairwise_distance(stream, metric, B2D_F32, x + r0 * ldx, ldx, ys, k, slab, L.ld, rows, n, k, 1, metric_arg,
                                   base + L.pw, L.pw_bytes);
    if (rc) return rc;
    sil_row_kernel<<<static_cast<unsigned>(rows), 256, static_cast<size_t>(n_labels) * 4, s>>>(
      slab, L.ld, r0, rows, labels, counts, offsets, where, n_labels, per_sample, total);
    B2D_CUDA(cudaGetLastError());
  }
  sil_finish_kernel<<<1, 32, 0, s>>>(total, score, n, bad);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

// raft::stats::trustworthiness_score on the kNN + distance engine (stats.cuh)
namespace {
struct TrustLayout {
  size_t penalty, emb_idx, emb_dist, slab, scratch, bytes;
  int64_t chunk, ld;
  size_t knn_bytes, pw_bytes;
};
TrustLayout trust_layout(int64_t n, int64_t m, int64_t d, int n_neighbors, int metric, int64_t batch_rows)
{
  TrustLayout L;
  L.ld = (n + 3) / 4 * 4;
  int64_t c = batch_rows > 0 ? batch_rows : (int64_t(1) << 30) / (4 * std::max<int64_t>(L.ld, 1));
  c         = std::max<int64_t>(128, c / 128 * 128);
  L.chunk   = std::min<int64_t>(c, std::max<int64_t>(n, 1));
  size_t off = 0;
  auto take  = [&](size_t b) { size_t o = off; off += align_up(b, 1024); return o; };
  L.penalty  = take(8);
  L.emb_idx  = take(static_cast<size_t>(n) * (n_neighbors + 1) * 8);
  L.emb_dist = take(static_cast<size_t>(n) * (n_neighbors + 1) * 4);
  L.slab     = take(static_cast<size_t>(L.chunk) * L.ld * 4);
  L.knn_bytes = b2d_knn_l2_workspace_bytes(n, n, d, n_neighbors + 1);
  L.pw_bytes  = b2d_pairwise_workspace_bytes(metric, B2D_F32, L.chunk, n, m);
  const size_t a = L.knn_bytes == static_cast<size_t>(-1) ? 0 : L.knn_bytes;
  const size_t b = L.pw_bytes == static_cast<size_t>(-1) ? 0 : L.pw_bytes;
  L.scratch  = take(std::max(a, b));
  L.bytes    = off;
  return L;
}
}  // namespace

size_t b2d_trustworthiness_score_workspace_bytes(int64_t n, int64_t m, int64_t d, int n_neighbors, int metric,
                                                 int64_t batch_rows)
{
  if (n < 0 || m < 0 || d < 0 || n_neighbors < 1 || n_neighbors + 1 > KNN_MAX_K || batch_rows < 0) return static_cast<size_t>(-1);
  if (b2d_pairwise_workspace_bytes(metric, B2D_F32, 1, 1, m) == static_cast<size_t>(-1)) return static_cast<size_t>(-1);
  return trust_layout(n, m, d, n_neighbors, metric, batch_rows).bytes;
}

int b2d_trustworthiness_score(void* stream, double* score, const float* x, int64_t ldx, const float* x_embedded,
                              int64_t lde, int64_t n, int64_t m, int64_t d, int n_neighbors, int metric,
                              int64_t batch_rows, void* workspace, size_t workspace_bytes)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n < 0 || m <= 0 || d <= 0 || batch_rows < 0) return fail(B2D_ERR_INVALID_ARG, "bad extent");
  if (!score || !x || !x_embedded) return fail(B2D_ERR_INVALID_ARG, "null score / x / x_embedded");
  // the reference uses distance_type for BOTH spaces (trustworthiness_score<math_t, distance_type>,
  // cpp/include/raft/stats/detail/trustworthiness_score.cuh:113-211): the embedded-space neighbours come from
  // the fused kNN, which covers the L2 and the cosine families
  int knn_metric = metric;
  if (metric == B2D_L2Unexpanded) knn_metric = B2D_L2Expanded;
  if (metric == B2D_L2SqrtUnexpanded) knn_metric = B2D_L2SqrtExpanded;
  if (knn_metric != B2D_L2Expanded && knn_metric != B2D_L2SqrtExpanded && knn_metric != B2D_CosineExpanded &&
      knn_metric != B2D_CorrelationExpanded)
    return fail(B2D_ERR_UNSUPPORTED, "trustworthiness_score supports the L2 metrics, CosineExpanded and CorrelationExpanded");
  if (n_neighbors < 1 || n_neighbors + 1 > KNN_MAX_K) return fail(B2D_ERR_UNSUPPORTED, "n_neighbors must be in [1, 63]");
  if (2 * n - 3 * static_cast<int64_t>(n_neighbors) - 1 <= 0 || n_neighbors + 1 > n)
    return fail(B2D_ERR_INVALID_ARG, "n_neighbors must be smaller than n / 2");
  if (ldx < m || lde < d) return fail(B2D_ERR_INVALID_ARG, "leading dimension smaller than the row length");
  const size_t need = b2d_trustworthiness_score_workspace_bytes(n, m, d, n_neighbors, metric, batch_rows);
  if (need == static_cast<size_t>(-1)) return fail(B2D_ERR_UNSUPPORTED, "metric " + std::to_string(metric) + " is not on the B200 distance path");
  if (!workspace || workspace_bytes < need)
    return fail(B2D_ERR_WORKSPACE, "workspace too small: need " + std::to_string(need) + " bytes");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return fail(B2D_ERR_INVALID_ARG, "workspace must be 256-byte aligned");
  const TrustLayout L = trust_layout(n, m, d, n_neighbors, metric, batch_rows);
  char* base = static_cast<char*>(workspace);
  unsigned long long* penalty = reinterpret_cast<unsigned long long*>(base + L.penalty);
  int64_t* emb_idx = reinterpret_cast<int64_t*>(base + L.emb_idx);
  float* emb_dist  = reinterpret_cast<float*>(base + L.emb_dist);
  float* slab      = reinterpret_cast<float*>(base + L.slab);
  const int kk1    = n_neighbors + 1;
  B2D_CUDA(cudaMemsetAsync(penalty, 0, 8, s));
  // neighbours in the embedded space (the sample itself included, as in the reference)
  int rc = b2d_knn(stream, emb_idx, emb_dist, knn_metric, x_embedded, lde, x_embedded, lde, n, n, d, kk1, base + L.scratch,
                   L.knn_bytes);
  if (rc) return rc;
  for (int64_t r0 = 0; r0 < n; r0 += L.chunk) {
    const int64_t rows = std::min<int64_t>(L.chunk, n - r0);
    rc = b2d_pairwise_distance(stream, metric, B2D_F32, x + r0 * ldx, ldx, x, ldx, slab, L.ld, rows, n, m, 1, 2.0f,
                               base + L.scratch, L.pw_bytes);
    if (rc) return rc;
    trust_rank_kernel<<<static_cast<unsigned>(rows), 256, 0, s>>>(slab, L.ld, r0, n, emb_idx, kk1, n_neighbors, penalty);
    B2D_CUDA(cudaGetLastError());
  }
  trust_finish_kernel<<<1, 1, 0, s>>>(penalty, score, n, n_neighbors);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

int b2d_row_argmin(void* stream, int32_t* out, const float* in, int64_t ld, int64_t rows, int64_t n)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows < 0 || n < 0 || ld < n) return fail(B2D_ERR_INVALID_ARG, "bad extents");
  if (n > 0x7fffffffll) return fail(B2D_ERR_INVALID_ARG, "row length exceeds 31 bits");
  if (rows == 0) return B2D_OK;
  if (!out || (!in && n > 0)) return fail(B2D_ERR_INVALID_ARG, "null out / in");
  if (n <= 2048)
    row_argmin_kernel<true><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(out, in, ld, rows, n);
  else
    row_argmin_kernel<false><<<static_cast<unsigned>(rows), 256, 0, s>>>(out, in, ld, rows, n);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

int b2d_row_norm(void* stream, float* out, const float* x, int64_t ldx, int64_t rows, int64_t k, int norm_type,
                 int do_sqrt)
{
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows < 0 || k < 0 || ldx < k) return fail(B2D_ERR_INVALID_ARG, "bad extents");
  if (norm_type < B2D_L0PseudoNorm || norm_type > B2D_LinfNorm) return fail(B2D_ERR_INVALID_ARG, "norm type");
  if (rows == 0) return B2D_OK;
  if (!out || !x) return fail(B2D_ERR_INVALID_ARG, "null out / x");
  row_norm_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(out, x, ldx, rows, static_cast<int>(k),
                                                                        norm_type, do_sqrt);
  B2D_CUDA(cudaGetLastError());
  return B2D_OK;
}

}  // extern "C"
