// K1 / K3: expanded-form metrics on tcgen05 tensor cores (sm_100a).
//
// Replaces the pair loop of raft::distance::pairwise_distance for L2Expanded / L2SqrtExpanded /
// CosineExpanded / CorrelationExpanded / InnerProduct and of raft::distance::fusedL2NN
// (SURVEY.md 8(a2),(a3),(a5); the reference's last implementation was a CUTLASS 3xTF32 mma.sync
// kernel, CHANGELOG.md:1443,1140 -- this is not a port of it).
//
// Numerics: fp32 inputs are pre-split (prep.cuh) into fp16 hi/lo with a power-of-two row scale;
// acc = hi*hi + hi*lo + lo*hi accumulates in fp32 in TMEM (products of fp16 pairs are exact in
// fp32), i.e. ~22 significant bits per operand -- fp32-grade, same idea as the reference's 3xTF32
// but at the 2x higher kind::f16 rate and with 4 B/element staged instead of 8.
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0     TMA producer    cp.async.bulk.tensor 2-D, SWIZZLE_128B, mbarrier complete_tx
//   warp 1     MMA issuer      one thread issues tcgen05.mma.cta_group::1.kind::f16, M128 N128 K16
//   warps 2-5  epilogue of the left  128 columns of every tile  (one warp per TMEM lane quarter)
//   warps 6-9  epilogue of the right 128 columns
// A 128x256 output tile is computed as two 128x128 halves; each half owns two fp32 accumulators
// in TMEM -- `main` (hi*hi) and `cross` (hi*lo + lo*hi) -- so the four 128-column slots fill TMEM's
// 512 columns and the MMA of one half overlaps the epilogue of the other.  Keeping the small
// cross terms out of the big accumulator matters: the tensor core aligns every product to the
// accumulator's exponent and truncates, so the error grows with the number of MMAs that touch a
// large accumulator; this layout leaves 2 per k-block instead of 6 (accuracy).
// Epilogue (thread == output row, tcgen05.ld 32x32b): d = (main+cross) * (a.x*b.x) + (a.y+b.y) on
// packed f32x2 pipes, then
//   EPI_STORE   clamp / sqrt, swizzled st.shared, one TMA tensor store (32x32 box) per warp per
//               32 columns -- whole 128-byte lines, clipped at the matrix edge by the hardware;
//               manual coalesced stores when dist is not 16-byte aligned
//   EPI_MINLOC  per-row running min / arg-min; one packed 64-bit atomicMin per row per half tile,
//               skipped when the row's current global key is already smaller.
// Work item = (256-column block of y, run of 128-row tiles of x).  With k <= 128 the y block
// (both halves, all of K: <= 128 KB) stays resident in shared memory for the whole run and only
// x tiles stream through a 4-stage ring, which cuts L2->SM operand traffic to ~21 B/clk/SM; for
// larger k both operands stream per k-block.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace b2d {

constexpr int TC_BM          = 128;
constexpr int TC_BN          = 256;
constexpr int TC_A_BYTES     = TC_BM * 128;  // one k-block of A: 128 rows x 128 B
constexpr int TC_B_BYTES     = TC_BN * 128;  // one k-block of B: 256 rows x 128 B
constexpr int TC_STAGES      = 4;
constexpr int TC_MAX_RES_KB  = 4;            // resident-B variant: k <= 128
constexpr int TC_EPI_WARPS   = 8;
constexpr int TC_THREADS     = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_STG_FLOATS  = 32 * 32;      // per epilogue warp transpose buffer (4 KB)

enum TcEpilogue : int { EPI_STORE = 0, EPI_MINLOC = 1 };
enum TcPost : int { POST_NONE = 0, POST_CLAMP = 1, POST_CLAMP_SQRT = 2 };

struct TcParams {
  int64_t m, n;
  int nkb;                // k-blocks of 32 source columns
  int tiles_m, tiles_n;   // ceil(m/128), ceil(n/256)
  int chunk;              // m-tiles per work item
  int chunks_m;           // ceil(tiles_m/chunk)
  int64_t n_items;        // tiles_n * chunks_m
  const float2* xvec;     // [m] (a.x, a.y)
  const float2* yvec;     // [n] (b.x, b.y)
  // EPI_STORE
  float* dist;
  int64_t ldd;
  int diag_zero;          // x and y alias: force d(i,i) = 0 (reference: CHANGELOG.md:1057,1213)
  int vec_ok;             // manual path: 16-byte aligned rows -> st.v4
  int st_policy;          // experiment knob: L2 policy of the tensor store (0 first, 1 normal, 2 last)
  // EPI_MINLOC
  long long* keys;        // [m] packed (ordered float bits << 32 | index)
  int64_t idx_offset;
};

constexpr size_t TC_SMEM_OPERANDS = (size_t)TC_MAX_RES_KB * TC_B_BYTES + (size_t)TC_STAGES * TC_A_BYTES;  // 192 KB
static_assert(TC_SMEM_OPERANDS == (size_t)TC_STAGES * (TC_A_BYTES + TC_B_BYTES), "both variants use the same carve");
constexpr size_t TC_SMEM_BYTES = TC_SMEM_OPERANDS + (size_t)TC_EPI_WARPS * TC_STG_FLOATS * 4 + 2 * TC_BN * 4 + 256;

// float -> int whose signed order equals the float order
__device__ __forceinline__ int ordered_bits(float v)
{
  int b = __float_as_int(v);
  return b < 0 ? (b ^ 0x7FFFFFFF) : b;
}

// packed f32x2 helpers (FADD2 / FMUL2 / FFMA2: two fp32 lanes per issue slot)
__device__ __forceinline__ uint64_t pk(float lo, float hi)
{
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk(uint32_t lo, uint32_t hi)
{
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpk(uint64_t v, float& lo, float& hi)
{
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
{
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float min3(float a, float b, float c)
{
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <bool kResident, int kEpi, int kPost, bool kTma>
__global__ void __launch_bounds__(TC_THREADS, 1)
expanded_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_d, const TcParams p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  // SWIZZLE_128B atoms need 1024-byte alignment; the dynamic window starts 1024-aligned (no static
  // shared memory in this kernel).  Checked, not assumed: a misaligned base traps.
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* b_base = smem;  // resident slabs, or per-stage B
  uint8_t* a_base = smem + (kResident ? TC_MAX_RES_KB * TC_B_BYTES : TC_STAGES * TC_B_BYTES);
  float* stg      = reinterpret_cast<float*>(smem + TC_SMEM_OPERANDS);
  float* col_cb   = stg + TC_EPI_WARPS * TC_STG_FLOATS;  // [256] b.x
  float* col_tb   = col_cb + TC_BN;                      // [256] b.y
  uint64_t* bars  = reinterpret_cast<uint64_t*>(col_tb + TC_BN);
  uint64_t* afull = bars;                  // [TC_STAGES]
  uint64_t* aempty = bars + TC_STAGES;     // [TC_STAGES]
  uint64_t* bfull = bars + 2 * TC_STAGES;  // [TC_MAX_RES_KB]
  uint64_t* bempty = bfull + TC_MAX_RES_KB;
  uint64_t* tfull = bempty + TC_MAX_RES_KB;  // [2]  one per half
  uint64_t* tempty = tfull + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { ptx::mbar_init(&afull[i], 1); ptx::mbar_init(&aempty[i], 1); }
    for (int i = 0; i < TC_MAX_RES_KB; ++i) { ptx::mbar_init(&bfull[i], 1); ptx::mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 4); }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (kEpi == EPI_STORE && kTma) ptx::prefetch_tmap(&tmap_d);
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkb = p.nkb;

  if (warp == 0) {
    // ================================ TMA producer =================================
    // The whole warp runs the loop with warp-uniform values and one elected lane issues: inside a
    // divergent `if (lane == 0)` the compiler cannot keep descriptors / barrier addresses in
    // uniform registers and wraps every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST waterfall
    // loop (measured: the MMA issue loop itself became the bottleneck).
    const uint64_t pol = ptx::policy_evict_last();
    uint32_t a_it = 0, it_local = 0;
    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it_local) {
      const int n_blk = static_cast<int>(item % p.tiles_n);
      const int ch    = static_cast<int>(item / p.tiles_n);
      const int mt0   = ch * p.chunk;
      const int mt1   = min(mt0 + p.chunk, p.tiles_m);
      for (int mt = mt0; mt < mt1; ++mt) {
        for (int kb = 0; kb < nkb; ++kb, ++a_it) {
          if (kResident && mt == mt0) {
            ptx::mbar_wait(&bempty[kb], (it_local & 1) ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&bfull[kb], TC_B_BYTES);
              ptx::tma_load_2d(b_base + kb * TC_B_BYTES, &tmap_b, &bfull[kb], kb * 64, n_blk * TC_BN, pol);
            }
          }
          const uint32_t s = a_it % TC_STAGES, ph = (a_it / TC_STAGES) & 1;
          ptx::mbar_wait(&aempty[s], ph ^ 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&afull[s], kResident ? TC_A_BYTES : TC_A_BYTES + TC_B_BYTES);
            ptx::tma_load_2d(a_base + s * TC_A_BYTES, &tmap_a, &afull[s], kb * 64, mt * TC_BM, pol);
            if (!kResident)
              ptx::tma_load_2d(b_base + s * TC_B_BYTES, &tmap_b, &afull[s], kb * 64, n_blk * TC_BN, pol);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    constexpr uint32_t idesc = ptx::umma_idesc_f16(TC_BM, TC_BN / 2);
    uint32_t a_it = 0, t_it = 0, it_local = 0;
    // six K=16 steps of one k-block into one half: cross terms -> d+128, hi*hi -> d
    auto mma_kblock = [&](uint32_t d_half, uint32_t a_addr, uint32_t b_addr, uint32_t acc) {
      const uint64_t da = ptx::umma_desc_sw128(a_addr);
      const uint64_t db = ptx::umma_desc_sw128(b_addr);
      // descriptor start-address units are 16 B; inside the 128-B swizzled row:
      //   hi k[0,16) +0, hi k[16,32) +2, lo k[0,16) +4, lo k[16,32) +6
      ptx::mma_f16_ss(d_half + 128, da + 4, db + 0, idesc, acc);  // lo0 * hi0
      ptx::mma_f16_ss(d_half + 128, da + 6, db + 2, idesc, 1u);   // lo1 * hi1
      ptx::mma_f16_ss(d_half + 128, da + 0, db + 4, idesc, 1u);   // hi0 * lo0
      ptx::mma_f16_ss(d_half + 128, da + 2, db + 6, idesc, 1u);   // hi1 * lo1
      ptx::mma_f16_ss(d_half, da + 0, db + 0, idesc, acc);        // hi0 * hi0
      ptx::mma_f16_ss(d_half, da + 2, db + 2, idesc, 1u);         // hi1 * hi1
    };
    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it_local) {
      const int ch  = static_cast<int>(item / p.tiles_n);
      const int mt0 = ch * p.chunk;
      const int mt1 = min(mt0 + p.chunk, p.tiles_m);
      for (int mt = mt0; mt < mt1; ++mt, ++t_it) {
        const uint32_t tph = t_it & 1;
        if (kResident) {
          // A tile (nkb <= TC_STAGES stages) is consumed twice: half 0, then half 1
          for (int h = 0; h < 2; ++h) {
            ptx::mbar_wait(&tempty[h], tph ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_half = tmem_base + h * 256;
            for (int kb = 0; kb < nkb; ++kb) {
              const uint32_t it = a_it + kb, s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
              if (h == 0) {
                if (mt == mt0) ptx::mbar_wait(&bfull[kb], it_local & 1);
                ptx::mbar_wait(&afull[s], ph);
                ptx::tc_fence_after();
              }
              if (ptx::elect_one()) {
                mma_kblock(d_half, ptx::smem_u32(a_base + s * TC_A_BYTES),
                           ptx::smem_u32(b_base + kb * TC_B_BYTES + h * (TC_B_BYTES / 2)), kb > 0 ? 1u : 0u);
                if (h == 1) {
                  ptx::mma_commit(&aempty[s]);
                  if (mt == mt1 - 1) ptx::mma_commit(&bempty[kb]);
                }
              }
              __syncwarp();
            }
            if (ptx::elect_one()) ptx::mma_commit(&tfull[h]);
            __syncwarp();
          }
          a_it += nkb;
        } else {
          ptx::mbar_wait(&tempty[0], tph ^ 1);
          ptx::mbar_wait(&tempty[1], tph ^ 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < nkb; ++kb, ++a_it) {
            const uint32_t s = a_it % TC_STAGES, ph = (a_it / TC_STAGES) & 1;
            ptx::mbar_wait(&afull[s], ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              for (int h = 0; h < 2; ++h)
                mma_kblock(tmem_base + h * 256, ptx::smem_u32(a_base + s * TC_A_BYTES),
                           ptx::smem_u32(b_base + s * TC_B_BYTES + h * (TC_B_BYTES / 2)), kb > 0 ? 1u : 0u);
              ptx::mma_commit(&aempty[s]);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) {
            ptx::mma_commit(&tfull[0]);
            ptx::mma_commit(&tfull[1]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ epilogue warps ===============================
    const int q        = warp & 3;          // TMEM lane quarter this warp may read
    const int ew       = warp - 2;          // 0..7
    const int h        = ew >> 2;           // which 128-column half of every tile this warp drains
    const int et       = threadIdx.x - 64;  // 0..255
    const int row_in_t = q * 32 + lane;
    float* my_stg      = stg + ew * TC_STG_FLOATS;
    const uint64_t pol_st = p.st_policy == 0 ? ptx::policy_evict_first()
                            : (p.st_policy == 1 ? ptx::policy_evict_normal() : ptx::policy_evict_last());
    uint32_t t_it      = 0;
    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int n_blk = static_cast<int>(item % p.tiles_n);
      const int ch    = static_cast<int>(item / p.tiles_n);
      const int mt0   = ch * p.chunk;
      const int mt1   = min(mt0 + p.chunk, p.tiles_m);
      // per-column epilogue pairs of this y block (shared by every tile of the item)
      ptx::bar_sync(1, 32 * TC_EPI_WARPS);
      {
        const int64_t gj = static_cast<int64_t>(n_blk) * TC_BN + et;
        float2 cv        = make_float2(0.f, kEpi == EPI_MINLOC ? __int_as_float(0x7f800000) : 0.f);
        if (gj < p.n) cv = __ldg(&p.yvec[gj]);
        col_cb[et] = cv.x;
        col_tb[et] = cv.y;
      }
      ptx::bar_sync(1, 32 * TC_EPI_WARPS);

      for (int mt = mt0; mt < mt1; ++mt, ++t_it) {
        const uint32_t tph = t_it & 1;
        const int64_t gi   = static_cast<int64_t>(mt) * TC_BM + row_in_t;
        float2 rv          = make_float2(0.f, 0.f);
        long long cur_key  = 0x7FFFFFFFFFFFFFFFll;
        if (gi < p.m) {
          rv = __ldg(&p.xvec[gi]);
          if (kEpi == EPI_MINLOC) cur_key = *reinterpret_cast<volatile long long*>(&p.keys[gi]);
        }
        const uint64_t ra2 = pk(rv.x, rv.x), ta2 = pk(rv.y, rv.y);
        float best_v = __int_as_float(0x7f800000);
        int best_j   = 0x7fffffff;

        ptx::mbar_wait(&tfull[h], tph);
        ptx::tc_fence_after();

        const uint32_t t_addr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 256;
        uint32_t r[32], rc[32];
        ptx::tmem_ld_32x32(t_addr0, r);
        ptx::tmem_ld_32x32(t_addr0 + 128, rc);
        // fast store path: the whole 32x32 block of this warp is inside the matrix
        const bool rows_in = static_cast<int64_t>(mt) * TC_BM + q * 32 + 31 < p.m;

#pragma unroll 1
        for (int chunk = 0; chunk < 4; ++chunk) {
          const int cbase = h * 128 + chunk * 32;  // first column of this chunk inside the tile
          ptx::tmem_ld_wait();
          if (chunk == 3) {
            // this half's accumulators are fully drained into registers: hand them back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[h]);
          }
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 cb = *reinterpret_cast<const float4*>(&col_cb[cbase + c]);
            const float4 tb = *reinterpret_cast<const float4*>(&col_tb[cbase + c]);
            const uint64_t s0 = add2(pk(r[c], r[c + 1]), pk(rc[c], rc[c + 1]));
            const uint64_t s1 = add2(pk(r[c + 2], r[c + 3]), pk(rc[c + 2], rc[c + 3]));
            const uint64_t m0 = mul2(ra2, pk(cb.x, cb.y));
            const uint64_t m1 = mul2(ra2, pk(cb.z, cb.w));
            uint64_t t0, t1;
            if (kEpi == EPI_STORE) {
              t0 = add2(ta2, pk(tb.x, tb.y));
              t1 = add2(ta2, pk(tb.z, tb.w));
            } else {  // the row-constant |x_i|^2 does not change the arg-min: added once at the end
              t0 = pk(tb.x, tb.y);
              t1 = pk(tb.z, tb.w);
            }
            unpk(fma2(s0, m0, t0), v[c], v[c + 1]);
            unpk(fma2(s1, m1, t1), v[c + 2], v[c + 3]);
          }
          if (chunk < 3) {  // r / rc are dead: prefetch the next 32 columns while this chunk is stored
            ptx::tmem_ld_32x32(t_addr0 + (chunk + 1) * 32, r);
            ptx::tmem_ld_32x32(t_addr0 + (chunk + 1) * 32 + 128, rc);
          }
          if (kEpi == EPI_STORE) {
            const int64_t gj0 = static_cast<int64_t>(n_blk) * TC_BN + cbase;
            if (kPost != POST_NONE) {
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
              if (p.diag_zero && gi >= gj0 && gi < gj0 + 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (gi == gj0 + c) v[c] = 0.f;
              }
              if (kPost == POST_CLAMP_SQRT) {
#pragma unroll
                for (int c = 0; c < 32; ++c) asm("sqrt.approx.f32 %0, %1;" : "=f"(v[c]) : "f"(v[c]));
              }
            }
            // stage the 32x32 block in shared memory, 16-byte chunks XOR-swizzled by (row & 7):
            // conflict-free, and exactly the SWIZZLE_128B pattern the TMA store descriptor expects
            if (kTma) {
              if (lane == 0) ptx::tma_store_wait_read();  // previous chunk's store has left smem
              __syncwarp();
            }
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
              *reinterpret_cast<float4*>(my_stg + lane * 32 + ((c4 ^ (lane & 7)) << 2)) =
                make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            if (kTma) {
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                ptx::tma_store_2d(&tmap_d, my_stg, static_cast<int32_t>(gj0), mt * TC_BM + q * 32, pol_st);
                ptx::tma_store_commit();
              }
            } else {
              __syncwarp();
              const int c4 = lane & 7;
              if (rows_in && p.vec_ok && gj0 + 31 < p.n) {
                float* dst = p.dist + (static_cast<int64_t>(mt) * TC_BM + q * 32 + (lane >> 3)) * p.ldd + gj0 + c4 * 4;
                const int64_t step = 4 * p.ldd;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int rr   = it * 4 + (lane >> 3);
                  const float4 o = *reinterpret_cast<const float4*>(my_stg + rr * 32 + ((c4 ^ (rr & 7)) << 2));
                  ptx::st_global_cs_v4(dst, o);
                  dst += step;
                }
              } else {
#pragma unroll 1
                for (int it = 0; it < 8; ++it) {
                  const int rr      = it * 4 + (lane >> 3);
                  const float4 o    = *reinterpret_cast<const float4*>(my_stg + rr * 32 + ((c4 ^ (rr & 7)) << 2));
                  const int64_t gi2 = static_cast<int64_t>(mt) * TC_BM + q * 32 + rr;
                  const int64_t gj  = gj0 + c4 * 4;
                  if (gi2 < p.m) {
                    float* dst = p.dist + gi2 * p.ldd + gj;
                    if (p.vec_ok && gj + 3 < p.n) {
                      ptx::st_global_cs_v4(dst, o);
                    } else {
                      if (gj < p.n) ptx::st_global_cs(dst, o.x);
                      if (gj + 1 < p.n) ptx::st_global_cs(dst + 1, o.y);
                      if (gj + 2 < p.n) ptx::st_global_cs(dst + 2, o.z);
                      if (gj + 3 < p.n) ptx::st_global_cs(dst + 3, o.w);
                    }
                  }
                }
              }
              __syncwarp();
            }
          } else {
            // min over the chunk with 3-input FMNMX, then (rarely) locate it: ascending scan with
            // '==' keeps the smallest column on ties, strict '<' against the running best keeps the
            // earliest chunk (raft::argmin_op, cpp/include/raft/core/operators.hpp:187-194)
            float mn = min3(v[0], v[1], v[2]);
#pragma unroll
            for (int c = 3; c < 31; c += 2) mn = min3(mn, v[c], v[c + 1]);
            mn = fminf(mn, v[31]);
            if (mn < best_v) {
              best_v = mn;
              int j  = 31;
#pragma unroll
              for (int c = 30; c >= 0; --c)
                if (v[c] == mn) j = c;
              best_j = cbase + j;
            }
          }
        }
        if (kEpi == EPI_MINLOC) {
          if (gi < p.m && best_j != 0x7fffffff) {
            const long long gj  = static_cast<long long>(n_blk) * TC_BN + best_j + p.idx_offset;
            const long long key = (static_cast<long long>(ordered_bits(best_v)) << 32) | (gj & 0xFFFFFFFFll);
            if (key < cur_key) atomicMin(&p.keys[gi], key);
          }
        }
      }
    }
    if (kEpi == EPI_STORE && kTma) {
      if (lane == 0) ptx::tma_store_wait_all();
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// small helper kernels for the NN path

__global__ void minloc_init_kernel(long long* keys, int64_t m)
{
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < m) keys[i] = 0x7FFFFFFFFFFFFFFFll;
}

struct KvpIF {
  int key;
  float value;
};

// packed key -> raft::KeyValuePair<int,float>{argmin, min distance}; adds the row-constant
// |x_i|^2 that the pair loop leaves out, clamps at 0, optional sqrt.
__global__ void minloc_finalize_kernel(KvpIF* out, const long long* keys, const float2* xvec,
                                       int64_t m, int do_sqrt, int merge_existing)
{
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long key = keys[i];
  int s               = static_cast<int>(key >> 32);
  int bits            = s < 0 ? (s ^ 0x7FFFFFFF) : s;
  float v             = __int_as_float(bits);
  float d             = fmaxf(xvec[i].y + v, 0.f);
  if (do_sqrt) d = sqrtf(d);
  KvpIF o;
  o.key   = static_cast<int>(key & 0xFFFFFFFFll);
  o.value = d;
  if (key == 0x7FFFFFFFFFFFFFFFll) {  // no candidate (n == 0)
    o.key   = 0x7fffffff;
    o.value = __int_as_float(0x7f7fffff);
  }
  if (merge_existing) {
    // initOutBuffer == false: reduce into what the caller already has (argmin_op)
    KvpIF e = out[i];
    if ((e.value < o.value) || (e.value == o.value && e.key < o.key)) o = e;
  }
  out[i] = o;
}

}  // namespace b2d
