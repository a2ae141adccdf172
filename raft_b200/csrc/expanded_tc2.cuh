// K1, 2-CTA form: expanded-metric pairwise distances for k <= 128 with the output written through full-width rows.
//
// Why a second kernel: measured (scripts/probes/store_width.cu) the output path of the 1-CTA kernel is
// bound per SM clock by the SHAPE of its store requests -- 32x32 swizzled TMA boxes (128-byte row pieces): 19.5 B/clk/SM;
// 8 full-width rows (8 x 1 KB) per tensor store: 22 -- and the full-width form needs ~66 KB of staging (8 private
// buffers of 8 padded rows), which a CTA holding a resident 256-row y block (128 KB at k = 128) cannot spare.
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256) splits the y block: each CTA keeps only ITS 128 y rows
// (64 KB), the tensor cores of both SMs read both halves, and every CTA still owns a 128 x 256 accumulator, i.e.
// complete 1 KB output rows.  Shared-memory operand reads per MMA drop as well (the y half is read once for both SMs).
//
//   cluster = (CTA 0 "leader", CTA 1), one cluster per SM pair (TPC), persistent: grid = 2 x #clusters
//   work item = (256-row block of y, run of PAIRS of 128-row x tiles); CTA r owns x tile 2*pair + r
//   384 threads = 3 warpgroups; setmaxnreg moves registers from the last one (40 each) to the two epilogue
//   warpgroups (232 each): the epilogue keeps its 64 column terms in registers (see below)
//   warp 8     TMA producer (each CTA loads its own x tile and its own half of the y block; the bytes are
//              signalled on the LEADER's mbarriers: cp.async.bulk.tensor ... cta_group::2)
//   warp 9     leader only: MMA issuer, tcgen05.mma.cta_group::2.kind::f16, M = 256 (128 per CTA), N = 256; commits
//              are multicast to the barriers of both CTAs.  (Both CTAs' warp 9 allocate / free the TMEM columns.)
//   warps 0-7  epilogue, per CTA: warp (q, h) owns tile rows 32q + 16h + [0,16) over all 256 columns; results go
//              through a private buffer of 8 padded full-width rows -- the smem image of a 3-D TMA box {132, 2, 8}:
//              [8 rows][2 chunks of 128 columns][132 floats], the 4 pad floats of a chunk lie outside dimension 0 of
//              the tensor map and are never written -- so the 8-byte stores from the 16x256b fragments are
//              bank-conflict free AND one tensor store moves 8 rows x 1 KB.
// Same numerics as expanded_tc.cuh (3-term fp16 split, cross terms first, fp32 accumulate in TMEM), same operand
// preparation, same per-row / per-column scale handling.  EPI_STORE only, dist 16-byte aligned, n % 4 == 0.
#pragma once
#include "expanded_tc.cuh"

namespace b2d {

constexpr int T2_THREADS   = 384;                     // 8 epilogue warps + producer + MMA issuer + 2 idle (warpgroup granularity)
constexpr int T2_STAGES    = 5;                       // x-tile k-block stages (16 KB each)
constexpr int T2_B_BYTES   = 128 * 128;               // one k-block of this CTA's half of the y block
constexpr int T2_STG_ROW   = 1024 + 32;               // staged row: 2 chunks x (128 + 4) floats
constexpr int T2_STG_BUF   = 8 * T2_STG_ROW;          // 8 rows
constexpr size_t T2_SMEM_B   = (size_t)TC_MAX_RES_KB * T2_B_BYTES;                 // 64 KB
constexpr size_t T2_SMEM_A   = (size_t)T2_STAGES * TC_A_BYTES;                     // 80 KB
constexpr size_t T2_SMEM_STG = (size_t)TC_EPI_WARPS * T2_STG_BUF;                  // 66 KB
constexpr size_t T2_SMEM_BYTES = T2_SMEM_B + T2_SMEM_A + T2_SMEM_STG + TC_BN * 4 + 256;
static_assert(T2_SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA limit");

namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank()
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync()
{
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank)
{
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
  // (default semantics, as CUTLASS's ClusterBarrier::arrive: an explicit .release.cluster costs a cluster-scope
  // fence per arrival -- 10 % of all warp stalls in the first profile; the tcgen05 fence before it orders the TMEM reads)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tiled load into THIS CTA's shared memory, bytes signalled on an mbarrier given by its shared::cluster address
// (the leader's barrier): the cta_group::2 form of the TMA load
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int32_t c0,
                                                int32_t c1, uint64_t policy)
{
  asm volatile(
    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
    " [%0], [%1, {%3, %4}], [%2], %5;"
    :
    : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
      "l"(policy)
    : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result)
{
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr)
{
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// D[tmem, both CTAs] (+)= A * B: M = 256 rows (128 per CTA, each from its own shared memory), B = N rows, half per CTA
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
    :
    : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
    : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar)
{
  asm volatile(
    "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
    "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(
      smem_u32(bar))
    : "memory");
}
}  // namespace ptx

// One (row group rr, column chunk cp) of the epilogue: 16 column pairs of this thread's row out of the two fragments
// r0 / r1 -> d = acc * c_row (* column factor) + (t_x + t_y) -> clamp / diagonal / sqrt -> the warp's staging buffer
// (kFull) or straight to global memory (a y block cut by the right edge of the matrix).  kFast = the block lies inside
// the matrix, no per-column factors, no x == y diagonal: the common case carries no predicated-off instructions
// (the first single-loop version issued twice the instructions: ncu smsp__inst_executed 5.1e8 vs 2.4e8).
template <int kPost, bool kFast, int kRr>
__device__ __forceinline__ void t2_emit(const uint32_t (&r0)[32], const uint32_t (&r1)[32], const uint64_t (&tb2)[16],
                                        uint64_t ta_rr, uint64_t cf_rr, uint8_t* rowp, float* orow, int64_t gi, int64_t gj,
                                        const float* col_sb_cp, bool blk_full, bool ynu, bool diag_zero, int64_t m, int64_t n)
{
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t* R = i < 8 ? r0 : r1;
    const int ii      = i & 7;
    uint64_t a        = kRr == 0 ? pk(R[4 * ii], R[4 * ii + 1]) : pk(R[4 * ii + 2], R[4 * ii + 3]);
    if (!kFast && ynu) {
      const float2 sb = *reinterpret_cast<const float2*>(&col_sb_cp[8 * i]);
      a               = mul2(a, pk(sb.x, sb.y));
    }
    float x0, x1;
    unpk(fma2(a, cf_rr, add2(ta_rr, tb2[i])), x0, x1);
    if (kPost != POST_NONE) {
      x0 = clamp0(x0);
      x1 = clamp0(x1);
      if (!kFast && diag_zero) {
        if (gi == gj + 8 * i) x0 = 0.f;
        if (gi == gj + 8 * i + 1) x1 = 0.f;
      }
      if (kPost == POST_CLAMP_SQRT) {
        asm("sqrt.approx.f32 %0, %1;" : "=f"(x0) : "f"(x0));
        asm("sqrt.approx.f32 %0, %1;" : "=f"(x1) : "f"(x1));
      }
    }
    if (kFast || blk_full) *reinterpret_cast<float2*>(rowp + 32 * i) = make_float2(x0, x1);
    else if (gi < m && gj + 8 * i < n) ptx::st_global_cs_v2(orow + gj + 8 * i, x0, x1);  // cut by the right edge
  }
}

template <int kPost>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
expanded_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_d, const TcParams p)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* b_base = smem;                              // [4] k-blocks of this CTA's 128 y rows
  uint8_t* a_base = smem + T2_SMEM_B;                  // [T2_STAGES] k-block stages of this CTA's x tile
  uint8_t* stg    = a_base + T2_SMEM_A;                // [8] private epilogue buffers
  float* col_sb   = reinterpret_cast<float*>(stg + T2_SMEM_STG);  // [256] per-column scale of this y block (only read when
                                                                  //       some row of y took its own exponent, prep.cuh)
  uint64_t* bars  = reinterpret_cast<uint64_t*>(col_sb + TC_BN);
  uint64_t* afull = bars;                        // [T2_STAGES]   (used in the leader: bytes of both CTAs)
  uint64_t* aempty = bars + T2_STAGES;           // [T2_STAGES]   (both CTAs: multicast commit)
  uint64_t* bfull = bars + 2 * T2_STAGES;        // [4]           (leader)
  uint64_t* bempty = bfull + TC_MAX_RES_KB;      // [4]           (both)
  uint64_t* tfull = bempty + TC_MAX_RES_KB;      // [2]           (both)
  uint64_t* tempty = tfull + 2;                  // [2]           (leader: 16 arrivals, 8 epilogue warps per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp      = threadIdx.x >> 5;
  const int lane      = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < T2_STAGES; ++i) { ptx::mbar_init(&afull[i], 1); ptx::mbar_init(&aempty[i], 1); }
    for (int i = 0; i < TC_MAX_RES_KB; ++i) { ptx::mbar_init(&bfull[i], 1); ptx::mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 2 * TC_EPI_WARPS); }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    ptx::prefetch_tmap(&tmap_d);
  }
  if (warp == 9) ptx::tmem_alloc_2sm<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // barriers of both CTAs initialised and visible before anything remote touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // register budget: 384 threads get 168 each at launch; the epilogue warpgroups need ~200 (64 fragment registers +
  // 64 column terms), the producer / issuer warpgroup next to nothing
  // (the setmaxnreg instructions sit at the top of the role branches below: ptxas allocates per region)

  const int nkb      = p.nkb;
  const int pairs_m  = (p.tiles_m + 1) >> 1;           // pairs of x tiles
  // p.chunk / p.chunks_m / p.n_items are in units of PAIRS here (set by the host for this kernel)

  if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 8) {
    // ================================ TMA producer (both CTAs) =================================
    const uint64_t pol = ptx::policy_evict_last();
    uint32_t a_it = 0, it_local = 0;
    for (int64_t item = cluster_id; item < p.n_items; item += n_clusters, ++it_local) {
      const int n_blk = static_cast<int>(item % p.tiles_n);
      const int ch    = static_cast<int>(item / p.tiles_n);
      const int pt0   = ch * p.chunk;
      const int pt1   = min(pt0 + p.chunk, pairs_m);
      for (int pt = pt0; pt < pt1; ++pt) {
        const int mt = 2 * pt + static_cast<int>(rank);
        for (int kb = 0; kb < nkb; ++kb, ++a_it) {
          if (pt == pt0) {
            ptx::mbar_wait(&bempty[kb], (it_local & 1) ^ 1);
            if (ptx::elect_one()) {
              if (rank == 0) ptx::mbar_expect_tx(&bfull[kb], 2 * T2_B_BYTES);
              ptx::tma_load_2d_2sm(b_base + kb * T2_B_BYTES, &tmap_b, ptx::mapa(ptx::smem_u32(&bfull[kb]), 0), kb * 64,
                                   n_blk * TC_BN + static_cast<int>(rank) * 128, pol);
            }
          }
          const uint32_t s = a_it % T2_STAGES, ph = (a_it / T2_STAGES) & 1;
          ptx::mbar_wait(&aempty[s], ph ^ 1);
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_expect_tx(&afull[s], 2 * TC_A_BYTES);
            ptx::tma_load_2d_2sm(a_base + s * TC_A_BYTES, &tmap_a, ptx::mapa(ptx::smem_u32(&afull[s]), 0), kb * 64,
                                 mt * TC_BM, pol);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 9) {
    if (rank == 0) {
      // ================================ MMA issuer (leader CTA) ===================================
      uint32_t a_it = 0, t_it = 0, it_local = 0;
      const bool has_lo = __ldg(p.has_lo) != 0u;
      constexpr uint32_t idesc = ptx::umma_idesc_f16(256, TC_BN);
      for (int64_t item = cluster_id; item < p.n_items; item += n_clusters, ++it_local) {
        const int ch  = static_cast<int>(item / p.tiles_n);
        const int pt0 = ch * p.chunk;
        const int pt1 = min(pt0 + p.chunk, pairs_m);
        for (int pt = pt0; pt < pt1; ++pt, ++t_it) {
          const uint32_t as = t_it & 1, aph = (t_it >> 1) & 1;
          ptx::mbar_wait(&tempty[as], aph ^ 1);
          ptx::tc_fence_after();
          const uint32_t d = tmem_base + as * TC_BN;
          // pass 1: every cross term of every k-block, while the accumulator is still tiny (expanded_tc.cuh)
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t it = a_it + kb, s = it % T2_STAGES, ph = (it / T2_STAGES) & 1;
            if (pt == pt0) ptx::mbar_wait(&bfull[kb], it_local & 1);
            ptx::mbar_wait(&afull[s], ph);
            ptx::tc_fence_after();
            if (has_lo && ptx::elect_one()) {
              const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(a_base + s * TC_A_BYTES));
              const uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(b_base + kb * T2_B_BYTES));
              ptx::mma_f16_ss_2sm(d, da + 4, db + 0, idesc, kb > 0 ? 1u : 0u);  // lo0 * hi0
              ptx::mma_f16_ss_2sm(d, da + 6, db + 2, idesc, 1u);               // lo1 * hi1
              ptx::mma_f16_ss_2sm(d, da + 0, db + 4, idesc, 1u);               // hi0 * lo0
              ptx::mma_f16_ss_2sm(d, da + 2, db + 6, idesc, 1u);               // hi1 * lo1
            }
            __syncwarp();
          }
          // pass 2: the hi*hi terms; each x stage is released (in both CTAs) as soon as its last MMA is queued
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t it = a_it + kb, s = it % T2_STAGES;
            if (ptx::elect_one()) {
              const uint64_t da = ptx::umma_desc_sw128(ptx::smem_u32(a_base + s * TC_A_BYTES));
              const uint64_t db = ptx::umma_desc_sw128(ptx::smem_u32(b_base + kb * T2_B_BYTES));
              ptx::mma_f16_ss_2sm(d, da + 0, db + 0, idesc, (has_lo || kb > 0) ? 1u : 0u);  // hi0 * hi0
              ptx::mma_f16_ss_2sm(d, da + 2, db + 2, idesc, 1u);                           // hi1 * hi1
              ptx::mma_commit_2sm(&aempty[s]);
              if (pt == pt1 - 1) ptx::mma_commit_2sm(&bempty[kb]);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx::mma_commit_2sm(&tfull[as]);
          __syncwarp();
          a_it += nkb;
        }
      }
    }
  } else if (warp < 8) {
    // ================================ epilogue warps (both CTAs) ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q    = warp & 3;          // TMEM lane quarter this warp may read
    const int g    = warp >> 2;         // 16-row half of the quarter: tile rows 32q + 16g + [0,16)
    const int et   = threadIdx.x;
    const int quad = lane >> 2, tq = lane & 3;
    uint32_t t_it  = 0;
    const float cf = __ldg(p.coef);
    const bool xnu = __ldg(&p.nonuni[0]) != 0u, ynu = __ldg(&p.nonuni[1]) != 0u;
    const uint64_t pol_st     = ptx::policy_evict_first();
    uint8_t* wbuf             = stg + warp * T2_STG_BUF;
    const uint32_t tempty_rem = ptx::mapa(ptx::smem_u32(&tempty[0]), 0);  // the leader's tempty[0] (tempty[1]: + 8 bytes)
    for (int64_t item = cluster_id; item < p.n_items; item += n_clusters) {
      const int n_blk = static_cast<int>(item % p.tiles_n);
      const int ch    = static_cast<int>(item / p.tiles_n);
      const int pt0   = ch * p.chunk;
      const int pt1   = min(pt0 + p.chunk, pairs_m);
      // Column terms of this y block in REGISTERS for the whole item (64 floats per thread: its 2 columns of every
      // 8-column group).  The first version re-read them from shared memory for every fragment: with all 8 row-quads
      // of a warp reading the same 32 bytes per instruction that was 2 x 1024 wavefronts per tile -- as much as the
      // tile's TMA store reads -- on a shared-memory data pipe that ncu showed 95 % busy (tensor-core operand reads
      // 26 %, LSU 69 %: profiles/r02_ncu_pw2cta_summary.txt).
      uint64_t tb2[2][16];
#pragma unroll
      for (int cp = 0; cp < 2; ++cp)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t gj = static_cast<int64_t>(n_blk) * TC_BN + 128 * cp + 8 * i + 2 * tq;
          float2 t         = make_float2(0.f, 0.f);
          if (gj < p.n) t = __ldg(reinterpret_cast<const float2*>(&p.yt[gj]));  // (n % 4 == 0, gj even: the pair is in or out)
          tb2[cp][i] = pk(t.x, t.y);
        }
      if (ynu) {  // rows of y with their own exponent (rare): per-column factors through shared memory
        ptx::bar_sync(1, 32 * TC_EPI_WARPS);
        if (et < TC_BN) {
          const int64_t gj = static_cast<int64_t>(n_blk) * TC_BN + et;
          col_sb[et]       = gj < p.n ? __ldg(&p.ysc[gj]) : 1.f;
        }
        ptx::bar_sync(1, 32 * TC_EPI_WARPS);
      }
      const bool blk_full = static_cast<int64_t>(n_blk + 1) * TC_BN <= p.n;  // else: cut by the right edge -> direct stores

      for (int pt = pt0; pt < pt1; ++pt, ++t_it) {
        const int mt          = 2 * pt + static_cast<int>(rank);
        const uint32_t tb_idx = t_it & 1, tph = (t_it >> 1) & 1;
        const int64_t rbase   = static_cast<int64_t>(mt) * TC_BM + q * 32 + 16 * g;   // first of this warp's 16 rows
        uint64_t ta2[2], cfr2[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int64_t gi = rbase + 8 * rr + quad;
          float rv = 0.f, cr = cf;
          if (gi < p.m) {
            rv = __ldg(&p.xt[gi]);
            if (xnu) cr *= __ldg(&p.xsc[gi]);
          }
          ta2[rr]  = pk(rv, rv);
          cfr2[rr] = pk(cr, cr);
        }
        ptx::mbar_wait(&tfull[tb_idx], tph);
        ptx::tc_fence_after();
        const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32 + 16 * g) << 16) + tb_idx * TC_BN;
        const bool fast = blk_full && !ynu && !p.diag_zero;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int64_t gi = rbase + 8 * rr + quad;
          float* orow      = p.dist + gi * p.ldd;
#pragma unroll
          for (int cp = 0; cp < 2; ++cp) {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_16x256_x8(t_base + 128 * cp, r0);
            ptx::tmem_ld_16x256_x8(t_base + 128 * cp + 64, r1);
            ptx::tmem_ld_wait();
            if (rr == 1 && cp == 1) {  // last read of the accumulator: hand it back to the leader's MMA warp
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(tempty_rem + 8 * tb_idx);
            }
            if (blk_full && cp == 0) {  // the previous store out of this buffer has left it
              if (lane == 0) ptx::tma_store_wait_read();
              __syncwarp();
            }
            const int64_t gj = static_cast<int64_t>(n_blk) * TC_BN + 128 * cp + 2 * tq;
            uint8_t* rowp    = wbuf + quad * T2_STG_ROW + cp * 528 + 8 * tq;
            const float* sbp = col_sb + 128 * cp + 2 * tq;
            const bool dz    = p.diag_zero != 0;
            if (fast) {   // (rr is a constant after unrolling: the dead branch disappears)
              if (rr == 0) t2_emit<kPost, true, 0>(r0, r1, tb2[cp], ta2[0], cfr2[0], rowp, orow, gi, gj, sbp, true, false, false, p.m, p.n);
              else t2_emit<kPost, true, 1>(r0, r1, tb2[cp], ta2[1], cfr2[1], rowp, orow, gi, gj, sbp, true, false, false, p.m, p.n);
            } else {
              if (rr == 0) t2_emit<kPost, false, 0>(r0, r1, tb2[cp], ta2[0], cfr2[0], rowp, orow, gi, gj, sbp, blk_full, ynu, dz, p.m, p.n);
              else t2_emit<kPost, false, 1>(r0, r1, tb2[cp], ta2[1], cfr2[1], rowp, orow, gi, gj, sbp, blk_full, ynu, dz, p.m, p.n);
            }
          }
          if (blk_full) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_3d(&tmap_d, wbuf, 0, 2 * n_blk, static_cast<int32_t>(rbase + 8 * rr), pol_st);
              ptx::tma_store_commit();
            }
          }
        }
      }
    }
    if (lane == 0) ptx::tma_store_wait_all();
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();   // no CTA leaves (or frees TMEM) while its peer may still address its barriers / shared memory
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<512>(tmem_base);
  }
}

}  // namespace b2d
