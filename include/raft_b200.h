/* raft_b200 -- C ABI of the B200-native pairwise-distance / fusedL2NN engine.
 *
 * This is the drop-in boundary: every entry point below is what a RAFT-side binding for the
 * distance path would call.  Reference interfaces replaced (all relative to /root/reference;
 * the distance sources themselves were deleted upstream in raft 26.02, CHANGELOG.md:59, so the
 * citations are the surviving call sites / patterns -- see SURVEY.md section 8(a),(b)):
 *
 *   b2d_pairwise_distance      raft::distance::pairwise_distance(handle, x, y, dist, m, n, k,
 *                              metric, isRowMajor, metric_arg)
 *                              call shape: cpp/include/raft/stats/detail/silhouette_score.cuh:205-206,
 *                              cpp/include/raft/stats/detail/trustworthiness_score.cuh:152-153;
 *                              runtime-ABI pattern: cpp/include/raft_runtime/random/
 *                              rmat_rectangular_generator.hpp:18-34
 *   b2d_fused_l2_nn            raft::distance::fusedL2NN / fusedL2NNMinReduce (out = KeyValuePair
 *                              <int,float>[m]); KVP type cpp/include/raft/core/kvp.hpp:20-62,
 *                              tie-break cpp/include/raft/core/operators.hpp:187-194
 *   b2d_fused_l2_nn_keys /     the per-GPU half and the post-exchange half of the multi-GPU
 *   b2d_fused_l2_nn_finalize   fusedL2NN (db row-sharded, packed min-loc all-reduce:
 *                              comms_t::allreduce(INT64, MIN), cpp/include/raft/core/comms.hpp:335,
 *                              cpp/include/raft/comms/detail/std_comms.hpp:365-374)
 *   b2d_row_norm               raft::linalg::rowNorm / norm<ALONG_ROWS>
 *                              cpp/include/raft/linalg/norm.cuh:50-58,118-147
 *
 * Conventions (same as the reference's, cpp/docs developer_guide.md:396-433): every call only
 * enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises, never
 * allocates device memory (temporaries come from the caller's `workspace`, sized by the
 * *_workspace_bytes query -- the counterpart of the handle's workspace memory resource,
 * cpp/include/raft/core/resource/device_memory_resource.hpp:100-129), is thread-safe, and
 * returns a status code instead of throwing.  There is no CPU fallback: on a machine without
 * an sm_100 GPU every compute entry point returns B2D_ERR_CUDA.
 */
#ifndef RAFT_B200_H_
#define RAFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
enum {
  B2D_OK              = 0,
  B2D_ERR_INVALID_ARG = 1, /* -> raft::logic_error  (cpp/include/raft/core/error.hpp:218-239) */
  B2D_ERR_CUDA        = 2, /* -> raft::cuda_error   (cpp/include/raft/util/cuda_rt_essentials.hpp:23-52) */
  B2D_ERR_UNSUPPORTED = 3,
  B2D_ERR_WORKSPACE   = 4
};

/* raft::distance::DistanceType values */
enum {
  B2D_L2Expanded          = 0,
  B2D_L2SqrtExpanded      = 1,
  B2D_CosineExpanded      = 2,
  B2D_L1                  = 3,
  B2D_L2Unexpanded        = 4,
  B2D_L2SqrtUnexpanded    = 5,
  B2D_InnerProduct        = 6,
  B2D_Linf                = 7,
  B2D_Canberra            = 8,
  B2D_LpUnexpanded        = 9,
  B2D_CorrelationExpanded = 10,
  B2D_JaccardExpanded     = 11, /* 1 - <x,y> / (|x|^2 + |y|^2 - <x,y>): Jaccard distance on indicator data */
  B2D_HellingerExpanded   = 12,
  B2D_BrayCurtis          = 14, /* sum |x - y| / sum |x + y| */
  B2D_JensenShannon       = 15,
  B2D_HammingUnexpanded   = 16,
  B2D_KLDivergence        = 17, /* 0.5 sum x log(x / y), 0 log 0 = 0; +inf where y = 0 < x (the mathematical value; pinned by
                                   tests/golden); operand roles follow (x, y) of the call in row- AND column-major order */
  B2D_RusselRaoExpanded   = 18,
  B2D_DiceExpanded        = 19  /* 1 - 2 <x,y> / (|x|^2 + |y|^2): Dice dissimilarity on indicator data */
};

/* element types of x / y: fp32 (dist fp32), fp16 (fp32 accumulate, dist fp32), fp64 (dist fp64: a SIMT double-precision
 * path for every metric; workspace = per-row statistics, 16 (m + n) bytes) */
enum { B2D_F32 = 0, B2D_F16 = 1, B2D_F64 = 2 };

/* raft::KeyValuePair<int,float> (cpp/include/raft/core/kvp.hpp:20-62) */
typedef struct {
  int32_t key;
  float value;
} b2d_kvp_if;

/* norm types of b2d_row_norm (raft::linalg::NormType, cpp/include/raft/linalg/norm_types.hpp:12-23) */
enum { B2D_L0PseudoNorm = 0, B2D_L1Norm = 1, B2D_L2Norm = 2, B2D_LinfNorm = 3 };

int b2d_version(void);
/* Process-wide tuning / test hooks (nothing on the hot path reads the environment):
 *   "nn_screen"  0 forces the exact arg-min kernel for fusedL2NN (default 1: screened search where it applies)
 *   "nn_tau"     candidates per row above which the trial pass calls screening off (default 6) */
int b2d_set_option(const char* name, double value);
/* Diagnostic -- SYNCHRONISES `stream`: control words of the last screened fusedL2NN chunk in `workspace`:
 * [0] list slots in use (one incumbent per row + the blocks of 64 the screen reserved), [1] list overflow, [2] go_screen,
 * [3] go_exact, [4] (unused), [5] candidates after the trial pass, [6] candidates found by the screen.  `out7`: 7 words. */
int b2d_debug_nn_stats(void* stream, const void* workspace, int64_t m, int64_t n, int64_t k, unsigned* out7);
/* thread-local description of the last non-zero status returned on this thread */
const char* b2d_last_error(void);

/* Bytes of device scratch b2d_pairwise_distance needs for this problem (0 for the unexpanded
 * metrics).  Returns (size_t)-1 for an unsupported metric / dtype. */
size_t b2d_pairwise_workspace_bytes(int metric, int dtype, int64_t m, int64_t n, int64_t k);

/* dist[i,j] = metric(x_i, y_j).  x:[m,k], y:[n,k], dist:[m,n]; row_major != 0: C order with
 * leading dimensions ldx, ldy, ldd (elements, >= k / k / n); row_major == 0: Fortran order
 * (ld >= m / n / m).  x and y may alias.  metric_arg = p of LpUnexpanded. */
int b2d_pairwise_distance(void* stream, int metric, int dtype, const void* x, int64_t ldx,
                          const void* y, int64_t ldy, void* dist /* float*, or double* for B2D_F64 */, int64_t ldd, int64_t m,
                          int64_t n, int64_t k, int row_major, float metric_arg, void* workspace,
                          size_t workspace_bytes);

size_t b2d_fused_l2_nn_workspace_bytes(int64_t m, int64_t n, int64_t k);

/* out[i] = {argmin_j, min_j} ||x_i - y_j||^2 (sqrt != 0: of the Euclidean distance); ties go to
 * the smaller j.  xn / yn: optional precomputed squared row norms (NULL: computed here).
 * init_out == 0: reduce into the existing contents of out (initOutBuffer=false). */
int b2d_fused_l2_nn(void* stream, b2d_kvp_if* out, const float* x, int64_t ldx, const float* y,
                    int64_t ldy, const float* xn, const float* yn, int64_t m, int64_t n, int64_t k,
                    int do_sqrt, int init_out, void* workspace, size_t workspace_bytes);

/* raft::distance::fusedDistanceNN[MinReduce]: the same fused arg-min for
 * metric in {L2Expanded, L2SqrtExpanded, CosineExpanded, CorrelationExpanded}; out[i].value is the
 * distance in that metric.  xn / yn (optional) are squared L2 row norms and are used by the L2 metrics.
 * All four metrics take the screened search described at b2d_fused_l2_nn_keys when 64 < k <= 128. */
int b2d_fused_distance_nn(void* stream, b2d_kvp_if* out, int metric, const float* x, int64_t ldx,
                          const float* y, int64_t ldy, const float* xn, const float* yn, int64_t m,
                          int64_t n, int64_t k, int init_out, void* workspace, size_t workspace_bytes);

/* Multi-GPU building blocks.  keys[i] = (order-preserving bits of ||x_i - y_j||^2 << 32)
 * | (j + idx_offset), reduced with signed 64-bit MIN: over this GPU's y shard here, then across
 * GPUs by the caller's all-reduce(INT64, MIN).  init_keys != 0 resets keys to +max first;
 * init_keys == 0 continues from the keys an earlier call or another GPU left (they also serve as
 * starting bounds of the screened search, so exchanging keys early makes the rest of a shard
 * cheaper -- INTEGRATION.md section 3).  b2d_fused_l2_nn_finalize unpacks (clamp at 0, optional sqrt).
 * For 64 < k <= 128 and n >= 16384 the result is produced by a screened search (y worked through in norm-ordered
 * chunks of 2^20 rows: exact tensor pass on 1/32 of a chunk, 1-product lower-bound screen on all of it, exact fp32
 * re-evaluation of the candidates, device-side fallback to the exact pass): same answer, about a third of the
 * work.  Row indices in the keys are always the caller's (source) rows. */
int b2d_fused_l2_nn_keys(void* stream, int64_t* keys, const float* x, int64_t ldx, const float* y,
                         int64_t ldy, const float* xn, const float* yn, int64_t m, int64_t n,
                         int64_t k, int64_t idx_offset, int init_keys, void* workspace,
                         size_t workspace_bytes);
int b2d_fused_l2_nn_finalize(void* stream, b2d_kvp_if* out, const int64_t* keys, int64_t m,
                             int do_sqrt, const void* workspace, size_t workspace_bytes);

/* Single-process multi-GPU fusedL2NN (the reference's SNMG pattern: one ncclComm_t per device from a
 * grouped ncclCommInitRank / ncclCommInitAll, cpp/include/raft/core/resource/nccl_comm.hpp:43-62,
 * cpp/include/raft/core/device_resources_snmg.hpp:35-154).  Device g holds the replicated queries x[g] [m,k], its
 * database row block y[g] [n_shard[g],k] (global index of its first row: idx_offset[g]), keys[g] [m] and a workspace of
 * b2d_fused_l2_nn_workspace_bytes(m, n_shard[g], k) bytes; streams[g] is a cudaStream_t of device devices[g] and
 * comms[g] the ncclComm_t of that device (NULL array allowed for ngpu == 1).  Work is only enqueued: per exchange
 * step one ncclAllReduce(int64, min) of the m packed keys per device inside a ncclGroupStart/End (a 32768-row head of
 * every shard, then sub-chunks growing x4: the screened search starts each from GLOBAL bounds), then out[g] [m] is
 * written on every device (identical results).  NCCL is resolved at run time from the libnccl.so.2 already loaded. */
int b2d_fused_l2_nn_multi(int ngpu, const int* devices, void* const* streams, void* const* comms, b2d_kvp_if* const* out,
                          const float* const* x, int64_t ldx, const float* const* y, int64_t ldy, const int64_t* n_shard,
                          const int64_t* idx_offset, int64_t m, int64_t k, int do_sqrt, int64_t* const* keys,
                          void* const* workspace, const size_t* workspace_bytes);

/* Fused brute-force kNN for the L2 metrics (replaces raft::neighbors::brute_force::knn /
 * fused_l2_knn, removed with the distance package -- CHANGELOG.md:59-60 -- and the pair
 * pairwise_distance + raft::matrix::select_k, cpp/include/raft/matrix/select_k.cuh:73-106):
 * out_idx[i, 0..n_neighbors) = rows of y nearest to x_i in ascending (distance, index) order,
 * out_dist the squared (do_sqrt != 0: Euclidean) distances.  The m x n matrix is never written.
 * n_neighbors <= 64, k <= 320.  Asynchronous like every other entry point: a pass whose per-row candidate lists overflowed
 * in a way that matters (database ordered by decreasing distance, ...) is repaired on the device.  A slot that
 * cannot be filled (fewer than n_neighbors comparable rows: NaN distances) holds index -1, distance +inf. */
size_t b2d_knn_l2_workspace_bytes(int64_t m, int64_t n, int64_t k, int64_t n_neighbors);
/* the same for metric in {L2Expanded, L2Unexpanded, L2SqrtExpanded, L2SqrtUnexpanded, CosineExpanded,
 * CorrelationExpanded} (workspace: b2d_knn_l2_workspace_bytes) */
int b2d_knn(void* stream, int64_t* out_idx, float* out_dist, int metric, const float* x, int64_t ldx, const float* y,
            int64_t ldy, int64_t m, int64_t n, int64_t k, int64_t n_neighbors, void* workspace,
            size_t workspace_bytes);
int b2d_knn_l2(void* stream, int64_t* out_idx, float* out_dist, const float* x, int64_t ldx, const float* y,
               int64_t ldy, int64_t m, int64_t n, int64_t k, int64_t n_neighbors, int do_sqrt,
               void* workspace, size_t workspace_bytes);

/* raft::stats::silhouette_score (cpp/include/raft/stats/detail/silhouette_score.cuh:186-328), a caller of
 * the distance path: *score (device) = mean over samples of (b - a) / max(a, b);
 * per_sample (device, [n]) optional.  labels: int32 in [0, n_labels).  metric: any metric of
 * b2d_pairwise_distance (the reference's default is L2Unexpanded).  The n x n matrix is produced and
 * consumed in [chunk_rows x n] slabs (0 = about 1 GiB per slab), like the reference's batched variant
 * (detail/batched/silhouette_score.cuh:213-243).  Never synchronises: labels are validated on the device, a label
 * outside [0, n_labels) makes *score (and the per-sample value of that row) NaN. */
size_t b2d_silhouette_score_workspace_bytes(int64_t n, int64_t k, int n_labels, int metric, int64_t chunk_rows);
int b2d_silhouette_score(void* stream, float* score, float* per_sample, const float* x, int64_t ldx,
                         const int* labels, int64_t n, int64_t k, int n_labels, int metric, float metric_arg,
                         int64_t chunk_rows, void* workspace, size_t workspace_bytes);

/* raft::stats::trustworthiness_score (cpp/include/raft/stats/detail/trustworthiness_score.cuh:113-211), the
 * other dangling caller: x [n, m] original space, x_embedded [n, d]; neighbours in the
 * embedded space from the fused kNN and ranks in the original space both under `metric` (the reference's
 * distance_type template parameter; L2 / L2Sqrt (Expanded or Unexpanded), CosineExpanded, CorrelationExpanded),
 * ranks counted over [batch_rows x n] slabs (0 = about 1 GiB).  *score is a DEVICE double; the call never
 * synchronises.  n_neighbors <= 63. */
size_t b2d_trustworthiness_score_workspace_bytes(int64_t n, int64_t m, int64_t d, int n_neighbors, int metric,
                                                 int64_t batch_rows);
int b2d_trustworthiness_score(void* stream, double* score, const float* x, int64_t ldx,
                              const float* x_embedded, int64_t lde, int64_t n, int64_t m, int64_t d,
                              int n_neighbors, int metric, int64_t batch_rows, void* workspace,
                              size_t workspace_bytes);

/* Measurement aid (bench.py's roofline): between b2d_profile_begin(capacity) and b2d_profile_end every
 * b2d_pairwise_distance call on the tensor path records a CUDA-event pair on ITS stream around its main
 * kernel launch(es) (operand preparation excluded).  Nothing synchronises until b2d_profile_end, which
 * waits for the recorded events and returns the durations in call order (milliseconds). */
int b2d_profile_begin(int capacity);
int b2d_profile_end(float* ms, int max_count, int* count);

/* out[r] = norm of row r of x:[rows,k] (L2Norm = sum of squares; do_sqrt applies sqrt_op as
 * fin_op, cpp/include/raft/linalg/norm.cuh:118-147). */
int b2d_row_norm(void* stream, float* out, const float* x, int64_t ldx, int64_t rows, int64_t k,
                 int norm_type, int do_sqrt);

/* raft::matrix::argmin (cpp/include/raft/matrix/argmin.cuh:25-37; cpp/include/raft/matrix/detail/math.cuh:290-343):
 * out[r] = column of the minimum of row r of in:[rows,n] (row pitch ld), ties -> smaller index, a row without any
 * value below +inf (all NaN / +inf) -> 0.  The separate-pass form of the arg-min that fusedL2NN fuses;
 * one read of the matrix. */
int b2d_row_argmin(void* stream, int32_t* out, const float* in, int64_t ld, int64_t rows, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* RAFT_B200_H_ */
