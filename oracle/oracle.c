/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the pairwise-distance / fusedL2NN path.
 *
 * PARITY UNPINNED (see oracle/__init__.py): raft::distance is absent from the reference
 * snapshot (CHANGELOG.md:59); this file restates the metric definitions of SURVEY.md 8(a3)/(a4)
 * -- identical to scipy.spatial.distance.cdist -- one scalar fp64 loop per metric, the same
 * shape as the reference's own "naive kernel in the test file" oracles
 * (cpp/tests/linalg/norm.cu:42-66).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Build: make -C oracle   (gcc -O3 -shared -fPIC) -> oracle/_build/liboracle.so
 * Scalar and single-threaded per call (this image's gcc has no libgomp); callers that want all
 * host cores split the row range [i0,i1) over Python threads (ctypes releases the GIL).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* enum raft::distance::DistanceType, SURVEY.md 8(a1) */
enum { L2Expanded = 0, L2SqrtExpanded = 1, CosineExpanded = 2, L1 = 3, L2Unexpanded = 4,
       L2SqrtUnexpanded = 5, InnerProduct = 6, Linf = 7, Canberra = 8, LpUnexpanded = 9,
       CorrelationExpanded = 10, HellingerExpanded = 12, JensenShannon = 15, HammingUnexpanded = 16,
       KLDivergence = 17, RusselRaoExpanded = 18 };

static double pair_metric(const float* a, const float* b, int64_t k, int metric, double p)
{
  double acc = 0.0;
  switch (metric) {
    case L2Expanded: case L2SqrtExpanded: case L2Unexpanded: case L2SqrtUnexpanded: {
      /* sum (x-y)^2 ; the expanded form xn+yn-2xy is the same number (8(a3)), clamped at 0 */
      for (int64_t i = 0; i < k; ++i) { double d = (double)a[i] - (double)b[i]; acc += d * d; }
      return (metric == L2SqrtExpanded || metric == L2SqrtUnexpanded) ? sqrt(acc) : acc;
    }
    case InnerProduct:
      for (int64_t i = 0; i < k; ++i) acc += (double)a[i] * (double)b[i];
      return acc;
    case CosineExpanded: {
      double an = 0, bn = 0;
      for (int64_t i = 0; i < k; ++i) {
        acc += (double)a[i] * b[i]; an += (double)a[i] * a[i]; bn += (double)b[i] * b[i];
      }
      return 1.0 - acc / (sqrt(an) * sqrt(bn));
    }
    case CorrelationExpanded: {
      double sa = 0, sb = 0, saa = 0, sbb = 0;
      for (int64_t i = 0; i < k; ++i) {
        acc += (double)a[i] * b[i]; sa += a[i]; sb += b[i];
        saa += (double)a[i] * a[i]; sbb += (double)b[i] * b[i];
      }
      double num = (double)k * acc - sa * sb;
      double den = sqrt(((double)k * saa - sa * sa) * ((double)k * sbb - sb * sb));
      return 1.0 - num / den;
    }
    case L1:
      for (int64_t i = 0; i < k; ++i) acc += fabs((double)a[i] - (double)b[i]);
      return acc;
    case Linf:
      for (int64_t i = 0; i < k; ++i) { double d = fabs((double)a[i] - (double)b[i]); if (d > acc) acc = d; }
      return acc;
    case Canberra:
      for (int64_t i = 0; i < k; ++i) {
        double d = fabs((double)a[i] - (double)b[i]), s = fabs((double)a[i]) + fabs((double)b[i]);
        acc += (s == 0.0) ? 0.0 : d / s;      /* 0/0 -> 0 */
      }
      return acc;
    case LpUnexpanded:
      for (int64_t i = 0; i < k; ++i) acc += pow(fabs((double)a[i] - (double)b[i]), p);
      return pow(acc, 1.0 / p);
    /* SURVEY.md 8(f) item 4 metrics ([RECALLED] definitions, see oracle.py) */
    case HellingerExpanded:
      for (int64_t i = 0; i < k; ++i) acc += sqrt((double)a[i]) * sqrt((double)b[i]);
      acc = 1.0 - acc;
      return sqrt(acc > 0.0 ? acc : 0.0);
    case RusselRaoExpanded:
      for (int64_t i = 0; i < k; ++i) acc += (double)a[i] * (double)b[i];
      return ((double)k - acc) / (double)k;
    case HammingUnexpanded:
      for (int64_t i = 0; i < k; ++i) acc += (a[i] != b[i]) ? 1.0 : 0.0;
      return acc / (double)k;
    case KLDivergence:
      for (int64_t i = 0; i < k; ++i)
        if (a[i] != 0.f) acc += (double)a[i] * (log((double)a[i]) - log((double)b[i]));
      return 0.5 * acc;
    case JensenShannon:
      for (int64_t i = 0; i < k; ++i) {
        double mm = 0.5 * ((double)a[i] + (double)b[i]);
        if (a[i] != 0.f) acc += (double)a[i] * (log((double)a[i]) - log(mm));
        if (b[i] != 0.f) acc += (double)b[i] * (log((double)b[i]) - log(mm));
      }
      acc *= 0.5;
      return sqrt(acc > 0.0 ? acc : 0.0);
    default: return NAN;
  }
}

/* dist[i*ldd + j] = metric(x_i, y_j); x:[m,k] y:[n,k] row-major (call shape:
 * cpp/include/raft/stats/detail/silhouette_score.cuh:205-206). */
int oracle_pairwise_distance(const float* x, const float* y, double* dist, int64_t m, int64_t n,
                             int64_t k, int metric, double metric_arg, int64_t i0, int64_t i1)
{
  (void)m;
  for (int64_t i = i0; i < i1; ++i)
    for (int64_t j = 0; j < n; ++j)
      dist[i * n + j] = pair_metric(x + i * k, y + j * k, k, metric, metric_arg);
  return 0;
}

/* fusedL2NN: (argmin_j, min_j) ||x_i-y_j||^2, ties -> smaller j
 * (tie-break law: cpp/include/raft/core/operators.hpp:187-194). */
int oracle_fused_l2_nn(const float* x, const float* y, int32_t* idx, double* val, int64_t m,
                       int64_t n, int64_t k, int do_sqrt, int64_t i0, int64_t i1)
{
  (void)m;
  for (int64_t i = i0; i < i1; ++i) {
    double best = INFINITY; int64_t bj = 0;
    for (int64_t j = 0; j < n; ++j) {
      double d = pair_metric(x + i * k, y + j * k, k, L2Unexpanded, 2.0);
      if (d < best) { best = d; bj = j; }
    }
    idx[i] = (int32_t)bj; val[i] = do_sqrt ? sqrt(best) : best;
  }
  return 0;
}

/* fp32 "reference-style" CPU path used as the timed cpu_baseline ("port"): expanded L2 with
 * fp32 accumulation, row norms precomputed, one thread per row block -- the arithmetic the
 * removed SIMT kernel performed, on host cores. */
int oracle_l2_expanded_f32(const float* x, const float* y, float* dist, int64_t m, int64_t n,
                           int64_t k, int do_sqrt, int64_t i0, int64_t i1)
{
  (void)m;
  for (int64_t i = i0; i < i1; ++i) {
    const float* a = x + i * k;
    float an = 0.f;
    for (int64_t t = 0; t < k; ++t) an += a[t] * a[t];
    for (int64_t j = 0; j < n; ++j) {
      const float* b = y + j * k;
      float dot = 0.f, bn = 0.f;
      for (int64_t t = 0; t < k; ++t) { dot += a[t] * b[t]; bn += b[t] * b[t]; }
      float d = an + bn - 2.f * dot;
      d = d < 0.f ? 0.f : d;
      dist[i * n + j] = do_sqrt ? sqrtf(d) : d;
    }
  }
  return 0;
}
