/* oracle/snag_oracle.c — CPU restatement of the SNAG_MMEA alignment-evaluation algorithm.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under snag_b200/ may import, link or execute this file; it is the
 * checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * It follows, operation by operation (citations relative to /root/reference/SNAG_MMEA):
 *   - pairwise_distances          src/utils.py:202-218   d = clamp((xn_i + yn_j) - 2 * x_i.y_j, 0)
 *   - csls_sim                    src/utils.py:417-435   nv1 = mean(topk(sim, k) rows), nv2 = same on sim.t(),
 *                                                        csls = (2*sim.t() - nv1).t() - nv2
 *   - Runner._test ranking loops  main.py:393-429        distance = 1 - csls_sim(1 - distance, k); per row /
 *                                                        column the position of the ground truth in an
 *                                                        ascending sort; top-3 retrieved ids per row
 * with the two places the reference leaves to the library pinned down so that results are reproducible
 * bit for bit on any machine:
 *   (1) the dot product x_i.y_j (torch.mm: accumulation order unspecified) is accumulated in fp64 in
 *       index order and rounded ONCE to fp32 (products of fp32 values are exact in fp64);
 *       ||x||^2 likewise (src/utils.py:210);
 *   (2) torch.mean over the k neighbours is the fp32 sum taken largest-first, divided by (float)k;
 *   (3) torch.sort(stable=False) ties are broken towards the lower index (= a stable sort, what
 *       torch's CPU sort does in practice).
 * Every other step is a single fp32 operation exactly as written in the reference.
 *
 * Parity is pinned against the reference itself: tests/golden/gen_golden.py imports the reference's
 * functions, runs them on seeded inputs and stores their outputs; tests/test_oracle.py checks this
 * file against those vectors.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define JB 256

int oracle_version(void) { return 1; }

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU legs of bench.py ask for all host threads explicitly */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* xn[i] = fl32( sum_k fp64 x_ik^2 )  in index order */
void oracle_norm2(const float* x, int64_t n, int64_t d, int64_t ld, float* xn) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double acc = 0.0;
    const float* r = x + i * ld;
    for (int64_t k = 0; k < d; ++k) acc += (double)r[k] * (double)r[k];
    xn[i] = (float)acc;
  }
}

/* yt[k*n2 + j] = y[j*ld + k] */
static float* transpose(const float* y, int64_t n2, int64_t d, int64_t ld) {
  float* yt = (float*)malloc(sizeof(float) * (size_t)n2 * (size_t)d);
  if (!yt) return NULL;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < d; ++k)
    for (int64_t j = 0; j < n2; ++j) yt[k * n2 + j] = y[j * ld + k];
  return yt;
}

/* one row of S: s[j] = fl32( sum_k fp64 x_k * yt[k][j] ), k ascending */
static void sim_row(const float* xr, const float* yt, int64_t n2, int64_t d, float* s) {
  for (int64_t jb = 0; jb < n2; jb += JB) {
    const int64_t w = (n2 - jb) < JB ? (n2 - jb) : JB;
    double acc[JB];
    for (int64_t j = 0; j < w; ++j) acc[j] = 0.0;
    for (int64_t k = 0; k < d; ++k) {
      const double xk = (double)xr[k];
      const float* yr = yt + k * n2 + jb;
      for (int64_t j = 0; j < w; ++j) acc[j] += xk * (double)yr[j];
    }
    for (int64_t j = 0; j < w; ++j) s[jb + j] = (float)acc[j];
  }
}

/* RB rows of S at once: the same per-element arithmetic as sim_row (one fp64 accumulator per (row, column), k
 * ascending, one rounding), register-blocked RB x CB so that the accumulators stay in registers and every converted
 * y value is used RB times. Purely a speed-up of the checker; results are bit-identical to sim_row. */
#define RB 4
#if defined(__AVX__)
#include <immintrin.h>
#define CB 12
static void sim_rows4(const float* x0, const float* x1, const float* x2, const float* x3, const float* yt, int64_t n2,
                      int64_t d, float* s0, float* s1, float* s2, float* s3) {
  const float* xr[RB] = {x0, x1, x2, x3};
  float* sr[RB] = {s0, s1, s2, s3};
  int64_t jb = 0;
  for (; jb + CB <= n2; jb += CB) {
    __m256d a[RB][3];
    for (int r = 0; r < RB; ++r) for (int c = 0; c < 3; ++c) a[r][c] = _mm256_setzero_pd();
    const float* yr = yt + jb;
    for (int64_t k = 0; k < d; ++k, yr += n2) {
      const __m256d y0 = _mm256_cvtps_pd(_mm_loadu_ps(yr));
      const __m256d y1 = _mm256_cvtps_pd(_mm_loadu_ps(yr + 4));
      const __m256d y2 = _mm256_cvtps_pd(_mm_loadu_ps(yr + 8));
      for (int r = 0; r < RB; ++r) {
        const __m256d p = _mm256_set1_pd((double)xr[r][k]);
        a[r][0] = _mm256_add_pd(a[r][0], _mm256_mul_pd(p, y0));     /* the product is exact in fp64: one rounding, in the add */
        a[r][1] = _mm256_add_pd(a[r][1], _mm256_mul_pd(p, y1));
        a[r][2] = _mm256_add_pd(a[r][2], _mm256_mul_pd(p, y2));
      }
    }
    for (int r = 0; r < RB; ++r)
      for (int c = 0; c < 3; ++c) _mm_storeu_ps(sr[r] + jb + 4 * c, _mm256_cvtpd_ps(a[r][c]));
  }
  for (; jb < n2; ++jb) {
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    for (int64_t k = 0; k < d; ++k) {
      const double yv = (double)yt[k * n2 + jb];
      b0 += (double)x0[k] * yv; b1 += (double)x1[k] * yv; b2 += (double)x2[k] * yv; b3 += (double)x3[k] * yv;
    }
    s0[jb] = (float)b0; s1[jb] = (float)b1; s2[jb] = (float)b2; s3[jb] = (float)b3;
  }
}
#else
static void sim_row(const float* xr, const float* yt, int64_t n2, int64_t d, float* s);
static void sim_rows4(const float* x0, const float* x1, const float* x2, const float* x3, const float* yt, int64_t n2,
                      int64_t d, float* s0, float* s1, float* s2, float* s3) {
  sim_row(x0, yt, n2, d, s0); sim_row(x1, yt, n2, d, s1); sim_row(x2, yt, n2, d, s2); sim_row(x3, yt, n2, d, s3);
}
#endif

/* rows [i0, i1) of S = x . y^T into out (row stride ldo); parallel over groups of RB rows */
static void sim_block(const float* x, int64_t ldx, int64_t i0, int64_t i1, const float* yt, int64_t n2, int64_t d,
                      float* out, int64_t ldo) {
  const int64_t groups = (i1 - i0 + RB - 1) / RB;
#pragma omp parallel for schedule(dynamic, 2)
  for (int64_t gidx = 0; gidx < groups; ++gidx) {
    const int64_t r = i0 + gidx * RB;
    if (r + RB <= i1) {
      sim_rows4(x + r * ldx, x + (r + 1) * ldx, x + (r + 2) * ldx, x + (r + 3) * ldx, yt, n2, d,
                out + (r - i0) * ldo, out + (r + 1 - i0) * ldo, out + (r + 2 - i0) * ldo, out + (r + 3 - i0) * ldo);
    } else {
      for (int64_t i = r; i < i1; ++i) sim_row(x + i * ldx, yt, n2, d, out + (i - i0) * ldo);
    }
  }
}

/* S[i,j] = x_i . y_j   (mode 0)   or   clamp((xn_i + yn_j) - 2 S, 0)   (mode 1, pairwise_distances) */
int oracle_pairwise(const float* x, const float* y, int64_t n1, int64_t n2, int64_t d, int64_t ldx, int64_t ldy,
                    int mode, float* out) {
  float* yt = transpose(y, n2, d, ldy);
  float *xn = NULL, *yn = NULL;
  if (!yt) return -1;
  if (mode == 1) {
    xn = (float*)malloc(sizeof(float) * (size_t)n1);
    yn = (float*)malloc(sizeof(float) * (size_t)n2);
    if (!xn || !yn) { free(yt); free(xn); free(yn); return -1; }
    oracle_norm2(x, n1, d, ldx, xn);
    oracle_norm2(y, n2, d, ldy, yn);
  }
  sim_block(x, ldx, 0, n1, yt, n2, d, out, n2);
  if (mode == 1) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n1; ++i) {
      float* o = out + i * n2;
      const float a = xn[i];
      for (int64_t j = 0; j < n2; ++j) {
        const float t = a + yn[j];
        float dd = t - 2.0f * o[j];
        o[j] = dd > 0.0f ? dd : 0.0f;
      }
    }
  }
  free(yt); free(xn); free(yn);
  return 0;
}

/* keep the k largest of v[0..n) in top[0..k), descending */
static void topk_desc(const float* v, int64_t n, int64_t stride, int k, float* top) {
  int m = 0;
  for (int64_t j = 0; j < n; ++j) {
    const float x = v[j * stride];
    if (m < k) {
      int p = m++;
      while (p > 0 && top[p - 1] < x) { top[p] = top[p - 1]; --p; }
      top[p] = x;
    } else if (x > top[k - 1]) {
      int p = k - 1;
      while (p > 0 && top[p - 1] < x) { top[p] = top[p - 1]; --p; }
      top[p] = x;
    }
  }
}

static float mean_desc(const float* top, int k) {
  float s = 0.0f;
  for (int t = 0; t < k; ++t) s = s + top[t];
  return s / (float)k;
}

/* csls_sim on a materialised similarity matrix (src/utils.py:417-435); out may alias sim */
int oracle_csls_sim(const float* sim, int64_t n1, int64_t n2, int k, float* out, float* nv1_out, float* nv2_out) {
  if (k < 1 || k > n1 || k > n2 || k > 4096) return -2;
  float* nv1 = (float*)malloc(sizeof(float) * (size_t)n1);
  float* nv2 = (float*)malloc(sizeof(float) * (size_t)n2);
  if (!nv1 || !nv2) { free(nv1); free(nv2); return -1; }
#pragma omp parallel
  {
    float* top = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n1; ++i) { topk_desc(sim + i * n2, n2, 1, k, top); nv1[i] = mean_desc(top, k); }
#pragma omp for schedule(static)
    for (int64_t j = 0; j < n2; ++j) { topk_desc(sim + j, n1, n2, k, top); nv2[j] = mean_desc(top, k); }
    free(top);
  }
  if (out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n1; ++i)
      for (int64_t j = 0; j < n2; ++j) {
        const float u = 2.0f * sim[i * n2 + j] - nv1[i];
        out[i * n2 + j] = u - nv2[j];
      }
  }
  if (nv1_out) memcpy(nv1_out, nv1, sizeof(float) * (size_t)n1);
  if (nv2_out) memcpy(nv2_out, nv2, sizeof(float) * (size_t)n2);
  free(nv1); free(nv2);
  return 0;
}

/* Full evaluation of n aligned pairs (x_i <-> y_i), the body of Runner._test from main.py:385 to :429.
 *   rank_l2r[i] = position of column i in the ascending stable sort of row i of `distance`
 *   rank_r2l[j] = position of row j in the ascending stable sort of column j
 *   top3[i][0..3) = indices of the 3 smallest entries of row i (ret1..ret3, main.py:411)
 * Optional outputs (may be NULL): nv1, nv2 [n], g [n] (= distance[i,i]), dist_out [n*n]. */
int oracle_align_eval(const float* x, const float* y, int64_t n, int64_t d, int64_t ldx, int64_t ldy, int use_csls, int k,
                      int32_t* rank_l2r, int32_t* rank_r2l, int32_t* top3, float* nv1_out, float* nv2_out, float* g_out,
                      float* dist_out) {
  float* dist = dist_out ? dist_out : (float*)malloc(sizeof(float) * (size_t)n * (size_t)n);
  if (!dist) return -1;
  int rc = oracle_pairwise(x, y, n, n, d, ldx, ldy, 1, dist);          /* main.py:386 */
  if (rc) goto done;
  if (use_csls) {                                                       /* main.py:392-393 */
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n * n; ++e) dist[e] = 1.0f - dist[e];
    rc = oracle_csls_sim(dist, n, n, k, dist, nv1_out, nv2_out);
    if (rc) goto done;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n * n; ++e) dist[e] = 1.0f - dist[e];
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {                                     /* main.py:400-411 */
    const float* r = dist + i * n;
    const float g = r[i];
    int32_t c = 0;
    for (int64_t j = 0; j < n; ++j) c += (r[j] < g) || (r[j] == g && j < i);
    rank_l2r[i] = c;
    if (g_out) g_out[i] = g;
    if (top3) {
      float bv[3] = {INFINITY, INFINITY, INFINITY};
      int32_t bi[3] = {-1, -1, -1};
      for (int64_t j = 0; j < n; ++j) {
        const float v = r[j];
        if (v < bv[2] || bi[2] < 0) {          /* strict <: an equal value with a higher index never displaces */
          int p = 2;
          while (p > 0 && (bi[p - 1] < 0 || v < bv[p - 1])) { bv[p] = bv[p - 1]; bi[p] = bi[p - 1]; --p; }
          bv[p] = v; bi[p] = (int32_t)j;
        }
      }
      top3[i * 3 + 0] = bi[0]; top3[i * 3 + 1] = bi[1]; top3[i * 3 + 2] = bi[2];
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {                                     /* main.py:422-429 */
    const float g = dist[j * n + j];
    int32_t c = 0;
    for (int64_t i = 0; i < n; ++i) { const float v = dist[i * n + j]; c += (v < g) || (v == g && i < j); }
    rank_r2l[j] = c;
  }
done:
  if (!dist_out) free(dist);
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Streaming forms of the same evaluation, for sizes whose n x n matrix does not fit in host memory. The arithmetic
 * of every element is the one of oracle_align_eval (same dot product, same fp32 chain); only the order in which the
 * elements are visited changes, and the selections / counts below do not depend on it:
 *   - the k largest values of a row / column form the same multiset whatever the visiting order, and mean_desc sums
 *     them largest first;
 *   - a rank is a count of elements with (value, index) below the ground truth's.
 * ---------------------------------------------------------------------------------------------------------------- */

/* insert x into a descending list of length k (the first *m entries are filled) */
static inline void topk_push(float* top, int* m, int k, float x) {
  if (*m < k) {
    int p = (*m)++;
    while (p > 0 && top[p - 1] < x) { top[p] = top[p - 1]; --p; }
    top[p] = x;
  } else if (x > top[k - 1]) {
    int p = k - 1;
    while (p > 0 && top[p - 1] < x) { top[p] = top[p - 1]; --p; }
    top[p] = x;
  }
}

/* c = 1 - clamp((xn + yn) - 2 s, 0)   (src/utils.py:210-218 then main.py:393's `1 - distance`) */
static inline float c_from_dot(float s, float xn, float yn) {
  const float t = xn + yn;
  float dd = t - 2.0f * s;
  dd = dd > 0.0f ? dd : 0.0f;
  return 1.0f - dd;
}
/* dist = 1 - ((2 c - nv1) - nv2)      (src/utils.py:433-434, main.py:393) */
static inline float dist_from_c(float c, float nv1, float nv2) {
  const float u = 2.0f * c - nv1;
  const float v = u - nv2;
  return 1.0f - v;
}

/* Runner._test (main.py:385-429) for n aligned pairs without materialising the matrix: rows are processed in blocks
 * of `block_rows`; pass 1 collects the CSLS neighbourhood means, pass 2 the ranks. Outputs as oracle_align_eval
 * (no top-3, no matrix). */
int oracle_align_eval_stream(const float* x, const float* y, int64_t n, int64_t d, int64_t ldx, int64_t ldy, int use_csls,
                             int k, int64_t block_rows, int32_t* rank_l2r, int32_t* rank_r2l, float* nv1_out,
                             float* nv2_out, float* g_out) {
  if (use_csls && (k < 1 || k > n || k > 4096)) return -2;
  if (block_rows < RB) block_rows = RB;
  float* yt = transpose(y, n, d, ldy);
  float* xn = (float*)malloc(sizeof(float) * (size_t)n);
  float* yn = (float*)malloc(sizeof(float) * (size_t)n);
  float* nv1 = (float*)calloc((size_t)n, sizeof(float));
  float* nv2 = (float*)calloc((size_t)n, sizeof(float));
  float* g = (float*)malloc(sizeof(float) * (size_t)n);
  float* blk = (float*)malloc(sizeof(float) * (size_t)block_rows * (size_t)n);
  float* ctop = NULL;
  int* cm = NULL;
  int rc = 0;
  if (!yt || !xn || !yn || !nv1 || !nv2 || !g || !blk) { rc = -1; goto done; }
  oracle_norm2(x, n, d, ldx, xn);
  oracle_norm2(y, n, d, ldy, yn);
  if (use_csls) {
    ctop = (float*)malloc(sizeof(float) * (size_t)n * (size_t)k);
    cm = (int*)calloc((size_t)n, sizeof(int));
    if (!ctop || !cm) { rc = -1; goto done; }
    for (int64_t i0 = 0; i0 < n; i0 += block_rows) {                       /* pass 1: neighbourhoods */
      const int64_t i1 = i0 + block_rows < n ? i0 + block_rows : n;
      sim_block(x, ldx, i0, i1, yt, n, d, blk, n);
#pragma omp parallel
      {
        float* top = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for schedule(static)
        for (int64_t i = i0; i < i1; ++i) {
          float* r = blk + (i - i0) * n;
          int m = 0;
          for (int64_t j = 0; j < n; ++j) { r[j] = c_from_dot(r[j], xn[i], yn[j]); topk_push(top, &m, k, r[j]); }
          nv1[i] = mean_desc(top, k);
        }
#pragma omp for schedule(static)
        for (int64_t j = 0; j < n; ++j) {
          float* t = ctop + j * k;
          int m = cm[j];
          for (int64_t i = i0; i < i1; ++i) topk_push(t, &m, k, blk[(i - i0) * n + j]);
          cm[j] = m;
        }
        free(top);
      }
    }
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n; ++j) nv2[j] = mean_desc(ctop + j * k, k);
  }
  /* ground-truth distances: the diagonal elements, with the arithmetic of every other element
   * (without CSLS the ranked quantity is the squared distance d itself, main.py:386,392) */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double acc = 0.0;
    const float* a = x + i * ldx;
    const float* b = y + i * ldy;
    for (int64_t kk = 0; kk < d; ++kk) acc += (double)a[kk] * (double)b[kk];
    if (use_csls) {
      g[i] = dist_from_c(c_from_dot((float)acc, xn[i], yn[i]), nv1[i], nv2[i]);
    } else {
      const float t = xn[i] + yn[i];
      const float dd = t - 2.0f * (float)acc;
      g[i] = dd > 0.0f ? dd : 0.0f;
    }
  }
  for (int64_t j = 0; j < n; ++j) rank_r2l[j] = 0;
  for (int64_t i0 = 0; i0 < n; i0 += block_rows) {                         /* pass 2: ranks */
    const int64_t i1 = i0 + block_rows < n ? i0 + block_rows : n;
    sim_block(x, ldx, i0, i1, yt, n, d, blk, n);
#pragma omp parallel
    {
#pragma omp for schedule(static)
      for (int64_t i = i0; i < i1; ++i) {
        float* r = blk + (i - i0) * n;
        const float gi = g[i];
        int32_t c = 0;
        for (int64_t j = 0; j < n; ++j) {
          float v;
          if (use_csls) v = dist_from_c(c_from_dot(r[j], xn[i], yn[j]), nv1[i], nv2[j]);
          else { const float t = xn[i] + yn[j]; v = t - 2.0f * r[j]; v = v > 0.0f ? v : 0.0f; }
          r[j] = v;
          c += (v < gi) || (v == gi && j < i);
        }
        rank_l2r[i] = c;
      }
#pragma omp for schedule(static)
      for (int64_t j = 0; j < n; ++j) {
        const float gj = g[j];
        int32_t c = 0;
        for (int64_t i = i0; i < i1; ++i) { const float v = blk[(i - i0) * n + j]; c += (v < gj) || (v == gj && i < j); }
        rank_r2l[j] += c;
      }
    }
  }
  if (nv1_out) memcpy(nv1_out, nv1, sizeof(float) * (size_t)n);
  if (nv2_out) memcpy(nv2_out, nv2, sizeof(float) * (size_t)n);
  if (g_out) memcpy(g_out, g, sizeof(float) * (size_t)n);
done:
  free(yt); free(xn); free(yn); free(nv1); free(nv2); free(g); free(blk); free(ctop); free(cm);
  return rc;
}

/* Audit of m selected entities of one side against ALL n entities of the other side (bench.py's sampled parity audit
 * at sizes where even the streaming evaluation would take hours on a CPU).
 *   xs [m, d]      the selected rows (sources when transposed = 0, targets when transposed = 1)
 *   y  [n, d]      the whole other side
 *   self_id [m]    pair id of every selected entity (its ground-truth partner is row self_id of y)
 *   nv_other [n]   CSLS neighbourhood means of the other side (nv2 when transposed = 0, nv1 when transposed = 1), as
 *                  produced by the implementation under audit; ignored without CSLS
 * Outputs: nv_self [m] (mean of the k largest c over all n partners), g [m] (ground-truth distance), rank [m]
 * (position of the ground truth in the stable ascending sort of the entity's row / column of `distance`).
 * The CSLS chain of main.py:393 / src/utils.py:433-434 is not symmetric in its two means — (2c - nv1_i) - nv2_j —
 * so `transposed` selects which of the two is the selected entity's own. */
int oracle_audit(const float* xs, int64_t m, const float* y, int64_t n, int64_t d, int64_t ldxs, int64_t ldy,
                 const int64_t* self_id, int use_csls, int k, const float* nv_other, int transposed, float* nv_self,
                 float* g, int32_t* rank) {
  if (use_csls && (k < 1 || k > n || k > 4096)) return -2;
  float* yt = transpose(y, n, d, ldy);
  float* xn = (float*)malloc(sizeof(float) * (size_t)m);
  float* yn = (float*)malloc(sizeof(float) * (size_t)n);
  const int64_t bm = 64;
  float* blk = (float*)malloc(sizeof(float) * (size_t)bm * (size_t)n);
  int rc = 0;
  if (!yt || !xn || !yn || !blk) { rc = -1; goto done; }
  oracle_norm2(xs, m, d, ldxs, xn);
  oracle_norm2(y, n, d, ldy, yn);
  for (int64_t i0 = 0; i0 < m; i0 += bm) {
    const int64_t i1 = i0 + bm < m ? i0 + bm : m;
    sim_block(xs, ldxs, i0, i1, yt, n, d, blk, n);
#pragma omp parallel
    {
      float* top = (float*)malloc(sizeof(float) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 1)
      for (int64_t i = i0; i < i1; ++i) {
        float* r = blk + (i - i0) * n;
        const int64_t me = self_id[i];
        float nvs = 0.0f;
        if (use_csls) {
          int mm = 0;
          /* fp32 addition is commutative, so xn + yn is the same whichever side is "x" */
          for (int64_t j = 0; j < n; ++j) { r[j] = c_from_dot(r[j], xn[i], yn[j]); topk_push(top, &mm, k, r[j]); }
          nvs = mean_desc(top, k);
          for (int64_t j = 0; j < n; ++j)
            r[j] = transposed ? dist_from_c(r[j], nv_other[j], nvs) : dist_from_c(r[j], nvs, nv_other[j]);
        } else {
          for (int64_t j = 0; j < n; ++j) { const float t = xn[i] + yn[j]; float v = t - 2.0f * r[j]; r[j] = v > 0.0f ? v : 0.0f; }
        }
        const float gi = r[me];
        int32_t c = 0;
        for (int64_t j = 0; j < n; ++j) c += (r[j] < gi) || (r[j] == gi && j < me);
        nv_self[i] = nvs;
        g[i] = gi;
        rank[i] = c;
      }
      free(top);
    }
  }
done:
  free(yt); free(xn); free(yn); free(blk);
  return rc;
}

/* --distance 1 (main.py:387-390): scipy.spatial.distance.cdist(..., metric="cityblock") on the fp32 embeddings, result
 * rounded to fp32 by torch.FloatTensor. out[i,j] = fl32( sum_k |(double)x_ik - (double)y_jk| ), k ascending. */
int oracle_l1_distance(const float* x, const float* y, int64_t n1, int64_t n2, int64_t d, int64_t ldx, int64_t ldy, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n1; ++i)
    for (int64_t j = 0; j < n2; ++j) {
      double acc = 0.0;
      const float* a = x + i * ldx;
      const float* b = y + j * ldy;
      for (int64_t k = 0; k < d; ++k) acc += fabs((double)a[k] - (double)b[k]);
      out[i * n2 + j] = (float)acc;
    }
  return 0;
}

/* ranks of the diagonal on a materialised distance matrix (the loops of main.py:400-411, 422-429 with a stable sort) */
int oracle_matrix_rank(const float* dist, int64_t n, int32_t* rank_l2r, int32_t* rank_r2l) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float g = dist[i * n + i];
    int32_t c = 0, cc = 0;
    for (int64_t j = 0; j < n; ++j) {
      const float v = dist[i * n + j];
      c += (v < g) || (v == g && j < i);
      const float w = dist[j * n + i];
      cc += (w < g) || (w == g && j < i);
    }
    rank_l2r[i] = c;
    rank_r2l[i] = cc;
  }
  return 0;
}
