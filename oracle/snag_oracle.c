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
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = 0; i < n1; ++i) {
    float* o = out + i * n2;
    sim_row(x + i * ldx, yt, n2, d, o);
    if (mode == 1) {
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
