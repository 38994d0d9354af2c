/* lq_oracle.c -- plain-C restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY:
 * used by tests/ (cross-check against oracle/lq_oracle.py and the golden fixtures) and by
 * bench.py's cpu_baseline / --impl reference arm.  The product never links or calls it.
 *
 * PARITY STATUS: see the header of oracle/lq_oracle.py -- unpinned against TensorFlow
 * binaries, pinned against the reference's own source executed on a TF-API shim.
 *
 * Like the reference (a chain of TensorFlow ops, each a separate pass over the tensor with
 * a materialised result), every reference op is ONE loop writing ONE temporary, in the
 * reference's order; loops are OpenMP-parallel the way TF's CPU kernels shard an
 * elementwise op over its thread pool.  Citations: NQ = IMAGENETTE/nested_quantization_layer/
 * custom_components/custom_layers.py, LF = IMAGENETTE/custom_loss_terms/custom_components/
 * custom_loss_functions.py (relative to the reference root).
 *
 * Layout: P viewed as [outer, C, inner], group(idx) = (idx / inner) % C, s has C entries.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS_F32 1.1920928955078125e-07f /* NQ:11 np.finfo(np.float32).eps */

int lqo_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline float scale_of(const float* s, int64_t idx, int64_t C, int64_t inner) {
  return s[(idx / inner) % C];
}

/* NQ:56-60: three ops, three passes. q and y are outputs (q is what TF keeps alive for the bwd). */
void lqo_fq_forward(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner,
                    float* nonrounded /* scratch N */, float* q, float* y) {
  const int64_t n = outer * C * inner;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) nonrounded[i] = P[i] / scale_of(s, i, C, inner); /* NQ:56-58 RealDiv */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) q[i] = floorf(nonrounded[i]);                    /* NQ:59 Floor */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = q[i] * scale_of(s, i, C, inner);          /* NQ:60 Mul */
}

/* NQ:62-118.  scratch: 4*N floats + 2*N bytes.  Returns ds[C]; dP is dy itself (NQ:118). */
void lqo_nested_backward(const float* dy, const float* q, const float* y, float lam,
                         int64_t outer, int64_t C, int64_t inner, float* ds, float* scratch) {
  const int64_t n = outer * C * inner;
  float* t0 = scratch;            /* non_zero_param, later ratio */
  float* t1 = scratch + n;        /* abs(dy), later inner-where result */
  float* t2 = scratch + 2 * n;    /* abs(q), later tanh chain */
  float* t3 = scratch + 3 * n;    /* scale_grads */
  uint8_t* ge = (uint8_t*)(scratch + 4 * n);
  uint8_t* allb = ge + n;
  float* maxvalue = (float*)calloc((size_t)C, sizeof(float));
  uint8_t* all_above = (uint8_t*)malloc((size_t)C);
  double* mean = (double*)calloc((size_t)C, sizeof(double));
  const float c0 = -1.0f * fabsf(tanhf(lam));                                     /* NQ:105 */

#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t0[i] = (y[i] == 0.0f) ? EPS_F32 : y[i];        /* NQ:63 Equal+Select */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t1[i] = fabsf(dy[i]);                           /* NQ:64 Abs */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t0[i] = fabsf(t0[i]);                           /* NQ:64 Abs */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t0[i] = t1[i] / t0[i];                          /* NQ:64 RealDiv -> ratio */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t2[i] = fabsf(q[i]);                            /* NQ:94 Abs */
  /* NQ:94 reduce_max over the axes where scale.shape == 1 */
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < C; ++g) {
    float m = -INFINITY;
    for (int64_t o = 0; o < outer; ++o) {
      const float* row = t2 + (o * C + g) * inner;
      for (int64_t i = 0; i < inner; ++i) m = row[i] > m ? row[i] : m;
    }
    maxvalue[g] = m;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) ge[i] = (t0[i] >= lam);                         /* NQ:97 GreaterEqual */
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < C; ++g) {                                               /* NQ:97 reduce_all */
    uint8_t a = 1;
    for (int64_t o = 0; o < outer; ++o) {
      const uint8_t* row = ge + (o * C + g) * inner;
      for (int64_t i = 0; i < inner; ++i) a &= row[i];
    }
    all_above[g] = a;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) allb[i] = all_above[(i / inner) % C];           /* NQ:98-99 BroadcastTo */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t2[i] = lam - t0[i];                            /* NQ:109 Sub */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t2[i] = tanhf(t2[i]);                           /* NQ:109 Tanh */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t2[i] = fabsf(t2[i]);                           /* NQ:109 Abs */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t2[i] = -1.0f * t2[i];                          /* NQ:109 Mul */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t1[i] = ge[i] ? 0.0f : t2[i];                   /* NQ:106-110 Select */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t3[i] = allb[i] ? c0 : t1[i];                   /* NQ:103-111 Select */
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < C; ++g) {                                               /* NQ:113 reduce_mean */
    double acc = 0.0;
    for (int64_t o = 0; o < outer; ++o) {
      const float* row = t3 + (o * C + g) * inner;
      for (int64_t i = 0; i < inner; ++i) acc += (double)row[i];
    }
    mean[g] = acc / ((double)outer * (double)inner);
  }
  for (int64_t g = 0; g < C; ++g) ds[g] = (float)mean[g] * maxvalue[g];           /* NQ:116 Mul */
  free(maxvalue); free(all_above); free(mean);
}

/* LF:75-116 for one tensor: mean over groups of max_j |P_j|/s, times #P. */
double lqo_maxbin_term(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner) {
  double sum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sum)
  for (int64_t g = 0; g < C; ++g) {
    float m = -INFINITY;
    for (int64_t o = 0; o < outer; ++o) {
      const float* row = P + (o * C + g) * inner;
      for (int64_t i = 0; i < inner; ++i) {
        const float d = fabsf(row[i]) / s[g];                                     /* LF:92 */
        m = d > m ? d : m;
      }
    }
    sum += (double)m;
  }
  return (double)(float)(sum / (double)C) * (double)(outer * C * inner);          /* LF:110 */
}

/* LF:161-195 for one tensor: mean |P - P/s| times #P. */
double lqo_difference_term(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner) {
  const int64_t n = outer * C * inner;
  double sum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sum)
  for (int64_t i = 0; i < n; ++i) {
    const float pq = P[i] / scale_of(s, i, C, inner);                             /* LF:175 */
    sum += (double)fabsf(P[i] - pq);                                              /* LF:178 */
  }
  return (double)(float)(sum / (double)n) * (double)n;                            /* LF:189 */
}

/* LF:240-275 for one tensor: mean 1/s (s==0 -> eps) times #P. */
double lqo_inverse_term(const float* s, int64_t C, int64_t n_elems) {
  double sum = 0.0;
  for (int64_t g = 0; g < C; ++g) {
    const float nz = (s[g] == 0.0f) ? EPS_F32 : s[g];                             /* LF:255 */
    sum += (double)(1.0f / nz);                                                   /* LF:258 */
  }
  return (double)(float)(sum / (double)C) * (double)n_elems;                      /* LF:269 */
}
