/* lq_demo.c -- the C-ABI of liblq.so used from plain C (no Python, no torch): fake-quant forward and
 * threshold backward of a [outer, C, inner] tensor, checked against a straightforward host loop that
 * follows IMAGENETTE/nested_quantization_layer/custom_components/custom_layers.py:55-118.
 *
 *   gcc -O2 -I../include -I/usr/local/cuda/include lq_demo.c -o lq_demo \
 *       -L../learned_quantization_b200/csrc -llq -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,'$ORIGIN/../learned_quantization_b200/csrc'
 *   ./lq_demo
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lq.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define LQ(x) do { int r_ = (x); if (r_ != LQ_OK) { fprintf(stderr, "%s: %d %s\n", #x, r_, lq_last_error()); return 3; } } while (0)

int main(void) {
  const int64_t outer = 9, C = 64, inner = 128, n = outer * C * inner;
  const float lam = 1e-3f, eps = 1.1920928955078125e-07f;
  float *P = malloc(n * 4), *dy = malloc(n * 4), *y = malloc(n * 4), *s = malloc(C * 4), *ds = malloc(C * 4);
  srand(7);
  for (int64_t i = 0; i < n; ++i) { P[i] = ((float)rand() / RAND_MAX - 0.5f) * 0.3f; dy[i] = ((float)rand() / RAND_MAX - 0.5f) * 2e-3f; }
  for (int64_t c = 0; c < C; ++c) s[c] = 5e-3f + 1.5e-2f * (float)rand() / RAND_MAX;

  float *dP, *ddy, *dyq, *dsv, *dds; void* ws; size_t ws_bytes = 0;
  lq_tensor_desc d; memset(&d, 0, sizeof d); d.outer = outer; d.C = C; d.inner = inner;
  LQ(lq_workspace_bytes(&d, 1, LQ_PHASE_BWD, &ws_bytes));
  CK(cudaMalloc((void**)&dP, n * 4)); CK(cudaMalloc((void**)&ddy, n * 4)); CK(cudaMalloc((void**)&dyq, n * 4));
  CK(cudaMalloc((void**)&dsv, C * 4)); CK(cudaMalloc((void**)&dds, C * 4)); CK(cudaMalloc(&ws, ws_bytes ? ws_bytes : 256));
  CK(cudaMemcpy(dP, P, n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ddy, dy, n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dsv, s, C * 4, cudaMemcpyHostToDevice));

  LQ(lq_fq_fwd(dP, dsv, dyq, NULL, LQ_CODES_NONE, outer, C, inner, LQ_ROUND_FLOOR, -INFINITY, INFINITY, NULL));
  LQ(lq_fq_bwd_threshold(ddy, dP, dsv, lam, NULL /* dX is dy itself */, dds, outer, C, inner,
                         LQ_ROUND_FLOOR, -INFINITY, INFINITY, ws, ws_bytes, NULL));
  CK(cudaMemcpy(y, dyq, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ds, dds, C * 4, cudaMemcpyDeviceToHost));

  /* host check: same formulas, element by element */
  int bad_y = 0; double worst = 0.0;
  for (int64_t c = 0; c < C; ++c) {
    double sum = 0.0; long below = 0; float maxq = 0.f;
    for (int64_t o = 0; o < outer; ++o)
      for (int64_t i = 0; i < inner; ++i) {
        const int64_t k = (o * C + c) * inner + i;
        const float q = floorf(P[k] / s[c]), yy = q * s[c];
        if (memcmp(&yy, &y[k], 4) != 0) ++bad_y;
        const float nz = (yy == 0.0f) ? eps : yy, r = fabsf(dy[k]) / fabsf(nz);
        if (fabsf(q) > maxq) maxq = fabsf(q);
        if (!(r >= lam)) { ++below; sum -= fabs(tanh((double)(lam - r))); }
      }
    const double mean = below ? sum / (double)(outer * inner) : -fabs(tanh((double)lam));
    const double ref = mean * maxq, err = fabs(ref - ds[c]) / (fabs(ref) > 1e-30 ? fabs(ref) : 1e-30);
    if (err > worst) worst = err;
  }
  printf("liblq %d: %lld elements, forward mismatches %d, worst scale-gradient error %.2e\n", lq_version(), (long long)n, bad_y, worst);
  return (bad_y == 0 && worst < 1e-5) ? 0 : 1;
}
