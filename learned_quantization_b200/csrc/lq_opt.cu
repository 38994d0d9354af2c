// lq_opt.cu -- SURVEY 8(f1): the step right after the hot path.  One multi-tensor launch applies
// the optimizer update to every parameter and learned scale of the model and, in the same pass,
// the L2-regulariser gradient (keras.regularizers.l2 on the residual convs,
// IMAGENETTE/nested_quantization_layer/experiment.py:638,645,661), the DP gradient scaling and
// MinValueConstraint (NQ:35-46: s <- max(s, 100*eps), applied by Keras after every update).
//
// Update rules restate Keras 2.11's default ("experimental") optimizers, which is what
// `from tensorflow.keras.optimizers import Adam, SGD` resolves to under requirements.txt:1
// (tensorflow==2.11.0; not vendored in the reference):
//   Adam.update_step:  alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
//                      m += (g - m) * (1 - b1);  v += (g*g - v) * (1 - b2)
//                      p -= (m * alpha) / (sqrt(v) + eps)
//   SGD.update_step:   m = -g * lr + m * momentum
//                      p += nesterov ? (-g * lr + m * momentum) : m        (momentum > 0)
//                      p += -g * lr                                        (momentum == 0)
//   weight_decay:      p -= p * wd * lr   before the update (Optimizer._apply_weight_decay)
// Call sites: IMAGENETTE/nested_quantization_layer/experiment.py:763-771 (Adam 1e-4),
// CIFAR-10/paper_implementation/experiment.py:130-141 (SGD 0.1 / 0.9 / nesterov / wd 5e-4).
#include <math.h>
#include "lq_common.cuh"

namespace lq {

constexpr int kOptMax = 56;          // tensors per launch (kernel-parameter table)
constexpr int kOptChunk = 4096;      // elements per block

struct OptJob {
  float* p; const float* g; float* m; float* v;
  long long n;
  float l2, min_value, wd;
  unsigned block_begin;
};
struct OptTable { int n; OptJob job[kOptMax]; };

struct AdamArgs { float lr, b1, b2, eps, alpha; };
struct SgdArgs { float lr, momentum; int nesterov; };

template <bool ADAM>
__device__ __forceinline__ float opt_one(float p, float g, float& m, float& v, const OptJob& j, float gscale,
                                         const AdamArgs& a, const SgdArgs& s) {
  g = __fmul_rn(g, gscale);
  if (j.l2 != 0.0f) g = __fadd_rn(g, __fmul_rn(__fmul_rn(2.0f, j.l2), p));       // d/dp l2 * sum(p^2)
  if (ADAM) {
    if (j.wd != 0.0f) p = __fsub_rn(p, __fmul_rn(__fmul_rn(p, j.wd), a.lr));
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.0f - a.b1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), 1.0f - a.b2));
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, a.alpha), __fadd_rn(sqrtf(v), a.eps)));
  } else {
    if (j.wd != 0.0f) p = __fsub_rn(p, __fmul_rn(__fmul_rn(p, j.wd), s.lr));
    if (s.momentum != 0.0f) {
      m = __fadd_rn(__fmul_rn(-g, s.lr), __fmul_rn(m, s.momentum));
      p = __fadd_rn(p, s.nesterov ? __fadd_rn(__fmul_rn(-g, s.lr), __fmul_rn(m, s.momentum)) : m);
    } else {
      p = __fadd_rn(p, __fmul_rn(-g, s.lr));
    }
  }
  return fmaxf(p, j.min_value);      // -inf: no constraint
}

// Same expression as the host side of lq_opt_adam (fp32 pow / sqrt like tf.pow / tf.sqrt).
__host__ __device__ inline float adam_alpha(float lr, float b1, float b2, float t) {
  return lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
}

template <bool ADAM>
__global__ void __launch_bounds__(256) opt_kernel(const __grid_constant__ OptTable tab, AdamArgs a, const SgdArgs s,
                                                  float gscale, const float* gscale_dev, const long long* step_dev) {
  if (ADAM && step_dev) {
    // step counter in device memory (CUDA-graph replays cannot change a host scalar): one thread
    // derives the bias-corrected step size for the block
    __shared__ float alpha_s;
    if (threadIdx.x == 0) alpha_s = adam_alpha(a.lr, a.b1, a.b2, (float)__ldg(step_dev));
    __syncthreads();
    a.alpha = alpha_s;
  }
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const OptJob& j = tab.job[lo];
  if (gscale_dev) gscale = __fmul_rn(gscale, __ldg(gscale_dev));
  const long long e0 = (long long)(blockIdx.x - j.block_begin) * kOptChunk;
  const long long e1 = (e0 + kOptChunk < j.n) ? e0 + kOptChunk : j.n;
  const bool has_m = j.m != nullptr, has_v = ADAM;
  const bool vec = (((uintptr_t)j.p | (uintptr_t)j.g | (uintptr_t)j.m | (uintptr_t)j.v) & 15u) == 0;
  if (vec) {
    for (long long e = e0 + 4ll * threadIdx.x; e + 3 < e1; e += 4ll * 256) {
      float4 p = *reinterpret_cast<const float4*>(j.p + e);
      const float4 g = *reinterpret_cast<const float4*>(j.g + e);
      float4 m = has_m ? *reinterpret_cast<const float4*>(j.m + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v = has_v ? *reinterpret_cast<const float4*>(j.v + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      p.x = opt_one<ADAM>(p.x, g.x, m.x, v.x, j, gscale, a, s);
      p.y = opt_one<ADAM>(p.y, g.y, m.y, v.y, j, gscale, a, s);
      p.z = opt_one<ADAM>(p.z, g.z, m.z, v.z, j, gscale, a, s);
      p.w = opt_one<ADAM>(p.w, g.w, m.w, v.w, j, gscale, a, s);
      *reinterpret_cast<float4*>(j.p + e) = p;
      if (has_m) *reinterpret_cast<float4*>(j.m + e) = m;
      if (has_v) *reinterpret_cast<float4*>(j.v + e) = v;
    }
    const long long tail = e0 + ((e1 - e0) & ~3ll);
    const long long e = tail + threadIdx.x;
    if (e < e1) {
      float m = has_m ? j.m[e] : 0.f, v = has_v ? j.v[e] : 0.f;
      j.p[e] = opt_one<ADAM>(j.p[e], j.g[e], m, v, j, gscale, a, s);
      if (has_m) j.m[e] = m;
      if (has_v) j.v[e] = v;
    }
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += 256) {
      float m = has_m ? j.m[e] : 0.f, v = has_v ? j.v[e] : 0.f;
      j.p[e] = opt_one<ADAM>(j.p[e], j.g[e], m, v, j, gscale, a, s);
      if (has_m) j.m[e] = m;
      if (has_v) j.v[e] = v;
    }
  }
}

template <bool ADAM>
static int opt_launch(const char* who, const lq_opt_desc* descs, int n, const AdamArgs& a, const SgdArgs& s,
                      float gscale, const float* gscale_dev, const long long* step_dev, cudaStream_t st) {
  if (n < 0) { set_error("%s: n < 0", who); return LQ_E_BADSHAPE; }
  if (n && !descs) { set_error("%s: null descs", who); return LQ_E_NULLPTR; }
  for (int i0 = 0; i0 < n; i0 += kOptMax) {
    OptTable tab{};
    unsigned total = 0;
    for (int i = i0; i < n && i < i0 + kOptMax; ++i) {
      const lq_opt_desc& d = descs[i];
      if (d.n < 0) { set_error("%s: desc %d has n < 0", who, i); return LQ_E_BADSHAPE; }
      if (d.n == 0) continue;
      if (!d.p || !d.g) { set_error("%s: desc %d has null p/g", who, i); return LQ_E_NULLPTR; }
      if (ADAM && (!d.m || !d.v)) { set_error("%s: desc %d has null m/v", who, i); return LQ_E_NULLPTR; }
      if (!ADAM && s.momentum != 0.0f && !d.m) { set_error("%s: desc %d has null momentum buffer", who, i); return LQ_E_NULLPTR; }
      const long long blocks = (d.n + kOptChunk - 1) / kOptChunk;
      if (blocks + total > 0x7fffffffll) { set_error("%s: too many elements for one launch", who); return LQ_E_BADSHAPE; }
      OptJob j{d.p, d.g, d.m, ADAM ? d.v : nullptr, d.n, d.l2, d.min_value, d.weight_decay, total};
      if (!ADAM && s.momentum == 0.0f) j.m = nullptr;
      tab.job[tab.n++] = j;
      total += (unsigned)blocks;
    }
    if (!total) continue;
    opt_kernel<ADAM><<<dim3(total), dim3(256), 0, st>>>(tab, a, s, gscale, gscale_dev, step_dev);
    int rc = cuda_fail(cudaGetLastError(), who);
    if (rc) return rc;
  }
  return LQ_OK;
}

}  // namespace lq

using namespace lq;

extern "C" int lq_opt_adam(const lq_opt_desc* descs, int n, float lr, float beta1, float beta2, float eps,
                           int64_t step, const int64_t* step_dev, float grad_scale, const float* grad_scale_dev,
                           lq_stream_t stream) {
  if (!step_dev && step < 1) { set_error("lq_opt_adam: step is 1-based, got %lld", (long long)step); return LQ_E_BADMODE; }
  AdamArgs a{lr, beta1, beta2, eps, 0.f};
  // alpha = lr * sqrt(1 - b2^t) / (1 - b1^t), in fp32 like tf.pow / tf.sqrt on fp32 variables
  if (!step_dev) a.alpha = adam_alpha(lr, beta1, beta2, (float)step);
  return opt_launch<true>("lq_opt_adam", descs, n, a, SgdArgs{}, grad_scale, grad_scale_dev,
                          reinterpret_cast<const long long*>(step_dev), (cudaStream_t)stream);
}

extern "C" int lq_opt_sgd(const lq_opt_desc* descs, int n, float lr, float momentum, int nesterov,
                          float grad_scale, const float* grad_scale_dev, lq_stream_t stream) {
  SgdArgs s{lr, momentum, nesterov ? 1 : 0};
  return opt_launch<false>("lq_opt_sgd", descs, n, AdamArgs{}, s, grad_scale, grad_scale_dev, nullptr, (cudaStream_t)stream);
}
