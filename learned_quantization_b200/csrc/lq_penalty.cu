// lq_penalty.cu -- the custom loss terms (K3): MaxBin (LF:75-116), Difference (LF:161-195),
// Inverse (LF:240-275) and the gradients TensorFlow's autodiff derives for them
// (math_grad.py rules restated in oracle/lq_oracle.py).  Every pass is one multi-tensor
// launch over all layers' kernels and biases.
#include <math.h>
#include "lq_traverse.cuh"

namespace lq {

// --------------------------------------------------------------- element operators
// MaxBin forward: per group max_j(|P_j| / s) and how many elements attain it (the tie
// count _MinOrMaxGrad divides by).  acc.c = max, acc.b = ties.
struct MaxBinStatOp : OpBase<MaxBinStatOp> {
  static constexpr bool kIn2 = false, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = -1.f; }
  __device__ static void combine(Acc& x, const Acc& y) {
    if (y.c > x.c) { x.c = y.c; x.b = y.b; }
    else if (y.c == x.c) x.b += y.b;
  }
  __device__ __forceinline__ void elem(float p, float, const PlainCtx& c, const Job&, Acc& acc, float&, float&) const {
    const float s = c.s;
    const float d = __fdiv_rn(fabsf(p), s);          // LF:92 tf.abs(kernel) / kernel_scale
    if (d > acc.c) { acc.c = d; acc.b = 1u; }
    else if (d == acc.c) acc.b += 1u;
  }
};

// Difference forward: sum_j |P_j - P_j / s|  (LF:175-178, un-floored quotient). acc.a = sum.
struct DiffStatOp : OpBase<DiffStatOp> {
  static constexpr bool kIn2 = false, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = 0.f; }
  __device__ static void combine(Acc& x, const Acc& y) { x.a += y.a; }
  __device__ __forceinline__ void elem(float p, float, const PlainCtx& c, const Job&, Acc& acc, float&, float&) const {
    const float s = c.s;
    acc.a += fabsf(__fsub_rn(p, __fdiv_rn(p, s)));
  }
};

__device__ __forceinline__ float sgn(float x) { return (float)((x > 0.f) - (x < 0.f)); }

// MaxBin backward w.r.t. P: aux = {mb_g, ties, coef_g = (1/ties) * g_mb}.
//   g_d = indicator(d == mb) * coef ; g_a = g_d / s ; g_P = g_a * sign(P)
template <bool ACCUM>
struct MaxBinBwdOp : OpBase<MaxBinBwdOp<ACCUM>> {
  const float* up;   // effective upstream (device scalar); already folded into aux.z
  static constexpr bool kIn2 = ACCUM, kOutIsY = false, kCodes = false, kReduce = false, kAux = true;
  __device__ static void init(Acc&) {}
  __device__ static void combine(Acc&, const Acc&) {}
  __device__ __forceinline__ void elem(float p, float old, const PlainCtx& c, const Job&, Acc&, float& dx, float&) const {
    const float s = c.s;
    const float4& aux = c.aux;
    const float d = __fdiv_rn(fabsf(p), s);
    const float g_d = (d == aux.x) ? aux.z : 0.f;
    const float g = __fmul_rn(__fdiv_rn(g_d, s), sgn(p));
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
};

// Difference backward: d = P - P/s ; g_d = g_abs * sign(d)
//   g_P = g_d + (-g_d)/s ;  ds_g += (-g_d) * ((-P / s) / s)      (job.t0 = g_abs)
template <bool ACCUM>
struct DiffBwdOp : OpBase<DiffBwdOp<ACCUM>> {
  const float* up;   // effective upstream (device scalar)
  static constexpr bool kIn2 = ACCUM, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = 0.f; }
  __device__ static void combine(Acc& x, const Acc& y) { x.a += y.a; }
  __device__ __forceinline__ void elem(float p, float old, const PlainCtx& c, const Job& j, Acc& acc, float& dx, float&) const {
    const float s = c.s;
    const float pq = __fdiv_rn(p, s);
    const float g_d = __fmul_rn(__fmul_rn(j.t0, __ldg(up)), sgn(__fsub_rn(p, pq)));
    const float g = __fadd_rn(g_d, __fdiv_rn(-g_d, s));
    acc.a += __fmul_rn(-g_d, __fdiv_rn(-pq, s));
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
};

template <class Op> static Op make_pen_op(const float* up) { Op op; op.up = up; return op; }

// --------------------------------------------------------------- second stages
enum { PFIN_MAXBIN = 1, PFIN_SUM = 2, PFIN_SUM_DS = 3 };

__device__ __forceinline__ void pen_finalize_group(const FinJob& f, long long g) {
  const long long K = (f.regime == REG_ROW) ? f.outer * f.nseg : f.nchunks * f.inner;
  double sum = 0.0;
  float mx = -1.f;
  unsigned long long ties = 0ull;
  for (long long k = threadIdx.x; k < K; k += kFinThreads) {
    const float4 v = f.part[partial_index(f, g, k)];
    sum += (double)v.x;
    const unsigned long long b = __float_as_uint(v.y);
    if (v.z > mx) { mx = v.z; ties = b; } else if (v.z == mx) ties += b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const unsigned long long ot = __shfl_xor_sync(0xffffffffu, ties, o);
    if (omx > mx) { mx = omx; ties = ot; } else if (omx == mx) ties += ot;
  }
  __shared__ double ssum[kFinThreads / 32];
  __shared__ unsigned long long sties[kFinThreads / 32];
  __shared__ float smx[kFinThreads / 32];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; sties[w] = ties; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; ties = sties[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < kFinThreads / 32; ++i) {
      sum += ssum[i];
      if (smx[i] > mx) { mx = smx[i]; ties = sties[i]; } else if (smx[i] == mx) ties += sties[i];
    }
    if (f.mode == PFIN_MAXBIN) {
      f.gstat[g] = make_float4(mx, __uint_as_float((unsigned)(ties > 0xffffffffull ? 0xffffffffull : ties)), 0.f, 0.f);
    } else if (f.mode == PFIN_SUM) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(sum);
      f.gstat[g] = make_float4(__uint_as_float((unsigned)(bits & 0xffffffffull)), __uint_as_float((unsigned)(bits >> 32)), 0.f, 0.f);
    } else {   // PFIN_SUM_DS
      const float v = (float)sum;
      f.ds[g] = (f.flags & LQ_F_ACCUM_DS) ? __fadd_rn(f.ds[g], v) : v;
    }
  }
}

__global__ void __launch_bounds__(kFinThreads) pen_finalize_kernel(const __grid_constant__ FinTable tab) {
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].group_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const FinJob& f = tab.job[lo];
  pen_finalize_group(f, blockIdx.x - f.group_begin);
}

// Per-tensor scalar stage: penalty = sum_t mean_t * dim_t / normalizer (LF:110-116 etc.)
struct ScalarJob { const float4* gstat; const float* s; long long C; double n; };
struct ScalarTable { int n, kind, first, last; double normalizer; ScalarJob job[kMaxMT]; };

__device__ __forceinline__ double gstat_double(const float4& v) {
  const unsigned long long bits = (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
  return __longlong_as_double((long long)bits);
}

__global__ void __launch_bounds__(256) pen_scalar_kernel(const __grid_constant__ ScalarTable tab, double* total, float* out) {
  __shared__ double red[256 / 32];
  __shared__ double acc;
  if (threadIdx.x == 0) acc = tab.first ? 0.0 : *total;
  __syncthreads();
  for (int t = 0; t < tab.n; ++t) {
    const ScalarJob& j = tab.job[t];
    double v = 0.0;
    for (long long g = threadIdx.x; g < j.C; g += 256) {
      if (tab.kind == LQ_PEN_MAXBIN) v += (double)j.gstat[g].x;
      else if (tab.kind == LQ_PEN_DIFFERENCE) v += gstat_double(j.gstat[g]);
      else {
        const float sv = j.s[g];
        const float nz = (sv == 0.0f) ? kEps : sv;     // LF:255-256
        v += (double)__fdiv_rn(1.0f, nz);              // LF:258
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < 256 / 32; ++i) s += red[i];
      // mean over groups (MaxBin LF:110, Inverse LF:258) or over elements (Difference LF:178), times #P
      const double denom = (tab.kind == LQ_PEN_DIFFERENCE) ? j.n : (double)j.C;
      const float mean = (float)(s / denom);
      acc += (double)__fmul_rn(mean, (float)j.n);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *total = acc;
    if (tab.last) *out = (float)(acc / tab.normalizer);
  }
}

// Per-group stage of the backward: scale gradients that need no element pass, and the
// per-group operands of the MaxBin element pass.
struct GroupJob { float4* gstat; const float* s; float* ds; long long C; float g; unsigned flags, group_begin; };
struct GroupTable { int n, kind; unsigned total; GroupJob job[kMaxMT]; };

__global__ void upstream_kernel(float* dst, float host_val, const float* dev_val) {
  *dst = dev_val ? host_val * (*dev_val) : host_val;
}

__global__ void __launch_bounds__(128) pen_group_kernel(const __grid_constant__ GroupTable tab, const float* up) {
  const unsigned gid = blockIdx.x * 128u + threadIdx.x;
  if (gid >= tab.total) return;
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].group_begin <= gid) lo = mid; else hi = mid - 1;
  }
  const GroupJob& j = tab.job[lo];
  const long long g = gid - j.group_begin;
  const float s = j.s[g];
  const float jg = __fmul_rn(j.g, *up);
  float grad;
  if (tab.kind == LQ_PEN_MAXBIN) {
    float4 st = j.gstat[g];
    const unsigned ties = __float_as_uint(st.y);
    const float coef = __fmul_rn(__fdiv_rn(1.0f, (float)ties), jg);    // indicators / num_selected * g_mb
    st.z = coef;
    j.gstat[g] = st;
    // sum over the tied elements of g_d * ((-a / s) / s), with a / s == mb for each of them
    grad = (float)((double)ties * (double)__fmul_rn(coef, __fdiv_rn(-st.x, s)));
  } else {   // LQ_PEN_INVERSE
    const float nz = (s == 0.0f) ? kEps : s;
    grad = (s == 0.0f) ? 0.0f : __fmul_rn(jg, __fdiv_rn(__fdiv_rn(-1.0f, nz), nz));
  }
  j.ds[g] = (j.flags & LQ_F_ACCUM_DS) ? __fadd_rn(j.ds[g], grad) : grad;
}

// --------------------------------------------------------------- host side
struct PenLayout { size_t part_off, gstat_off; };

static size_t pen_slots(const lq_tensor_desc& d) {
  const Geom a = make_geom(d.outer, d.C, d.inner, true);
  const Geom b = make_geom(d.outer, d.C, d.inner, false);
  return (size_t)(a.npartials > b.npartials ? a.npartials : b.npartials);
}

// Must mirror lq_workspace_bytes(LQ_PHASE_PENALTY_*).
static size_t pen_layout(const lq_tensor_desc* descs, int n, PenLayout* lay) {
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    if (descs[i].outer * descs[i].C * descs[i].inner == 0) { lay[i] = {off, off}; continue; }
    lay[i].part_off = off;
    off += align256(pen_slots(descs[i]) * sizeof(float4));
    lay[i].gstat_off = off;
    off += align256((size_t)descs[i].C * sizeof(float4));
  }
  return off;   // scalar slot (256 B) follows
}

static int pen_check(const char* who, int kind, const lq_tensor_desc* descs, int n, void* ws, size_t ws_bytes,
                     PenLayout** lay_out, size_t* scalar_off, double* normalizer) {
  if (kind < LQ_PEN_MAXBIN || kind > LQ_PEN_INVERSE) { set_error("%s: bad kind %d", who, kind); return LQ_E_BADMODE; }
  if (n <= 0) { set_error("%s: need at least one tensor", who); return LQ_E_BADSHAPE; }
  if (!descs) { set_error("%s: null descs", who); return LQ_E_NULLPTR; }
  double nrm = 0.0;
  for (int i = 0; i < n; ++i) {
    int rc = check_dims(descs[i].outer, descs[i].C, descs[i].inner);
    if (rc) return rc;
    const long long cnt = descs[i].outer * descs[i].C * descs[i].inner;
    if (cnt && (!descs[i].P || !descs[i].s)) { set_error("%s: desc %d has null P/s", who, i); return LQ_E_NULLPTR; }
    nrm += (double)cnt;
  }
  if (nrm == 0.0) { set_error("%s: all tensors are empty", who); return LQ_E_BADSHAPE; }
  PenLayout* lay = new PenLayout[n];
  const size_t need = pen_layout(descs, n, lay) + 256;
  if (!ws || ws_bytes < need || !aligned16(ws)) {
    delete[] lay;
    set_error("%s: workspace %zu B (16-B aligned) needed, got %zu", who, need, ws_bytes);
    return LQ_E_WORKSPACE;
  }
  *lay_out = lay; *scalar_off = need - 256; *normalizer = nrm;
  return LQ_OK;
}

}  // namespace lq

using namespace lq;

extern "C" int lq_penalty_fwd(int kind, const lq_tensor_desc* descs, int n, float* out_penalty,
                              void* ws, size_t ws_bytes, lq_stream_t stream) {
  PenLayout* lay = nullptr; size_t soff = 0; double nrm = 0.0;
  int rc = pen_check("lq_penalty_fwd", kind, descs, n, ws, ws_bytes, &lay, &soff, &nrm);
  if (rc) return rc;
  if (!out_penalty) { delete[] lay; set_error("lq_penalty_fwd: null out_penalty"); return LQ_E_NULLPTR; }
  cudaStream_t st = (cudaStream_t)stream;
  char* base = static_cast<char*>(ws);
  double* total = reinterpret_cast<double*>(base + soff);
  for (int i0 = 0; i0 < n && !rc; i0 += kMaxMT) {
    const int i1 = (i0 + kMaxMT < n) ? i0 + kMaxMT : n;
    JobTable tab{}; FinTable fin{}; ScalarTable sc{};
    unsigned groups = 0;
    sc.kind = kind; sc.first = (i0 == 0); sc.last = (i1 == n); sc.normalizer = nrm;
    for (int i = i0; i < i1; ++i) {
      const lq_tensor_desc& d = descs[i];
      const long long cnt = d.outer * d.C * d.inner;
      if (cnt == 0) continue;
      float4* part = reinterpret_cast<float4*>(base + lay[i].part_off);
      float4* gstat = reinterpret_cast<float4*>(base + lay[i].gstat_off);
      if (kind != LQ_PEN_INVERSE) {
        Geom g = make_geom(d.outer, d.C, d.inner, aligned16(d.P));
        if (g.nblocks > 0x7fffffffll) { rc = LQ_E_BADSHAPE; set_error("tensor too large"); break; }
        Job j = make_job(g, d.P, d.s, nullptr, nullptr, nullptr, nullptr, 0, part);
        tab.job[tab.n++] = j;
        FinJob f{};
        f.part = part; f.gstat = gstat; f.s = d.s; f.ds = nullptr;
        f.C = g.C; f.inner = g.inner; f.W = g.W; f.outer = g.outer; f.nseg = g.nseg; f.nchunks = g.nchunks;
        f.regime = g.regime; f.mode = (kind == LQ_PEN_MAXBIN) ? PFIN_MAXBIN : PFIN_SUM;
        f.group_begin = groups; groups += (unsigned)d.C;
        fin.job[fin.n++] = f;
      }
      sc.job[sc.n++] = ScalarJob{gstat, d.s, d.C, (double)cnt};
    }
    if (rc) break;
    if (tab.n) {
      if (kind == LQ_PEN_MAXBIN) rc = launch_mt(tab, MaxBinStatOp{}, st);
      else rc = launch_mt(tab, DiffStatOp{}, st);
      if (rc) break;
      pen_finalize_kernel<<<dim3(groups), dim3(kFinThreads), 0, st>>>(fin);
      if ((rc = cuda_fail(cudaGetLastError(), "penalty finalize launch"))) break;
    }
    pen_scalar_kernel<<<1, 256, 0, st>>>(sc, total, out_penalty);
    rc = cuda_fail(cudaGetLastError(), "penalty scalar launch");
  }
  delete[] lay;
  return rc;
}

extern "C" int lq_penalty_bwd(int kind, const lq_tensor_desc* descs, int n, float upstream,
                              const float* upstream_dev, void* ws, size_t ws_bytes, lq_stream_t stream) {
  PenLayout* lay = nullptr; size_t soff = 0; double nrm = 0.0;
  int rc = pen_check("lq_penalty_bwd", kind, descs, n, ws, ws_bytes, &lay, &soff, &nrm);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = static_cast<char*>(ws);
  const float g_total = 1.0f / (float)nrm;                     // LF:116 / LF:195 / LF:275 (times *up on device)
  float* up = reinterpret_cast<float*>(base + soff + 128);
  upstream_kernel<<<1, 1, 0, st>>>(up, upstream, upstream_dev);
  rc = cuda_fail(cudaGetLastError(), "upstream launch");
  int accum = -1;
  for (int i = 0; i < n; ++i) {
    const lq_tensor_desc& d = descs[i];
    if (d.outer * d.C * d.inner == 0) continue;
    if (!d.ds) { set_error("lq_penalty_bwd: desc %d has null ds", i); rc = LQ_E_NULLPTR; break; }
    if (kind != LQ_PEN_INVERSE) {
      if (!d.dX) { set_error("lq_penalty_bwd: desc %d has null dX", i); rc = LQ_E_NULLPTR; break; }
      const int a = (d.flags & LQ_F_ACCUM_DX) ? 1 : 0;
      if (accum < 0) accum = a;
      else if (accum != a) { set_error("lq_penalty_bwd: LQ_F_ACCUM_DX must agree across descs"); rc = LQ_E_BADMODE; break; }
    }
  }
  for (int i0 = 0; i0 < n && !rc; i0 += kMaxMT) {
    const int i1 = (i0 + kMaxMT < n) ? i0 + kMaxMT : n;
    JobTable tab{}; FinTable fin{}; GroupTable gt{};
    unsigned groups = 0;
    gt.kind = kind;
    for (int i = i0; i < i1; ++i) {
      const lq_tensor_desc& d = descs[i];
      const long long cnt = d.outer * d.C * d.inner;
      if (cnt == 0) continue;
      float4* part = reinterpret_cast<float4*>(base + lay[i].part_off);
      float4* gstat = reinterpret_cast<float4*>(base + lay[i].gstat_off);
      const float g_mean = g_total * (float)cnt;                // d(mean * dim)/d mean
      if (kind == LQ_PEN_MAXBIN || kind == LQ_PEN_INVERSE) {
        GroupJob gj{gstat, d.s, d.ds, d.C, g_mean / (float)d.C, d.flags, groups};
        gt.job[gt.n++] = gj;
      }
      if (kind != LQ_PEN_INVERSE) {
        Geom g = make_geom(d.outer, d.C, d.inner, aligned16(d.P) && aligned16(d.dX));
        if (g.nblocks > 0x7fffffffll) { rc = LQ_E_BADSHAPE; set_error("tensor too large"); break; }
        Job j = make_job(g, d.P, d.s, d.dX, nullptr, d.dX, nullptr, 0, part);
        j.write_dx = 1;
        j.gaux = gstat;
        j.t0 = g_mean / (float)cnt;                             // _MeanGrad over elements (Difference)
        tab.job[tab.n++] = j;
        if (kind == LQ_PEN_DIFFERENCE) {
          FinJob f{};
          f.part = part; f.gstat = gstat; f.s = d.s; f.ds = d.ds;
          f.C = g.C; f.inner = g.inner; f.W = g.W; f.outer = g.outer; f.nseg = g.nseg; f.nchunks = g.nchunks;
          f.regime = g.regime; f.mode = PFIN_SUM_DS; f.flags = d.flags;
          f.group_begin = groups;
          fin.job[fin.n++] = f;
        }
      }
      groups += (unsigned)d.C;
    }
    if (rc) break;
    if (gt.n) {
      gt.total = groups;
      pen_group_kernel<<<dim3((groups + 127) / 128), dim3(128), 0, st>>>(gt, up);
      if ((rc = cuda_fail(cudaGetLastError(), "penalty group launch"))) break;
    }
    if (tab.n) {
      if (kind == LQ_PEN_MAXBIN) rc = accum ? launch_mt(tab, make_pen_op<MaxBinBwdOp<true>>(up), st) : launch_mt(tab, make_pen_op<MaxBinBwdOp<false>>(up), st);
      else rc = accum ? launch_mt(tab, make_pen_op<DiffBwdOp<true>>(up), st) : launch_mt(tab, make_pen_op<DiffBwdOp<false>>(up), st);
      if (rc) break;
      if (fin.n) {
        pen_finalize_kernel<<<dim3(groups), dim3(kFinThreads), 0, st>>>(fin);
        rc = cuda_fail(cudaGetLastError(), "penalty ds finalize launch");
      }
    }
  }
  delete[] lay;
  return rc;
}
