// lq_penalty.cu -- the custom loss terms (K3): MaxBin (LF:75-116), Difference (LF:161-195),
// Inverse (LF:240-275) and the gradients TensorFlow's autodiff derives for them
// (math_grad.py rules restated in oracle/lq_oracle.py).  Every pass is one multi-tensor
// launch over all layers' kernels and biases.
#include <math.h>
#include "lq_traverse.cuh"

namespace lq {

// --------------------------------------------------------------- element operators
// All quotients by the group's scale go through the hoisted reciprocal (div_by / div_fast,
// lq_common.cuh): the same IEEE results as __fdiv_rn at a third of the instructions.  Quotients whose
// dividend is a per-group constant (the incoming gradient) are computed once in prep().
struct PenCtx { Recip rc; float4 aux; float k0, k1, k2; };

__device__ __forceinline__ float sgn(float x) { return (float)((x > 0.f) - (x < 0.f)); }

// MaxBin forward: per group max_j(|P_j| / s) and how many elements attain it (the tie
// count _MinOrMaxGrad divides by).  acc.c = max, acc.b = ties.
struct MaxBinStatOp : OpBase<MaxBinStatOp> {
  using Ctx = PenCtx;
  static constexpr bool kIn2 = false, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = -1.f; }
  __device__ static void combine(Acc& x, const Acc& y) {
    if (y.c > x.c) { x.c = y.c; x.b = y.b; }
    else if (y.c == x.c) x.b += y.b;
  }
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job&) const { return PenCtx{make_recip(s), aux, 0.f, 0.f, 0.f}; }
  // branch-free (a thread sees only 16 elements per unit, so "is this a new maximum" is true far too
  // often for a branch: the branchy version ran with 24 of 32 lanes live); a NaN quotient is ignored
  // exactly as by the comparisons of combine()
  __device__ __forceinline__ void take(float d, Acc& acc) const {
    const unsigned eq = (d == acc.c) ? 1u : 0u;
    acc.b = (d > acc.c) ? 1u : acc.b + eq;
    acc.c = fmaxf(acc.c, d);
  }
  __device__ __forceinline__ void elem(float p, float, const PenCtx& c, const Job&, Acc& acc, float&, float&) const {
    take(div_by(fabsf(p), c.rc), acc);                 // LF:92 tf.abs(kernel) / kernel_scale
  }
  __device__ __forceinline__ void quad(const float4& p, const float4&, const PenCtx& c, const Job&, Acc& acc,
                                       float4&, float4&, Defer&) const {
    float4 d;
    if (c.rc.fast && fast_range4(p)) {
      d.x = div_fast(fabsf(p.x), c.rc); d.y = div_fast(fabsf(p.y), c.rc);
      d.z = div_fast(fabsf(p.z), c.rc); d.w = div_fast(fabsf(p.w), c.rc);
    } else {
      d.x = __fdiv_rn(fabsf(p.x), c.rc.s); d.y = __fdiv_rn(fabsf(p.y), c.rc.s);
      d.z = __fdiv_rn(fabsf(p.z), c.rc.s); d.w = __fdiv_rn(fabsf(p.w), c.rc.s);
    }
    // four at once: the new maximum first, then the ties with it (the count of elements equal to the final
    // maximum does not depend on the order they were seen in); NaNs drop out of fmaxf and of == alike
    const float m4 = fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w));
    const float mx = fmaxf(acc.c, m4);
    unsigned cnt = (m4 > acc.c) ? 0u : acc.b;
    cnt += (d.x == mx) ? 1u : 0u; cnt += (d.y == mx) ? 1u : 0u;
    cnt += (d.z == mx) ? 1u : 0u; cnt += (d.w == mx) ? 1u : 0u;
    acc.b = cnt; acc.c = mx;
  }
};

// Difference forward: sum_j |P_j - P_j / s|  (LF:175-178, un-floored quotient). acc.a = sum.
struct DiffStatOp : OpBase<DiffStatOp> {
  using Ctx = PenCtx;
  static constexpr bool kIn2 = false, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = 0.f; }
  __device__ static void combine(Acc& x, const Acc& y) { x.a += y.a; }
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job&) const { return PenCtx{make_recip(s), aux, 0.f, 0.f, 0.f}; }
  __device__ __forceinline__ void elem(float p, float, const PenCtx& c, const Job&, Acc& acc, float&, float&) const {
    acc.a += fabsf(__fsub_rn(p, div_by(p, c.rc)));
  }
  __device__ __forceinline__ void quad(const float4& p, const float4&, const PenCtx& c, const Job&, Acc& acc,
                                       float4&, float4&, Defer&) const {
    if (c.rc.fast && fast_range4(p)) {
      acc.a += fabsf(__fsub_rn(p.x, div_fast(p.x, c.rc))); acc.a += fabsf(__fsub_rn(p.y, div_fast(p.y, c.rc)));
      acc.a += fabsf(__fsub_rn(p.z, div_fast(p.z, c.rc))); acc.a += fabsf(__fsub_rn(p.w, div_fast(p.w, c.rc)));
    } else {
      acc.a += fabsf(__fsub_rn(p.x, __fdiv_rn(p.x, c.rc.s))); acc.a += fabsf(__fsub_rn(p.y, __fdiv_rn(p.y, c.rc.s)));
      acc.a += fabsf(__fsub_rn(p.z, __fdiv_rn(p.z, c.rc.s))); acc.a += fabsf(__fsub_rn(p.w, __fdiv_rn(p.w, c.rc.s)));
    }
  }
};

// MaxBin backward w.r.t. P: aux = {mb_g, ties, coef_g = (1/ties) * g_mb}.
//   g_d = indicator(d == mb) * coef ; g_a = g_d / s ; g_P = g_a * sign(P)
// g_d takes two values per group, so both quotients g_d / s are prepared once: k0 = coef / s, k1 = 0 / s.
template <bool ACCUM>
struct MaxBinBwdOp : OpBase<MaxBinBwdOp<ACCUM>> {
  using Ctx = PenCtx;
  const float* up;   // effective upstream (device scalar); already folded into aux.z
  static constexpr bool kIn2 = ACCUM, kOutIsY = false, kCodes = false, kReduce = false, kAux = true;
  __device__ static void init(Acc&) {}
  __device__ static void combine(Acc&, const Acc&) {}
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job&) const {
    return PenCtx{make_recip(s), aux, __fdiv_rn(aux.z, s), __fdiv_rn(0.0f, s), 0.f};
  }
  __device__ __forceinline__ void elem(float p, float old, const PenCtx& c, const Job&, Acc&, float& dx, float&) const {
    const float d = div_by(fabsf(p), c.rc);
    const float g = __fmul_rn((d == c.aux.x) ? c.k0 : c.k1, sgn(p));
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
};

// Difference backward: d = P - P/s ; g_d = g_abs * sign(d)
//   g_P = g_d + (-g_d)/s ;  ds_g += (-g_d) * ((-P / s) / s)
// With g_abs = t0 * upstream constant per tensor: g_P = sign(d) * (g_abs + (-g_abs)/s) exactly (RN is
// symmetric under negation), prepared once per group as k1; k0 = g_abs.
template <bool ACCUM>
struct DiffBwdOp : OpBase<DiffBwdOp<ACCUM>> {
  using Ctx = PenCtx;
  const float* up;   // effective upstream (device scalar)
  static constexpr bool kIn2 = ACCUM, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = 0.f; }
  __device__ static void combine(Acc& x, const Acc& y) { x.a += y.a; }
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job& j) const {
    const float gabs = __fmul_rn(j.t0, __ldg(up));
    return PenCtx{make_recip(s), aux, gabs, __fadd_rn(gabs, __fdiv_rn(-gabs, s)), 0.f};
  }
  __device__ __forceinline__ void elem(float p, float old, const PenCtx& c, const Job&, Acc& acc, float& dx, float&) const {
    const float pq = div_by(p, c.rc);
    const float sg = sgn(__fsub_rn(p, pq));
    const float g = __fmul_rn(sg, c.k1);
    acc.a += __fmul_rn(__fmul_rn(-sg, c.k0), div_by(-pq, c.rc));
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
  __device__ __forceinline__ void one_fast(float p, float old, const PenCtx& c, Acc& acc, float& dx) const {
    const float pq = div_fast(p, c.rc);
    const float sg = sgn(__fsub_rn(p, pq));
    acc.a += __fmul_rn(__fmul_rn(-sg, c.k0), div_fast(-pq, c.rc));
    const float g = __fmul_rn(sg, c.k1);
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
  __device__ __forceinline__ void quad(const float4& p, const float4& old, const PenCtx& c, const Job& j, Acc& acc,
                                       float4& dx, float4& q, Defer&) const {
    // both dividends inside the window: |p| in [2^-80, 2^80] and, with |s| in [2^-40, 2^40], |p / s| in
    // [2^-120, 2^120] -- the second quotient is checked on its own dividend
    bool fast = c.rc.fast && fast_range4(p);
    if (fast) {
      const float4 pq = make_float4(div_fast(p.x, c.rc), div_fast(p.y, c.rc), div_fast(p.z, c.rc), div_fast(p.w, c.rc));
      fast = fast_range4(pq);
    }
    if (fast) {
      one_fast(p.x, old.x, c, acc, dx.x); one_fast(p.y, old.y, c, acc, dx.y);
      one_fast(p.z, old.z, c, acc, dx.z); one_fast(p.w, old.w, c, acc, dx.w);
    } else {
      elem(p.x, old.x, c, j, acc, dx.x, q.x); elem(p.y, old.y, c, j, acc, dx.y, q.y);
      elem(p.z, old.z, c, j, acc, dx.z, q.z); elem(p.w, old.w, c, j, acc, dx.w, q.w);
    }
  }
};

template <class Op> static Op make_pen_op(const float* up) { Op op; op.up = up; return op; }

// --------------------------------------------------------------- second stages
enum { PFIN_MAXBIN = 1, PFIN_SUM = 2, PFIN_SUM_DS = 3 };

// Intermediate level for groups with very many partials (a per-tensor scale on a large tensor):
// block (g, m) folds partials [m*kFinChunk, (m+1)*kFinChunk) of group g -- sum, and max with its tie
// count -- into lin[g*nlin + m]; the last level then reads REG_LIN partials.  Fixed order.
__global__ void __launch_bounds__(256) pen_fold_kernel(const __grid_constant__ FinJob f) {
  cudaGridDependencySynchronize();
  const long long g = blockIdx.x / f.nlin, m = blockIdx.x - g * f.nlin;
  const long long K = partials_per_group(f);
  const long long k0 = m * kFinChunk;
  double sum = 0.0;
  float mx = -1.f;
  unsigned long long ties = 0ull;
  float4 v[kFinChunk / 256];
#pragma unroll
  for (int u = 0; u < kFinChunk / 256; ++u) {
    const long long k = k0 + u * 256 + threadIdx.x;
    v[u] = (k < K) ? f.part[partial_index(f, g, k)] : make_float4(0.f, 0.f, -2.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < kFinChunk / 256; ++u) {
    sum += (double)v[u].x;
    const unsigned long long b = __float_as_uint(v[u].y);
    if (v[u].z > mx) { mx = v[u].z; ties = b; } else if (v[u].z == mx) ties += b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const unsigned long long ot = __shfl_xor_sync(0xffffffffu, ties, o);
    if (omx > mx) { mx = omx; ties = ot; } else if (omx == mx) ties += ot;
  }
  __shared__ double ssum[8];
  __shared__ unsigned long long sties[8];
  __shared__ float smx[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; sties[w] = ties; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; ties = sties[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      sum += ssum[i];
      if (smx[i] > mx) { mx = smx[i]; ties = sties[i]; } else if (smx[i] == mx) ties += sties[i];
    }
    f.lin[blockIdx.x] = make_float4((float)sum, __uint_as_float((unsigned)(ties > 0xffffffffull ? 0xffffffffull : ties)), mx, 0.f);
  }
}

__device__ __forceinline__ void pen_finalize_group(const FinJob& f, long long g) {
  const long long K = partials_per_group(f);
  double sum = 0.0;
  float mx = -1.f;
  unsigned long long ties = 0ull;
  // four independent 16-byte loads per thread per trip, and no 64-bit division on the layouts whose
  // partials sit at a constant stride (one load at a time behind partial_index() cost 17 us for the
  // 2048 partials per group of an interleaved 64 Mi tensor); thread t still folds t, t+128, ... in order
  const bool simple = (f.regime == REG_ROW && f.nseg == 1) || f.regime == REG_LIN || (f.regime == REG_COL && f.inner == 1);
  const long long stride = (f.regime == REG_ROW) ? f.C : (f.regime == REG_COL ? f.W : 1);
  const float4* base = f.part + ((f.regime == REG_LIN) ? g * f.nlin : g);
  for (long long k0 = threadIdx.x; k0 < K; k0 += 4 * kFinThreads) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = k0 + (long long)u * kFinThreads;
      v[u] = (k < K) ? (simple ? base[k * stride] : f.part[partial_index(f, g, k)]) : make_float4(0.f, 0.f, -2.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sum += (double)v[u].x;
      const unsigned long long b = __float_as_uint(v[u].y);
      if (v[u].z > mx) { mx = v[u].z; ties = b; } else if (v[u].z == mx) ties += b;
    }
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
  cudaGridDependencySynchronize();
  pen_finalize_group(f, blockIdx.x - f.group_begin);
}

// Per-tensor scalar stage: penalty = sum_t mean_t * dim_t / normalizer (LF:110-116 etc.)
struct ScalarJob { const float4* gstat; const float* s; long long C; double n; };
struct ScalarTable { int n, kind, first, last; double normalizer; ScalarJob job[kMaxMT]; };

__device__ __forceinline__ double gstat_double(const float4& v) {
  const unsigned long long bits = (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
  return __longlong_as_double((long long)bits);
}

// Warp w takes tensors w, w+8, ...: the per-tensor term needs one warp reduction, the terms are then
// added in tensor order by one thread (a single block used to walk the tensors one after the other
// with two block barriers each: 30 us for the 40 tensors of the IMAGENETTE model).
__global__ void __launch_bounds__(256) pen_scalar_kernel(const __grid_constant__ ScalarTable tab, double* total, float* out) {
  cudaGridDependencySynchronize();
  __shared__ double term[kMaxMT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < tab.n; t += 8) {
    const ScalarJob& j = tab.job[t];
    double v = 0.0;
    for (long long g = lane; g < j.C; g += 32) {
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
    if (lane == 0) {
      // mean over groups (MaxBin LF:110, Inverse LF:258) or over elements (Difference LF:178), times #P
      const double denom = (tab.kind == LQ_PEN_DIFFERENCE) ? j.n : (double)j.C;
      const float mean = (float)(v / denom);
      term[t] = (double)__fmul_rn(mean, (float)j.n);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = tab.first ? 0.0 : *total;
    for (int t = 0; t < tab.n; ++t) acc += term[t];
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
  cudaGridDependencySynchronize();
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
// Tensors whose groups have more than kFinSplit partials get the intermediate fold; their entry in
// the finalize table then points at the folded partials.
static int pen_fold_large(FinTable& fin, cudaStream_t st) {
  for (int i = 0; i < fin.n; ++i) {
    FinJob& f = fin.job[i];
    if (f.nlin <= 0 || f.regime == REG_LIN) continue;
    if (f.C * f.nlin > 0x7fffffffll) { set_error("too many intermediate partials"); return LQ_E_BADSHAPE; }
    int rc = launch_after(pen_fold_kernel, dim3((unsigned)(f.C * f.nlin)), dim3(256), st, "penalty fold launch", f);
    if (rc) return rc;
    f.part = f.lin; f.regime = REG_LIN;
  }
  return LQ_OK;
}

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
        f.nlin = g.nlin; f.lin = part + g.npart1;
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
      if ((rc = pen_fold_large(fin, st))) break;
      if ((rc = launch_after(pen_finalize_kernel, dim3(groups), dim3(kFinThreads), st, "penalty finalize launch", fin))) break;
    }
    rc = launch_after(pen_scalar_kernel, dim3(1), dim3(256), st, "penalty scalar launch", sc, total, out_penalty);
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
          f.nlin = g.nlin; f.lin = part + g.npart1;
          f.group_begin = groups;
          fin.job[fin.n++] = f;
        }
      }
      groups += (unsigned)d.C;
    }
    if (rc) break;
    if (gt.n) {
      gt.total = groups;
      if ((rc = launch_after(pen_group_kernel, dim3((groups + 127) / 128), dim3(128), st, "penalty group launch", gt, up))) break;
    }
    if (tab.n) {
      if (kind == LQ_PEN_MAXBIN) rc = accum ? launch_mt(tab, make_pen_op<MaxBinBwdOp<true>>(up), st) : launch_mt(tab, make_pen_op<MaxBinBwdOp<false>>(up), st);
      else rc = accum ? launch_mt(tab, make_pen_op<DiffBwdOp<true>>(up), st) : launch_mt(tab, make_pen_op<DiffBwdOp<false>>(up), st);
      if (rc) break;
      if (fin.n) {
        if ((rc = pen_fold_large(fin, st))) break;
        rc = launch_after(pen_finalize_kernel, dim3(groups), dim3(kFinThreads), st, "penalty ds finalize launch", fin);
      }
    }
  }
  delete[] lay;
  return rc;
}
