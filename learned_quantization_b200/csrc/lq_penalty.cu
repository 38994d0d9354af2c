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

// is this element one of the group's maxima?  lo >= 0: |p| >= lo (fast regime); otherwise compare quotients
__device__ __forceinline__ unsigned tie_hit(float p, float lo, float mb, float sc) {
  const float ap = fabsf(p);
  return (lo >= 0.f) ? (ap >= lo ? 1u : 0u) : (__fdiv_rn(ap, sc) == mb ? 1u : 0u);
}

// ---- the fast regime of the loss-term passes (what makes them HBM-bound instead of issue-bound)
// Both reductions can be carried out on |P| alone when the group's scale is an ordinary positive number:
//  * MaxBin (LF:90-94): division by s > 0 is monotone, so max_j RN(|P_j| / s) = RN(max_j |P_j| / s): the element
//    pass keeps max |P| (ONE instruction per element) and the second stage divides once per group.  The tie
//    count TF's _MinOrMaxGrad needs is the number of elements whose QUOTIENT equals the maximum; those are the
//    elements with |P_j| >= lo, lo = the smallest float whose quotient still rounds to the maximum (at most a
//    few ulps below max |P|, found in the second stage by stepping down).  They are counted by pen_ties_kernel,
//    which re-reads only the units whose own maximum reaches lo (normally one unit per group).
//  * Difference (LF:175-178): |P - RN(P/s)| = |P| |1 - 1/s| (1 + e) with |e| <= 3 * 2^-24 whenever s <= 1/2 or
//    s >= 2 (the two roundings are each 2^-24 relative to |P/s| and |d|, and |P/s| <= 2 |d| there), so the
//    element pass keeps sum |P| and the second stage multiplies by |1 - 1/s| in double: inside the 1e-5 the
//    loss terms are held to, with no division per element.  sign(d_j) = -sign(P_j) (s <= 1/2) or +sign(P_j)
//    (s >= 2) makes the backward division-free too, and its scale gradient closed-form:
//    sum_j g_d ((P_j / s) / s) = g_abs * sigma * sum |P| / s^2.
// Scales outside these sets (negative, zero, tiny, huge, NaN, or within a factor 2 of 1 for Difference) and
// elements that are not ordinary finite numbers take the exact per-element formulation below, as before.
constexpr float kPenLo = 9.094947017729282e-13f /*2^-40*/, kPenHi = 1099511627776.0f /*2^40*/;
__host__ __device__ __forceinline__ bool maxbin_fast_scale(float s) { return s >= kPenLo && s <= kPenHi; }
__host__ __device__ __forceinline__ bool diff_fast_scale(float s) { return (s >= kPenLo && s <= 0.5f) || (s >= 2.0f && s <= kPenHi); }

// MaxBin forward.  fast: acc.c = max |P|.  exact: acc.c = max_j(|P_j| / s), acc.b = how many attain it.
struct MaxBinStatOp : OpBase<MaxBinStatOp> {
  using Ctx = PenCtx;
  static constexpr bool kIn2 = false, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = -1.f; }
  __device__ static void combine(Acc& x, const Acc& y) {
    if (y.c > x.c) { x.c = y.c; x.b = y.b; }
    else if (y.c == x.c) x.b += y.b;
  }
  // k0 = 1: fast regime
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job&) const {
    return PenCtx{make_recip(s), aux, maxbin_fast_scale(s) ? 1.f : 0.f, 0.f, 0.f};
  }
  // branch-free (a thread sees only 16 elements per unit, so "is this a new maximum" is true far too
  // often for a branch: the branchy version ran with 24 of 32 lanes live); a NaN quotient is ignored
  // exactly as by the comparisons of combine()
  __device__ __forceinline__ void take(float d, Acc& acc) const {
    const unsigned eq = (d == acc.c) ? 1u : 0u;
    acc.b = (d > acc.c) ? 1u : acc.b + eq;
    acc.c = fmaxf(acc.c, d);
  }
  __device__ __forceinline__ void elem(float p, float, const PenCtx& c, const Job&, Acc& acc, float&, float&) const {
    if (c.k0 != 0.f) { acc.c = fmaxf(acc.c, fabsf(p)); return; }
    take(div_by(fabsf(p), c.rc), acc);                 // LF:92 tf.abs(kernel) / kernel_scale
  }
  __device__ __forceinline__ void quad(const float4& p, const float4&, const PenCtx& c, const Job&, Acc& acc,
                                       float4&, float4&, Defer&) const {
    if (c.k0 != 0.f) {                                  // NaNs drop out of fmaxf as they do out of the comparisons
      acc.c = fmaxf(fmaxf(acc.c, fabsf(p.x)), fmaxf(fabsf(p.y), fmaxf(fabsf(p.z), fabsf(p.w))));
      return;
    }
    exact4(p, c.rc.s, acc);
  }
  __device__ __forceinline__ static void exact4(const float4& p, float s, Acc& acc) {
    float4 d;
    d.x = __fdiv_rn(fabsf(p.x), s); d.y = __fdiv_rn(fabsf(p.y), s);
    d.z = __fdiv_rn(fabsf(p.z), s); d.w = __fdiv_rn(fabsf(p.w), s);
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

// Difference forward: sum_j |P_j - P_j / s|  (LF:175-178, un-floored quotient).
// fast: acc.a = sum |P| over the ordinary elements; acc.c = exact terms of the others.  exact: acc.c = all terms.
struct DiffStatOp : OpBase<DiffStatOp> {
  using Ctx = PenCtx;
  static constexpr bool kIn2 = false, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = 0.f; }
  __device__ static void combine(Acc& x, const Acc& y) { x.a += y.a; x.c += y.c; }
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job&) const {
    return PenCtx{make_recip(s), aux, diff_fast_scale(s) ? 1.f : 0.f, 0.f, 0.f};
  }
  __device__ __forceinline__ float exact(float p, const PenCtx& c) const { return fabsf(__fsub_rn(p, __fdiv_rn(p, c.rc.s))); }
  __device__ __forceinline__ void elem(float p, float, const PenCtx& c, const Job&, Acc& acc, float&, float&) const {
    const float a = fabsf(p);
    if (c.k0 != 0.f && a <= kFastHi) acc.a += a; else acc.c += exact(p, c);   // inf / NaN: exact (NaN fails the compare)
  }
  __device__ __forceinline__ void quad(const float4& p, const float4&, const PenCtx& c, const Job&, Acc& acc,
                                       float4&, float4&, Defer&) const {
    const float t = (fabsf(p.x) + fabsf(p.y)) + (fabsf(p.z) + fabsf(p.w));
    if (c.k0 != 0.f && t <= kFastHi) { acc.a += t; return; }                 // one compare covers inf and NaN
    acc.c += exact4(p, c.rc.s);
  }
  __device__ __forceinline__ static float exact4(const float4& p, float s) {
    return ((fabsf(__fsub_rn(p.x, __fdiv_rn(p.x, s))) + fabsf(__fsub_rn(p.y, __fdiv_rn(p.y, s)))) +
            fabsf(__fsub_rn(p.z, __fdiv_rn(p.z, s)))) + fabsf(__fsub_rn(p.w, __fdiv_rn(p.w, s)));
  }
  __device__ __forceinline__ Acc warp_reduce(const Acc& x) const {
    Acc r; r.a = warp_sum(x.a); r.b = 0u; r.c = warp_sum(x.c); return r;
  }
};

// MaxBin backward w.r.t. P: aux = {mb_g, ties, coef_g = (1/ties) * g_mb, lo_g}.
//   g_d = indicator(d == mb) * coef ; g_a = g_d / s ; g_P = g_a * sign(P)
// g_d takes two values per group, so both quotients g_d / s are prepared once: k0 = coef / s, k1 = 0 / s.
// The indicator is |P| >= lo in the fast regime (lo >= 0) and the quotient comparison otherwise (lo < 0).
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
    const bool hit = (c.aux.w >= 0.f) ? (fabsf(p) >= c.aux.w) : (div_by(fabsf(p), c.rc) == c.aux.x);
    const float g = __fmul_rn(hit ? c.k0 : c.k1, sgn(p));
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
};

// Difference backward: d = P - P/s ; g_d = g_abs * sign(d)
//   g_P = g_d + (-g_d)/s ;  ds_g += (-g_d) * ((-P / s) / s)
// With g_abs = t0 * upstream constant per tensor: g_P = sign(d) * (g_abs + (-g_abs)/s) exactly (RN is
// symmetric under negation), prepared once per group as k1; k0 = g_abs.
// fast regime (k2 = sigma = -1 / +1): sign(d) = sigma * sign(P), and the scale gradient of the ordinary elements
// is closed-form from the forward's sum |P| (pen_finalize adds it); only the other elements accumulate here.
template <bool ACCUM>
struct DiffBwdOp : OpBase<DiffBwdOp<ACCUM>> {
  using Ctx = PenCtx;
  const float* up;   // effective upstream (device scalar)
  static constexpr bool kIn2 = ACCUM, kOutIsY = false, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = 0u; a.c = 0.f; }
  __device__ static void combine(Acc& x, const Acc& y) { x.a += y.a; }
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job& j) const {
    const float gabs = __fmul_rn(j.t0, __ldg(up));
    const float sigma = diff_fast_scale(s) ? (s <= 0.5f ? -1.f : 1.f) : 0.f;
    return PenCtx{make_recip(s), aux, gabs, __fadd_rn(gabs, __fdiv_rn(-gabs, s)), sigma};
  }
  __device__ __forceinline__ void exact(float p, float old, const PenCtx& c, Acc& acc, float& dx) const {
    const float pq = __fdiv_rn(p, c.rc.s);
    const float sg = sgn(__fsub_rn(p, pq));
    const float g = __fmul_rn(sg, c.k1);
    acc.a += __fmul_rn(__fmul_rn(-sg, c.k0), __fdiv_rn(-pq, c.rc.s));
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
  __device__ __forceinline__ void fast1(float p, float old, const PenCtx& c, float& dx) const {
    const float g = __fmul_rn(__fmul_rn(c.k2, sgn(p)), c.k1);
    dx = ACCUM ? __fadd_rn(old, g) : g;
  }
  __device__ __forceinline__ void elem(float p, float old, const PenCtx& c, const Job&, Acc& acc, float& dx, float&) const {
    if (c.k2 != 0.f && fabsf(p) <= kFastHi) fast1(p, old, c, dx); else exact(p, old, c, acc, dx);
  }
  __device__ __forceinline__ void quad(const float4& p, const float4& old, const PenCtx& c, const Job&, Acc& acc,
                                       float4& dx, float4&, Defer&) const {
    const float t = (fabsf(p.x) + fabsf(p.y)) + (fabsf(p.z) + fabsf(p.w));     // the forward's split, bit for bit
    if (c.k2 != 0.f && t <= kFastHi) {
      fast1(p.x, old.x, c, dx.x); fast1(p.y, old.y, c, dx.y); fast1(p.z, old.z, c, dx.z); fast1(p.w, old.w, c, dx.w);
    } else {
      exact(p.x, old.x, c, acc, dx.x); exact(p.y, old.y, c, acc, dx.y);
      exact(p.z, old.z, c, acc, dx.z); exact(p.w, old.w, c, acc, dx.w);
    }
  }
};

template <class Op> static Op make_pen_op(const float* up) { Op op; op.up = up; return op; }

// --------------------------------------------------------------- lean forward kernels (fast regime)
// The forward statistics only READ: to run at HBM speed a read-only pass needs about twice the bytes in flight of
// the read+write kernels, and the generic traversal + operator (four 16-byte loads per lane, 44-64 registers with
// the exact formulation inlined) sat at 56-71 % DRAM throughput with 42-50 % of the warps resident.  These kernels
// keep EIGHT 16-byte loads per lane in flight, do one FMNMX (MaxBin) or one FADD (Difference) per element, and hand
// anything that is not the fast regime -- a scale outside the fast set, unaligned or odd-sized rows, a unit that
// holds an inf / NaN (Difference) -- to the generic traversal, out of line.  Same geometry, same partials.
template <int KIND> struct StatKind;
template <> struct StatKind<LQ_PEN_MAXBIN> { using Op = MaxBinStatOp; };
template <> struct StatKind<LQ_PEN_DIFFERENCE> { using Op = DiffStatOp; };

// ---- the fused epilogue's control words (see "fused epilogue" below).  They live in the workspace's scalar slot
// (kPenSlotBytes at the end of the workspace: the running total, these counters, the backward's upstream, one term per
// tensor) and are zeroed by block 0 of every statistics launch of the same call, i.e. before the second stage runs.
constexpr size_t kPenCtlOff = 16, kPenUpOff = 240, kPenTermOff = 256;     // total: [0, 8); terms: kMaxMT doubles
enum { kCtlTensors = 0, kCtlPending = 1, kCtlTensor0 = 2, kCtlWords = 2 + kMaxMT };
static_assert(kPenCtlOff + 4 * kCtlWords <= kPenUpOff && kPenTermOff + 8 * kMaxMT <= kPenSlotBytes, "scalar slot layout");

__device__ __forceinline__ void zero_ctl(unsigned* ctl) {
  if (ctl && blockIdx.x == 0 && threadIdx.x < kCtlWords) ctl[threadIdx.x] = 0u;
}

// K1 with the loss term's statistics as by-products of the same pass ("penalty forward fused into K1:
// 0 extra bytes"): y = floor(P / s) * s is written exactly as FwdOp (lq_fq.cu) writes it, and the unit's partial is the
// statistics operator's, so the second stage, the tie counting and the backward are the stand-alone pass's.
template <int KIND, bool EXT>
struct FwdPenOp : OpBase<FwdPenOp<KIND, EXT>> {
  using Stat = typename StatKind<KIND>::Op;
  using Ctx = PenCtx;
  QMode qm;
  static constexpr bool kIn2 = false, kOutIsY = true, kCodes = false, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { Stat::init(a); }
  __device__ static void combine(Acc& x, const Acc& y) { Stat::combine(x, y); }
  __device__ __forceinline__ PenCtx prep(float s, const float4& aux, const Job& j) const { return Stat().prep(s, aux, j); }
  __device__ __forceinline__ Acc warp_reduce(const Acc& x) const { return Stat().warp_reduce(x); }
  __device__ __forceinline__ void elem(float p, float, const PenCtx& c, const Job& j, Acc& acc, float& y, float& q) const {
    q = round_code<EXT>(div_by(p, c.rc), qm);                    // NQ:56-57
    y = __fmul_rn(q, c.rc.s);                                    // NQ:58
    float o0, o1;
    Stat().elem(p, 0.f, c, j, acc, o0, o1);
  }
  __device__ __forceinline__ void quad(const float4& p, const float4& d, const PenCtx& c, const Job& j, Acc& acc,
                                       float4& y, float4& q, Defer& df) const {
    if (c.rc.fast && fast_range4(p)) {        // one guard for the four elements, as in FwdOp::quad
      q.x = round_code<EXT>(div_fast(p.x, c.rc), qm); q.y = round_code<EXT>(div_fast(p.y, c.rc), qm);
      q.z = round_code<EXT>(div_fast(p.z, c.rc), qm); q.w = round_code<EXT>(div_fast(p.w, c.rc), qm);
    } else {
      q.x = round_code<EXT>(__fdiv_rn(p.x, c.rc.s), qm); q.y = round_code<EXT>(__fdiv_rn(p.y, c.rc.s), qm);
      q.z = round_code<EXT>(__fdiv_rn(p.z, c.rc.s), qm); q.w = round_code<EXT>(__fdiv_rn(p.w, c.rc.s), qm);
    }
    y.x = __fmul_rn(q.x, c.rc.s); y.y = __fmul_rn(q.y, c.rc.s); y.z = __fmul_rn(q.z, c.rc.s); y.w = __fmul_rn(q.w, c.rc.s);
    float4 o0, o1;
    Stat().quad(p, d, c, j, acc, o0, o1, df);
  }
};

template <int KIND>
__device__ __noinline__ void stat_row_generic(const Job& j, long long unit) {
  typename StatKind<KIND>::Op op;
  using IO = LocalIO<false>;
  if (j.vec) run_row<typename StatKind<KIND>::Op, false, true>(j, 0u, op, nullptr, IO(), unit);
  else run_row<typename StatKind<KIND>::Op, false, false>(j, 0u, op, nullptr, IO(), unit);
}

template <int KIND>
__global__ void __launch_bounds__(kThreads, 4) stat_row_kernel(const __grid_constant__ JobTable tab, unsigned* ctl) {
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const Job& j = tab.job[lo];
  const unsigned blk = blockIdx.x - j.block_begin;
  cudaGridDependencySynchronize();
  zero_ctl(ctl);
  const int lane = threadIdx.x & 31;
  const long long unit = (long long)blk * kWarps + (threadIdx.x >> 5);
  if (unit >= j.units) return;
  long long row, grp;
  int sg;
  if (j.fast32) {
    const unsigned r = fd_div((unsigned)unit, j.fd_nseg);
    sg = (int)((unsigned)unit - r * (unsigned)j.nseg);
    grp = r - fd_div(r, j.fd_C) * (unsigned)j.C;
    row = r;
  } else {
    row = unit / j.nseg;
    sg = (int)(unit - row * j.nseg);
    grp = row % j.C;
  }
  if (!j.vec || j.seg > 1024) { stat_row_generic<KIND>(j, unit); return; }                // per tensor
  const long long base = row * j.inner + (long long)sg * j.seg;
  const long long rem = j.inner - (long long)sg * j.seg;
  const int nvec = (rem < j.seg ? (int)rem : j.seg) >> 2;                                 // <= 256: one trip
  const float4* P4 = reinterpret_cast<const float4*>(j.P + base);
  float4 p[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int v = u * 32 + lane;
    if (v < nvec) p[u] = j.stream_hint ? __ldcs(P4 + v) : P4[v];
  }
  // the scale is looked at AFTER the unit's loads have been issued: with the regime test in front of them every warp
  // spent one round trip to memory on four bytes before it asked for its 4 KB
  const float s = __ldg(j.s + grp);
  const bool fast = (KIND == LQ_PEN_MAXBIN) ? maxbin_fast_scale(s) : diff_fast_scale(s);
  float m = -1.f, a = 0.f;
  bool bad = false;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    if (u * 32 + lane < nvec) {
      if (KIND == LQ_PEN_MAXBIN) {
        m = fmaxf(fmaxf(m, fabsf(p[u].x)), fmaxf(fabsf(p[u].y), fmaxf(fabsf(p[u].z), fabsf(p[u].w))));
      } else {
        const float t = (fabsf(p[u].x) + fabsf(p[u].y)) + (fabsf(p[u].z) + fabsf(p[u].w));
        bad |= !(t <= kFastHi);                     // an inf or a NaN among the four
        a += t;
      }
    }
  }
  if (!fast || (KIND == LQ_PEN_DIFFERENCE && __any_sync(0xffffffffu, bad))) { stat_row_generic<KIND>(j, unit); return; }   // warp-uniform
  if (KIND == LQ_PEN_MAXBIN) {
    // m is -1 or >= 0 and never NaN: as signed integers the bit patterns order like the values
    m = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(m)));
    if (lane == 0) j.part[unit] = make_float4(0.f, __uint_as_float(0u), m, 0.f);
  } else {
    a = warp_sum(a);
    if (lane == 0) j.part[unit] = make_float4(a, __uint_as_float(0u), 0.f, 0.f);
  }
}

// Short rows (segments below 1024 elements: four loads per lane already cover the row): the generic ROW traversal,
// eight blocks per SM (letting a warp of the lean kernel take two 512-element rows was slower: 60 us against 47).
template <int KIND>
__global__ void __launch_bounds__(kThreads, LQ_STAT_MINBLK) stat_short_kernel(const __grid_constant__ JobTable tab, unsigned* ctl) {
  using Op = typename StatKind<KIND>::Op;
  __shared__ float2 queue[1];
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const Job& j = tab.job[lo];
  const unsigned blk = blockIdx.x - j.block_begin;
  cudaGridDependencySynchronize();
  zero_ctl(ctl);
  Op op;
  if (j.vec) run_row<Op, false, true>(j, blk, op, queue);
  else run_row<Op, false, false>(j, blk, op, queue);
}

template <int KIND>
__device__ __noinline__ void stat_col_generic(const Job& j, unsigned blk, Acc* smem) {
  typename StatKind<KIND>::Op op;
  if (j.vec) run_col<typename StatKind<KIND>::Op, false, true>(j, blk, op, smem, nullptr);
  else run_col<typename StatKind<KIND>::Op, false, false>(j, blk, op, smem, nullptr);
}

// COL: a thread owns four adjacent columns (one 16-byte load per row) and walks down its rows, eight rows in flight.
template <int KIND>
__global__ void __launch_bounds__(kThreads, 4) stat_col_kernel(const __grid_constant__ JobTable tab, unsigned* ctl) {
  __shared__ Acc smem[4 * kThreads];
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const Job& j = tab.job[lo];
  const unsigned blk = blockIdx.x - j.block_begin;
  cudaGridDependencySynchronize();
  zero_ctl(ctl);
#if LQ_COL_ORDER
  const long long ncb_ = (j.ncol + j.bx - 1) / j.bx;     // column blocks of the same rows are neighbours in launch order
  const long long cb = blk % ncb_;
  const long long chunk = blk / ncb_;
#else
  const long long chunk = blk % j.nchunks;
  const long long cb = blk / j.nchunks;
#endif
  const int t = threadIdx.x;
  const int ty = t / j.bx;
  const int tx = t - ty * j.bx;
  const long long cv = cb * j.bx + tx;
  const bool active = (ty < j.by) && (cv < j.ncol);
  const long long col0 = cv * 4;
  const long long r0 = chunk * j.rpc;
  const long long r1 = (r0 + j.rpc < j.outer) ? r0 + j.rpc : j.outer;
  const int rows_blk = (int)((r1 - r0 + j.by - 1) / j.by);
  const int rows_mine = (active && j.vec) ? (int)((r1 - r0 - ty + j.by - 1) / j.by) : 0;
  const long long stride = (long long)j.by * j.W;
  const float* Pp = j.P + (active ? (r0 + ty) * j.W + col0 : 0);
  auto ldrow = [&](const float* q) {
    return j.stream_hint ? __ldcs(reinterpret_cast<const float4*>(q)) : *reinterpret_cast<const float4*>(q);
  };
  // the first eight rows are asked for BEFORE the scales are looked at (the regime test in front of the loop cost every
  // block one round trip to memory on sixteen bytes); a block that turns out not to be in the fast regime drops them
  float4 p[8];
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (u < rows_mine) p[u] = ldrow(Pp + u * stride);
  bool fast = j.vec != 0;
  if (fast && active) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float s = __ldg(j.s + (col0 + k) / j.inner);
      fast = fast && ((KIND == LQ_PEN_MAXBIN) ? maxbin_fast_scale(s) : diff_fast_scale(s));
    }
  }
  if (!__syncthreads_and(fast ? 1 : 0)) { stat_col_generic<KIND>(j, blk, smem); return; }   // block-uniform
  float v[4];                                       // MaxBin: max |P| ; Difference: sum |P|, per column
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (KIND == LQ_PEN_MAXBIN) ? -1.f : 0.f;
  bool bad = false;
#pragma unroll 1
  for (int i0 = 0; i0 < rows_blk; i0 += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (i0 + u < rows_mine) {
        if (KIND == LQ_PEN_MAXBIN) {
          v[0] = fmaxf(v[0], fabsf(p[u].x)); v[1] = fmaxf(v[1], fabsf(p[u].y));
          v[2] = fmaxf(v[2], fabsf(p[u].z)); v[3] = fmaxf(v[3], fabsf(p[u].w));
        } else {
          v[0] += fabsf(p[u].x); v[1] += fabsf(p[u].y); v[2] += fabsf(p[u].z); v[3] += fabsf(p[u].w);
        }
      }
    }
    Pp += 8 * stride;
#pragma unroll
    for (int u = 0; u < 8; ++u)                        // the next eight rows
      if (i0 + 8 + u < rows_mine) p[u] = ldrow(Pp + u * stride);
  }
  if (KIND == LQ_PEN_DIFFERENCE) {
    bad = !(((v[0] + v[1]) + (v[2] + v[3])) <= kFastHi);       // an inf or a NaN anywhere in the thread's columns
    if (__syncthreads_or(bad ? 1 : 0)) { stat_col_generic<KIND>(j, blk, smem); return; }
  }
  // fixed-shape tree over the threads that share a column (same tx, ty = 0..by-1)
  float* red = reinterpret_cast<float*>(smem);
  if (j.by > 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red[k * kThreads + t] = v[k];
    int span = 1;
    while (span < j.by) span <<= 1;
    for (int st = span >> 1; st >= 1; st >>= 1) {
      __syncthreads();
      if (ty < st && ty + st < j.by) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float o = red[k * kThreads + t + st * j.bx];
          red[k * kThreads + t] = (KIND == LQ_PEN_MAXBIN) ? fmaxf(red[k * kThreads + t], o) : red[k * kThreads + t] + o;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = red[k * kThreads + t];
  }
  if (ty == 0 && cv < j.ncol) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      j.part[chunk * j.W + col0 + k] = (KIND == LQ_PEN_MAXBIN) ? make_float4(0.f, __uint_as_float(0u), v[k], 0.f)
                                                                : make_float4(v[k], __uint_as_float(0u), 0.f, 0.f);
  }
}

// Which kernel takes which tensor (measured at 64 Mi): rows of a full 1024-element segment -> the lean ROW kernel
// (6.5 TB/s); shorter rows -> the generic ROW traversal, whose four loads per lane already cover the whole row (letting
// a warp take two 512-element rows was slower: 60 us against 47); interleaved layouts -> the lean COL kernel.
template <int KIND>
static int launch_stat(JobTable& tab, cudaStream_t st, unsigned* ctl) {
  JobTable part;
  for (int cls = 0; cls < 3; ++cls) {              // 0: ROW long rows (lean), 1: ROW short rows (generic), 2: COL (lean)
    part.n = 0;
    unsigned total = 0;
    for (int i = 0; i < tab.n; ++i) {
      const Job& tj = tab.job[i];
      if (tj.nblocks == 0) continue;
      const int c = (tj.regime == REG_COL) ? 2 : ((tj.vec && tj.seg == kSegMax) ? 0 : 1);
      if (c != cls) continue;
      part.job[part.n] = tj;
      part.job[part.n].block_begin = total;
      total += tj.nblocks;
      ++part.n;
    }
    if (total == 0) continue;
    int rc;
    if (cls == 0) rc = launch_after(stat_row_kernel<KIND>, dim3(total), dim3(kThreads), st, "loss-term statistics ROW launch", part, ctl);
    else if (cls == 2) rc = launch_after(stat_col_kernel<KIND>, dim3(total), dim3(kThreads), st, "loss-term statistics COL launch", part, ctl);
    else rc = launch_after(stat_short_kernel<KIND>, dim3(total), dim3(kThreads), st, "loss-term statistics ROW (short rows) launch", part, ctl);
    if (rc) return rc;
  }
  return LQ_OK;
}

// Per-tensor scalar stage: penalty = sum_t mean_t * dim_t / normalizer (LF:110-116 etc.)
struct ScalarJob { const float4* gstat; const float* s; long long C; double n; };
struct ScalarTable { int n, kind, first, last; double normalizer; ScalarJob job[kMaxMT]; };

__device__ __forceinline__ double gstat_double(const float4& v) {
  const unsigned long long bits = (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
  return __longlong_as_double((long long)bits);
}

__device__ __forceinline__ double scalar_group_value(int kind, const ScalarJob& j, long long g) {
  if (kind == LQ_PEN_MAXBIN) return (double)__ldcg(&j.gstat[g].x);          // written by other blocks of this launch:
  if (kind == LQ_PEN_DIFFERENCE) return gstat_double(__ldcg(&j.gstat[g]));    // read past L1
  const float sv = j.s[g];
  const float nz = (sv == 0.0f) ? kEps : sv;     // LF:255-256
  return (double)__fdiv_rn(1.0f, nz);            // LF:258
}
// mean over groups (MaxBin LF:110, Inverse LF:258) or over elements (Difference LF:178), times #P
__device__ __forceinline__ double scalar_term(int kind, const ScalarJob& j, double v) {
  const double denom = (kind == LQ_PEN_DIFFERENCE) ? j.n : (double)j.C;
  const float mean = (float)(v / denom);
  return (double)__fmul_rn(mean, (float)j.n);
}
__device__ __forceinline__ void scalar_finish(const ScalarTable& tab, const double* term, double* total, float* out) {
  double acc = tab.first ? 0.0 : *total;
  for (int t = 0; t < tab.n; ++t) acc += term[t];              // tensor order
  *total = acc;
  if (tab.last) *out = (float)(acc / tab.normalizer);
}

// Stand-alone form (Inverse, which has no element pass): warp w takes tensors w, w+8, ...: the per-tensor term needs one
// warp reduction, the terms are then added in tensor order by one thread (a single block used to walk the tensors one
// after the other with two block barriers each: 30 us for the 40 tensors of the IMAGENETTE model).
__global__ void __launch_bounds__(256) pen_scalar_kernel(const __grid_constant__ ScalarTable tab, double* total, float* out) {
  cudaGridDependencySynchronize();
  __shared__ double term[kMaxMT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < tab.n; t += 8) {
    const ScalarJob& j = tab.job[t];
    double v = 0.0;
    for (long long g = lane; g < j.C; g += 32) v += scalar_group_value(tab.kind, j, g);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) term[t] = scalar_term(tab.kind, j, v);
  }
  __syncthreads();
  if (threadIdx.x == 0) scalar_finish(tab, term, total, out);
}

// ---- fused epilogue (LQ_PEN_FUSED).  MaxBin / Difference used to end in up to three small dependent launches after the
// second stage (pending ties, scalar stage), each a drain and a ramp of the whole GPU; the scalar stage walked the tensors
// with one warp each.  Now the second-stage blocks of a tensor count themselves off (ctl[kCtlTensor0 + t]); the LAST block
// of tensor t -- all its groups' statistics are then in memory -- reduces them to the tensor's term with the whole block
// (thread i folds groups i, i + 128, ... in order, then a fixed tree: deterministic whichever block comes last), and the
// last tensor to finish adds the terms in tensor order.  The forward VALUE does not depend on the tie counts, so it is
// complete when the second stage ends; the tail kernel (MaxBin) only serves groups whose ties were left pending and
// returns at once when there are none (ctl[kCtlPending]).
#ifndef LQ_PEN_FUSED
#define LQ_PEN_FUSED 1
#endif

// true in every thread of the block that arrives last on *counter (of `expected` blocks)
__device__ __forceinline__ bool last_block(unsigned* counter, unsigned expected) {
  __shared__ unsigned ticket;
  __syncthreads();                                        // the block's own writes are done
  if (threadIdx.x == 0) { __threadfence(); ticket = atomicAdd(counter, 1u); }
  __syncthreads();
  const bool last = ticket == expected - 1u;
  if (last) __threadfence();
  return last;
}

// called by every second-stage block of tensor t (kFinThreads threads) after its group is final
__device__ __forceinline__ void pen_fused_epilogue(const ScalarTable& sc, int t, unsigned* ctl, double* term, double* total, float* out) {
  const ScalarJob& j = sc.job[t];
  if (!last_block(&ctl[kCtlTensor0 + t], (unsigned)j.C)) return;
  double v = 0.0;
  for (long long g0 = threadIdx.x; g0 < j.C; g0 += 4 * kFinThreads) {       // four loads in flight
    double x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long g = g0 + (long long)u * kFinThreads;
      x[u] = (g < j.C) ? scalar_group_value(sc.kind, j, g) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) v += x[u];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double wsum[kFinThreads / 32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    v = wsum[0];
#pragma unroll
    for (int i = 1; i < kFinThreads / 32; ++i) v += wsum[i];
    term[t] = scalar_term(sc.kind, j, v);
    __threadfence();
    if (atomicAdd(&ctl[kCtlTensors], 1u) == (unsigned)sc.n - 1u) {          // every tensor's term is in memory
      __threadfence();
      double acc = sc.first ? 0.0 : *total;
      for (int i = 0; i < sc.n; ++i) acc += __ldcg(&term[i]);
      *total = acc;
      if (sc.last) *out = (float)(acc / sc.normalizer);
    }
  }
}

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
  const bool sums = f.mode != PFIN_MAXBIN;        // Difference: .x and .z are both sums
  double sum = 0.0, sum_c = 0.0;
  float mx = -1.f;
  unsigned long long ties = 0ull;
  float4 v[kFinChunk / 256];
#pragma unroll
  for (int u = 0; u < kFinChunk / 256; ++u) {
    const long long k = k0 + u * 256 + threadIdx.x;
    v[u] = (k < K) ? f.part[partial_index(f, g, k)] : make_float4(0.f, 0.f, sums ? 0.f : -2.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < kFinChunk / 256; ++u) {
    sum += (double)v[u].x;
    if (sums) { sum_c += (double)v[u].z; continue; }
    const unsigned long long b = __float_as_uint(v[u].y);
    if (v[u].z > mx) { mx = v[u].z; ties = b; } else if (v[u].z == mx) ties += b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sum_c += __shfl_xor_sync(0xffffffffu, sum_c, o);
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const unsigned long long ot = __shfl_xor_sync(0xffffffffu, ties, o);
    if (omx > mx) { mx = omx; ties = ot; } else if (omx == mx) ties += ot;
  }
  __shared__ double ssum[8], ssumc[8];
  __shared__ unsigned long long sties[8];
  __shared__ float smx[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; ssumc[w] = sum_c; sties[w] = ties; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; sum_c = ssumc[0]; ties = sties[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      sum += ssum[i]; sum_c += ssumc[i];
      if (smx[i] > mx) { mx = smx[i]; ties = sties[i]; } else if (smx[i] == mx) ties += sties[i];
    }
    f.lin[blockIdx.x] = make_float4((float)sum, __uint_as_float((unsigned)(ties > 0xffffffffull ? 0xffffffffull : ties)),
                                    sums ? (float)sum_c : mx, 0.f);
  }
}

// MaxBin, fast regime: the block that has just reduced group g counts the elements whose quotient equals the
// group's maximum.  Only first-level partials (a ROW unit, or one column of a COL row chunk) whose own max |P|
// reaches lo can hold such elements -- normally one per group: they are collected in a small shared list and
// re-read by the whole block.  Integer counts: the result does not depend on any order.  Groups whose candidates
// do not fit the list, and large groups that must compare quotients (lo = -2), are left to pen_ties_kernel.
constexpr int kTieList = 256;

__device__ __forceinline__ long long partial_index1(const FinJob& f, long long g, long long k) {
  if (f.regime1 == REG_ROW) {
    const long long o = k / f.nseg, sg = k - o * f.nseg;
    return (o * f.C + g) * f.nseg + sg;
  }
  const long long ch = k / f.inner, ci = k - ch * f.inner;
  return ch * f.W + g * f.inner + ci;
}

__device__ void pen_count_ties(const FinJob& f, long long g, float lo, float mb, unsigned* ctl) {
  __shared__ long long list[kTieList];
  __shared__ int nlist, nlin_list;
  __shared__ unsigned total;
  __shared__ long long lin_list[32];
  const long long K1 = (f.regime1 == REG_ROW) ? f.outer * f.nseg : f.nchunks * f.inner;   // first-level partials of the group
  const bool all = lo == -2.f;
  if (threadIdx.x == 0) { nlist = 0; nlin_list = 0; total = 0u; }
  __syncthreads();
  bool overflow = all && K1 > kTieList;
  if (!overflow) {
    if (f.regime == REG_LIN) {
      // two levels: the intermediate partials that reach lo first, then the first-level partials behind them
      for (long long k = threadIdx.x; k < f.nlin; k += kFinThreads)
        if (all || f.part[g * f.nlin + k].z >= lo) { const int i = atomicAdd(&nlin_list, 1); if (i < 32) lin_list[i] = k; }
      __syncthreads();
      const int nl = nlin_list;
      if (nl > 32) overflow = true;
      for (int i = 0; i < nl && !overflow; ++i) {
        const long long k0 = lin_list[i] * kFinChunk;
        for (long long k = k0 + threadIdx.x; k < k0 + kFinChunk && k < K1; k += kFinThreads)
          if (all || f.part1[partial_index1(f, g, k)].z >= lo) { const int n = atomicAdd(&nlist, 1); if (n < kTieList) list[n] = k; }
      }
    } else {
      // same walk as the reduction above (four loads in flight; no index arithmetic where partials sit at a stride)
      const bool simple = (f.regime1 == REG_ROW && f.nseg == 1) || (f.regime1 == REG_COL && f.inner == 1);
      const long long stride = (f.regime1 == REG_ROW) ? f.C : f.W;
      for (long long k0 = threadIdx.x; k0 < K1; k0 += 4 * kFinThreads) {
        float z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long k = k0 + (long long)u * kFinThreads;
          z[u] = (k < K1) ? (simple ? f.part1[g + k * stride].z : f.part1[partial_index1(f, g, k)].z) : -1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long k = k0 + (long long)u * kFinThreads;
          if (k < K1 && (all || z[u] >= lo)) { const int n = atomicAdd(&nlist, 1); if (n < kTieList) list[n] = k; }
        }
      }
    }
    __syncthreads();
    if (nlist > kTieList) overflow = true;
  }
  if (overflow) {                               // block-uniform: mark the group pending (.z = 1) for pen_ties_kernel
    if (threadIdx.x == 0) { f.gstat[g].z = 1.f; if (ctl) atomicAdd(&ctl[kCtlPending], 1u); }
    return;
  }
  const float sc = all ? f.s[g] : 1.f;
  unsigned cnt = 0;
  const int n = nlist;
  for (int i = 0; i < n; ++i) {
    const long long k = list[i];
    if (f.regime1 == REG_ROW) {
      const long long o = k / f.nseg, sg = k - o * f.nseg;
      const long long base = (o * f.C + g) * f.inner + sg * f.seg;
      const long long rem = f.inner - sg * f.seg;
      const int len = rem < f.seg ? (int)rem : f.seg;
      // all of the unit's loads in flight at once (the tensor has been streamed through: these come from HBM, and one
      // load per trip made this the longest step of the second stage)
      for (int e0 = threadIdx.x; e0 < len; e0 += 8 * kFinThreads) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int e = e0 + u * kFinThreads; v[u] = (e < len) ? f.P[base + e] : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) if (e0 + u * kFinThreads < len) cnt += tie_hit(v[u], lo, mb, sc);
      }
    } else {
      const long long ch = k / f.inner, ci = k - ch * f.inner;
      const long long col = g * f.inner + ci;
      const long long r0 = ch * f.rpc;
      const long long r1 = (r0 + f.rpc < f.outer) ? r0 + f.rpc : f.outer;
      for (long long q0 = r0 + threadIdx.x; q0 < r1; q0 += 8 * kFinThreads) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const long long r = q0 + u * kFinThreads; v[u] = (r < r1) ? f.P[r * f.W + col] : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) if (q0 + u * kFinThreads < r1) cnt += tie_hit(v[u], lo, mb, sc);
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&total, cnt);
  __syncthreads();
  if (threadIdx.x == 0) f.gstat[g].y = __uint_as_float(total);
}

__device__ __forceinline__ void pen_finalize_group(const FinJob& f, long long g, unsigned* ctl) {
  const long long K = partials_per_group(f);
  double sum = 0.0, sum_c = 0.0;
  float mx = -1.f;
  unsigned long long ties = 0ull;
  // four independent 16-byte loads per thread per trip, and no 64-bit division on the layouts whose
  // partials sit at a constant stride (one load at a time behind partial_index() cost 17 us for the
  // 2048 partials per group of an interleaved 64 Mi tensor); thread t still folds t, t+128, ... in order
  const bool simple = (f.regime == REG_ROW && f.nseg == 1) || f.regime == REG_LIN || (f.regime == REG_COL && f.inner == 1);
  const long long stride = (f.regime == REG_ROW) ? f.C : (f.regime == REG_COL ? f.W : 1);
  const float4* base = f.part + ((f.regime == REG_LIN) ? g * f.nlin : g);
  const bool sums = f.mode != PFIN_MAXBIN;
  for (long long k0 = threadIdx.x; k0 < K; k0 += 4 * kFinThreads) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = k0 + (long long)u * kFinThreads;
      v[u] = (k < K) ? (simple ? base[k * stride] : f.part[partial_index(f, g, k)]) : make_float4(0.f, 0.f, sums ? 0.f : -2.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sum += (double)v[u].x;
      if (sums) { sum_c += (double)v[u].z; continue; }      // Difference: .z is the sum of the exactly evaluated terms
      const unsigned long long b = __float_as_uint(v[u].y);
      if (v[u].z > mx) { mx = v[u].z; ties = b; } else if (v[u].z == mx) ties += b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sum_c += __shfl_xor_sync(0xffffffffu, sum_c, o);
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const unsigned long long ot = __shfl_xor_sync(0xffffffffu, ties, o);
    if (omx > mx) { mx = omx; ties = ot; } else if (omx == mx) ties += ot;
  }
  __shared__ double ssum[kFinThreads / 32], ssumc[kFinThreads / 32];
  __shared__ unsigned long long sties[kFinThreads / 32];
  __shared__ float smx[kFinThreads / 32];
  __shared__ float tie_lo, tie_mb;            // MaxBin fast regime: what the block counts after the reduction
  if (threadIdx.x == 0) tie_lo = -1.f;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; ssumc[w] = sum_c; sties[w] = ties; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; sum_c = ssumc[0]; ties = sties[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < kFinThreads / 32; ++i) {
      sum += ssum[i]; sum_c += ssumc[i];
      if (smx[i] > mx) { mx = smx[i]; ties = sties[i]; } else if (smx[i] == mx) ties += sties[i];
    }
    const float sc = f.s[g];
    if (f.mode == PFIN_MAXBIN) {
      const unsigned t32 = (unsigned)(ties > 0xffffffffull ? 0xffffffffull : ties);
      if (!maxbin_fast_scale(sc) || mx < 0.f) {
        // exact regime (mx is the largest quotient, ties counted by the element pass), or nothing but NaNs
        f.gstat[g] = make_float4(mx, __uint_as_float(t32), 0.f, -1.f);
      } else {
        // fast regime: mx is max |P| and the maximum is RN(max |P| / s) (division by s > 0 is monotone): ONE
        // division per group.  The tied elements are those with |P| >= lo, lo = the smallest float whose quotient
        // still rounds to the maximum (a handful of ulps below mx at most); when max |P| is zero, denormal-small,
        // huge or infinite they are found by comparing quotients instead (lo = -2).  Counted below by the block.
        const float mb = __fdiv_rn(mx, sc);
        float lo = -2.f;
        if (mx >= kFastLo && mx <= kFastHi) {
          lo = mx;
          for (int i = 0; i < 16; ++i) {
            const float a = __uint_as_float(__float_as_uint(lo) - 1u);
            if (__fdiv_rn(a, sc) != mb) break;
            lo = a;
          }
        }
        tie_lo = lo; tie_mb = mb;
        f.gstat[g] = make_float4(mb, __uint_as_float(0u), 0.f, lo);
      }
    } else if (f.mode == PFIN_SUM) {
      // Difference: fast part sum |P| * |1 - 1/s| (double) + the exactly evaluated terms
      const double k = diff_fast_scale(sc) ? fabs(1.0 - 1.0 / (double)sc) : 0.0;
      const double val = sum * k + sum_c;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(val);
      const unsigned long long abits = (unsigned long long)__double_as_longlong(sum);
      f.gstat[g] = make_float4(__uint_as_float((unsigned)(bits & 0xffffffffull)), __uint_as_float((unsigned)(bits >> 32)),
                               __uint_as_float((unsigned)(abits & 0xffffffffull)), __uint_as_float((unsigned)(abits >> 32)));
    } else {   // PFIN_SUM_DS: exactly evaluated terms + the closed form of the fast part (see DiffBwdOp)
      double v = sum;
      if (diff_fast_scale(sc)) {
        const float4 st = f.gstat[g];
        const unsigned long long abits = (unsigned long long)__float_as_uint(st.z) | ((unsigned long long)__float_as_uint(st.w) << 32);
        const double S = __longlong_as_double((long long)abits);
        const double gabs = (double)__fmul_rn(f.p0, *f.up);
        v += gabs * (sc <= 0.5f ? -1.0 : 1.0) * S / ((double)sc * (double)sc);
      }
      const float r = (float)v;
      f.ds[g] = (f.flags & LQ_F_ACCUM_DS) ? __fadd_rn(f.ds[g], r) : r;
    }
  }
  if (f.mode != PFIN_MAXBIN) return;
  __syncthreads();
  const float lo = tie_lo, mb = tie_mb;
  if (lo == -1.f) return;                      // exact regime / nothing but NaNs: already final
  pen_count_ties(f, g, lo, mb, ctl);
}

__global__ void __launch_bounds__(kFinThreads) pen_finalize_kernel(const __grid_constant__ FinTable tab,
                                                                   const __grid_constant__ ScalarTable sc, unsigned* ctl,
                                                                   double* term, double* total, float* out) {
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].group_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const FinJob& f = tab.job[lo];
  cudaGridDependencySynchronize();
  pen_finalize_group(f, blockIdx.x - f.group_begin, ctl);
  if (ctl) pen_fused_epilogue(sc, lo, ctl, term, total, out);     // (null: the backward's second stage, or the unfused chain)
}

// The groups pen_count_ties left pending (gstat.z == 1: more candidate units than its list holds -- many equal
// maxima -- or a large group that must compare quotients, gstat.w == -2): thread <-> first-level partial, a warp
// re-reads the flagged ones.  Rare by construction; a fixed small grid walks all partials.
struct TieTable { int n; unsigned begin[kMaxMT]; };
constexpr unsigned kTieBlocksPerTensor = 64;

__global__ void __launch_bounds__(256) pen_ties_kernel(const __grid_constant__ JobTable tab, const __grid_constant__ TieTable tt,
                                                       const unsigned* ctl) {
  if (ctl) {
    cudaGridDependencySynchronize();
    if (ctl[kCtlPending] == 0u) return;                   // the usual case: the second stage has counted every group's ties
  }
  int lo_i = 0, hi_i = tt.n - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (tt.begin[mid] <= blockIdx.x) lo_i = mid; else hi_i = mid - 1;
  }
  const Job& j = tab.job[lo_i];
  cudaGridDependencySynchronize();
  float4* gstat = const_cast<float4*>(j.gaux);
  const long long npart = (j.regime == REG_ROW) ? j.units : j.nchunks * j.W;
  const unsigned nblk = ((lo_i + 1 < tt.n) ? tt.begin[lo_i + 1] : gridDim.x) - tt.begin[lo_i];
  const int lane = threadIdx.x & 31;
  for (long long s0 = (long long)(blockIdx.x - tt.begin[lo_i]) * 256; s0 < npart; s0 += 256ll * nblk) {
    const long long slot = s0 + threadIdx.x;
    long long a = 0, b = 0, grp = 0;      // ROW: a = row, b = segment;  COL: a = chunk, b = column
    float lo = -1.f, mb = 0.f;
    bool flagged = false;
    if (slot < npart) {
      if (j.regime == REG_ROW) { a = slot / j.nseg; b = slot - a * j.nseg; grp = a % j.C; }
      else { a = slot / j.W; b = slot - a * j.W; grp = b / j.inner; }
      const float4 st = gstat[grp];
      lo = st.w; mb = st.x;
      flagged = st.z == 1.f && (lo == -2.f || j.part[slot].z >= lo);
    }
    unsigned todo = __ballot_sync(0xffffffffu, flagged);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const long long ua = __shfl_sync(0xffffffffu, a, src), ub = __shfl_sync(0xffffffffu, b, src);
      const long long ug = __shfl_sync(0xffffffffu, grp, src);
      const float ulo = __shfl_sync(0xffffffffu, lo, src), umb = __shfl_sync(0xffffffffu, mb, src);
      const float sc = (ulo == -2.f) ? __ldg(j.s + ug) : 1.f;
      unsigned cnt = 0;
      if (j.regime == REG_ROW) {
        const long long base = ua * j.inner + ub * j.seg;
        const long long rem = j.inner - ub * j.seg;
        const int len = rem < j.seg ? (int)rem : j.seg;
        for (int e = lane; e < len; e += 32) cnt += tie_hit(j.P[base + e], ulo, umb, sc);
      } else {
        const long long r0 = ua * j.rpc;
        const long long r1 = (r0 + j.rpc < j.outer) ? r0 + j.rpc : j.outer;
        for (long long r = r0 + lane; r < r1; r += 32) cnt += tie_hit(j.P[r * j.W + ub], ulo, umb, sc);
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0 && cnt) atomicAdd(reinterpret_cast<unsigned*>(&gstat[ug].y), cnt);
    }
  }
}

// --------------------------------------------------------------- lean backward kernel, interleaved layouts
// The generic COL traversal carries a full operator context per column (reciprocal, operands, three constants: ten
// registers times four columns) and ran the Difference backward at 4.8 TB/s on [131072, 512] (64 registers, spilling).
// In the fast regime the element gradient needs one constant per column (Difference: g_P = sign(P) * (sigma * k1);
// MaxBin: two, the threshold |P| >= lo and coef / s), so this kernel keeps eight rows in flight per thread like the
// forward statistics, on the same geometry (LONG row chunks: the Difference partials shrink from 6 % of the traffic to
// under 1 %, and forward and backward cut the tensor into the same quads, so they agree on which elements are "ordinary").
// Quads that are not ordinary (an inf / NaN / huge value among four rows of a column) take the exact per-element form
// and accumulate into the column's partial exactly as the generic traversal does; blocks whose scales are outside the
// fast set go to the generic traversal, out of line.
template <int KIND, bool ACCUM> struct BwdKind;
template <bool ACCUM> struct BwdKind<LQ_PEN_MAXBIN, ACCUM> { using Op = MaxBinBwdOp<ACCUM>; };
template <bool ACCUM> struct BwdKind<LQ_PEN_DIFFERENCE, ACCUM> { using Op = DiffBwdOp<ACCUM>; };

// Difference, lean kernel below: up to four rows of a thread's four columns that are NOT all ordinary (an inf / NaN /
// huge value in a quad), or the last 1-3 rows of a chunk -- re-read from memory and processed exactly as
// DiffBwdOp::quad (full = four rows: the quad's split) / ::elem (per element) do.  Returns the four columns' terms of
// the scale gradient.  Rare and out of line, so that the streaming loop keeps its registers.
template <bool ACCUM>
__device__ __noinline__ float4 diff_bwd_rows_slow(const Job& j, const float* Pp, float* Xp, long long stride, int nrows,
                                                  long long col0, float gabs) {
  float accv[4] = {0.f, 0.f, 0.f, 0.f};
  const bool full = nrows == 4;
  for (int k = 0; k < 4; ++k) {
    const float s = __ldg(j.s + (col0 + k) / j.inner);
    const float k1 = __fadd_rn(gabs, __fdiv_rn(-gabs, s));
    const float sigma = diff_fast_scale(s) ? (s <= 0.5f ? -1.f : 1.f) : 0.f;
    float pv[4];
    for (int u = 0; u < 4; ++u) pv[u] = (u < nrows) ? Pp[u * stride + k] : 0.f;
    const float tq = (fabsf(pv[0]) + fabsf(pv[1])) + (fabsf(pv[2]) + fabsf(pv[3]));
    for (int u = 0; u < nrows; ++u) {
      const bool ord = sigma != 0.f && (full ? (tq <= kFastHi) : (fabsf(pv[u]) <= kFastHi));
      float g;
      if (ord) {
        g = __fmul_rn(__fmul_rn(sigma, sgn(pv[u])), k1);
      } else {
        const float pq = __fdiv_rn(pv[u], s);
        const float sg = sgn(__fsub_rn(pv[u], pq));
        g = __fmul_rn(sg, k1);
        accv[k] += __fmul_rn(__fmul_rn(-sg, gabs), __fdiv_rn(-pq, s));
      }
      float* x = Xp + u * stride + k;
      *x = ACCUM ? __fadd_rn(*x, g) : g;
    }
  }
  return make_float4(accv[0], accv[1], accv[2], accv[3]);
}

// four rows x four columns: is every column's quad ordinary (DiffBwdOp::quad's test, all four at once)?
__device__ __forceinline__ bool diff_quads_ordinary(const float4& a, const float4& b, const float4& c, const float4& d) {
  const float t0 = (fabsf(a.x) + fabsf(b.x)) + (fabsf(c.x) + fabsf(d.x));
  const float t1 = (fabsf(a.y) + fabsf(b.y)) + (fabsf(c.y) + fabsf(d.y));
  const float t2 = (fabsf(a.z) + fabsf(b.z)) + (fabsf(c.z) + fabsf(d.z));
  const float t3 = (fabsf(a.w) + fabsf(b.w)) + (fabsf(c.w) + fabsf(d.w));
  return t0 <= kFastHi && t1 <= kFastHi && t2 <= kFastHi && t3 <= kFastHi;      // a NaN fails its compare
}

template <int KIND, bool ACCUM>
__device__ __noinline__ void pen_bwd_col_generic(const Job& j, unsigned blk, Acc* smem, const float* up) {
  using Op = typename BwdKind<KIND, ACCUM>::Op;
  Op op; op.up = up;
  if (j.vec) run_col<Op, false, true>(j, blk, op, smem, nullptr);
  else run_col<Op, false, false>(j, blk, op, smem, nullptr);
}

#ifndef LQ_PENB_MINBLK
#define LQ_PENB_MINBLK 3
#endif
template <int KIND, bool ACCUM>
__global__ void __launch_bounds__(kThreads, LQ_PENB_MINBLK) pen_bwd_col_kernel(const __grid_constant__ JobTable tab, const float* up) {
  constexpr bool kDiff = KIND == LQ_PEN_DIFFERENCE;
  constexpr int U = ACCUM ? 4 : 8;                  // rows in flight per thread (two arrays when accumulating)
  __shared__ Acc smem[kDiff ? 4 * kThreads : 1];
  int lo_i = 0, hi_i = tab.n - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo_i = mid; else hi_i = mid - 1;
  }
  const Job& j = tab.job[lo_i];
  const unsigned blk = blockIdx.x - j.block_begin;
  cudaGridDependencySynchronize();
#if LQ_COL_ORDER
  const long long ncb_ = (j.ncol + j.bx - 1) / j.bx;
  const long long cb = blk % ncb_;
  const long long chunk = blk / ncb_;
#else
  const long long chunk = blk % j.nchunks;
  const long long cb = blk / j.nchunks;
#endif
  const int t = threadIdx.x;
  const int ty = t / j.bx;
  const int tx = t - ty * j.bx;
  const long long cv = cb * j.bx + tx;
  const bool active = (ty < j.by) && (cv < j.ncol);
  const long long col0 = cv * 4;
  const long long r0 = chunk * j.rpc;
  const long long r1 = (r0 + j.rpc < j.outer) ? r0 + j.rpc : j.outer;
  const int rows_blk = (int)((r1 - r0 + j.by - 1) / j.by);
  const int rows_mine = (active && j.vec) ? (int)((r1 - r0 - ty + j.by - 1) / j.by) : 0;
  const long long stride = (long long)j.by * j.W;
  const long long base = active ? (r0 + ty) * j.W + col0 : 0;
  const float* Pp = j.P + base;
  float* Xp = j.dX + base;
  // the first trip's rows are asked for BEFORE the per-column constants (several IEEE divisions) and the regime vote
  float4 p[U], old[ACCUM ? U : 1];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < rows_mine) {
      p[u] = *reinterpret_cast<const float4*>(Pp + u * stride);
      if (ACCUM) old[ACCUM ? u : 0] = *reinterpret_cast<const float4*>(Xp + u * stride);
    }
  }
  // per-column constants.  Difference: kk = sigma * k1 (g_P = sign(d) * k1, sign(d) = sigma * sign(P); the product
  // (sigma * sign(P)) * k1 of DiffBwdOp::fast1 has the same bits: sigma is +-1).  MaxBin: lo, k0 = coef / s, k1 = 0 / s.
  bool fast = j.vec != 0;
  float kk[4], thr[4];                                // Difference: kk;  MaxBin: kk = coef / s, thr = lo
#pragma unroll
  for (int k = 0; k < 4; ++k) { kk[k] = 0.f; thr[k] = 0.f; }
  const float gabs = kDiff ? __fmul_rn(j.t0, __ldg(up)) : 0.f;
  if (fast && active) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long grp = (col0 + k) / j.inner;
      const float s = __ldg(j.s + grp);
      if (kDiff) {
        fast = fast && diff_fast_scale(s);
        kk[k] = __fmul_rn(s <= 0.5f ? -1.f : 1.f, __fadd_rn(gabs, __fdiv_rn(-gabs, s)));
      } else {
        const float4 aux = __ldg(j.gaux + grp);
        fast = fast && aux.w >= 0.f;                  // the group's ties are the elements with |P| >= lo; s > 0 there,
        thr[k] = aux.w;                               // so MaxBinBwdOp's other constant, 0 / s, is +0
        kk[k] = __fdiv_rn(aux.z, s);
      }
    }
  }
  if (!__syncthreads_and(fast ? 1 : 0)) { pen_bwd_col_generic<KIND, ACCUM>(j, blk, smem, up); return; }   // block-uniform
  float acc[4] = {0.f, 0.f, 0.f, 0.f};               // Difference: the exactly evaluated terms of the scale gradient
  bool any_exact = false, redo = false;
  const float* const Pp0 = Pp;
  float* const Xp0 = Xp;
#pragma unroll 1
  for (int i0 = 0; i0 < rows_blk; i0 += U) {
#pragma unroll
    for (int q0 = 0; q0 < U; q0 += 4) {               // the generic traversal's quads: four rows of one column
      const int left = rows_mine - (i0 + q0);         // this thread's rows from here on
      if (left <= 0) continue;
      if (kDiff) {
        if (left < 4) continue;                       // the chunk's last 1-3 rows: after the loop
        if (!diff_quads_ordinary(p[q0], p[q0 + 1], p[q0 + 2], p[q0 + 3])) { redo = true; continue; }   // after the loop
#pragma unroll
        for (int u = q0; u < q0 + 4; ++u) {
          float4 g;
          g.x = __fmul_rn(sgn(p[u].x), kk[0]); g.y = __fmul_rn(sgn(p[u].y), kk[1]);
          g.z = __fmul_rn(sgn(p[u].z), kk[2]); g.w = __fmul_rn(sgn(p[u].w), kk[3]);
          if (ACCUM) {
            const float4 o4 = old[ACCUM ? u : 0];
            g.x = __fadd_rn(o4.x, g.x); g.y = __fadd_rn(o4.y, g.y); g.z = __fadd_rn(o4.z, g.z); g.w = __fadd_rn(o4.w, g.w);
          }
          *reinterpret_cast<float4*>(Xp + u * stride) = g;
        }
      } else {
#pragma unroll
        for (int u = q0; u < q0 + 4; ++u) {
          if (u - q0 < left) {
            float4 g;
            g.x = __fmul_rn((fabsf(p[u].x) >= thr[0]) ? kk[0] : 0.0f, sgn(p[u].x));
            g.y = __fmul_rn((fabsf(p[u].y) >= thr[1]) ? kk[1] : 0.0f, sgn(p[u].y));
            g.z = __fmul_rn((fabsf(p[u].z) >= thr[2]) ? kk[2] : 0.0f, sgn(p[u].z));
            g.w = __fmul_rn((fabsf(p[u].w) >= thr[3]) ? kk[3] : 0.0f, sgn(p[u].w));
            if (ACCUM) {
              const float4 o4 = old[ACCUM ? u : 0];
              g.x = __fadd_rn(o4.x, g.x); g.y = __fadd_rn(o4.y, g.y); g.z = __fadd_rn(o4.z, g.z); g.w = __fadd_rn(o4.w, g.w);
            }
            *reinterpret_cast<float4*>(Xp + u * stride) = g;
          }
        }
      }
    }
    Pp += U * stride;
    Xp += U * stride;
#pragma unroll
    for (int u = 0; u < U; ++u) {                     // the next trip's rows
      if (i0 + U + u < rows_mine) {
        p[u] = *reinterpret_cast<const float4*>(Pp + u * stride);
        if (ACCUM) old[ACCUM ? u : 0] = *reinterpret_cast<const float4*>(Xp + u * stride);
      }
    }
  }
  if (!kDiff) return;
  // What the streaming loop left out (no call in there: it would spill the registers that hold loads in flight): the
  // chunk's last 1-3 rows, and -- only if this thread met one -- the quads that are not ordinary, found again by
  // re-reading the thread's rows.
  if (rows_mine & 3) {
    const long long off = (long long)(rows_mine & ~3) * stride;
    const float4 a4 = diff_bwd_rows_slow<ACCUM>(j, Pp0 + off, Xp0 + off, stride, rows_mine & 3, col0, gabs);
    acc[0] += a4.x; acc[1] += a4.y; acc[2] += a4.z; acc[3] += a4.w;
    any_exact = true;
  }
  if (redo) {
    for (int q = 0; q + 4 <= rows_mine; q += 4) {
      const float* pq = Pp0 + (long long)q * stride;
      if (diff_quads_ordinary(*reinterpret_cast<const float4*>(pq), *reinterpret_cast<const float4*>(pq + stride),
                              *reinterpret_cast<const float4*>(pq + 2 * stride), *reinterpret_cast<const float4*>(pq + 3 * stride)))
        continue;
      const float4 a4 = diff_bwd_rows_slow<ACCUM>(j, pq, Xp0 + (long long)q * stride, stride, 4, col0, gabs);
      acc[0] += a4.x; acc[1] += a4.y; acc[2] += a4.z; acc[3] += a4.w;
    }
    any_exact = true;
  }
  // the column partials of the scale gradient (zero unless some quad took the exact form): same tree as run_col
  if (__syncthreads_or(any_exact ? 1 : 0) && j.by > 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { Acc a; a.a = acc[k]; a.b = 0u; a.c = 0.f; smem[k * kThreads + t] = a; }
    int span = 1;
    while (span < j.by) span <<= 1;
    for (int st = span >> 1; st >= 1; st >>= 1) {
      __syncthreads();
      if (ty < st && ty + st < j.by) {
#pragma unroll
        for (int k = 0; k < 4; ++k) smem[k * kThreads + t].a += smem[k * kThreads + t + st * j.bx].a;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = smem[k * kThreads + t].a;
  }
  if (ty == 0 && cv < j.ncol) {
#pragma unroll
    for (int k = 0; k < 4; ++k) j.part[chunk * j.W + col0 + k] = make_float4(acc[k], __uint_as_float(0u), 0.f, 0.f);
  }
}

// ROW tensors through the generic multi-tensor traversal, interleaved ones through the lean kernel above
template <int KIND, bool ACCUM>
static int launch_pen_bwd(JobTable& tab, const float* up, cudaStream_t st) {
  using Op = typename BwdKind<KIND, ACCUM>::Op;
  JobTable rows, cols;
  rows.n = cols.n = 0;
  unsigned total = 0;
  for (int i = 0; i < tab.n; ++i) {
    const Job& tj = tab.job[i];
    if (tj.nblocks == 0) continue;
    // (Difference on short interleaved tensors -- a model's kernels, a few rows per thread -- stays on the generic
    // traversal: there the lean kernel's per-block constants and partial tree outweigh its loop, 35 us against 31 for
    // the 40-tensor inventory)
    if (tj.regime == REG_COL && tj.vec && (KIND == LQ_PEN_MAXBIN || tj.rpc / tj.by >= 16)) {
      cols.job[cols.n] = tj;
      cols.job[cols.n].block_begin = total;
      total += tj.nblocks;
      ++cols.n;
    } else {
      rows.job[rows.n++] = tj;
    }
  }
  if (rows.n) { const int rc = launch_mt(rows, make_pen_op<Op>(up), st); if (rc) return rc; }
  if (total) return launch_after(pen_bwd_col_kernel<KIND, ACCUM>, dim3(total), dim3(kThreads), st, "loss-term backward COL launch", cols, up);
  return LQ_OK;
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

// Geometry of the forward statistics on interleaved layouts: narrow column blocks (a warp reads 512 contiguous bytes
// per row), eight threads deep, LONG row chunks -- the second stage walks chunks-per-group partials, and with the
// element kernels' 64-row chunks a [131072, 512] tensor had 2048 of them per group (25 us of second stage for a
// 45 us pass).  Never more partials than the default geometry the workspace is sized for.
#ifndef LQ_STAT_BX
#define LQ_STAT_BX 32
#endif
#ifndef LQ_STAT_BPS
#define LQ_STAT_BPS 8
#endif
// target: blocks a multi-tensor launch grants this tensor (mt_col_targets; 0 = a launch of its own): the chunks only grow.
static Geom make_stat_geom(long long outer, long long C, long long inner, bool aligned, long long target = 0) {
  Geom g = make_geom(outer, C, inner, aligned);
  if (g.regime != REG_COL || !g.vec_ok) return g;
  const int bx = g.ncol < LQ_STAT_BX ? (int)g.ncol : LQ_STAT_BX;
  const int by = kThreads / bx;
  const long long ncolblk = cdiv(g.ncol, bx);
  long long rpc = cdiv(cdiv(outer * ncolblk, 148ll * LQ_STAT_BPS), by) * by;
  if (rpc < by) rpc = by;
  if (rpc > 1024) rpc = 1024 / by * by;
  const long long shared = share_rpc(outer, ncolblk, target, by, (long long)bx * 4);
  if (shared > rpc && shared > g.rpc) rpc = shared;    // the launch's 40 tensors share one grid (see mt_col_targets)
  if (rpc > outer) rpc = outer;
  if (rpc <= g.rpc) return g;                       // short tensors: the default chunks are already the longer ones
  g.bx = bx; g.by = by; g.rpc = (int)rpc;
  g.nchunks = cdiv(outer, rpc);
  g.ncolblk = ncolblk;
  g.nblocks = g.nchunks * ncolblk;
  g.npart1 = g.nchunks * g.W;
  g.kper = g.nchunks * inner;
  g.nlin = g.kper > kFinSplit ? cdiv(g.kper, kFinChunk) : 0;
  g.npartials = g.npart1 + C * g.nlin;
  return g;
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
  return off;   // scalar slot (kPenSlotBytes) follows
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
  const size_t need = pen_layout(descs, n, lay) + kPenSlotBytes;
  if (!ws || ws_bytes < need || !aligned16(ws)) {
    delete[] lay;
    set_error("%s: workspace %zu B (16-B aligned) needed, got %zu", who, need, ws_bytes);
    return LQ_E_WORKSPACE;
  }
  *lay_out = lay; *scalar_off = need - kPenSlotBytes; *normalizer = nrm;
  return LQ_OK;
}

}  // namespace lq

using namespace lq;

// fwd != nullptr: the element pass is the fake-quant forward itself (FwdPenOp: writes desc.y, statistics as by-products)
static int pen_fwd_impl(const char* who, int kind, const lq_tensor_desc* descs, int n, float* out_penalty,
                        void* ws, size_t ws_bytes, cudaStream_t st, const QMode* fwd) {
  PenLayout* lay = nullptr; size_t soff = 0; double nrm = 0.0;
  int rc = pen_check(who, kind, descs, n, ws, ws_bytes, &lay, &soff, &nrm);
  if (rc) return rc;
  if (!out_penalty) { delete[] lay; set_error("%s: null out_penalty", who); return LQ_E_NULLPTR; }
  if (fwd) {
    if (kind == LQ_PEN_INVERSE) { delete[] lay; set_error("%s: the Inverse term reads no tensor: use lq_mt_fq_fwd + lq_penalty_fwd", who); return LQ_E_BADMODE; }
    for (int i = 0; i < n; ++i)
      if (descs[i].outer * descs[i].C * descs[i].inner != 0 && !descs[i].y) {
        delete[] lay; set_error("%s: desc %d has null y", who, i); return LQ_E_NULLPTR;
      }
  }
  char* base = static_cast<char*>(ws);
  double* total = reinterpret_cast<double*>(base + soff);
  for (int i0 = 0; i0 < n && !rc; i0 += kMaxMT) {
    const int i1 = (i0 + kMaxMT < n) ? i0 + kMaxMT : n;
    JobTable tab{}; FinTable fin{}; ScalarTable sc{}; TieTable ties{};
    unsigned groups = 0, tie_blocks = 0;
    sc.kind = kind; sc.first = (i0 == 0); sc.last = (i1 == n); sc.normalizer = nrm;
    long long target[kMaxMT];
    mt_col_targets(descs, i0, i1, target);
    for (int i = i0; i < i1; ++i) {
      const lq_tensor_desc& d = descs[i];
      const long long cnt = d.outer * d.C * d.inner;
      if (cnt == 0) continue;
      float4* part = reinterpret_cast<float4*>(base + lay[i].part_off);
      float4* gstat = reinterpret_cast<float4*>(base + lay[i].gstat_off);
      if (kind != LQ_PEN_INVERSE) {
        Geom g = make_stat_geom(d.outer, d.C, d.inner, aligned16(d.P) && (!fwd || aligned16(d.y)), target[i - i0]);
        if (g.nblocks > 0x7fffffffll) { rc = LQ_E_BADSHAPE; set_error("tensor too large"); break; }
        Job j = make_job(g, d.P, d.s, nullptr, fwd ? d.y : nullptr, nullptr, nullptr, 0, part);
        j.gaux = gstat;                                           // pen_ties_kernel: the group's {max, ties, -, lo}
        tab.job[tab.n++] = j;
        long long nb = cdiv(g.npart1, 256);
        if (nb > kTieBlocksPerTensor) nb = kTieBlocksPerTensor;
        if ((long long)tie_blocks + nb > 0x7fffffffll) { rc = LQ_E_BADSHAPE; set_error("tensor too large"); break; }
        ties.begin[ties.n++] = tie_blocks;
        tie_blocks += (unsigned)nb;
        FinJob f{};
        f.part = part; f.gstat = gstat; f.s = d.s; f.ds = nullptr;
        f.C = g.C; f.inner = g.inner; f.W = g.W; f.outer = g.outer; f.nseg = g.nseg; f.nchunks = g.nchunks;
        f.regime = g.regime; f.mode = (kind == LQ_PEN_MAXBIN) ? PFIN_MAXBIN : PFIN_SUM;
        f.P = d.P; f.part1 = part; f.regime1 = g.regime; f.seg = g.seg; f.rpc = g.rpc;
        f.nlin = g.nlin; f.lin = part + g.npart1;
        f.group_begin = groups; groups += (unsigned)d.C;
        fin.job[fin.n++] = f;
      }
      sc.job[sc.n++] = ScalarJob{gstat, d.s, d.C, (double)cnt};
    }
    if (rc) break;
    unsigned* ctl = LQ_PEN_FUSED ? reinterpret_cast<unsigned*>(base + soff + kPenCtlOff) : nullptr;
    double* term = reinterpret_cast<double*>(base + soff + kPenTermOff);
    if (tab.n) {
      if (fwd) {
        // the generic multi-tensor traversal knows nothing of the epilogue's counters: zeroed by a memset node here
        if (ctl && (rc = cuda_fail(cudaMemsetAsync(ctl, 0, sizeof(unsigned) * kCtlWords, st), "zero epilogue counters"))) break;
        if (kind == LQ_PEN_MAXBIN) {
          if (fwd->ext) { FwdPenOp<LQ_PEN_MAXBIN, true> op; op.qm = *fwd; rc = launch_mt(tab, op, st); }
          else { FwdPenOp<LQ_PEN_MAXBIN, false> op; op.qm = *fwd; rc = launch_mt(tab, op, st); }
        } else {
          if (fwd->ext) { FwdPenOp<LQ_PEN_DIFFERENCE, true> op; op.qm = *fwd; rc = launch_mt(tab, op, st); }
          else { FwdPenOp<LQ_PEN_DIFFERENCE, false> op; op.qm = *fwd; rc = launch_mt(tab, op, st); }
        }
      } else if (kind == LQ_PEN_MAXBIN) rc = launch_stat<LQ_PEN_MAXBIN>(tab, st, ctl);
      else rc = launch_stat<LQ_PEN_DIFFERENCE>(tab, st, ctl);
      if (rc) break;
      if ((rc = pen_fold_large(fin, st))) break;
      // second stage; with the fused epilogue its last blocks also produce the per-tensor terms and the penalty
      if ((rc = launch_after(pen_finalize_kernel, dim3(groups), dim3(kFinThreads), st, "penalty finalize launch", fin, sc, ctl,
                             term, total, out_penalty))) break;
      if (kind == LQ_PEN_MAXBIN && tie_blocks &&
          (rc = launch_after(pen_ties_kernel, dim3(tie_blocks), dim3(256), st, "penalty tie launch", tab, ties,
                             (const unsigned*)ctl))) break;
      if (ctl) continue;
    }
    rc = launch_after(pen_scalar_kernel, dim3(1), dim3(256), st, "penalty scalar launch", sc, total, out_penalty);
  }
  delete[] lay;
  return rc;
}

extern "C" int lq_penalty_fwd(int kind, const lq_tensor_desc* descs, int n, float* out_penalty,
                              void* ws, size_t ws_bytes, lq_stream_t stream) {
  return pen_fwd_impl("lq_penalty_fwd", kind, descs, n, out_penalty, ws, ws_bytes, (cudaStream_t)stream, nullptr);
}

extern "C" int lq_mt_fq_fwd_penalty(int kind, const lq_tensor_desc* descs, int n, int round_mode, float qmin, float qmax,
                                    float* out_penalty, void* ws, size_t ws_bytes, lq_stream_t stream) {
  int rc = 0;
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  return pen_fwd_impl("lq_mt_fq_fwd_penalty", kind, descs, n, out_penalty, ws, ws_bytes, (cudaStream_t)stream, &qm);
}

extern "C" int lq_penalty_bwd(int kind, const lq_tensor_desc* descs, int n, float upstream,
                              const float* upstream_dev, void* ws, size_t ws_bytes, lq_stream_t stream) {
  PenLayout* lay = nullptr; size_t soff = 0; double nrm = 0.0;
  int rc = pen_check("lq_penalty_bwd", kind, descs, n, ws, ws_bytes, &lay, &soff, &nrm);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = static_cast<char*>(ws);
  const float g_total = 1.0f / (float)nrm;                     // LF:116 / LF:195 / LF:275 (times *up on device)
  float* up = reinterpret_cast<float*>(base + soff + kPenUpOff);
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
    long long target[kMaxMT];
    mt_col_targets(descs, i0, i1, target);               // the forward's cut of every tensor, again
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
        Geom g = make_stat_geom(d.outer, d.C, d.inner, aligned16(d.P) && aligned16(d.dX), target[i - i0]);   // the forward's cut of the tensor
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
          f.p0 = j.t0; f.up = up;                                  // the closed-form part needs g_abs = t0 * upstream
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
      if (kind == LQ_PEN_MAXBIN) rc = accum ? launch_pen_bwd<LQ_PEN_MAXBIN, true>(tab, up, st) : launch_pen_bwd<LQ_PEN_MAXBIN, false>(tab, up, st);
      else rc = accum ? launch_pen_bwd<LQ_PEN_DIFFERENCE, true>(tab, up, st) : launch_pen_bwd<LQ_PEN_DIFFERENCE, false>(tab, up, st);
      if (rc) break;
      if (fin.n) {
        if ((rc = pen_fold_large(fin, st))) break;
        rc = launch_after(pen_finalize_kernel, dim3(groups), dim3(kFinThreads), st, "penalty ds finalize launch", fin,
                          ScalarTable{}, (unsigned*)nullptr, (double*)nullptr, (double*)nullptr, (float*)nullptr);
      }
    }
  }
  delete[] lay;
  return rc;
}
