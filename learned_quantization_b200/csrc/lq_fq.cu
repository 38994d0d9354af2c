// lq_fq.cu -- fake-quant forward (K1), threshold backward (K2), second-stage reduce (K4)
// and their multi-tensor forms (K5).  Reference semantics: NQ:49-120 (see include/lq.h).
#include <math.h>
#include "lq_fq_ops.cuh"
#include "lq_col_tma.cuh"

namespace lq {

// ------------------------------------------------------------------ element operators
// K1  NQ:55-60   y = floor(P / s) * s.  The IEEE quotient comes from div_by(): the scale's
// refined reciprocal is computed once per (thread, group), see lq_common.cuh.
template <bool EXT>
struct FwdOp : OpBase<FwdOp<EXT>> {
  QMode qm;
  using Ctx = Recip;
  static constexpr bool kIn2 = false, kOutIsY = true, kCodes = true, kReduce = false, kAux = false;
  __device__ static void init(Acc&) {}
  __device__ static void combine(Acc&, const Acc&) {}
  __device__ __forceinline__ Recip prep(float s, const float4&, const Job&) const { return make_recip(s); }
  __device__ __forceinline__ void elem(float p, float, const Recip& c, const Job&, Acc&, float& y, float& q) const {
    q = round_code<EXT>(div_by(p, c), qm);
    y = __fmul_rn(q, c.s);
  }
  __device__ __forceinline__ void quad(const float4& p, const float4&, const Recip& c, const Job&, Acc&,
                                       float4& y, float4& q, Defer&) const {
    if (c.fast && fast_range4(p)) {          // one guard for the four elements
      q.x = round_code<EXT>(div_fast(p.x, c), qm); q.y = round_code<EXT>(div_fast(p.y, c), qm);
      q.z = round_code<EXT>(div_fast(p.z, c), qm); q.w = round_code<EXT>(div_fast(p.w, c), qm);
    } else {
      q.x = round_code<EXT>(__fdiv_rn(p.x, c.s), qm); q.y = round_code<EXT>(__fdiv_rn(p.y, c.s), qm);
      q.z = round_code<EXT>(__fdiv_rn(p.z, c.s), qm); q.w = round_code<EXT>(__fdiv_rn(p.w, c.s), qm);
    }
    y.x = __fmul_rn(q.x, c.s); y.y = __fmul_rn(q.y, c.s); y.z = __fmul_rn(q.z, c.s); y.w = __fmul_rn(q.w, c.s);
  }
};

// K1 with the per-group statistics of SURVEY 8b as by-products of the same pass: min P, max P (what the callbacks'
// max |q| and the MaxBin term are functions of: floor and division by s > 0 are monotone) and sum |P - P / s| (the
// Difference term, LF:175-178, from the quotient the forward has just computed).  acc.a = sum, acc.b = bits of
// max(-P) (the minimum, negated, so that both extremes fold with fmaxf), acc.c = max P.  NaNs drop out of the extremes
// (fmaxf) and poison the sum, as in the loss terms.
template <bool EXT>
struct FwdStatsOp : OpBase<FwdStatsOp<EXT>> {
  QMode qm;
  using Ctx = Recip;
  static constexpr bool kIn2 = false, kOutIsY = true, kCodes = true, kReduce = true, kAux = false;
  __device__ static void init(Acc& a) { a.a = 0.f; a.b = __float_as_uint(-INFINITY); a.c = -INFINITY; }
  __device__ static void combine(Acc& x, const Acc& y) {
    x.a += y.a;
    x.b = __float_as_uint(fmaxf(__uint_as_float(x.b), __uint_as_float(y.b)));
    x.c = fmaxf(x.c, y.c);
  }
  __device__ __forceinline__ Recip prep(float s, const float4&, const Job&) const { return make_recip(s); }
  __device__ __forceinline__ void elem(float p, float, const Recip& c, const Job&, Acc& acc, float& y, float& q) const {
    const float t = div_by(p, c);
    q = round_code<EXT>(t, qm);
    y = __fmul_rn(q, c.s);
    acc.a += fabsf(__fsub_rn(p, t));
    acc.b = __float_as_uint(fmaxf(__uint_as_float(acc.b), -p));
    acc.c = fmaxf(acc.c, p);
  }
};

// Second stage of the forward statistics: one block per group, fixed order, fp64 sum.  stats[3g .. 3g+2] =
// {min P, max P, sum |P - P/s|} (+inf, -inf, 0 for a group without ordinary elements).
__global__ void __launch_bounds__(256) fwd_stats_finalize_kernel(const __grid_constant__ FinJob f) {
  cudaGridDependencySynchronize();
  const long long g = blockIdx.x;
  const long long K = partials_per_group(f);
  double sum = 0.0;
  float nmin = -INFINITY, mx = -INFINITY;
  for (long long k0 = threadIdx.x; k0 < K; k0 += 4 * 256) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = k0 + u * 256;
      v[u] = (k < K) ? f.part[partial_index(f, g, k)] : make_float4(0.f, -INFINITY, -INFINITY, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { sum += (double)v[u].x; nmin = fmaxf(nmin, v[u].y); mx = fmaxf(mx, v[u].z); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    nmin = fmaxf(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double ssum[8];
  __shared__ float smin[8], smx[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; smin[w] = nmin; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) { sum += ssum[i]; nmin = fmaxf(nmin, smin[i]); mx = fmaxf(mx, smx[i]); }
    f.ds[3 * g] = -nmin; f.ds[3 * g + 1] = mx; f.ds[3 * g + 2] = (float)sum;
  }
}

// ------------------------------------------------------------------ K4 second stage
// One block per group.  Walks the group's partials in index order with a fixed
// thread->partial assignment, then a fixed shuffle/shared-memory tree: deterministic.
enum { FIN_THRESHOLD = 0, FIN_STATS = 1 };   // STATS: write raw {sum, count, max|q|} as doubles (sharded tensors)

// Intermediate level for groups with very many first-level partials (per-tensor scales on
// large tensors: one group, N/2048 partials): block (g, m) folds partials [m*kFinChunk,
// (m+1)*kFinChunk) of group g into lin[g*nlin + m] in a fixed order; the last level then
// reads REG_LIN partials.  Same {sum, count, max} format, fp64 accumulation inside the block.
__global__ void __launch_bounds__(256) fold_partials_kernel(const __grid_constant__ FinJob f) {
  cudaGridDependencySynchronize();
  const long long g = blockIdx.x / f.nlin, m = blockIdx.x - g * f.nlin;
  const long long K = partials_per_group(f);
  const long long k0 = m * kFinChunk;
  double sum = 0.0;
  unsigned cnt = 0u;
  float mx = 0.f;
  float4 v[kFinChunk / 256];
#pragma unroll
  for (int u = 0; u < kFinChunk / 256; ++u) {
    const long long k = k0 + u * 256 + threadIdx.x;
    v[u] = (k < K) ? f.part[partial_index(f, g, k)] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < kFinChunk / 256; ++u) {
    sum += (double)v[u].x; cnt += __float_as_uint(v[u].y); mx = fmaxf(mx, v[u].z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double ssum[8];
  __shared__ unsigned scnt[8];
  __shared__ float smx[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; scnt[w] = cnt; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; cnt = scnt[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { sum += ssum[i]; cnt += scnt[i]; mx = fmaxf(mx, smx[i]); }
    f.lin[blockIdx.x] = make_float4((float)sum, __uint_as_float(cnt), mx, 0.f);
  }
}

__device__ __forceinline__ void finalize_group(const FinJob& f, long long g) {
  const long long K = partials_per_group(f);
  double sum = 0.0;
  unsigned long long cnt = 0ull;
  float mx = 0.f;
  for (long long k = threadIdx.x; k < K; k += kFinThreads) {
    const float4 v = f.part[partial_index(f, g, k)];
    sum += (double)v.x;
    cnt += (unsigned long long)__float_as_uint(v.y);
    mx = fmaxf(mx, v.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double ssum[kFinThreads / 32];
  __shared__ unsigned long long scnt[kFinThreads / 32];
  __shared__ float smx[kFinThreads / 32];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { ssum[w] = sum; scnt[w] = cnt; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; cnt = scnt[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < kFinThreads / 32; ++i) { sum += ssum[i]; cnt += scnt[i]; mx = fmaxf(mx, smx[i]); }
    // NQ:103-116: all above -> every element contributes c0 = -|tanh(lambda)| (mean = c0),
    // otherwise the mean of the below-threshold terms over ALL N_g elements; times max|q|.
    if (f.mode == FIN_STATS) {
      double* st = reinterpret_cast<double*>(f.ds) + 3 * g;
      st[0] = sum; st[1] = (double)cnt; st[2] = (double)mx;
      return;
    }
    const double n_g = (double)f.outer * (double)f.inner;
    const float mean = (cnt == 0ull) ? f.p1 : (float)(sum / n_g);
    f.ds[g] = __fmul_rn(mean, mx);
  }
}

// One WARP per group: the whole group's partials are in flight at once (4 independent 16-byte
// loads per lane per trip, no shared memory, no block barrier) -- the block-per-group version above
// spent most of its 5 us on two dependent load rounds and a 64-bit division per partial.
// Lane l folds partials l, l+32, ... in that order, then a fixed shuffle tree: deterministic.
__device__ __forceinline__ void finish_group_value(const FinJob& f, long long g, double sum, unsigned long long cnt, float mx) {
  if (f.mode == FIN_STATS) {
    double* st = reinterpret_cast<double*>(f.ds) + 3 * g;
    st[0] = sum; st[1] = (double)cnt; st[2] = (double)mx;
    return;
  }
  const double n_g = (double)f.outer * (double)f.inner;
  const float mean = (cnt == 0ull) ? f.p1 : (float)(sum / n_g);
  f.ds[g] = __fmul_rn(mean, mx);
}

// WPG = warps per group: 1 (eight groups per block) when a group has at most 256 partials, 8 (one
// block per group) above that -- a per-tensor scale on a 4 Mi-element tensor has 2048 partials in its
// single group, which one warp would walk in 16 dependent trips.
template <int WPG>
__global__ void __launch_bounds__(256) finalize_warp_kernel(const __grid_constant__ FinJob f) {
  cudaGridDependencySynchronize();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g = (WPG == 1) ? (long long)blockIdx.x * 8 + warp : (long long)blockIdx.x;
  if (g >= f.C) return;
  const int sub = (WPG == 1) ? 0 : warp;                   // which slice of the group's partials this warp folds
  const long long K = partials_per_group(f);
  double sum = 0.0;
  unsigned long long cnt = 0ull;
  float mx = 0.f;
  const bool simple = (f.regime == REG_ROW && f.nseg == 1) || f.regime == REG_LIN || (f.regime == REG_COL && f.inner == 1);
  // simple layouts: partial k of group g sits at base + k * stride
  const long long stride = (f.regime == REG_ROW) ? f.C : (f.regime == REG_COL ? f.W : 1);
  const float4* base = f.part + ((f.regime == REG_LIN) ? g * f.nlin : g);
  for (long long k0 = sub * 32 + lane; k0 < K; k0 += 128ll * WPG) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = k0 + 32ll * WPG * u;
      v[u] = (k < K) ? (simple ? base[k * stride] : f.part[partial_index(f, g, k)]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sum += (double)v[u].x; cnt += (unsigned long long)__float_as_uint(v[u].y); mx = fmaxf(mx, v[u].z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (WPG == 1) {
    if (lane == 0) finish_group_value(f, g, sum, cnt, mx);
    return;
  }
  __shared__ double ssum[8];
  __shared__ unsigned long long scnt[8];
  __shared__ float smx[8];
  if (lane == 0) { ssum[warp] = sum; scnt[warp] = cnt; smx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sum = ssum[0]; cnt = scnt[0]; mx = smx[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { sum += ssum[i]; cnt += scnt[i]; mx = fmaxf(mx, smx[i]); }
    finish_group_value(f, g, sum, cnt, mx);
  }
}

__global__ void __launch_bounds__(kFinThreads) mt_finalize_kernel(const __grid_constant__ FinTable tab) {
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].group_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const FinJob& f = tab.job[lo];
  cudaGridDependencySynchronize();
  finalize_group(f, blockIdx.x - f.group_begin);
}

// ------------------------------------------------------------------ host side
Job make_job(const Geom& g, const float* P, const float* s, const float* dy, float* y, float* dX,
             void* codes, int codes_dtype, float4* part) {
  Job j{};
  j.P = P; j.s = s; j.dy = dy; j.y = y; j.dX = dX; j.codes = codes; j.part = part;
  j.outer = g.outer; j.C = g.C; j.inner = g.inner; j.W = g.W; j.units = g.units;
  j.nchunks = g.nchunks; j.ncol = g.ncol;
  j.regime = g.regime; j.vec = g.vec_ok; j.seg = g.seg; j.nseg = g.nseg;
  j.bx = g.bx; j.by = g.by; j.rpc = g.rpc; j.codes_dtype = codes_dtype;
  j.stream_hint = (g.n * 4 > kStreamBytes) ? 1 : 0;
  j.n = g.n;
  j.write_dx = (dX != nullptr && dX != dy) ? 1 : 0;
  j.block_begin = 0; j.nblocks = (unsigned)g.nblocks;
  j.fast32 = (g.regime == REG_ROW && g.units < (1ll << 31) && g.C < (1ll << 31)) ? 1 : 0;
  if (j.fast32) { j.fd_nseg = make_fastdiv((unsigned long long)g.nseg); j.fd_C = make_fastdiv((unsigned long long)g.C); }
  return j;
}

FinJob make_fin(const Geom& g, const float4* part, float* ds, const float* s, float lam) {
  FinJob f{};
  f.part = part; f.ds = ds; f.s = s;
  f.C = g.C; f.inner = g.inner; f.W = g.W; f.outer = g.outer; f.nseg = g.nseg; f.nchunks = g.nchunks;
  f.nlin = g.nlin; f.lin = const_cast<float4*>(part) + g.npart1;
  f.regime = g.regime; f.mode = FIN_THRESHOLD;
  f.p0 = lam;
  f.p1 = -fabsf(tanhf(lam));   // c0 of NQ:79 / NQ:105
  return f;
}

// The forward writes no partials, so its COL row chunks need not be the backward's: short-lived blocks (a few trips
// per thread) leave a finer tail at the end of the grid.  Only the block count changes; vector forms only.
// min_rpc: a multi-tensor launch's share-based chunk (mt_col_rpc below), which wins over the short chunks.
Geom fwd_geom(long long outer, long long C, long long inner, bool al, long long min_rpc = 0) {
  Geom g = make_geom(outer, C, inner, al);
  if (g.regime == REG_COL && g.vec_ok) {
    long long rpc = g.rpc;
    if (LQ_COL_FWD_TRIPS > 0 && (long long)g.by * LQ_COL_U1 * LQ_COL_FWD_TRIPS < rpc) rpc = (long long)g.by * LQ_COL_U1 * LQ_COL_FWD_TRIPS;
    if (min_rpc > rpc) rpc = min_rpc;
    if (rpc > g.outer) rpc = g.outer;
    if (rpc != g.rpc) {
      g.rpc = (int)rpc;
      g.nchunks = cdiv(g.outer, rpc);
      g.nblocks = g.nchunks * g.ncolblk;
    }
  }
  return g;
}

// Internal launch helpers shared with lq_host.cu.
int fq_fwd_launch(const Geom& g, const float* P, const float* s, float* y, void* codes, int codes_dtype,
                  const QMode& qm, cudaStream_t st) {
  Job j = make_job(g, P, s, nullptr, y, nullptr, codes, codes_dtype, nullptr);
  if (qm.ext) { FwdOp<true> op; op.qm = qm; return launch_job(j, op, st); }
  FwdOp<false> op; op.qm = qm;
  return launch_job(j, op, st);
}

// Geometry of the threshold backward for one tensor: TMA-staged COL kernel when the shape and the
// pointers allow it, otherwise ROW / register COL.  `tma` tells fq_bwd_elements_launch which.
Geom bwd_geom(long long outer, long long C, long long inner, const float* dy, const float* P, const float* dX, bool* tma) {
  const bool wr = (dX != nullptr && dX != dy);
  const bool al = aligned16(P) && aligned16(dy) && (!wr || aligned16(dX));
  *tma = al && col_tma_shape_ok(outer, C, inner);
  if (*tma) return make_col_tma_geom(outer, C, inner);
  // register COL: the backward takes the one-column-per-thread traversal (run_col1)
  return make_geom(outer, C, inner, al && inner >= kRowMinInner);
}

int fq_bwd_col_tma_launch(const Geom& g, const float* dy, const float* P, const float* s, float lam,
                          float* dX, float4* part, const QMode& qm, cudaStream_t st) {
  if (qm.ext) return col_tma_launch(g, dy, P, s, dX, part, make_bwd_op<true>(qm, lam), st);
  return col_tma_launch(g, dy, P, s, dX, part, make_bwd_op<false>(qm, lam), st);
}

int fq_bwd_elements_launch(const Geom& g, const float* dy, const float* P, const float* s, float lam,
                           float* dX, float4* part, const QMode& qm, cudaStream_t st) {
  Job j = make_job(g, P, s, dy, nullptr, dX, nullptr, 0, part);
  if (qm.ext) return launch_job(j, make_bwd_op<true>(qm, lam), st);
  return launch_job(j, make_bwd_op<false>(qm, lam), st);
}

// Combine the raw statistics of `nshards` shards of one tensor (fixed shard order: every rank
// that holds the same gathered buffer computes bit-identical ds) and apply NQ:103-116.
__global__ void __launch_bounds__(128) finish_shards_kernel(const double* stats, int nshards, long long C, float c0,
                                                            double n_per_group, float* ds) {
  cudaGridDependencySynchronize();     // launched with launch_after: wait for the gather before it in the stream
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= C) return;
  double sum = 0.0, cnt = 0.0, mx = 0.0;
  for (int r = 0; r < nshards; ++r) {
    const double* st = stats + ((long long)r * C + g) * 3;
    sum += st[0]; cnt += st[1]; mx = fmax(mx, st[2]);
  }
  const float mean = (cnt == 0.0) ? c0 : (float)(sum / n_per_group);
  ds[g] = __fmul_rn(mean, (float)mx);
}

int fq_finalize_launch(const Geom& g, const float4* part, float* ds, const float* s, float lam, cudaStream_t st, int mode) {
  FinJob f = make_fin(g, part, ds, s, lam);
  f.mode = mode;
  if (g.nlin) {   // two-level second stage
    if (g.C * g.nlin > 0x7fffffffll) { set_error("too many intermediate partials"); return LQ_E_BADSHAPE; }
    int rc = launch_after(fold_partials_kernel, dim3((unsigned)(g.C * g.nlin)), dim3(256), st, "fold launch", f);
    if (rc) return rc;
    f.part = f.lin; f.regime = REG_LIN;
  }
  const long long K = f.regime == REG_LIN ? f.nlin : g.kper;
  if (K > 256)
    return launch_after(finalize_warp_kernel<8>, dim3((unsigned)g.C), dim3(256), st, "finalize launch", f);
  return launch_after(finalize_warp_kernel<1>, dim3((unsigned)((g.C + 7) / 8)), dim3(256), st, "finalize launch", f);
}

// Second stage of a multi-tensor backward: one block per group over all tensors of `fin`.  Groups with more
// than kFinSplit first-level partials (a per-tensor scale on a large tensor) get the intermediate fold first,
// exactly as in the single-tensor path, so that no block ever walks more than kFinSplit partials.
int mt_finalize_launch(FinTable& fin, unsigned groups, cudaStream_t st) {
  if (fin.n == 0 || groups == 0) return LQ_OK;
  for (int i = 0; i < fin.n; ++i) {
    FinJob& f = fin.job[i];
    if (f.nlin <= 0 || f.regime == REG_LIN) continue;
    if (f.C * f.nlin > 0x7fffffffll) { set_error("too many intermediate partials"); return LQ_E_BADSHAPE; }
    int rc = launch_after(fold_partials_kernel, dim3((unsigned)(f.C * f.nlin)), dim3(256), st, "mt fold launch", f);
    if (rc) return rc;
    f.part = f.lin; f.regime = REG_LIN;
  }
  return launch_after(mt_finalize_kernel, dim3(groups), dim3(kFinThreads), st, "mt finalize launch", fin);
}

// CL:61-62 / PI:62-64 for a whole table in ONE launch: ds = 0 for every tensor and, where a separate dX was
// asked for, dX = dy.  Block -> tensor by binary search like the element kernels.
struct SteJob { const float* dy; float* dX; float* ds; long long n, C; unsigned block_begin; };
struct SteTable { int n; SteJob job[kMaxMT]; };
constexpr long long kSteBlockElems = 256 * 4 * 8;   // elements one block copies

__global__ void __launch_bounds__(256) mt_ste_kernel(const __grid_constant__ SteTable tab) {
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const SteJob& j = tab.job[lo];
  const long long blk = blockIdx.x - j.block_begin;
  cudaGridDependencySynchronize();
  const long long e0 = blk * kSteBlockElems;
  for (long long g = e0 + threadIdx.x; g < j.C && g < e0 + kSteBlockElems; g += 256) j.ds[g] = 0.0f;
  if (j.dX == nullptr) return;
  const long long e1 = (e0 + kSteBlockElems < j.n) ? e0 + kSteBlockElems : j.n;
  if (aligned16_dev(j.dy) && aligned16_dev(j.dX)) {
    const float4* s4 = reinterpret_cast<const float4*>(j.dy + e0);
    float4* d4 = reinterpret_cast<float4*>(j.dX + e0);
    const long long nv = (e1 - e0) >> 2;
    for (long long v = threadIdx.x; v < nv; v += 256) d4[v] = s4[v];
    for (long long e = e0 + (nv << 2) + threadIdx.x; e < e1; e += 256) j.dX[e] = j.dy[e];
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += 256) j.dX[e] = j.dy[e];
  }
}

int check_blocks(const Geom& g) {
  if (g.nblocks > 0x7fffffffll) { set_error("tensor too large for one grid (%lld blocks)", g.nblocks); return LQ_E_BADSHAPE; }
  return LQ_OK;
}

}  // namespace lq

using namespace lq;

extern "C" int lq_fq_fwd(const float* P, const float* s, float* y, void* codes, int codes_dtype,
                         int64_t outer, int64_t C, int64_t inner, int round_mode, float qmin, float qmax,
                         lq_stream_t stream) {
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  if (outer * C * inner == 0) return LQ_OK;
  if (!P || !s || !y) { set_error("lq_fq_fwd: null P/s/y"); return LQ_E_NULLPTR; }
  if (codes_dtype < LQ_CODES_NONE || codes_dtype > LQ_CODES_I8 || (codes && codes_dtype == LQ_CODES_NONE)) {
    set_error("lq_fq_fwd: bad codes_dtype %d", codes_dtype); return LQ_E_BADMODE;
  }
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  const bool al = aligned16(P) && aligned16(y) && (!codes || codes_dtype == LQ_CODES_I8 || aligned16(codes));
  Geom g = fwd_geom(outer, C, inner, al);
  if ((rc = check_blocks(g))) return rc;
  return fq_fwd_launch(g, P, s, y, codes_dtype == LQ_CODES_NONE ? nullptr : codes, codes_dtype, qm, (cudaStream_t)stream);
}

extern "C" int lq_fq_fwd_stats(const float* P, const float* s, float* y, void* codes, int codes_dtype,
                               int64_t outer, int64_t C, int64_t inner, int round_mode, float qmin, float qmax,
                               float* stats, void* ws, size_t ws_bytes, lq_stream_t stream) {
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  if (!stats) { set_error("lq_fq_fwd_stats: null stats (use lq_fq_fwd)"); return LQ_E_NULLPTR; }
  cudaStream_t st = (cudaStream_t)stream;
  if (outer * C * inner == 0) {                    // empty groups: +inf, -inf, 0
    if (C > 0) {
      set_error("lq_fq_fwd_stats: empty tensor with C > 0 groups has no statistics"); return LQ_E_BADSHAPE;
    }
    return LQ_OK;
  }
  if (!P || !s || !y) { set_error("lq_fq_fwd_stats: null P/s/y"); return LQ_E_NULLPTR; }
  if (codes_dtype < LQ_CODES_NONE || codes_dtype > LQ_CODES_I8 || (codes && codes_dtype == LQ_CODES_NONE)) {
    set_error("lq_fq_fwd_stats: bad codes_dtype %d", codes_dtype); return LQ_E_BADMODE;
  }
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  const bool al = aligned16(P) && aligned16(y) && (!codes || codes_dtype == LQ_CODES_I8 || aligned16(codes));
  Geom g = make_geom(outer, C, inner, al);
  if ((rc = check_blocks(g))) return rc;
  const size_t need = align256((size_t)g.npart1 * sizeof(float4));
  if (!ws || ws_bytes < need || !aligned16(ws)) {
    set_error("lq_fq_fwd_stats: workspace %zu B (16-B aligned) needed, got %zu", need, ws_bytes);
    return LQ_E_WORKSPACE;
  }
  Job j = make_job(g, P, s, nullptr, y, nullptr, codes_dtype == LQ_CODES_NONE ? nullptr : codes, codes_dtype,
                   reinterpret_cast<float4*>(ws));
  if (qm.ext) { FwdStatsOp<true> op; op.qm = qm; rc = launch_job(j, op, st); }
  else { FwdStatsOp<false> op; op.qm = qm; rc = launch_job(j, op, st); }
  if (rc) return rc;
  FinJob f = make_fin(g, reinterpret_cast<const float4*>(ws), stats, s, 0.f);
  f.nlin = 0;
  return launch_after(fwd_stats_finalize_kernel, dim3((unsigned)g.C), dim3(256), st, "forward statistics finalize launch", f);
}

extern "C" int lq_fq_bwd_threshold(const float* dy, const float* P, const float* s, float penalty_threshold,
                                   float* dX, float* ds, int64_t outer, int64_t C, int64_t inner,
                                   int round_mode, float qmin, float qmax,
                                   void* ws, size_t ws_bytes, lq_stream_t stream) {
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  if (!dy || !P || !s || !ds) { set_error("lq_fq_bwd_threshold: null dy/P/s/ds"); return LQ_E_NULLPTR; }
  if (outer * C * inner == 0) return LQ_OK;
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  bool tma = false;
  Geom g = bwd_geom(outer, C, inner, dy, P, dX, &tma);
  if ((rc = check_blocks(g))) return rc;
  const size_t need = align256((size_t)g.npartials * sizeof(float4));
  if (!ws || ws_bytes < need || !aligned16(ws)) {
    set_error("lq_fq_bwd_threshold: workspace %zu B (16-B aligned) needed, got %zu", need, ws_bytes);
    return LQ_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = (tma ? fq_bwd_col_tma_launch : fq_bwd_elements_launch)(g, dy, P, s, penalty_threshold, dX, reinterpret_cast<float4*>(ws), qm, st))) return rc;
  return fq_finalize_launch(g, reinterpret_cast<const float4*>(ws), ds, s, penalty_threshold, st, FIN_THRESHOLD);
}

extern "C" int lq_fq_bwd_threshold_stats(const float* dy, const float* P, const float* s, float penalty_threshold,
                                         float* dX, double* stats, int64_t outer, int64_t C, int64_t inner,
                                         int round_mode, float qmin, float qmax,
                                         void* ws, size_t ws_bytes, lq_stream_t stream) {
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  if (!dy || !P || !s || !stats) { set_error("lq_fq_bwd_threshold_stats: null dy/P/s/stats"); return LQ_E_NULLPTR; }
  if (reinterpret_cast<uintptr_t>(stats) & 7u) { set_error("lq_fq_bwd_threshold_stats: stats must be 8-byte aligned"); return LQ_E_BADSHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (outer * C * inner == 0) {
    if (C > 0) return cuda_fail(cudaMemsetAsync(stats, 0, (size_t)C * 3 * sizeof(double), st), "zero stats");
    return LQ_OK;
  }
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  bool tma = false;
  Geom g = bwd_geom(outer, C, inner, dy, P, dX, &tma);
  if ((rc = check_blocks(g))) return rc;
  const size_t need = align256((size_t)g.npartials * sizeof(float4));
  if (!ws || ws_bytes < need || !aligned16(ws)) {
    set_error("lq_fq_bwd_threshold_stats: workspace %zu B (16-B aligned) needed, got %zu", need, ws_bytes);
    return LQ_E_WORKSPACE;
  }
  if ((rc = (tma ? fq_bwd_col_tma_launch : fq_bwd_elements_launch)(g, dy, P, s, penalty_threshold, dX, reinterpret_cast<float4*>(ws), qm, st))) return rc;
  return fq_finalize_launch(g, reinterpret_cast<const float4*>(ws), reinterpret_cast<float*>(stats), s, penalty_threshold, st, FIN_STATS);
}

extern "C" int lq_fq_bwd_threshold_finish(const double* stats, int nshards, int64_t C, float penalty_threshold,
                                          double n_per_group, float* ds, lq_stream_t stream) {
  if (nshards < 1 || C < 0 || !(n_per_group > 0.0)) { set_error("lq_fq_bwd_threshold_finish: bad nshards / C / n_per_group"); return LQ_E_BADSHAPE; }
  if (C == 0) return LQ_OK;
  if (!stats || !ds) { set_error("lq_fq_bwd_threshold_finish: null stats/ds"); return LQ_E_NULLPTR; }
  const float c0 = -fabsf(tanhf(penalty_threshold));
  return launch_after(finish_shards_kernel, dim3((unsigned)((C + 127) / 128)), dim3(128), (cudaStream_t)stream,
                      "finish launch", stats, nshards, (long long)C, c0, n_per_group, ds);
}

extern "C" int lq_fq_bwd_ste(const float* dy, float* dX, float* ds, int64_t n_elems, int64_t C,
                             lq_stream_t stream) {
  if (n_elems < 0 || C < 0) { set_error("lq_fq_bwd_ste: negative size"); return LQ_E_BADSHAPE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (dX && dX != dy && n_elems > 0) {
    if (!dy) { set_error("lq_fq_bwd_ste: null dy"); return LQ_E_NULLPTR; }
    int rc = cuda_fail(cudaMemcpyAsync(dX, dy, (size_t)n_elems * 4, cudaMemcpyDeviceToDevice, st), "dX = dy");
    if (rc) return rc;
  }
  if (ds && C > 0) return cuda_fail(cudaMemsetAsync(ds, 0, (size_t)C * 4, st), "ds = 0");   // CL:62
  return LQ_OK;
}

// ------------------------------------------------------------------ multi-tensor forms
extern "C" int lq_mt_fq_fwd(const lq_tensor_desc* descs, int n, int round_mode, float qmin, float qmax,
                            lq_stream_t stream) {
  if (n < 0) { set_error("lq_mt_fq_fwd: n < 0"); return LQ_E_BADSHAPE; }
  if (n && !descs) { set_error("lq_mt_fq_fwd: null descs"); return LQ_E_NULLPTR; }
  int rc = 0;
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n; i0 += kMaxMT) {
    JobTable tab{};
    const int i1 = (i0 + kMaxMT < n) ? i0 + kMaxMT : n;
    for (int i = i0; i < i1; ++i)
      if ((rc = check_dims(descs[i].outer, descs[i].C, descs[i].inner))) return rc;
    long long share_rpc[kMaxMT];
    mt_col_rpc(descs, i0, i1, LQ_COL_U1, true, share_rpc);
    for (int i = i0; i < i1; ++i) {
      const lq_tensor_desc& d = descs[i];
      if (d.outer * d.C * d.inner == 0) continue;
      if (!d.P || !d.s || !d.y) { set_error("lq_mt_fq_fwd: desc %d has null P/s/y", i); return LQ_E_NULLPTR; }
      Geom g = fwd_geom(d.outer, d.C, d.inner, aligned16(d.P) && aligned16(d.y), share_rpc[i - i0]);
      if ((rc = check_blocks(g))) return rc;
      tab.job[tab.n++] = make_job(g, d.P, d.s, nullptr, d.y, nullptr, nullptr, 0, nullptr);
    }
    if (qm.ext) { FwdOp<true> op; op.qm = qm; rc = launch_mt(tab, op, st); }
    else { FwdOp<false> op; op.qm = qm; rc = launch_mt(tab, op, st); }
    if (rc) return rc;
  }
  return LQ_OK;
}

extern "C" int lq_mt_fq_bwd(const lq_tensor_desc* descs, int n, float penalty_threshold, int scale_grad_mode,
                            int round_mode, float qmin, float qmax,
                            void* ws, size_t ws_bytes, lq_stream_t stream) {
  if (n < 0) { set_error("lq_mt_fq_bwd: n < 0"); return LQ_E_BADSHAPE; }
  if (n && !descs) { set_error("lq_mt_fq_bwd: null descs"); return LQ_E_NULLPTR; }
  if (scale_grad_mode != LQ_SGRAD_NONE && scale_grad_mode != LQ_SGRAD_THRESHOLD) {
    set_error("lq_mt_fq_bwd: bad scale_grad_mode %d", scale_grad_mode); return LQ_E_BADMODE;
  }
  int rc = 0;
  QMode qm = make_qmode(round_mode, qmin, qmax, &rc);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (scale_grad_mode == LQ_SGRAD_NONE) {     // CL:61-62 / PI:62-64 for every tensor: one launch per 48 tensors
    for (int i0 = 0; i0 < n; i0 += kMaxMT) {
      SteTable tab{};
      unsigned total = 0;
      for (int i = i0; i < n && i < i0 + kMaxMT; ++i) {
        const lq_tensor_desc& d = descs[i];
        if ((rc = check_dims(d.outer, d.C, d.inner))) return rc;
        const long long cnt = d.outer * d.C * d.inner;
        const bool copy = d.dX != nullptr && d.dX != d.dy && cnt > 0;
        if (copy && !d.dy) { set_error("lq_mt_fq_bwd: desc %d has null dy", i); return LQ_E_NULLPTR; }
        if (!copy && !(d.ds && d.C > 0)) continue;
        const long long work = copy ? (cnt > d.C ? cnt : d.C) : d.C;
        const long long nb = cdiv(work, kSteBlockElems);
        if (nb + total > 0x7fffffffll) { set_error("lq_mt_fq_bwd: too many blocks"); return LQ_E_BADSHAPE; }
        tab.job[tab.n++] = SteJob{d.dy, copy ? d.dX : nullptr, d.ds, cnt, d.ds ? d.C : 0, total};
        total += (unsigned)nb;
      }
      if (total && (rc = launch_after(mt_ste_kernel, dim3(total), dim3(256), st, "mt STE launch", tab))) return rc;
    }
    return LQ_OK;
  }
  if (ws && !aligned16(ws)) { set_error("lq_mt_fq_bwd: workspace must be 16-byte aligned"); return LQ_E_WORKSPACE; }
  size_t off = 0;
  for (int i0 = 0; i0 < n; i0 += kMaxMT) {
    JobTable tab{};
    FinTable fin{};
    unsigned groups = 0;
    const int i1 = (i0 + kMaxMT < n) ? i0 + kMaxMT : n;
    for (int i = i0; i < i1; ++i)
      if ((rc = check_dims(descs[i].outer, descs[i].C, descs[i].inner))) return rc;
    long long share_rpc[kMaxMT];
    mt_col_rpc(descs, i0, i1, 8, false, share_rpc);            // run_col1: one column per thread, eight rows per trip
    for (int i = i0; i < i1; ++i) {
      const lq_tensor_desc& d = descs[i];
      if (d.outer * d.C * d.inner == 0) continue;
      if (!d.P || !d.s || !d.dy || !d.ds) { set_error("lq_mt_fq_bwd: desc %d has null P/s/dy/ds", i); return LQ_E_NULLPTR; }
      const bool wr = (d.dX != nullptr && d.dX != d.dy);
      Geom g = make_geom(d.outer, d.C, d.inner, aligned16(d.P) && aligned16(d.dy) && (!wr || aligned16(d.dX)) &&
                                                d.inner >= kRowMinInner, (int)share_rpc[i - i0]);
      if ((rc = check_blocks(g))) return rc;
      const size_t need = align256((size_t)g.npartials * sizeof(float4));
      if (!ws || off + need > ws_bytes) {
        set_error("lq_mt_fq_bwd: workspace too small (need > %zu B, got %zu)", off + need, ws_bytes);
        return LQ_E_WORKSPACE;
      }
      float4* part = reinterpret_cast<float4*>(static_cast<char*>(ws) + off);
      off += need;
      tab.job[tab.n++] = make_job(g, d.P, d.s, d.dy, nullptr, d.dX, nullptr, 0, part);
      FinJob f = make_fin(g, part, d.ds, d.s, penalty_threshold);
      f.group_begin = groups;
      groups += (unsigned)d.C;
      fin.job[fin.n++] = f;
    }
    if (tab.n == 0) continue;
    if (qm.ext) rc = launch_mt(tab, make_bwd_op<true>(qm, penalty_threshold), st);
    else rc = launch_mt(tab, make_bwd_op<false>(qm, penalty_threshold), st);
    if (rc) return rc;
    if ((rc = mt_finalize_launch(fin, groups, st))) return rc;
  }
  return LQ_OK;
}
