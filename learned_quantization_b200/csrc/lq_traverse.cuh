// lq_traverse.cuh -- the two traversals (ROW / COL, see lq_common.cuh) as templates over an
// element operator.  An operator provides
//   struct Acc { float a; unsigned b; float c; }  semantics via  init / combine
//   static constexpr bool kIn2      second input array (Job::dy) is read
//   static constexpr bool kOutIsY   out0 goes to Job::y (else to Job::dX when Job::write_dx)
//   static constexpr bool kCodes    out1 is the integer code, written to Job::codes if set
//   static constexpr bool kReduce   per-unit partials are produced
//   static constexpr bool kAux      per-group float4 Job::gaux[g] is fetched next to the scale
//   Ctx prep(s, aux, job)            per-(thread, group) constants, hoisted out of the element loop
//   elem(p, in2, ctx, job, acc, out0, out1)
//   quad(p4, in2_4, ctx, job, acc, out0_4, out1_4, defer)   four consecutive elements of one
//                                    group (ROW, 128-bit path); OpBase supplies it as 4 x elem
//   static constexpr int kDeferSlots  per-thread queue entries (float2) in shared memory that quad()
//                                    may push rare-case work to; flush(acc, defer) drains them once per
//                                    16 elements with the warp converged (0 = unused)
#pragma once
#include "lq_common.cuh"

namespace lq {

struct Acc {
  float a; unsigned b; float c;
};

__device__ __forceinline__ float4 pack_acc(const Acc& x) {
  return make_float4(x.a, __uint_as_float(x.b), x.c, 0.0f);
}
__device__ __forceinline__ Acc unpack_acc(const float4& v) {
  Acc x; x.a = v.x; x.b = __float_as_uint(v.y); x.c = v.z; return x;
}

template <class Op>
__device__ __forceinline__ Acc warp_reduce_acc(Acc x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Acc y;
    y.a = __shfl_xor_sync(0xffffffffu, x.a, o);
    y.b = __shfl_xor_sync(0xffffffffu, x.b, o);
    y.c = __shfl_xor_sync(0xffffffffu, x.c, o);
    Op::combine(x, y);
  }
  return x;
}

// Default per-group context (scale + optional aux operands) and default quad.
struct PlainCtx { float s; float4 aux; };
// A thread's private rare-case queue: entry k lives at slot[k * 32] (lane-interleaved inside the
// warp's region, so a warp-wide push or pop of the same k is bank-conflict free).  Only the
// owning thread ever touches its entries: no synchronisation is involved.
struct Defer { float2* slot; int n; };
template <class D>
struct OpBase {
  using Ctx = PlainCtx;
  static constexpr int kDeferSlots = 0;
  __device__ __forceinline__ PlainCtx prep(float s, const float4& aux, const Job&) const { return PlainCtx{s, aux}; }
  __device__ __forceinline__ void flush(Acc&, Defer&) const {}
  // warp-level reduction of a unit's accumulator (fixed tree: deterministic); operators whose fields
  // are plain sum / integer sum / max override it with single-instruction integer reductions
  __device__ __forceinline__ Acc warp_reduce(const Acc& x) const { return warp_reduce_acc<D>(x); }
  // COL traversal: element of the thread's k-th column (its own group context and accumulator)
  template <class C>
  __device__ __forceinline__ void elem_k(float p, float d, const C& c, const Job& j, Acc* acc, int k,
                                         float& o, float& q, Defer&) const {
    static_cast<const D*>(this)->elem(p, d, c, j, acc[k], o, q);
  }
  __device__ __forceinline__ void flush_k(Acc*, Defer&) const {}
  template <class C>
  __device__ __forceinline__ void quad(const float4& p, const float4& d, const C& c, const Job& j, Acc& acc,
                                       float4& o, float4& q, Defer&) const {
    const D& self = *static_cast<const D*>(this);
    self.elem(p.x, d.x, c, j, acc, o.x, q.x);
    self.elem(p.y, d.y, c, j, acc, o.y, q.y);
    self.elem(p.z, d.z, c, j, acc, o.z, q.z);
    self.elem(p.w, d.w, c, j, acc, o.w, q.w);
  }
};


__device__ __forceinline__ void store_code(void* codes, int dtype, long long idx, float q) {
  if (dtype == LQ_CODES_F32) reinterpret_cast<float*>(codes)[idx] = q;
  else if (dtype == LQ_CODES_I32) reinterpret_cast<int*>(codes)[idx] = (int)q;
  else reinterpret_cast<signed char*>(codes)[idx] = (signed char)(int)q;  // wraps like .astype(int8)
}

// ------------------------------------------------------------------------- where dy / dX / partials live
// The traversals read the second input (dy), write the output and the unit's partial through an IO policy.
// LocalIO: this device's memory (every single-device kernel).  lq_dpx.cu supplies PeerIO: dy is the mean over
// the same address in every rank's peer-mapped arena, outputs and partials are stored to every rank.
template <bool STREAM>
struct LocalIO {
  __device__ __forceinline__ float4 ld_in2_4(const float4* p) const { return ld4<STREAM>(p); }
  __device__ __forceinline__ float ld_in2_1(const float* p) const { return ld1<STREAM>(p); }
  __device__ __forceinline__ void st_out4(float4* p, const float4& v) const { st4<STREAM>(p, v); }
  __device__ __forceinline__ void st_out1(float* p, float v) const { st1<STREAM>(p, v); }
  __device__ __forceinline__ void st_part(float4* p, const float4& v) const { *p = v; }
};

#ifndef LQ_COL_U1
#define LQ_COL_U1 8
#endif
#ifndef LQ_COL_ORDER
#define LQ_COL_ORDER 1
#endif
// COL forward (no partials, so the row chunk is free): trips of LQ_COL_U1 rows a thread makes; 0 = the backward's chunks
#ifndef LQ_COL_FWD_TRIPS
#define LQ_COL_FWD_TRIPS 2
#endif
#ifndef LQ_STAT_UNROLL
#define LQ_STAT_UNROLL 4
#endif
#ifndef LQ_STAT_MINBLK
#define LQ_STAT_MINBLK 8
#endif
// resident blocks per SM asked of the compiler for the ROW multi-tensor kernel: the read-only reductions are
// latency-bound (more warps in flight = more bytes in flight), everything else keeps its registers
template <class Op>
constexpr int kRowMinBlocks = (!Op::kIn2 && Op::kReduce && !Op::kOutIsY) ? LQ_STAT_MINBLK : 1;

template <class Op>
struct RowUnroll { static constexpr int value = (!Op::kIn2 && Op::kReduce && !Op::kOutIsY) ? LQ_STAT_UNROLL : 4; };

// ------------------------------------------------------------------------- ROW regime
template <class Op, bool STREAM, bool VEC, class IO = LocalIO<STREAM>>
__device__ __forceinline__ void run_row(const Job& j, unsigned blk, const Op& op, float2* queue, const IO& io = IO(),
                                        long long unit_override = -1) {
  const int lane = threadIdx.x & 31;
  const long long unit = unit_override >= 0 ? unit_override : (long long)blk * kWarps + (threadIdx.x >> 5);
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
  const typename Op::Ctx ctx = op.prep(__ldg(j.s + grp),
                                       Op::kAux ? __ldg(j.gaux + grp) : make_float4(0.f, 0.f, 0.f, 0.f), j);
  const long long base = row * j.inner + (long long)sg * j.seg;
  const long long rem = j.inner - (long long)sg * j.seg;
  const int len = rem < j.seg ? (int)rem : j.seg;

  Acc acc; Op::init(acc);
  float* out = Op::kOutIsY ? j.y : j.dX;
  const bool wr = Op::kOutIsY ? true : (j.write_dx != 0);
  Defer df;
  df.slot = queue + (threadIdx.x >> 5) * (Op::kDeferSlots * 32) + lane;
  df.n = 0;

  if constexpr (VEC) {
    const int nvec = len >> 2;
    const float4* P4 = reinterpret_cast<const float4*>(j.P + base);
    const float4* D4 = reinterpret_cast<const float4*>(j.dy + base);
    float4* O4 = reinterpret_cast<float4*>(out + base);
    // loads in flight per lane and trip: 4 x 16 B per array (8 was measured for the read-only reductions:
    // 73 registers, 31 % warps active, slower)
    constexpr int UR = RowUnroll<Op>::value;
    for (int v0 = 0; v0 < nvec; v0 += 32 * UR) {
      float4 p[UR], d[Op::kIn2 ? UR : 1];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int v = v0 + u * 32 + lane;
        if (v < nvec) {
          p[u] = ld4<STREAM>(P4 + v);
          if (Op::kIn2) d[Op::kIn2 ? u : 0] = io.ld_in2_4(D4 + v);
        }
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int v = v0 + u * 32 + lane;
        if (v < nvec) {
          float4 o, c;
          op.quad(p[u], Op::kIn2 ? d[Op::kIn2 ? u : 0] : make_float4(0.f, 0.f, 0.f, 0.f), ctx, j, acc, o, c, df);
          if (wr) io.st_out4(O4 + v, o);
          if (Op::kCodes && j.codes) {
            const long long e = base + 4ll * v;
            store_code(j.codes, j.codes_dtype, e + 0, c.x);
            store_code(j.codes, j.codes_dtype, e + 1, c.y);
            store_code(j.codes, j.codes_dtype, e + 2, c.z);
            store_code(j.codes, j.codes_dtype, e + 3, c.w);
          }
        }
      }
      if (Op::kDeferSlots) op.flush(acc, df);   // warp is converged here (v0 loop is warp-uniform)
    }
  } else {
    for (int e0 = 0; e0 < len; e0 += 128) {
      float p[4], d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * 32 + lane;
        if (e < len) {
          p[u] = ld1<STREAM>(j.P + base + e);
          if (Op::kIn2) d[u] = io.ld_in2_1(j.dy + base + e);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * 32 + lane;
        if (e < len) {
          float o, c;
          op.elem(p[u], Op::kIn2 ? d[u] : 0.f, ctx, j, acc, o, c);
          if (wr) io.st_out1(out + base + e, o);
          if (Op::kCodes && j.codes) store_code(j.codes, j.codes_dtype, base + e, c);
        }
      }
    }
  }
  if (Op::kReduce && j.part != nullptr) {
    acc = op.warp_reduce(acc);
    if (lane == 0) io.st_part(j.part + unit, pack_acc(acc));
  }
}

// ------------------------------------------------------------------------- COL regime
// COL traversal, one column per thread: a warp reads 32 consecutive columns of a row (one 128-byte
// line), a thread walks DOWN its column, so -- exactly as in ROW -- it serves a single group: one
// scale/reciprocal, one accumulator, the 4-element guard and the untagged rare-path queue of the
// ROW operators apply unchanged to four consecutive rows.  8 rows (x 2 arrays) are in flight per
// thread.  This is the path of every two-input reducing operator (the threshold backward): with
// four columns per thread their per-thread state for four groups cost 50 % more instructions and
// ran at 4.2 TB/s where ROW reaches 6.3.
template <class Op, bool STREAM, class IO = LocalIO<STREAM>>
__device__ __forceinline__ void run_col1(const Job& j, unsigned blk, const Op& op, Acc* smem, float2* queue, const IO& io = IO()) {
  constexpr int U = 8;
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
  const long long col = cb * j.bx + tx;
  const bool active = (ty < j.by) && (col < j.ncol);
  const long long r0 = chunk * j.rpc;
  const long long r1 = (r0 + j.rpc < j.outer) ? r0 + j.rpc : j.outer;
  const int rows_blk = (int)((r1 - r0 + j.by - 1) / j.by);                 // trips are block-uniform
  const int rows_mine = active ? (int)((r1 - r0 - ty + j.by - 1) / j.by) : 0;   // rows r0+ty, +by, ...

  const long long grp = active ? col / j.inner : 0;
  const typename Op::Ctx ctx = op.prep(active ? __ldg(j.s + grp) : 1.0f,
                                       (Op::kAux && active) ? __ldg(j.gaux + grp) : make_float4(0.f, 0.f, 0.f, 0.f), j);
  Acc acc; Op::init(acc);
  float* out = Op::kOutIsY ? j.y : j.dX;
  const bool wr = Op::kOutIsY ? true : (j.write_dx != 0);
  Defer df;
  df.slot = queue + (t >> 5) * (Op::kDeferSlots * 32) + (t & 31);
  df.n = 0;

  const long long stride = (long long)j.by * j.W;
  const long long base = active ? (r0 + ty) * j.W + col : 0;
  const float* Pp = j.P + base;
  const float* Dp = Op::kIn2 ? j.dy + base : nullptr;
  float* Op_ = out + base;

#pragma unroll 1
  for (int i0 = 0; i0 < rows_blk; i0 += U) {
    float p[U], d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u < rows_mine) {
        p[u] = ld1<STREAM>(Pp + u * stride);
        if (Op::kIn2) d[u] = io.ld_in2_1(Dp + u * stride);
      }
    }
#pragma unroll
    for (int qd = 0; qd < U / 4; ++qd) {
      const int i = i0 + 4 * qd;
      if (i + 3 < rows_mine) {
        float4 o, c;
        op.quad(make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]),
                Op::kIn2 ? make_float4(d[4 * qd], d[4 * qd + 1], d[4 * qd + 2], d[4 * qd + 3]) : make_float4(0.f, 0.f, 0.f, 0.f),
                ctx, j, acc, o, c, df);
        if (wr) {
          io.st_out1(Op_ + (4 * qd) * stride, o.x); io.st_out1(Op_ + (4 * qd + 1) * stride, o.y);
          io.st_out1(Op_ + (4 * qd + 2) * stride, o.z); io.st_out1(Op_ + (4 * qd + 3) * stride, o.w);
        }
        if (Op::kCodes && j.codes) {
          const long long e = base + (long long)i * stride;
          store_code(j.codes, j.codes_dtype, e, c.x); store_code(j.codes, j.codes_dtype, e + stride, c.y);
          store_code(j.codes, j.codes_dtype, e + 2 * stride, c.z); store_code(j.codes, j.codes_dtype, e + 3 * stride, c.w);
        }
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) {          // the column's last 1-3 rows
          if (i + v < rows_mine) {
            float o, c;
            op.elem(p[4 * qd + v], Op::kIn2 ? d[4 * qd + v] : 0.f, ctx, j, acc, o, c);
            if (wr) io.st_out1(Op_ + (4 * qd + v) * stride, o);
            if (Op::kCodes && j.codes) store_code(j.codes, j.codes_dtype, base + (long long)(i + v) * stride, c);
          }
        }
      }
    }
    if (Op::kDeferSlots) op.flush(acc, df);     // block-uniform trip count: the warp is converged
    Pp += U * stride;
    if (Op::kIn2) Dp += U * stride;
    Op_ += U * stride;
  }

  if (Op::kReduce) {
    if (j.by > 1) {
      smem[t] = acc;
      int span = 1;
      while (span < j.by) span <<= 1;
      for (int st = span >> 1; st >= 1; st >>= 1) {
        __syncthreads();
        if (ty < st && ty + st < j.by) {
          Acc mine = smem[t];
          Op::combine(mine, smem[t + st * j.bx]);
          smem[t] = mine;
        }
      }
      __syncthreads();
      acc = smem[t];
    }
    if (ty == 0 && col < j.ncol) io.st_part(j.part + chunk * j.W + col, pack_acc(acc));
  }
}

template <class Op, bool STREAM, bool VEC>
__device__ __forceinline__ void run_col(const Job& j, unsigned blk, const Op& op, Acc* smem, float2* queue) {
  if constexpr (!VEC) { run_col1<Op, STREAM>(j, blk, op, smem, queue); return; }
  constexpr int V = VEC ? 4 : 1;
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
  const long long col0 = cv * V;
  const long long r0 = chunk * j.rpc;
  const long long r1 = (r0 + j.rpc < j.outer) ? r0 + j.rpc : j.outer;

  typename Op::Ctx ctx[V];
  Acc acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    Op::init(acc[k]);
    const long long grp = active ? (col0 + k) / j.inner : 0;
    ctx[k] = op.prep(active ? __ldg(j.s + grp) : 1.0f,
                     (Op::kAux && active) ? __ldg(j.gaux + grp) : make_float4(0.f, 0.f, 0.f, 0.f), j);
  }
  float* out = Op::kOutIsY ? j.y : j.dX;
  const bool wr = Op::kOutIsY ? true : (j.write_dx != 0);
  Defer df;
  df.slot = queue + (t >> 5) * (Op::kDeferSlots * 32) + (t & 31);
  df.n = 0;

  // The row loop is block-uniform (every thread runs the same trips; rows past the end and
  // idle threads are predicated off), so the warp is converged at flush_k.
  // Rows per trip: 4 for the light operators; 2 for two-input reducing operators (the accumulating
  // Difference backward), whose per-thread state for four groups would otherwise push the kernel past
  // 64 registers and halve the resident warps.  (The threshold backward does not come here: it takes
  // run_col1 or the TMA-staged kernel.)
  // ... and LQ_COL_U1 rows for the one-input operators without a reduction (the forward): ncu had it waiting on loads
  // (long-scoreboard stall 14 per issue, 70 % DRAM throughput, 35 % issue-active) with four 16-byte loads per thread in flight.
  constexpr int U = (Op::kIn2 && Op::kReduce) ? 2 : ((!Op::kIn2 && !Op::kReduce) ? LQ_COL_U1 : 4);
  // Addresses: one pointer per array, advanced by a constant stride (the per-row 64-bit multiply of
  // row * W + column and its bound check used to be a third of the instruction stream); rows are
  // counted, not compared: this thread owns rows r0+ty, r0+ty+by, ... (rows_mine of them).
  const int rows_blk = (int)((r1 - r0 + j.by - 1) / j.by);                       // block-uniform trips
  const int rows_mine = active ? (int)((r1 - r0 - ty + j.by - 1) / j.by) : 0;
  const long long stride = (long long)j.by * j.W;
  const long long base = active ? (r0 + ty) * j.W + col0 : 0;
  const float* Pp = j.P + base;
  const float* Dp = Op::kIn2 ? j.dy + base : nullptr;
  float* Op_ = out + base;
  // A full trip of four rows hands each of the thread's columns to the operator as one quad (four
  // elements of ONE group: one range guard, one branch), exactly like a ROW thread's 16 bytes.
  constexpr bool kQuads = (U % 4 == 0) && (V == 4) && (Op::kDeferSlots == 0);

#pragma unroll 1
  for (int i0 = 0; i0 < rows_blk; i0 += U) {
    float p[U][V], d[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u < rows_mine) {
        if constexpr (VEC) {
          const float4 pv = ld4<STREAM>(reinterpret_cast<const float4*>(Pp + u * stride));
          p[u][0] = pv.x; p[u][1] = pv.y; p[u][2] = pv.z; p[u][3] = pv.w;
          if (Op::kIn2) {
            const float4 dv = ld4<STREAM>(reinterpret_cast<const float4*>(Dp + u * stride));
            d[u][0] = dv.x; d[u][1] = dv.y; d[u][2] = dv.z; d[u][3] = dv.w;
          }
        } else {
          p[u][0] = ld1<STREAM>(Pp + u * stride);
          if (Op::kIn2) d[u][0] = ld1<STREAM>(Dp + u * stride);
        }
      }
    }
    // Four rows at a time: a full group of four hands each of the thread's columns to the operator as one quad; the
    // last 1-3 rows of a chunk, and launches that export codes, take the per-element form (same operations per element,
    // same accumulation order: the results do not depend on which form ran).
    constexpr int G = kQuads ? 4 : U;
#pragma unroll
    for (int q0 = 0; q0 < U; q0 += G) {
      bool done = false;
      if constexpr (kQuads) {
        if (i0 + q0 + 4 <= rows_mine && !(Op::kCodes && j.codes)) {
#pragma unroll
          for (int k = 0; k < V; ++k) {
            float4 oq, cq;
            op.quad(make_float4(p[q0][k], p[q0 + 1][k], p[q0 + 2][k], p[q0 + 3][k]),
                    Op::kIn2 ? make_float4(d[q0][k], d[q0 + 1][k], d[q0 + 2][k], d[q0 + 3][k]) : make_float4(0.f, 0.f, 0.f, 0.f),
                    ctx[k], j, acc[k], oq, cq, df);
            p[q0][k] = oq.x; p[q0 + 1][k] = oq.y; p[q0 + 2][k] = oq.z; p[q0 + 3][k] = oq.w;   // outputs take the inputs' registers
          }
          if (wr) {
#pragma unroll
            for (int u = q0; u < q0 + 4; ++u)
              st4<STREAM>(reinterpret_cast<float4*>(Op_ + u * stride), make_float4(p[u][0], p[u][1], p[u][2], p[u][3]));
          }
          done = true;
        }
      }
      if (!done) {
#pragma unroll
        for (int u = q0; u < q0 + G; ++u) {
          if (i0 + u < rows_mine) {
            float o[V], c[V];
#pragma unroll
            for (int k = 0; k < V; ++k)
              op.elem_k(p[u][k], Op::kIn2 ? d[u][k] : 0.f, ctx[k], j, acc, k, o[k], c[k], df);
            if (wr) {
              if constexpr (VEC) st4<STREAM>(reinterpret_cast<float4*>(Op_ + u * stride), make_float4(o[0], o[1], o[2], o[3]));
              else st1<STREAM>(Op_ + u * stride, o[0]);
            }
            if (Op::kCodes && j.codes) {
              const long long e = base + (long long)(i0 + u) * stride;
#pragma unroll
              for (int k = 0; k < V; ++k) store_code(j.codes, j.codes_dtype, e + k, c[k]);
            }
          }
        }
      }
    }
    if (Op::kDeferSlots) op.flush_k(acc, df);     // block-uniform trip count: the warp is converged
    Pp += U * stride;
    if (Op::kIn2) Dp += U * stride;
    Op_ += U * stride;
  }

  if (Op::kReduce) {
    // fixed-shape tree over the threads that share a column (same tx, ty = 0..by-1)
    if (j.by > 1) {
#pragma unroll
      for (int k = 0; k < V; ++k) smem[k * kThreads + t] = acc[k];
      int span = 1;
      while (span < j.by) span <<= 1;
      for (int stride = span >> 1; stride >= 1; stride >>= 1) {
        __syncthreads();
        if (ty < stride && ty + stride < j.by && ty < j.by) {
#pragma unroll
          for (int k = 0; k < V; ++k) {
            Acc mine = smem[k * kThreads + t];
            Op::combine(mine, smem[k * kThreads + t + stride * j.bx]);
            smem[k * kThreads + t] = mine;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] = smem[k * kThreads + t];
    }
    if (ty == 0 && cv < j.ncol) {
#pragma unroll
      for (int k = 0; k < V; ++k) j.part[chunk * j.W + col0 + k] = pack_acc(acc[k]);
    }
  }
}

// Single-tensor kernels: the Job travels in the kernel parameters.  One instantiation per
// regime so that each keeps only its own registers (the ROW kernel is the bandwidth path).
// COL kernels of the light operators are held to 64 registers (four blocks per SM): grids of ~3.5 blocks
// per SM (a 4 Mi-element tensor) otherwise need a second wave.  The two-input operators keep
// what they need (they carry per-thread state for the rare path / four groups).
#ifndef LQ_COL_FWD_MINBLK
#define LQ_COL_FWD_MINBLK 3
#endif
template <class Op>
constexpr int kColMinBlocks = Op::kIn2 ? 1 : ((!Op::kReduce) ? LQ_COL_FWD_MINBLK : 4);

template <class Op>
struct SmemPlan {
  static constexpr int kQueueEntries = Op::kDeferSlots ? Op::kDeferSlots * kThreads : 1;
  static constexpr int kColBytes = Op::kReduce ? 4 * kThreads * (int)sizeof(Acc) : 16;
  static constexpr int kColEntries = Op::kReduce ? 4 * kThreads : 1;
};

template <class Op, bool STREAM, bool VEC>
__global__ void __launch_bounds__(kThreads) row_kernel(const __grid_constant__ Job j, const Op op) {
  __shared__ float2 queue[SmemPlan<Op>::kQueueEntries];
  cudaGridDependencySynchronize();   // launched with launch_after: ordering with the previous kernel is kept here
  run_row<Op, STREAM, VEC>(j, blockIdx.x, op, queue);
}
template <class Op, bool STREAM, bool VEC>
__global__ void __launch_bounds__(kThreads, kColMinBlocks<Op>) col_kernel(const __grid_constant__ Job j, const Op op) {
  __shared__ Acc smem[SmemPlan<Op>::kColEntries];
  __shared__ float2 queue[SmemPlan<Op>::kQueueEntries];
  cudaGridDependencySynchronize();
  run_col<Op, STREAM, VEC>(j, blockIdx.x, op, smem, queue);
}

// Multi-tensor kernel: block -> tensor by binary search over block_begin.  One instantiation per
// regime (a launch holds the ROW tensors or the COL tensors of a step, never both), so that the ROW
// units keep the ROW kernel's register budget: the combined kernel needed 114 registers.
template <class Op, int REGIME>
__global__ void __launch_bounds__(kThreads, REGIME == REG_COL ? kColMinBlocks<Op> : kRowMinBlocks<Op>) mt_kernel(const __grid_constant__ JobTable tab, const Op op) {
  __shared__ Acc smem[REGIME == REG_COL ? SmemPlan<Op>::kColEntries : 1];
  __shared__ float2 queue[SmemPlan<Op>::kQueueEntries];
  int lo = 0, hi = tab.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.job[mid].block_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const Job& j = tab.job[lo];
  const unsigned blk = blockIdx.x - j.block_begin;
  cudaGridDependencySynchronize();
  if constexpr (REGIME == REG_ROW) {
    if (j.vec) run_row<Op, false, true>(j, blk, op, queue);
    else run_row<Op, false, false>(j, blk, op, queue);
  } else {
    if (j.vec) run_col<Op, false, true>(j, blk, op, smem, queue);
    else run_col<Op, false, false>(j, blk, op, smem, queue);
  }
}

constexpr int kFinThreads = 128;

__device__ __forceinline__ long long partials_per_group(const FinJob& f) {
  return f.regime == REG_ROW ? f.outer * f.nseg : (f.regime == REG_COL ? f.nchunks * f.inner : f.nlin);
}
__device__ __forceinline__ long long partial_index(const FinJob& f, long long g, long long k) {
  if (f.regime == REG_LIN) return g * f.nlin + k;
  if (f.regime == REG_ROW) {
    const long long o = k / f.nseg, sg = k - o * f.nseg;
    return (o * f.C + g) * f.nseg + sg;
  }
  const long long ch = k / f.inner, ci = k - ch * f.inner;
  return ch * f.W + g * f.inner + ci;
}

// ---- host side
int fq_fwd_launch(const Geom& g, const float* P, const float* s, float* y, void* codes, int codes_dtype,
                  const QMode& qm, cudaStream_t st);
int fq_bwd_elements_launch(const Geom& g, const float* dy, const float* P, const float* s, float lam,
                           float* dX, float4* part, const QMode& qm, cudaStream_t st);
int fq_finalize_launch(const Geom& g, const float4* part, float* ds, const float* s, float lam, cudaStream_t st, int mode = 0);
int check_blocks(const Geom& g);
FinJob make_fin(const Geom& g, const float4* part, float* ds, const float* s, float lam);
int mt_finalize_launch(FinTable& fin, unsigned groups, cudaStream_t st);
Job make_job(const Geom& g, const float* P, const float* s, const float* dy, float* y, float* dX,
             void* codes, int codes_dtype, float4* part);

// Launch a second-stage kernel with programmatic dependent launch: its launch latency overlaps
// the tail of the preceding kernel in the stream; the kernel itself must call
// cudaGridDependencySynchronize() before touching that kernel's output.
template <class... KArgs, class... Args>
int launch_after(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, const char* what, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cuda_fail(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...), what);
}

template <class Op>
int launch_job(const Job& j, const Op& op, cudaStream_t st) {
  if (j.nblocks == 0) return LQ_OK;
  dim3 grid(j.nblocks), block(kThreads);
#define LQ_LAUNCH(KERNEL)                                                            \
  do {                                                                               \
    if (j.stream_hint) {                                                             \
      if (j.vec) return launch_after(KERNEL<Op, true, true>, grid, block, st, "kernel launch", j, op);   \
      return launch_after(KERNEL<Op, true, false>, grid, block, st, "kernel launch", j, op);             \
    } else {                                                                         \
      if (j.vec) return launch_after(KERNEL<Op, false, true>, grid, block, st, "kernel launch", j, op);  \
      return launch_after(KERNEL<Op, false, false>, grid, block, st, "kernel launch", j, op);            \
    }                                                                                \
  } while (0)
  if (j.regime == REG_ROW) LQ_LAUNCH(row_kernel); else LQ_LAUNCH(col_kernel);
#undef LQ_LAUNCH
  return cuda_fail(cudaGetLastError(), "kernel launch");
}

// Row chunks of the interleaved (COL) tensors of ONE multi-tensor launch.  make_geom's rule gives every tensor enough
// blocks to fill the GPU by itself (~4 per SM); with the 40 tensors of a model in one launch that made 2700-3700 blocks
// of 8 rows each (16 KB of work behind a prologue of four reciprocals per thread): 29 us forward / 38 us backward for
// 45 MB that sit in L2.  Here the launch as a whole gets ~3 blocks per SM (LQ_MT_BPS) and a tensor its share by size; a tensor never
// gets shorter chunks than make_geom's (a launch of one large tensor is unchanged).  rpc[i - i0] = 0: keep the default.
// The workspace is sized for the default geometry (lq_workspace_bytes), an upper bound: longer chunks, fewer partials.
#ifndef LQ_MT_BPS
#define LQ_MT_BPS 3.0
#endif
// (3 measured best of 3 / 5 / 8 / 12 / 20 on the 40-tensor inventory: forward 18.0 / 19.3 / 19.5 / 20.6 / 20.5 us)
// target[i - i0] = blocks the launch's budget (~LQ_MT_BPS per SM) grants interleaved tensor i by its size; 0 = not such a tensor
inline void mt_col_targets(const lq_tensor_desc* descs, int i0, int i1, long long* target) {
  double n_col = 0.0;
  auto is_col = [](const lq_tensor_desc& d) { return d.outer > 0 && d.C > 1 && d.inner > 0 && d.inner < kRowMinInner; };
  for (int i = i0; i < i1; ++i) {
    target[i - i0] = 0;
    if (is_col(descs[i])) n_col += (double)descs[i].outer * (double)descs[i].C * (double)descs[i].inner;
  }
  int count = 0;
  for (int i = i0; i < i1; ++i) count += is_col(descs[i]) ? 1 : 0;
  if (count < 2) return;                              // a tensor by itself keeps the geometry tuned for it
  for (int i = i0; i < i1; ++i) {
    if (!is_col(descs[i])) continue;
    const double n = (double)descs[i].outer * (double)descs[i].C * (double)descs[i].inner;
    const long long t = (long long)(148.0 * LQ_MT_BPS * n / n_col);
    target[i - i0] = t < 1 ? 1 : t;
  }
}
// the row chunk that gives a tensor of `outer` rows and `ncolblk` column blocks of `cols_per_block` columns about `target`
// blocks: a multiple of `round_to`, and never more than ~64 Ki elements per block (the rule exists for the small tensors
// of a model; large ones keep blocks of the size their own geometry was tuned to)
inline long long share_rpc(long long outer, long long ncolblk, long long target, long long round_to, long long cols_per_block) {
  if (target <= 0) return 0;
  long long r = cdiv(cdiv(outer * ncolblk, target), round_to) * round_to;
  long long cap = (65536 / (cols_per_block > 0 ? cols_per_block : 1)) / round_to * round_to;
  if (cap < round_to) cap = round_to;
  return r < cap ? r : cap;
}
inline void mt_col_rpc(const lq_tensor_desc* descs, int i0, int i1, int rows_per_trip, bool vec_form, long long* rpc) {
  long long target[kMaxMT];
  mt_col_targets(descs, i0, i1, target);
  for (int i = i0; i < i1; ++i) {
    rpc[i - i0] = 0;
    if (target[i - i0] == 0) continue;
    const lq_tensor_desc& d = descs[i];
    const Geom g = make_geom(d.outer, d.C, d.inner, vec_form);      // bx / by / ncolblk of the form that will run
    if (g.regime != REG_COL) continue;
    const long long r = share_rpc(d.outer, g.ncolblk, target[i - i0], (long long)g.by * rows_per_trip, (long long)g.bx * (g.vec_ok ? 4 : 1));
    if (r > g.rpc) rpc[i - i0] = r;
  }
}

template <class Op>
int launch_mt(JobTable& tab, const Op& op, cudaStream_t st) {
  JobTable part;
  for (int regime = REG_ROW; regime <= REG_COL; ++regime) {
    part.n = 0;
    unsigned total = 0;
    for (int i = 0; i < tab.n; ++i) {
      if (tab.job[i].regime != regime || tab.job[i].nblocks == 0) continue;
      part.job[part.n] = tab.job[i];
      part.job[part.n].block_begin = total;
      total += tab.job[i].nblocks;
      ++part.n;
    }
    if (total == 0) continue;
    const int rc = (regime == REG_ROW)
        ? launch_after(mt_kernel<Op, REG_ROW>, dim3(total), dim3(kThreads), st, "multi-tensor ROW launch", part, op)
        : launch_after(mt_kernel<Op, REG_COL>, dim3(total), dim3(kThreads), st, "multi-tensor COL launch", part, op);
    if (rc) return rc;
  }
  return LQ_OK;
}

}  // namespace lq
