// lq_debug.cu -- on-device self-check of the hoisted-reciprocal division used by K1/K2
// (div_fast in lq_common.cuh) against __fdiv_rn, over whole mantissa ranges.
#include "lq_traverse.cuh"

namespace lq {

// Block b handles scale index b / 128 and the a-mantissa slice b % 128 (65536 mantissas);
// each thread walks 256 of them.  For every pair inside the guard window the fast quotient,
// its floor and the floor's product with s must equal the __fdiv_rn versions bit for bit.
__global__ void __launch_bounds__(256) divcheck_kernel(const uint32_t* s_bits, uint32_t a_hi_bits,
                                                        unsigned long long* out /* [0]=mismatches [1]=checked [2]=first a|s */) {
  const uint32_t si = blockIdx.x >> 7;
  const uint32_t slice = blockIdx.x & 127u;
  const float s = __uint_as_float(s_bits[si]);
  const Recip rc = make_recip(s);
  unsigned bad = 0, seen = 0;
  for (int k = 0; k < 256; ++k) {
    const uint32_t mant = (slice << 16) | ((uint32_t)k << 8) | threadIdx.x;
    const float a = __uint_as_float(a_hi_bits | mant);
    if (!(rc.fast && fast_range(a))) continue;
    ++seen;
    const float qf = div_fast(a, rc);
    const float qr = __fdiv_rn(a, s);
    if (__float_as_uint(qf) != __float_as_uint(qr)) {
      if (bad == 0) atomicCAS(out + 2, 0ull, ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(s));
      ++bad;
    }
  }
  bad = warp_sum_u(bad);
  seen = warp_sum_u(seen);
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(out, (unsigned long long)bad);
    atomicAdd(out + 1, (unsigned long long)seen);
  }
}

}  // namespace lq

namespace lq { Geom fwd_geom(long long outer, long long C, long long inner, bool al, long long min_rpc); }   // lq_fq.cu

using namespace lq;

extern "C" int lq_debug_divcheck(const uint32_t* s_bits, int64_t n_s, uint32_t a_sign_exp_bits,
                                 unsigned long long* out3, lq_stream_t stream) {
  if (!s_bits || !out3) { set_error("lq_debug_divcheck: null argument"); return LQ_E_NULLPTR; }
  if (n_s < 0 || n_s > (1ll << 24)) { set_error("lq_debug_divcheck: n_s out of range"); return LQ_E_BADSHAPE; }
  if (a_sign_exp_bits & 0x007fffffu) { set_error("lq_debug_divcheck: mantissa bits set in a_sign_exp_bits"); return LQ_E_BADMODE; }
  if (n_s == 0) return LQ_OK;
  divcheck_kernel<<<dim3((unsigned)(n_s * 128)), dim3(256), 0, (cudaStream_t)stream>>>(s_bits, a_sign_exp_bits, out3);
  return cuda_fail(cudaGetLastError(), "divcheck launch");
}

// Host-only: how lq_mt_fq_fwd (phase LQ_PHASE_FWD) / lq_mt_fq_bwd (LQ_PHASE_BWD) would cut each tensor of a table --
// out[4 i .. 4 i + 3] = {regime (0 ROW, 1 COL), rows per chunk (COL) or segment length (ROW), blocks, partial slots}.
// No device work, no pointers read: for tests of the geometry rules (a tensor by itself keeps its own geometry; in a
// table the interleaved tensors share one grid; chunks only grow; the workspace bound holds).
extern "C" int lq_debug_mt_geometry(const lq_tensor_desc* descs, int n, int phase, int64_t* out) {
  if (n < 0 || (n && (!descs || !out))) { set_error("lq_debug_mt_geometry: bad arguments"); return LQ_E_NULLPTR; }
  if (phase != LQ_PHASE_FWD && phase != LQ_PHASE_BWD) { set_error("lq_debug_mt_geometry: phase must be FWD or BWD"); return LQ_E_BADMODE; }
  int rc = 0;
  for (int i0 = 0; i0 < n; i0 += kMaxMT) {
    const int i1 = (i0 + kMaxMT < n) ? i0 + kMaxMT : n;
    for (int i = i0; i < i1; ++i)
      if ((rc = check_dims(descs[i].outer, descs[i].C, descs[i].inner))) return rc;
    long long share[kMaxMT];
    if (phase == LQ_PHASE_FWD) mt_col_rpc(descs, i0, i1, LQ_COL_U1, true, share);
    else mt_col_rpc(descs, i0, i1, 8, false, share);
    for (int i = i0; i < i1; ++i) {
      const lq_tensor_desc& d = descs[i];
      int64_t* o = out + 4 * (int64_t)i;
      if (d.outer * d.C * d.inner == 0) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
      const Geom g = (phase == LQ_PHASE_FWD) ? fwd_geom(d.outer, d.C, d.inner, true, share[i - i0])
                                             : make_geom(d.outer, d.C, d.inner, d.inner >= kRowMinInner, (int)share[i - i0]);
      o[0] = g.regime; o[1] = (g.regime == REG_COL) ? g.rpc : g.seg; o[2] = g.nblocks; o[3] = g.npartials;
    }
  }
  return LQ_OK;
}

