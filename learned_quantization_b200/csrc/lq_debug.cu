// lq_debug.cu -- on-device self-check of the hoisted-reciprocal division used by K1/K2
// (div_fast in lq_common.cuh) against __fdiv_rn, over whole mantissa ranges.
#include "lq_common.cuh"

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
