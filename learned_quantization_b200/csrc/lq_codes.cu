// lq_codes.cu -- SURVEY 8(f2): integer-code statistics on the device.
//
// The reference recomputes floor(P / s) on the HOST every epoch (weights pulled with .numpy()),
// then np.unique(..., return_counts=True) and max|q| (custom_callbacks.py:84-129), and at the end
// exports floor(P / s).astype(np.int8) (utils/log_scripts.py:74-79), which silently wraps when a
// code does not fit int8 (codes reach +-2e4 with the initial scale).  Here:
//   lq_code_range      min / max code of a tensor                       (2 floats back to the host)
//   lq_code_histogram  counts per integer code in a window [lo, lo+nbins) + how many fall outside
// np.unique's (values, counts) are the non-zero bins; the number of codes outside [-128, 127] is
// the int8 saturation check; the int8 payload itself is lq_fq_fwd(codes_dtype = LQ_CODES_I8).
// All accumulation is integer (or min/max): deterministic.
#include <math.h>
#include "lq_common.cuh"

namespace lq {

__device__ __forceinline__ unsigned ordered(float f) {          // monotone float -> uint map
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float unordered(unsigned u) {
  const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(v);
#else
  float f; memcpy(&f, &v, 4); return f;
#endif
}

struct CodeJob { const float* P; const float* s; long long n, C, inner; QMode qm; };

template <bool EXT>
__device__ __forceinline__ float code_at(const CodeJob& j, long long idx) {
  const long long g = (idx / j.inner) % j.C;
  return round_code<EXT>(__fdiv_rn(j.P[idx], __ldg(j.s + g)), j.qm);
}

// out[0] = ordered(min q), out[1] = ordered(max q); NaN codes are skipped (np.unique would list them;
// they are reported through the histogram's "outside" count instead).
template <bool EXT>
__global__ void __launch_bounds__(256) code_range_kernel(const __grid_constant__ CodeJob j, unsigned* out) {
  float mn = INFINITY, mx = -INFINITY;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += stride) {
    const float q = code_at<EXT>(j, i);
    mn = fminf(mn, q); mx = fmaxf(mx, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mn <= mx) { atomicMin(out, ordered(mn)); atomicMax(out + 1, ordered(mx)); }
  }
}

__global__ void code_range_init_kernel(unsigned* out) { out[0] = 0xffffffffu; out[1] = 0u; }
__global__ void code_range_decode_kernel(const unsigned* in, float* out) {
  out[0] = (in[0] == 0xffffffffu && in[1] == 0u) ? NAN : unordered(in[0]);
  out[1] = (in[0] == 0xffffffffu && in[1] == 0u) ? NAN : unordered(in[1]);
}

constexpr int kSmemBins = 8192;

template <bool EXT, bool SMEM>
__global__ void __launch_bounds__(256) code_hist_kernel(const __grid_constant__ CodeJob j, long long lo, long long nbins,
                                                        unsigned* hist, unsigned long long* outside) {
  __shared__ unsigned sh[SMEM ? kSmemBins : 1];
  if (SMEM) {
    for (int b = threadIdx.x; b < nbins; b += 256) sh[b] = 0u;
    __syncthreads();
  }
  unsigned out_cnt = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += stride) {
    const float q = code_at<EXT>(j, i);
    const float rel = q - (float)lo;
    if (rel >= 0.0f && rel < (float)nbins) {          // false for NaN
      const long long b = (long long)rel;
      if (SMEM) atomicAdd(&sh[b], 1u); else atomicAdd(&hist[b], 1u);
    } else {
      ++out_cnt;
    }
  }
  out_cnt = warp_sum_u(out_cnt);
  if ((threadIdx.x & 31) == 0 && out_cnt) atomicAdd(outside, (unsigned long long)out_cnt);
  if (SMEM) {
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += 256) if (sh[b]) atomicAdd(&hist[b], sh[b]);
  }
}

static int code_job(const char* who, const float* P, const float* s, long long outer, long long C, long long inner,
                    int round_mode, float qmin, float qmax, CodeJob* j) {
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  if (outer * C * inner != 0 && (!P || !s)) { set_error("%s: null P/s", who); return LQ_E_NULLPTR; }
  j->P = P; j->s = s; j->n = outer * C * inner; j->C = C; j->inner = inner;
  j->qm = make_qmode(round_mode, qmin, qmax, &rc);
  return rc;
}

static unsigned code_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148ll * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace lq

using namespace lq;

extern "C" int lq_code_range(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner,
                             int round_mode, float qmin, float qmax, float* out_min_max,
                             void* ws, size_t ws_bytes, lq_stream_t stream) {
  CodeJob j;
  int rc = code_job("lq_code_range", P, s, outer, C, inner, round_mode, qmin, qmax, &j);
  if (rc) return rc;
  if (!out_min_max) { set_error("lq_code_range: null out"); return LQ_E_NULLPTR; }
  if (!ws || ws_bytes < 8 || (reinterpret_cast<uintptr_t>(ws) & 3u)) { set_error("lq_code_range: 8-byte workspace needed"); return LQ_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* tmp = static_cast<unsigned*>(ws);
  code_range_init_kernel<<<1, 1, 0, st>>>(tmp);
  if (j.n) {
    if (j.qm.ext) code_range_kernel<true><<<code_grid(j.n), 256, 0, st>>>(j, tmp);
    else code_range_kernel<false><<<code_grid(j.n), 256, 0, st>>>(j, tmp);
  }
  code_range_decode_kernel<<<1, 1, 0, st>>>(tmp, out_min_max);
  return cuda_fail(cudaGetLastError(), "code range launch");
}

extern "C" int lq_code_histogram(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner,
                                 int round_mode, float qmin, float qmax, int64_t code_lo, int64_t nbins,
                                 uint32_t* hist, unsigned long long* outside, lq_stream_t stream) {
  CodeJob j;
  int rc = code_job("lq_code_histogram", P, s, outer, C, inner, round_mode, qmin, qmax, &j);
  if (rc) return rc;
  if (nbins <= 0 || nbins > (1ll << 28)) { set_error("lq_code_histogram: nbins out of range"); return LQ_E_BADSHAPE; }
  if (code_lo < -(1ll << 30) || code_lo + nbins > (1ll << 30)) { set_error("lq_code_histogram: window outside +-2^30"); return LQ_E_BADSHAPE; }
  if (!hist || !outside) { set_error("lq_code_histogram: null hist/outside"); return LQ_E_NULLPTR; }
  if (!j.n) return LQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = code_grid(j.n);
  if (nbins <= kSmemBins) {
    if (j.qm.ext) code_hist_kernel<true, true><<<grid, 256, 0, st>>>(j, code_lo, nbins, hist, outside);
    else code_hist_kernel<false, true><<<grid, 256, 0, st>>>(j, code_lo, nbins, hist, outside);
  } else {
    if (j.qm.ext) code_hist_kernel<true, false><<<grid, 256, 0, st>>>(j, code_lo, nbins, hist, outside);
    else code_hist_kernel<false, false><<<grid, 256, 0, st>>>(j, code_lo, nbins, hist, outside);
  }
  return cuda_fail(cudaGetLastError(), "code histogram launch");
}
