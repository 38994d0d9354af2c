// lq_api.cu -- error reporting, argument checks, workspace arithmetic and the small
// stand-alone entry points of liblq.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <mutex>
#include "lq_common.cuh"
#include "lq_col_tma.cuh"

namespace lq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return LQ_OK;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return LQ_E_CUDA;
}

int check_dims(long long outer, long long C, long long inner) {
  if (outer < 0 || C < 0 || inner < 0) { set_error("negative dimension [%lld,%lld,%lld]", outer, C, inner); return LQ_E_BADSHAPE; }
  if (outer == 0 || C == 0 || inner == 0) return LQ_OK;
  if (C > 0x7fffffffll) { set_error("C = %lld exceeds 2^31-1 groups", C); return LQ_E_BADSHAPE; }
  const long double n = (long double)outer * (long double)C * (long double)inner;
  if (n > 4.0e18L) { set_error("element count overflows int64"); return LQ_E_BADSHAPE; }
  return LQ_OK;
}

QMode make_qmode(int round_mode, float qmin, float qmax, int* err) {
  QMode m{};
  *err = LQ_OK;
  if (round_mode != LQ_ROUND_FLOOR && round_mode != LQ_ROUND_RNE) {
    set_error("bad round_mode %d", round_mode);
    *err = LQ_E_BADMODE;
    return m;
  }
  if (isnan(qmin) || isnan(qmax) || qmin > qmax) {
    set_error("bad clamp range [%g, %g]", (double)qmin, (double)qmax);
    *err = LQ_E_BADMODE;
    return m;
  }
  m.rne = (round_mode == LQ_ROUND_RNE);
  m.qmin = qmin; m.qmax = qmax;
  m.ext = m.rne || !(isinf(qmin) && qmin < 0) || !(isinf(qmax) && qmax > 0);
  return m;
}

__global__ void constraint_kernel(float* s, long long n, float min_value) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) s[i] = fmaxf(s[i], min_value);   // NQ:43 tf.maximum(w, min_value)
}

}  // namespace lq

using namespace lq;

extern "C" int lq_version(void) { return LQ_VERSION; }

extern "C" const char* lq_last_error(void) { return g_err; }

extern "C" int lq_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cuda_fail(cudaGetDevice(&dev), "cudaGetDevice")) return 0;
  if (cuda_fail(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return 0;
  return sms;
}

extern "C" int lq_workspace_bytes(const lq_tensor_desc* descs, int n, int phase, size_t* out_bytes) {
  if (!out_bytes) { set_error("lq_workspace_bytes: null out_bytes"); return LQ_E_NULLPTR; }
  if (n < 0) { set_error("lq_workspace_bytes: n < 0"); return LQ_E_BADSHAPE; }
  if (n && !descs) { set_error("lq_workspace_bytes: null descs"); return LQ_E_NULLPTR; }
  if (phase < LQ_PHASE_FWD || phase > LQ_PHASE_PENALTY_BWD) { set_error("bad phase %d", phase); return LQ_E_BADMODE; }
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    int rc = check_dims(descs[i].outer, descs[i].C, descs[i].inner);
    if (rc) return rc;
    if (descs[i].outer * descs[i].C * descs[i].inner == 0) continue;
    // Partial slots depend on the regime only through the shape, except that the COL
    // regime's slot count is nchunks*W regardless of the vector/scalar path; take the
    // larger of the two possible layouts so any pointer alignment fits.
    const Geom a = make_geom(descs[i].outer, descs[i].C, descs[i].inner, true);
    const Geom b = make_geom(descs[i].outer, descs[i].C, descs[i].inner, false);
    long long slots = a.npartials > b.npartials ? a.npartials : b.npartials;
    if (col_tma_shape_ok(descs[i].outer, descs[i].C, descs[i].inner)) {
      const long long t = make_col_tma_geom(descs[i].outer, descs[i].C, descs[i].inner).npartials;
      if (t > slots) slots = t;
    }
    if (phase == LQ_PHASE_BWD) total += align256((size_t)slots * sizeof(float4));
    else if (phase == LQ_PHASE_PENALTY_FWD || phase == LQ_PHASE_PENALTY_BWD)
      total += align256((size_t)slots * sizeof(float4)) + align256((size_t)descs[i].C * sizeof(float4));
  }
  if (phase == LQ_PHASE_PENALTY_FWD || phase == LQ_PHASE_PENALTY_BWD) total += kPenSlotBytes;   // scalar slot
  *out_bytes = total;
  return LQ_OK;
}

extern "C" int lq_scale_constraint(float* s, int64_t n, float min_value, lq_stream_t stream) {
  if (n < 0) { set_error("lq_scale_constraint: n < 0"); return LQ_E_BADSHAPE; }
  if (n == 0) return LQ_OK;
  if (!s) { set_error("lq_scale_constraint: null s"); return LQ_E_NULLPTR; }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  constraint_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(s, n, min_value);
  return cuda_fail(cudaGetLastError(), "constraint launch");
}
