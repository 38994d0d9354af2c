// lq_dp.cu -- data-parallel plumbing: one NCCL communicator per process, flat fp32
// gradient bucket allreduce over NVLink 5 / NVSwitch.  (The reference is single-device:
// IMAGENETTE/nested_quantization_layer/experiment.py:774-788; this is new work.)
//
// NCCL is resolved with dlopen at first use so that (a) liblq.so loads on machines without
// NCCL (the CPU test tier), and (b) inside a torch process we bind to the libnccl.so.2 that
// torch has already mapped instead of a second copy.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "lq_common.cuh"

namespace lq {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2, ncclAvg = 4 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static char g_nccl_dlerr[256] = "";

static void load_nccl() {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    const char* e = dlerror();                 // read ONCE: the call clears the error
    snprintf(g_nccl_dlerr, sizeof(g_nccl_dlerr), "%s", e ? e : "libnccl.so.2 not found");
    return;
  }
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllGather");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather;
}

static int need_nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) { set_error("NCCL not available: %s", g_nccl_dlerr[0] ? g_nccl_dlerr : "a required symbol is missing"); return LQ_E_NCCL; }
  return LQ_OK;
}

static int nccl_fail(int r, const char* what) {
  if (r == ncclSuccess) return LQ_OK;
  set_error("%s: NCCL error %d (%s)", what, r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return LQ_E_NCCL;
}

__global__ void scale_kernel(float* v, long long n, float scale, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    float4* v4 = reinterpret_cast<float4*>(v);
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += stride) {
      float4 x = v4[i];
      x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
      v4[i] = x;
    }
    const long long t = (n4 << 2) + tid;
    if (t < n) v[t] *= scale;
  } else {
    for (long long i = tid; i < n; i += stride) v[i] *= scale;
  }
}

}  // namespace lq

using namespace lq;

struct lq_dp {
  ncclComm_t comm;
  int rank, world;
  int device;      // the device the communicator was created on: every collective must be issued with it current
};

// A communicator is bound to the device that was current at lq_dp_init; a collective issued from another device
// would enqueue on the wrong GPU's stream.  Refuse instead of guessing.
static int dp_check_device(const lq_dp* dp, const char* who) {
  int dev = -1;
  int rc = cuda_fail(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  if (dev != dp->device) {
    set_error("%s: the communicator belongs to device %d but device %d is current", who, dp->device, dev);
    return LQ_E_BADMODE;
  }
  return LQ_OK;
}

extern "C" int lq_dp_unique_id(void* out_id128) {
  if (!out_id128) { set_error("lq_dp_unique_id: null out"); return LQ_E_NULLPTR; }
  int rc = need_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  return nccl_fail(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(out_id128)), "ncclGetUniqueId");
}

extern "C" int lq_dp_init(const void* id128, int rank, int world, lq_dp_t** out) {
  if (!id128 || !out) { set_error("lq_dp_init: null argument"); return LQ_E_NULLPTR; }
  if (world < 1 || rank < 0 || rank >= world) { set_error("lq_dp_init: bad rank %d / world %d", rank, world); return LQ_E_BADSHAPE; }
  int rc = need_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  if ((rc = nccl_fail(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank"))) return rc;
  int dev = 0;
  if ((rc = cuda_fail(cudaGetDevice(&dev), "cudaGetDevice"))) { g_nccl.CommDestroy(comm); return rc; }
  lq_dp* dp = new lq_dp{comm, rank, world, dev};
  *out = dp;
  return LQ_OK;
}

extern "C" int lq_dp_allreduce(lq_dp_t* dp, float* flat, int64_t n, float scale, lq_stream_t stream) {
  if (!dp) { set_error("lq_dp_allreduce: null communicator"); return LQ_E_NULLPTR; }
  if (int drc = dp_check_device(dp, "lq_dp_allreduce")) return drc;
  if (n < 0) { set_error("lq_dp_allreduce: n < 0"); return LQ_E_BADSHAPE; }
  if (n == 0) return LQ_OK;
  if (!flat) { set_error("lq_dp_allreduce: null buffer"); return LQ_E_NULLPTR; }
  cudaStream_t st = (cudaStream_t)stream;
  // scale == 1/world is the data-parallel mean: NCCL's own average op does it inside the
  // collective (one kernel, NVLS-capable); any other scale is a sum followed by one scaling pass.
  const float mean = 1.0f / (float)dp->world;
  if (scale == mean && dp->world > 1)
    return nccl_fail(g_nccl.AllReduce(flat, flat, (size_t)n, ncclFloat32, ncclAvg, dp->comm, st), "ncclAllReduce(avg)");
  int rc = nccl_fail(g_nccl.AllReduce(flat, flat, (size_t)n, ncclFloat32, ncclSum, dp->comm, st), "ncclAllReduce");
  if (rc) return rc;
  if (scale != 1.0f) {
    const int vec = aligned16(flat) ? 1 : 0;
    long long blocks = ((vec ? n / 4 : n) + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(flat, n, scale, vec);
    if ((rc = cuda_fail(cudaGetLastError(), "scale launch"))) return rc;
  }
  return LQ_OK;
}

extern "C" int lq_dp_allreduce_max(lq_dp_t* dp, float* flat, int64_t n, lq_stream_t stream) {
  if (!dp) { set_error("lq_dp_allreduce_max: null communicator"); return LQ_E_NULLPTR; }
  if (int drc = dp_check_device(dp, "lq_dp_allreduce_max")) return drc;
  if (n < 0) { set_error("lq_dp_allreduce_max: n < 0"); return LQ_E_BADSHAPE; }
  if (n == 0) return LQ_OK;
  if (!flat) { set_error("lq_dp_allreduce_max: null buffer"); return LQ_E_NULLPTR; }
  return nccl_fail(g_nccl.AllReduce(flat, flat, (size_t)n, ncclFloat32, ncclMax, dp->comm, (cudaStream_t)stream), "ncclAllReduce(max)");
}

extern "C" int lq_dp_allgather_f64(lq_dp_t* dp, const double* send, double* recv, int64_t n, lq_stream_t stream) {
  if (!dp) { set_error("lq_dp_allgather_f64: null communicator"); return LQ_E_NULLPTR; }
  if (int drc = dp_check_device(dp, "lq_dp_allgather_f64")) return drc;
  if (n < 0) { set_error("lq_dp_allgather_f64: n < 0"); return LQ_E_BADSHAPE; }
  if (n == 0) return LQ_OK;
  if (!send || !recv) { set_error("lq_dp_allgather_f64: null buffer"); return LQ_E_NULLPTR; }
  return nccl_fail(g_nccl.AllGather(send, recv, (size_t)n, ncclFloat64, dp->comm, (cudaStream_t)stream), "ncclAllGather");
}

extern "C" int lq_dp_destroy(lq_dp_t* dp) {
  if (!dp) return LQ_OK;
  int rc = g_nccl.ok ? nccl_fail(g_nccl.CommDestroy(dp->comm), "ncclCommDestroy") : LQ_OK;
  delete dp;
  return rc;
}
