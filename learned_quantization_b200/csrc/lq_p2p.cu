// lq_p2p.cu -- the sharded threshold backward's exchange fused with its finish, over NVLink peer
// memory.  After the per-shard second stage has produced {sum, n_below, max|q|} per
// group, ONE kernel per rank
//   1. stores its statistics straight into every peer's receive slot (P2P stores through NVSwitch;
//      24 bytes per group, coalesced),
//   2. publishes a per-block epoch flag on every peer (system-scope release),
//   3. waits for the peers' flags for the same epoch (system-scope acquire), and
//   4. folds the world's statistics in RANK ORDER out of its own memory and applies NQ:103-116,
// so every rank ends with the same bits, with no NCCL call and no second launch.  Slots and flags
// are double-buffered by epoch parity: a rank can be at most one step ahead of a peer that is still
// reading.  The receive buffer is the library's own cudaMalloc allocation (IPC-exportable; the
// caller's allocator may not be); handles are exchanged by the host program.
//
// Every rank must be its own process on its own GPU (the kernels of different ranks wait on one
// another).  A rank that waits longer than ~2 s gives up, writes NaN into ds and raises the
// object's error flag, which the next host call reports: a crashed peer cannot hang the GPU.
#include <string.h>
#include "lq_common.cuh"

namespace lq {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PBlock = 128;        // groups per block

struct P2PView {
  double* slots[kP2PMaxWorld];                    // base of each rank's receive buffer (peer-mapped)
  unsigned long long* flags[kP2PMaxWorld];
  int rank, world;
  long long max_c, nblocks_max;
};

// receive layout per rank:  slots[parity][src_rank][max_c][3] doubles ; flags[parity][src_rank][nblocks_max]
__device__ __forceinline__ double* slot_ptr(double* base, const P2PView& v, int parity, int src, long long g) {
  return base + (((long long)parity * v.world + src) * v.max_c + g) * 3;
}
__device__ __forceinline__ unsigned long long* flag_ptr(unsigned long long* base, const P2PView& v, int parity, int src, int blk) {
  return base + ((long long)parity * v.world + src) * v.nblocks_max + blk;
}

__global__ void __launch_bounds__(kP2PBlock) finish_p2p_kernel(const __grid_constant__ P2PView v, const double* local_stats,
                                                               long long C, unsigned long long epoch, float c0,
                                                               double n_per_group, float* ds, int* err) {
  cudaGridDependencySynchronize();
  const int parity = (int)(epoch & 1ull);
  const long long g = (long long)blockIdx.x * kP2PBlock + threadIdx.x;
  // 1. push my statistics into every rank's slot for me (my own included)
  if (g < C) {
    const double a = local_stats[3 * g], b = local_stats[3 * g + 1], c = local_stats[3 * g + 2];
    for (int p = 0; p < v.world; ++p) {
      double* dst = slot_ptr(v.slots[p], v, parity, v.rank, g);
      dst[0] = a; dst[1] = b; dst[2] = c;
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait
  bool ok = true;
  if (threadIdx.x < v.world) {
    const int p = threadIdx.x;
    unsigned long long* f = flag_ptr(v.flags[p], v, parity, v.rank, blockIdx.x);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    const unsigned long long* mine = flag_ptr(v.flags[v.rank], v, parity, p, blockIdx.x);
    unsigned long long seen = 0, t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
      if (seen >= epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 2000000000ull) { ok = false; break; }
      __nanosleep(100);
    }
  }
  const int all_ok = __syncthreads_and(ok ? 1 : 0);
  if (!all_ok) {
    if (threadIdx.x == 0) atomicExch(err, 1);
    if (g < C) ds[g] = __int_as_float(0x7fc00000);
    return;
  }
  // 4. fold in rank order out of my own memory
  if (g < C) {
    double sum = 0.0, cnt = 0.0, mx = 0.0;
    for (int r = 0; r < v.world; ++r) {
      const double* st = slot_ptr(v.slots[v.rank], v, parity, r, g);
      // L2 is the coherence point for the peers' stores: bypass this SM's L1
      sum += __ldcg(st); cnt += __ldcg(st + 1); mx = fmax(mx, __ldcg(st + 2));
    }
    const float mean = (cnt == 0.0) ? c0 : (float)(sum / n_per_group);
    ds[g] = __fmul_rn(mean, (float)mx);
  }
}

}  // namespace lq

using namespace lq;

struct lq_p2p {
  int rank, world;
  long long max_c, nblocks_max;
  void* buf;                 // my receive buffer: slots then flags, then the error word
  size_t slots_bytes, flags_bytes;
  void* peer[kP2PMaxWorld];  // mapped bases (peer[rank] == buf)
  bool opened[kP2PMaxWorld];
  unsigned long long epoch;
  bool connected;
};

extern "C" int lq_p2p_create(int rank, int world, int64_t max_groups, lq_p2p_t** out, void* out_handle64) {
  if (!out || !out_handle64) { set_error("lq_p2p_create: null argument"); return LQ_E_NULLPTR; }
  if (world < 1 || world > kP2PMaxWorld || rank < 0 || rank >= world || max_groups < 1) {
    set_error("lq_p2p_create: bad rank %d / world %d (max %d) / max_groups", rank, world, kP2PMaxWorld); return LQ_E_BADSHAPE;
  }
  lq_p2p* p = new lq_p2p();
  p->rank = rank; p->world = world; p->max_c = max_groups;
  p->nblocks_max = (max_groups + kP2PBlock - 1) / kP2PBlock;
  p->slots_bytes = align256((size_t)2 * world * max_groups * 3 * sizeof(double));
  p->flags_bytes = align256((size_t)2 * world * p->nblocks_max * sizeof(unsigned long long));
  const size_t total = p->slots_bytes + p->flags_bytes + 256;
  int rc = cuda_fail(cudaMalloc(&p->buf, total), "lq_p2p_create: cudaMalloc");
  if (rc) { delete p; return rc; }
  if ((rc = cuda_fail(cudaMemset(p->buf, 0, total), "lq_p2p_create: cudaMemset"))) { cudaFree(p->buf); delete p; return rc; }
  cudaIpcMemHandle_t h;
  if ((rc = cuda_fail(cudaIpcGetMemHandle(&h, p->buf), "cudaIpcGetMemHandle"))) { cudaFree(p->buf); delete p; return rc; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out_handle64, &h, 64);
  for (int i = 0; i < kP2PMaxWorld; ++i) { p->peer[i] = nullptr; p->opened[i] = false; }
  p->peer[rank] = p->buf;
  p->epoch = 0;
  p->connected = (world == 1);
  *out = p;
  return LQ_OK;
}

extern "C" int lq_p2p_connect(lq_p2p_t* p, const void* handles64_all) {
  if (!p || !handles64_all) { set_error("lq_p2p_connect: null argument"); return LQ_E_NULLPTR; }
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank || p->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles64_all) + 64 * r, 64);
    int rc = cuda_fail(cudaIpcOpenMemHandle(&p->peer[r], h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    if (rc) return rc;
    p->opened[r] = true;
  }
  p->connected = true;
  return LQ_OK;
}

extern "C" int lq_p2p_connect_ptrs(lq_p2p_t* p, const void* const* bufs) {
  if (!p || !bufs) { set_error("lq_p2p_connect_ptrs: null argument"); return LQ_E_NULLPTR; }
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    if (!bufs[r]) { set_error("lq_p2p_connect_ptrs: null buffer for rank %d", r); return LQ_E_NULLPTR; }
    p->peer[r] = const_cast<void*>(bufs[r]);
  }
  p->connected = true;
  return LQ_OK;
}

extern "C" int lq_p2p_buffer(lq_p2p_t* p, void** out) {
  if (!p || !out) { set_error("lq_p2p_buffer: null argument"); return LQ_E_NULLPTR; }
  *out = p->buf;
  return LQ_OK;
}

extern "C" int lq_fq_bwd_threshold_finish_p2p(lq_p2p_t* p, const double* local_stats, int64_t C, float penalty_threshold,
                                              double n_per_group, float* ds, lq_stream_t stream) {
  if (!p || !local_stats || !ds) { set_error("lq_fq_bwd_threshold_finish_p2p: null argument"); return LQ_E_NULLPTR; }
  if (!p->connected) { set_error("lq_fq_bwd_threshold_finish_p2p: lq_p2p_connect has not been called"); return LQ_E_BADMODE; }
  if (C < 1 || C > p->max_c || !(n_per_group > 0.0)) { set_error("lq_fq_bwd_threshold_finish_p2p: bad C / n_per_group"); return LQ_E_BADSHAPE; }
  int* err = reinterpret_cast<int*>(static_cast<char*>(p->buf) + p->slots_bytes + p->flags_bytes);
  P2PView v{};
  for (int r = 0; r < p->world; ++r) {
    v.slots[r] = static_cast<double*>(p->peer[r]);
    v.flags[r] = reinterpret_cast<unsigned long long*>(static_cast<char*>(p->peer[r]) + p->slots_bytes);
  }
  v.rank = p->rank; v.world = p->world; v.max_c = p->max_c; v.nblocks_max = p->nblocks_max;
  p->epoch += 1;
  const float c0 = -fabsf(tanhf(penalty_threshold));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((C + kP2PBlock - 1) / kP2PBlock)); cfg.blockDim = dim3(kP2PBlock); cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cuda_fail(cudaLaunchKernelEx(&cfg, finish_p2p_kernel, v, local_stats, (long long)C, p->epoch, c0, n_per_group, ds, err),
                   "finish_p2p launch");
}

extern "C" int lq_p2p_error(lq_p2p_t* p, int* out_flag) {
  if (!p || !out_flag) { set_error("lq_p2p_error: null argument"); return LQ_E_NULLPTR; }
  const int* err = reinterpret_cast<const int*>(static_cast<char*>(p->buf) + p->slots_bytes + p->flags_bytes);
  int rc = cuda_fail(cudaDeviceSynchronize(), "lq_p2p_error: cudaDeviceSynchronize");   // non-blocking streams included
  if (rc) return rc;
  return cuda_fail(cudaMemcpy(out_flag, err, sizeof(int), cudaMemcpyDeviceToHost), "lq_p2p_error: cudaMemcpy");
}

extern "C" int lq_p2p_destroy(lq_p2p_t* p) {
  if (!p) return LQ_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < p->world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->peer[r]);
  cudaFree(p->buf);
  delete p;
  return LQ_OK;
}
