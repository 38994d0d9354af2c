// lq_dpx.cu -- the data-parallel gradient exchange fused with the threshold backward, over NVLink
// peer memory (the reference is single-device, experiment.py:774-788).
//
// Under data parallelism every rank holds the same parameters and a local gradient bucket
// [dy of the quantised tensors | plain gradients].  The exact-mode step needs, on every rank,
//     dy_mean = mean over ranks of dy            (the all-reduce)
//     ds      = threshold backward of dy_mean    (NQ:62-118, K2)
// A library all-reduce followed by K2 moves the bucket through HBM three times and cannot live inside
// a CUDA graph on this stack.  Here ONE persistent kernel per rank does both:
//   * every rank's bucket and partial workspace live in an IPC-exported arena mapped by all peers;
//   * the step's work items (the multi-tensor K2 grid: ROW units / COL tiles, plus flat ranges for the
//     plain gradients) are dealt round-robin to the ranks.  The owner of an item LOADS its dy from
//     every rank's arena (P2P loads through NVSwitch), averages in rank order, runs the K2 element
//     operator on the averaged values with its local (replicated) P and s, and STORES the averaged
//     dy and the item's {sum, n_below, max|q|} partial into every rank's arena: reduce-scatter,
//     compute and all-gather of one tile in one pass over registers;
//   * two flag barriers (system-scope release / acquire on words in the peers' arenas) bracket the
//     data phase: "every rank's local gradients are complete" and "every rank has finished storing".
//     Epochs live in device memory, so a CUDA-graph replay of the launch advances them by itself.
// After the kernel every rank holds identical averaged gradients and identical partials; the ordinary
// second stage (mt_finalize) then yields bit-identical ds everywhere with no further exchange.
// A rank that waits ~2 s for a peer raises the error flag and gives up: a dead peer cannot hang the GPU.
#include <stdlib.h>
#include <string.h>
#include "lq_fq_ops.cuh"
#include "lq_col_tma.cuh"

namespace lq {

constexpr int kDpxMaxWorld = 8;
constexpr int kDpxGridMax = 148 * 2;          // persistent blocks per rank (two resident per SM)
constexpr int kDpxChannels = 4;               // exchanges that may be in flight at once (one per stream)
constexpr size_t kDpxCtlBytes = 192 << 10;    // control region in front of the user arena

// control region layout (bytes from the arena base), per channel ch at ch * kChanBytes:
//   [0, 37 888)        flags[2 phases][kDpxMaxWorld src][kDpxGridMax] u64   written by the peers
//   [38 912, 41 280)   blk_epoch[kDpxGridMax] u64                           local only
// then, once:  [kOffErr] err (int) ; [kOffOne] one (float 1.0f)             local only
constexpr size_t kChanBytes = 44 << 10, kOffEpoch = 38 << 10;
constexpr size_t kOffErr = kDpxChannels * kChanBytes, kOffOne = kOffErr + 64;
static_assert(2 * kDpxMaxWorld * kDpxGridMax * 8 <= kOffEpoch && kOffEpoch + kDpxGridMax * 8 <= kChanBytes &&
              kOffOne + 4 <= kDpxCtlBytes, "control region layout");

struct DpxCtl {
  long long delta[kDpxMaxWorld];     // byte distance from a local arena address to rank r's mapping of it
  unsigned long long* flags;         // local control region
  unsigned long long* blk_epoch;
  int* err;
  int rank, world;
  unsigned total;                    // work items (global blocks) of this launch
  float inv;                         // 1 / world
};

__device__ __forceinline__ float4 ld_sys4(const void* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys1(const void* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// W = world size known at compile time (all loads of a trip in flight at once); W == 0: run-time world.
template <int W>
struct PeerIO {
  const DpxCtl* c;
  __device__ __forceinline__ int world() const { return W ? W : c->world; }
  __device__ __forceinline__ float4 ld_in2_4(const float4* p) const {
    const char* a = reinterpret_cast<const char*>(p);
    if constexpr (W > 0) {
      float4 v[W];
#pragma unroll
      for (int r = 0; r < W; ++r) v[r] = ld_sys4(a + c->delta[r]);
      float4 t = v[0];
#pragma unroll
      for (int r = 1; r < W; ++r) { t.x = __fadd_rn(t.x, v[r].x); t.y = __fadd_rn(t.y, v[r].y); t.z = __fadd_rn(t.z, v[r].z); t.w = __fadd_rn(t.w, v[r].w); }
      return make_float4(__fmul_rn(t.x, c->inv), __fmul_rn(t.y, c->inv), __fmul_rn(t.z, c->inv), __fmul_rn(t.w, c->inv));
    } else {
      float4 t = ld_sys4(a + c->delta[0]);
      for (int r = 1; r < c->world; ++r) {
        const float4 v = ld_sys4(a + c->delta[r]);
        t.x = __fadd_rn(t.x, v.x); t.y = __fadd_rn(t.y, v.y); t.z = __fadd_rn(t.z, v.z); t.w = __fadd_rn(t.w, v.w);
      }
      return make_float4(__fmul_rn(t.x, c->inv), __fmul_rn(t.y, c->inv), __fmul_rn(t.z, c->inv), __fmul_rn(t.w, c->inv));
    }
  }
  __device__ __forceinline__ float ld_in2_1(const float* p) const {
    const char* a = reinterpret_cast<const char*>(p);
    if constexpr (W > 0) {
      float v[W];
#pragma unroll
      for (int r = 0; r < W; ++r) v[r] = ld_sys1(a + c->delta[r]);
      float t = v[0];
#pragma unroll
      for (int r = 1; r < W; ++r) t = __fadd_rn(t, v[r]);
      return __fmul_rn(t, c->inv);
    } else {
      float t = ld_sys1(a + c->delta[0]);
      for (int r = 1; r < c->world; ++r) t = __fadd_rn(t, ld_sys1(a + c->delta[r]));
      return __fmul_rn(t, c->inv);
    }
  }
  __device__ __forceinline__ void st_out4(float4* p, const float4& v) const {
    char* a = reinterpret_cast<char*>(p);
#pragma unroll
    for (int r = 0; r < (W ? W : kDpxMaxWorld); ++r)
      if (W || r < c->world) *reinterpret_cast<float4*>(a + c->delta[r]) = v;
  }
  __device__ __forceinline__ void st_out1(float* p, float v) const {
    char* a = reinterpret_cast<char*>(p);
#pragma unroll
    for (int r = 0; r < (W ? W : kDpxMaxWorld); ++r)
      if (W || r < c->world) *reinterpret_cast<float*>(a + c->delta[r]) = v;
  }
  __device__ __forceinline__ void st_part(float4* p, const float4& v) const { st_out4(p, v); }
};

// K2 on the quantised tensors; ranges flagged `plain` (BatchNorm / unquantised Dense gradients) are
// only averaged.
struct PeerBwdOp : BwdThresholdOp<false> {
  using Base = BwdThresholdOp<false>;
  __device__ __forceinline__ void elem(float p, float dy, const Recip& c, const Job& j, Acc& acc, float& dx, float& q) const {
    if (j.plain) { dx = dy; return; }
    Base::elem(p, dy, c, j, acc, dx, q);
  }
  __device__ __forceinline__ void quad(const float4& p, const float4& d, const Recip& c, const Job& j, Acc& acc,
                                       float4& dx, float4& q, Defer& df) const {
    if (j.plain) { dx = d; return; }
    Base::quad(p, d, c, j, acc, dx, q, df);
  }
};

__device__ __forceinline__ unsigned long long* dpx_flag(unsigned long long* flags, int phase, int src, unsigned blk) {
  return flags + ((size_t)phase * kDpxMaxWorld + src) * kDpxGridMax + blk;
}

// Block b of every rank meets block b of every other rank.  Returns false after ~2 s without a peer.
__device__ __forceinline__ bool dpx_barrier(const DpxCtl& c, int phase, unsigned long long epoch, bool wait) {
  bool ok = true;
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    unsigned long long* f = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<char*>(dpx_flag(c.flags, phase, c.rank, blockIdx.x)) + c.delta[p]);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
    if (wait) {
      const unsigned long long* mine = dpx_flag(c.flags, phase, p, blockIdx.x);
      unsigned long long seen = 0, t0 = 0, now = 0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 2000000000ull) { ok = false; break; }
        __nanosleep(64);
      }
    }
  }
  return __syncthreads_and(ok ? 1 : 0) != 0;
}

template <int W>
__global__ void __launch_bounds__(kThreads, 2) dpx_kernel(const __grid_constant__ JobTable tab, const PeerBwdOp op,
                                                          const __grid_constant__ DpxCtl c) {
  __shared__ Acc smem[kThreads];
  __shared__ float2 queue[PeerBwdOp::kDeferSlots * kThreads];
  __shared__ unsigned long long s_epoch;
  cudaGridDependencySynchronize();            // the local gradients are produced by the kernels before this one
  if (threadIdx.x == 0) s_epoch = c.blk_epoch[blockIdx.x] + 1ull;
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  bool ok = dpx_barrier(c, 0, epoch, true);   // every rank's local gradients are complete
  if (ok) {
    PeerIO<W> io{&c};
    const unsigned world = (unsigned)c.world;
    for (unsigned li = blockIdx.x;; li += gridDim.x) {
      const unsigned long long gb64 = (unsigned long long)li * world + (unsigned)c.rank;
      if (gb64 >= c.total) break;
      const unsigned gb = (unsigned)gb64;
      int lo = 0, hi = tab.n - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab.job[mid].block_begin <= gb) lo = mid; else hi = mid - 1;
      }
      const Job& j = tab.job[lo];
      const unsigned blk = gb - j.block_begin;
      if (j.regime == REG_ROW) {
        if (j.vec) run_row<PeerBwdOp, false, true, PeerIO<W>>(j, blk, op, queue, io);
        else run_row<PeerBwdOp, false, false, PeerIO<W>>(j, blk, op, queue, io);
      } else {
        run_col1<PeerBwdOp, false, PeerIO<W>>(j, blk, op, smem, queue, io);
      }
      __syncthreads();                        // smem / queue are reused by the next item
    }
  } else if (threadIdx.x == 0) {
    atomicExch(c.err, 1);
  }
  __threadfence_system();                     // my stores into the peers' arenas before my flag
  __syncthreads();
  const bool ok2 = dpx_barrier(c, 1, epoch, ok);   // every rank has finished storing into my arena
  if (!ok2 && ok && threadIdx.x == 0) atomicExch(c.err, 1);
  if (threadIdx.x == 0) c.blk_epoch[blockIdx.x] = epoch;
}

// ------------------------------------------------------------------ flat mean all-reduce on the copy engine
// The exchange that runs WHILE the backward is still computing must not take SMs away from the convolutions:
// an NVLink round trip is several microseconds, so a kernel that keeps its bytes in flight in registers needs
// hundreds of resident blocks to fill the links (measured: ~2 GB/s per 256-thread block).  Here the bytes in
// flight live in shared memory instead: a block owns a ring of stages, one thread issues bulk asynchronous
// copies (cp.async.bulk, the TMA engine) of one 8 KB tile from EVERY rank's arena into a stage, four warps add
// the W tiles in rank order out of shared memory, and the mean goes back to every rank by bulk stores straight
// from shared memory.  ~190 KB in flight per block at any world size: a couple of dozen blocks fill the links.
// Tile t of the range belongs to rank t % W (reduce-scatter + all-gather of one tile in one pass); the same
// two flag barriers as above bracket the data phase.
constexpr int kArTile = 2048;                       // floats per tile (8 KB)
constexpr int kArConsumers = 128;                   // 4 consumer warps; warp 0 is the producer
constexpr int kArThreads = 32 + kArConsumers;
constexpr int kArSmemBudget = 216 << 10;
constexpr int kArMaxStages = 8;

struct ArJob { float* flat; long long n; int stages; };

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// (mbar_arrive, bulk_wait_read1: lq_col_tma.cuh)

__global__ void __launch_bounds__(kArThreads, 1) dpx_allreduce_kernel(const __grid_constant__ ArJob job,
                                                                      const __grid_constant__ DpxCtl c) {
  extern __shared__ __align__(128) unsigned char ar_smem[];
  __shared__ unsigned long long full[kArMaxStages], empty[kArMaxStages];
  __shared__ unsigned long long s_epoch;
  const int W = c.world, S = job.stages;
  const int t = threadIdx.x;
  float* stage0 = reinterpret_cast<float*>(ar_smem);
  const size_t stage_floats = (size_t)(W + 1) * kArTile;         // W input tiles, then the output tile

  cudaGridDependencySynchronize();
  if (t == 0) {
    s_epoch = c.blk_epoch[blockIdx.x] + 1ull;
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const bool ok = dpx_barrier(c, 0, epoch, true);                 // every rank's local gradients are complete
  if (ok) {
    const long long ntiles = (job.n + kArTile - 1) / kArTile;
    // my tiles: t = rank + W * m, m = blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long n_mine_all = (ntiles - c.rank + W - 1) / W;                    // tiles of this rank (ntiles > rank)
    const long long n_my = (n_mine_all > (long long)blockIdx.x)
        ? (n_mine_all - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;               // ... of this block
    auto tile_of = [&](long long i) { return (long long)c.rank + (long long)W * ((long long)blockIdx.x + i * gridDim.x); };
    auto tile_floats = [&](long long tl) { const long long r = job.n - tl * kArTile; return (int)(r < kArTile ? r : kArTile); };
    if (t == 0) {
      // producer: W bulk loads per tile, as far ahead as the ring allows
      asm volatile("fence.proxy.async;" ::: "memory");
      for (long long i = 0; i < n_my; ++i) {
        const int st = (int)(i % S);
        if (i >= S) mbar_wait(&empty[st], (unsigned)(((i / S) - 1) & 1));
        const long long tl = tile_of(i);
        const unsigned bytes = (unsigned)tile_floats(tl) * 4u;
        mbar_expect_tx(&full[st], bytes * (unsigned)W);
        float* dst = stage0 + (size_t)st * stage_floats;
        const char* src = reinterpret_cast<const char*>(job.flat + tl * kArTile);
        for (int r = 0; r < W; ++r) bulk_load(dst + (size_t)r * kArTile, src + c.delta[r], bytes, &full[st]);
      }
    } else if (t >= 32) {
      const int ct = t - 32;
      for (long long i = 0; i < n_my; ++i) {
        const int st = (int)(i % S);
        mbar_wait(&full[st], (unsigned)((i / S) & 1));
        const long long tl = tile_of(i);
        const int nf = tile_floats(tl);
        float* in = stage0 + (size_t)st * stage_floats;
        float* out = in + (size_t)W * kArTile;
#pragma unroll
        for (int u = 0; u < kArTile / 4 / kArConsumers; ++u) {
          const int v = ct + u * kArConsumers;                     // float4 index inside the tile
          if (4 * v < nf) {
            float4 a = reinterpret_cast<const float4*>(in)[v];
            for (int r = 1; r < W; ++r) {
              const float4 b = reinterpret_cast<const float4*>(in + (size_t)r * kArTile)[v];
              a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
            }
            reinterpret_cast<float4*>(out)[v] = make_float4(__fmul_rn(a.x, c.inv), __fmul_rn(a.y, c.inv),
                                                            __fmul_rn(a.z, c.inv), __fmul_rn(a.w, c.inv));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // my smem writes, before the bulk stores read them
        asm volatile("bar.sync 1, %0;" ::"n"(kArConsumers) : "memory");
        if (ct == 0) {
          char* gdst = reinterpret_cast<char*>(job.flat + tl * kArTile);
          for (int r = 0; r < W; ++r) bulk_store(gdst + c.delta[r], out, (unsigned)nf * 4u);
          bulk_commit();
          if (i >= 1) { bulk_wait_read1(); mbar_arrive(&empty[(int)((i - 1) % S)]); }   // tile i-1's stores have read their stage
        }
      }
      if (ct == 0) {
        bulk_wait0();                                             // every store of mine is complete
        asm volatile("fence.proxy.async;" ::: "memory");
      }
    }
  } else if (t == 0) {
    atomicExch(c.err, 1);
  }
  __threadfence_system();
  __syncthreads();
  const bool ok2 = dpx_barrier(c, 1, epoch, ok);                  // every rank has finished storing into my arena
  if (!ok2 && ok && t == 0) atomicExch(c.err, 1);
  if (t == 0) c.blk_epoch[blockIdx.x] = epoch;
}

}  // namespace lq

using namespace lq;

struct lq_dpx {
  int rank, world, device;
  void* base;                 // control region + user arena
  size_t arena_bytes;         // user part
  void* peer[kDpxMaxWorld];   // mapped bases (peer[rank] == base)
  bool opened[kDpxMaxWorld];
  bool connected;
};

extern "C" int lq_dpx_create(int rank, int world, size_t arena_bytes, lq_dpx_t** out, void* out_handle64, void** out_arena) {
  if (!out || !out_handle64 || !out_arena) { set_error("lq_dpx_create: null argument"); return LQ_E_NULLPTR; }
  if (world < 1 || world > kDpxMaxWorld || rank < 0 || rank >= world || arena_bytes == 0) {
    set_error("lq_dpx_create: bad rank %d / world %d (max %d) / arena size", rank, world, kDpxMaxWorld); return LQ_E_BADSHAPE;
  }
  lq_dpx* x = new lq_dpx();
  x->rank = rank; x->world = world; x->arena_bytes = align256(arena_bytes);
  int rc = cuda_fail(cudaGetDevice(&x->device), "lq_dpx_create: cudaGetDevice");
  const size_t total = kDpxCtlBytes + x->arena_bytes;
  if (!rc) rc = cuda_fail(cudaMalloc(&x->base, total), "lq_dpx_create: cudaMalloc");
  if (rc) { delete x; return rc; }
  if (!rc) rc = cuda_fail(cudaMemset(x->base, 0, total), "lq_dpx_create: cudaMemset");
  const float one = 1.0f;
  if (!rc) rc = cuda_fail(cudaMemcpy(static_cast<char*>(x->base) + kOffOne, &one, 4, cudaMemcpyHostToDevice), "lq_dpx_create: cudaMemcpy");
  cudaIpcMemHandle_t h;
  if (!rc) rc = cuda_fail(cudaIpcGetMemHandle(&h, x->base), "cudaIpcGetMemHandle");
  if (rc) { cudaFree(x->base); delete x; return rc; }
  memcpy(out_handle64, &h, 64);
  for (int i = 0; i < kDpxMaxWorld; ++i) { x->peer[i] = nullptr; x->opened[i] = false; }
  x->peer[rank] = x->base;
  x->connected = (world == 1);
  *out = x;
  *out_arena = static_cast<char*>(x->base) + kDpxCtlBytes;
  return LQ_OK;
}

extern "C" int lq_dpx_connect(lq_dpx_t* x, const void* handles64_all) {
  if (!x || !handles64_all) { set_error("lq_dpx_connect: null argument"); return LQ_E_NULLPTR; }
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank || x->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles64_all) + 64 * r, 64);
    int rc = cuda_fail(cudaIpcOpenMemHandle(&x->peer[r], h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    if (rc) return rc;
    x->opened[r] = true;
  }
  x->connected = true;
  return LQ_OK;
}

extern "C" int lq_dpx_connect_ptrs(lq_dpx_t* x, const void* const* arenas) {
  if (!x || !arenas) { set_error("lq_dpx_connect_ptrs: null argument"); return LQ_E_NULLPTR; }
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank) continue;
    if (!arenas[r]) { set_error("lq_dpx_connect_ptrs: null arena for rank %d", r); return LQ_E_NULLPTR; }
    x->peer[r] = const_cast<char*>(static_cast<const char*>(arenas[r])) - kDpxCtlBytes;
  }
  x->connected = true;
  return LQ_OK;
}

static bool in_arena(const lq_dpx* x, const void* p, size_t bytes) {
  const char* a = static_cast<const char*>(x->base) + kDpxCtlBytes;
  const char* q = static_cast<const char*>(p);
  return q >= a && q + bytes <= a + x->arena_bytes;
}

template <int W>
static int dpx_launch(unsigned grid, const JobTable& tab, const PeerBwdOp& op, const DpxCtl& c, cudaStream_t st) {
  dpx_kernel<W><<<dim3(grid), dim3(kThreads), 0, st>>>(tab, op, c);
  return cuda_fail(cudaGetLastError(), "dpx launch");
}

extern "C" int lq_dpx_exchange(lq_dpx_t* x, const lq_tensor_desc* descs, int n, float* plain, int64_t n_plain,
                               float penalty_threshold, int scale_grad_mode, int channel, int max_blocks,
                               void* ws, size_t ws_bytes, lq_stream_t stream) {
  if (!x) { set_error("lq_dpx_exchange: null handle"); return LQ_E_NULLPTR; }
  if (channel < 0 || channel >= kDpxChannels || max_blocks < 0) {
    set_error("lq_dpx_exchange: channel must be in [0, %d), max_blocks >= 0", kDpxChannels); return LQ_E_BADMODE;
  }
  if (!x->connected) { set_error("lq_dpx_exchange: lq_dpx_connect has not been called"); return LQ_E_BADMODE; }
  if (n < 0 || n_plain < 0) { set_error("lq_dpx_exchange: negative count"); return LQ_E_BADSHAPE; }
  if ((n && !descs) || (n_plain && !plain)) { set_error("lq_dpx_exchange: null descs / plain"); return LQ_E_NULLPTR; }
  if (scale_grad_mode != LQ_SGRAD_NONE && scale_grad_mode != LQ_SGRAD_THRESHOLD) {
    set_error("lq_dpx_exchange: bad scale_grad_mode %d", scale_grad_mode); return LQ_E_BADMODE;
  }
  if (n + 1 > kMaxMT) { set_error("lq_dpx_exchange: at most %d tensors per call", kMaxMT - 1); return LQ_E_BADSHAPE; }
  const bool k2 = scale_grad_mode == LQ_SGRAD_THRESHOLD;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  QMode qm = make_qmode(LQ_ROUND_FLOOR, -INFINITY, INFINITY, &rc);
  if (rc) return rc;
  char* ctl = static_cast<char*>(x->base);
  const float* one = reinterpret_cast<const float*>(ctl + kOffOne);
  JobTable tab{};
  FinTable fin{};
  unsigned total = 0, groups = 0;
  size_t off = 0;
  auto add_plain = [&](float* p, long long cnt) -> int {
    if (cnt == 0) return LQ_OK;
    if (!in_arena(x, p, (size_t)cnt * 4)) { set_error("lq_dpx_exchange: a gradient range lies outside the arena"); return LQ_E_BADSHAPE; }
    Geom g = make_geom(1, 1, cnt, aligned16(p));
    int r = check_blocks(g);
    if (r) return r;
    Job j = make_job(g, p, one, p, nullptr, p, nullptr, 0, nullptr);
    j.write_dx = 1; j.plain = 1; j.block_begin = total;
    if ((unsigned long long)total + j.nblocks > 0x7fffffffull) { set_error("lq_dpx_exchange: too many blocks"); return LQ_E_BADSHAPE; }
    total += j.nblocks;
    tab.job[tab.n++] = j;
    return LQ_OK;
  };
  for (int i = 0; i < n; ++i)
    if ((rc = check_dims(descs[i].outer, descs[i].C, descs[i].inner))) return rc;
  // (make_geom's own row chunks here, not the shared grid of lq_mt_fq_bwd: this kernel waits on NVLink, and the many short
  // blocks are what keeps bytes in flight -- with the shared grid the exchange alone took 0.34 ms instead of 0.22 at N = 8.
  // The two paths therefore add a group's partials in different orders: same ds within rounding, not the same bits.)
  for (int i = 0; i < n; ++i) {
    const lq_tensor_desc& d = descs[i];
    const long long cnt = d.outer * d.C * d.inner;
    if (cnt == 0) continue;
    if (!d.dy) { set_error("lq_dpx_exchange: desc %d has null dy", i); return LQ_E_NULLPTR; }
    if (d.dX && d.dX != d.dy) { set_error("lq_dpx_exchange: desc %d: dX must be NULL or dy (the mean is written in place)", i); return LQ_E_BADMODE; }
    float* dy = const_cast<float*>(d.dy);
    if (!k2) { if ((rc = add_plain(dy, cnt))) return rc; continue; }
    if (!d.P || !d.s || !d.ds) { set_error("lq_dpx_exchange: desc %d has null P/s/ds", i); return LQ_E_NULLPTR; }
    if (!in_arena(x, dy, (size_t)cnt * 4)) { set_error("lq_dpx_exchange: desc %d: dy lies outside the arena", i); return LQ_E_BADSHAPE; }
    Geom g = make_geom(d.outer, d.C, d.inner, aligned16(d.P) && aligned16(dy) && d.inner >= kRowMinInner);
    if ((rc = check_blocks(g))) return rc;
    const size_t need = align256((size_t)g.npartials * sizeof(float4));
    if (!ws || !aligned16(ws) || off + need > ws_bytes || !in_arena(x, static_cast<char*>(ws) + off, need)) {
      set_error("lq_dpx_exchange: workspace too small or outside the arena (need > %zu B, got %zu)", off + need, ws_bytes);
      return LQ_E_WORKSPACE;
    }
    float4* part = reinterpret_cast<float4*>(static_cast<char*>(ws) + off);
    off += need;
    Job j = make_job(g, d.P, d.s, dy, nullptr, dy, nullptr, 0, part);
    j.write_dx = 1; j.block_begin = total;
    if ((unsigned long long)total + j.nblocks > 0x7fffffffull) { set_error("lq_dpx_exchange: too many blocks"); return LQ_E_BADSHAPE; }
    total += j.nblocks;
    tab.job[tab.n++] = j;
    FinJob f = make_fin(g, part, d.ds, d.s, penalty_threshold);
    f.group_begin = groups;
    groups += (unsigned)d.C;
    fin.job[fin.n++] = f;
  }
  if ((rc = add_plain(plain, n_plain))) return rc;
  if (total) {
    DpxCtl c{};
    for (int r = 0; r < x->world; ++r) c.delta[r] = static_cast<char*>(x->peer[r]) - ctl;
    c.flags = reinterpret_cast<unsigned long long*>(ctl + (size_t)channel * kChanBytes);
    c.blk_epoch = reinterpret_cast<unsigned long long*>(ctl + (size_t)channel * kChanBytes + kOffEpoch);
    c.err = reinterpret_cast<int*>(ctl + kOffErr);
    c.rank = x->rank; c.world = x->world; c.total = total;
    c.inv = 1.0f / (float)x->world;
    const unsigned mine = (total + (unsigned)x->world - 1) / (unsigned)x->world;
    unsigned cap = (max_blocks > 0 && max_blocks < kDpxGridMax) ? (unsigned)max_blocks : (unsigned)kDpxGridMax;
    if (const char* e = getenv("LQ_DPX_GRID")) {          // tuning knob (tools/dpx_probe.py); must agree on every rank
      const long v = strtol(e, nullptr, 10);
      if (v >= 1 && v <= (long)cap) cap = (unsigned)v;
    }
    const unsigned grid = mine < cap ? (mine ? mine : 1u) : cap;
    PeerBwdOp op;
    static_cast<BwdThresholdOp<false>&>(op) = make_bwd_op<false>(qm, penalty_threshold);
    switch (x->world) {
      case 1: rc = dpx_launch<1>(grid, tab, op, c, st); break;
      case 2: rc = dpx_launch<2>(grid, tab, op, c, st); break;
      case 4: rc = dpx_launch<4>(grid, tab, op, c, st); break;
      case 8: rc = dpx_launch<8>(grid, tab, op, c, st); break;
      default: rc = dpx_launch<0>(grid, tab, op, c, st); break;
    }
    if (rc) return rc;
  }
  if (k2) return mt_finalize_launch(fin, groups, st);
  // CL:61-62 / PI:62-64: the operator's scale gradient is zeros
  for (int i = 0; i < n; ++i)
    if (descs[i].ds && descs[i].C > 0)
      if ((rc = cuda_fail(cudaMemsetAsync(descs[i].ds, 0, (size_t)descs[i].C * 4, st), "ds = 0"))) return rc;
  return LQ_OK;
}

static DpxCtl make_ctl(const lq_dpx* x, int channel) {
  char* ctl = static_cast<char*>(x->base);
  DpxCtl c{};
  for (int r = 0; r < x->world; ++r) c.delta[r] = static_cast<char*>(x->peer[r]) - ctl;
  c.flags = reinterpret_cast<unsigned long long*>(ctl + (size_t)channel * kChanBytes);
  c.blk_epoch = reinterpret_cast<unsigned long long*>(ctl + (size_t)channel * kChanBytes + kOffEpoch);
  c.err = reinterpret_cast<int*>(ctl + kOffErr);
  c.rank = x->rank; c.world = x->world;
  c.inv = 1.0f / (float)x->world;
  return c;
}

extern "C" int lq_dpx_allreduce(lq_dpx_t* x, float* flat, int64_t n, int channel, int max_blocks, lq_stream_t stream) {
  if (!x) { set_error("lq_dpx_allreduce: null handle"); return LQ_E_NULLPTR; }
  if (!x->connected) { set_error("lq_dpx_allreduce: lq_dpx_connect has not been called"); return LQ_E_BADMODE; }
  if (channel < 0 || channel >= kDpxChannels || max_blocks < 0) {
    set_error("lq_dpx_allreduce: channel must be in [0, %d), max_blocks >= 0", kDpxChannels); return LQ_E_BADMODE;
  }
  if (n < 0) { set_error("lq_dpx_allreduce: n < 0"); return LQ_E_BADSHAPE; }
  if (n == 0) return LQ_OK;
  if (!flat) { set_error("lq_dpx_allreduce: null buffer"); return LQ_E_NULLPTR; }
  if (!aligned16(flat) || (n & 3)) { set_error("lq_dpx_allreduce: the range must be 16-byte aligned and a multiple of 4 floats"); return LQ_E_BADSHAPE; }
  if (!in_arena(x, flat, (size_t)n * 4)) { set_error("lq_dpx_allreduce: the range lies outside the arena"); return LQ_E_BADSHAPE; }
  ArJob job{};
  job.flat = flat; job.n = n;
  const size_t stage_bytes = (size_t)(x->world + 1) * kArTile * 4;
  job.stages = (int)(kArSmemBudget / stage_bytes);
  if (job.stages > kArMaxStages) job.stages = kArMaxStages;
  if (job.stages < 2) { set_error("lq_dpx_allreduce: world too large for the shared-memory ring"); return LQ_E_BADSHAPE; }
  const size_t smem = (size_t)job.stages * stage_bytes;
  DpxCtl c = make_ctl(x, channel);
  const long long ntiles = (n + kArTile - 1) / kArTile;
  const long long per_rank = (ntiles + x->world - 1) / x->world;
  long long grid = max_blocks > 0 ? max_blocks : 32;
  if (grid > kDpxGridMax) grid = kDpxGridMax;
  if (grid > per_rank) grid = per_rank;
  if (grid < 1) grid = 1;
  cudaError_t e = cudaFuncSetAttribute(dpx_allreduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "dpx_allreduce smem attribute");
  dpx_allreduce_kernel<<<dim3((unsigned)grid), dim3(kArThreads), smem, (cudaStream_t)stream>>>(job, c);
  return cuda_fail(cudaGetLastError(), "dpx_allreduce launch");
}

extern "C" int lq_dpx_error(lq_dpx_t* x, int* out_flag) {
  if (!x || !out_flag) { set_error("lq_dpx_error: null argument"); return LQ_E_NULLPTR; }
  int rc = cuda_fail(cudaDeviceSynchronize(), "lq_dpx_error: cudaDeviceSynchronize");   // non-blocking streams included
  if (rc) return rc;
  return cuda_fail(cudaMemcpy(out_flag, static_cast<char*>(x->base) + kOffErr, sizeof(int), cudaMemcpyDeviceToHost),
                   "lq_dpx_error: cudaMemcpy");
}

extern "C" int lq_dpx_destroy(lq_dpx_t* x) {
  if (!x) return LQ_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (x->opened[r]) cudaIpcCloseMemHandle(x->peer[r]);
  cudaFree(x->base);
  delete x;
  return LQ_OK;
}
