// lq_col_tma.cuh -- COL-regime threshold backward with TMA-staged tiles (sm_100a).
//
// Interleaved layouts ([outer, W = C*inner] with inner < 128: dense column-wise scales,
// channelwise scales on kernels with few output channels, [N/C, C, 1] microbench tensors) put
// consecutive elements in DIFFERENT groups.  Walking down a column from registers costs a 64-bit
// address computation per element (a run-time row stride), walking four columns per thread costs
// per-thread state for four groups: both versions issued ~50 % more instructions than ROW and ran at
// 4.2 TB/s (profiles/README.md).  Here the copy engine does the strided part:
//   * a block owns a tile column of TC (32/64/128) columns and a chunk of rows; a ninth, producer warp issues
//     2-D tensor-map TMA loads (cp.async.bulk.tensor.2d, one [rows x TC] box = 8 KB per array) of P
//     and dy into a 3-stage shared memory ring, completion counted on a "full" mbarrier per stage; the eight
//     consumer warps hand a stage back through its "empty" mbarrier (one arrival per warp), so no block-wide
//     barrier sits in the tile loop and the warps drift apart by up to the ring's depth (round 2: 133 -> 130 us
//     at 64 Mi; with the barrier, thread 0 doubling as the producer, the kernel was issue-limited at 35 % occupancy).
//     Column blocks of the same rows are neighbours in launch order (what is in flight covers whole rows:
//     140 -> 136 us), and the kernel is launched with programmatic stream serialization like the others.
//     (A first version issued one 1-D bulk copy per row: 1.5 M copies of 256-512 bytes per launch
//     were bound by the copy engine's request rate -- 238 us where the register version took 189.)
//   * dX = dy (NQ:118) never touches a register: the dy tile is written back to dX by a TMA tensor
//     store straight from shared memory; out-of-range rows / columns are clipped by the hardware;
//   * every thread then walks 8 rows of ONE column out of shared memory (immediate offsets, bank
//     conflict free), i.e. serves a single group exactly like a ROW thread: the operator's 4-element
//     quad(), its rare-path queue and flush() are reused unchanged.
// Partials: one float4 per (row chunk, column), same layout as the register COL traversal, so the
// second stage is shared.
#pragma once
#include <cuda.h>
#include "lq_traverse.cuh"

namespace lq {

#ifndef LQ_TMA_STAGES
#define LQ_TMA_STAGES 3
#endif
constexpr int kTmaStages = LQ_TMA_STAGES;
#ifndef LQ_TMA_BPS
#define LQ_TMA_BPS 8
#endif
#ifndef LQ_COL_ORDER
#define LQ_COL_ORDER 1
#endif
// LQ_TMA_QSLOTS: entries of the per-thread rare-path queue (8: drained once per tile; 4: once per four rows, half the
// shared memory).  LQ_TMA_WS: a ninth warp issues the copies and the consumer warps hand stages back through "empty"
// mbarriers, so that no block-wide barrier sits in the tile loop.
#ifndef LQ_TMA_QSLOTS
#define LQ_TMA_QSLOTS 8
#endif
#ifndef LQ_TMA_WS
#define LQ_TMA_WS 1
#endif
// three blocks per SM are resident (shared memory); said explicitly, because with 288 threads and no second argument
// ptxas held the kernel to 40 registers and spilled
#ifndef LQ_TMA_MINBLK
#define LQ_TMA_MINBLK 3
#endif
#define LQ_TMA_BOUNDS __launch_bounds__(kTmaBlockThreads, LQ_TMA_MINBLK)
constexpr int kTmaQSlots = LQ_TMA_QSLOTS;
constexpr int kTmaBlockThreads = LQ_TMA_WS ? kThreads + 32 : kThreads;
constexpr int kTmaRowsPerThread = 8;
constexpr int kTmaTileFloats = kTmaRowsPerThread * kThreads;   // 2048 floats = 8 KB per array per stage

struct ColTmaJob {
  const float* s; float4* part;
  long long outer, W, inner, nchunks, ncolblk;
  int rpc, tc, write_dx;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "LQ_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra LQ_DONE_%=;\n\t"
      "bra LQ_WAIT_%=;\n\t"
      "LQ_DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(map), "r"(x), "r"(y), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct ColTmaSmem {
  float P[kTmaStages][kTmaTileFloats];
  float D[kTmaStages][kTmaTileFloats];
  float2 queue[kTmaQSlots * kThreads];
  Acc tree[kThreads];
  unsigned long long full[kTmaStages];
  unsigned long long empty[kTmaStages];
};

template <class Op>
__global__ void LQ_TMA_BOUNDS col_tma_kernel(const __grid_constant__ ColTmaJob j, const Op op,
                                                                   const __grid_constant__ CUtensorMap mapP,
                                                                   const __grid_constant__ CUtensorMap mapD,
                                                                   const __grid_constant__ CUtensorMap mapX) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ColTmaSmem& sm = *reinterpret_cast<ColTmaSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tc = j.tc;                         // 32, 64 or 128 columns per block
  const int nh = kThreads / tc;                // row groups per tile
  const int tile_rows_max = kTmaRowsPerThread * nh;
  const int c = t % tc, h = t / tc;            // (producer warp: h >= nh, never a consumer index)
#if LQ_COL_ORDER
  // column blocks of the same rows are neighbours in launch order: what is in flight at any time covers whole rows
  const long long cb = blockIdx.x % j.ncolblk;
  const long long chunk = blockIdx.x / j.ncolblk;
#else
  const long long chunk = blockIdx.x % j.nchunks;
  const long long cb = blockIdx.x / j.nchunks;
#endif
  const long long col0 = cb * tc;
  const int ncols = (int)((j.W - col0 < tc) ? (j.W - col0) : tc);       // multiple of 4 (W % 4 == 0)
  const long long r0 = chunk * j.rpc;
  const long long r1 = (r0 + j.rpc < j.outer) ? r0 + j.rpc : j.outer;
  const int nrows = (int)(r1 - r0);
  const int ntiles = (nrows + tile_rows_max - 1) / tile_rows_max;
  const bool consumer = t < kThreads;
  const bool col_ok = consumer && c < ncols;

  if (t == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // launched with programmatic stream serialization: everything above overlaps the tail of the kernel before this one
  // in the stream; nothing of that kernel's output (dy, P, s) is touched before this point
  cudaGridDependencySynchronize();

  // A box is always transferred whole (rows / columns outside the tensor arrive as zeros), so the
  // transaction count per stage is a constant.  rpc is a multiple of the tile height: only the last
  // chunk of the tensor has a short tile, and there the hardware clips.
  const unsigned stage_bytes = 2u * (unsigned)(tile_rows_max * tc) * 4u;
  auto load_tile = [&](int tile) {             // one thread only
    const int st = tile % kTmaStages;
    mbar_expect_tx(&sm.full[st], stage_bytes);
    const int y = (int)(r0 + (long long)tile * tile_rows_max);
    tma_load_2d(&sm.P[st][0], &mapP, (int)col0, y, &sm.full[st]);
    tma_load_2d(&sm.D[st][0], &mapD, (int)col0, y, &sm.full[st]);
  };
  auto store_tile = [&](int tile) {            // dX = dy, straight from the staged tile (one thread only)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma_store_2d(&mapX, (int)col0, (int)(r0 + (long long)tile * tile_rows_max), &sm.D[tile % kTmaStages][0]);
    bulk_commit();
  };

#if LQ_TMA_WS
  if (!consumer) {
    // producer warp: lane 0 keeps kTmaStages - 1 tiles in flight and writes each dy tile back as dX once it has landed
    if (lane == 0) {
      for (int tile = 0; tile < kTmaStages - 1 && tile < ntiles; ++tile) load_tile(tile);
      for (int tile = 0; tile < ntiles; ++tile) {
        const int nxt = tile + kTmaStages - 1;                 // goes into the stage tile - 1 was consumed from
        if (nxt < ntiles) {
          if (tile >= 1) {
            mbar_wait(&sm.empty[(tile - 1) % kTmaStages], (unsigned)(((tile - 1) / kTmaStages) & 1));   // all consumer warps are done with it
            bulk_wait_read0();                                 // and so is the dX store that read its dy
          }
          load_tile(nxt);
        }
        if (j.write_dx) {
          mbar_wait(&sm.full[tile % kTmaStages], (unsigned)((tile / kTmaStages) & 1));
          store_tile(tile);
        }
      }
      bulk_wait0();                            // outstanding dX stores
    }
  }
#else
  auto issue = [&](int tile) {                 // thread 0 only
    bulk_wait_read0();                         // the dX store out of this stage's previous tile has read it
    load_tile(tile);
  };
  if (t == 0) {
    for (int tile = 0; tile < kTmaStages - 1 && tile < ntiles; ++tile) issue(tile);
  }
#endif

  const long long grp = col_ok ? (col0 + c) / j.inner : 0;
  Job dummy{};
  const typename Op::Ctx ctx = op.prep(col_ok ? __ldg(j.s + grp) : 1.0f, make_float4(0.f, 0.f, 0.f, 0.f), dummy);
  Acc acc; Op::init(acc);
  Defer df;
  df.slot = sm.queue + (consumer ? warp : 0) * (kTmaQSlots * 32) + lane;
  df.n = 0;

  if (consumer) {
    for (int tile = 0; tile < ntiles; ++tile) {
      const int st = tile % kTmaStages;
#if !LQ_TMA_WS
      if (t == 0 && tile + kTmaStages - 1 < ntiles) issue(tile + kTmaStages - 1);
#endif
      mbar_wait(&sm.full[st], (unsigned)((tile / kTmaStages) & 1));
      const int rows = (nrows - tile * tile_rows_max < tile_rows_max) ? nrows - tile * tile_rows_max : tile_rows_max;
#if !LQ_TMA_WS
      if (t == 0 && j.write_dx) store_tile(tile);
#endif
      const int mine = col_ok ? rows - h * kTmaRowsPerThread : 0;         // this thread's valid rows (may be <= 0)
      const float* ps = &sm.P[st][(h * kTmaRowsPerThread) * tc + c];
      const float* ds = &sm.D[st][(h * kTmaRowsPerThread) * tc + c];
#pragma unroll
      for (int qd = 0; qd < kTmaRowsPerThread / 4; ++qd) {
        if (4 * qd + 3 < mine) {
          float4 o, cc;
          op.quad(make_float4(ps[(4 * qd) * tc], ps[(4 * qd + 1) * tc], ps[(4 * qd + 2) * tc], ps[(4 * qd + 3) * tc]),
                  make_float4(ds[(4 * qd) * tc], ds[(4 * qd + 1) * tc], ds[(4 * qd + 2) * tc], ds[(4 * qd + 3) * tc]),
                  ctx, dummy, acc, o, cc, df);
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (4 * qd + v < mine) {
              float o, cc;
              op.elem(ps[(4 * qd + v) * tc], ds[(4 * qd + v) * tc], ctx, dummy, acc, o, cc);
            }
          }
        }
        if (kTmaQSlots < kTmaRowsPerThread) op.flush(acc, df);   // the loop is warp-uniform: the warp is converged
      }
      if (kTmaQSlots >= kTmaRowsPerThread) op.flush(acc, df);
#if LQ_TMA_WS
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[st]);   // this warp is done reading stage st
#else
      __syncthreads();                           // everyone is done reading stage st
#endif
    }
  }
#if !LQ_TMA_WS
  if (t == 0) bulk_wait0();                    // outstanding dX stores
#endif

  // combine the nh row groups of each column in a fixed tree, one partial per (chunk, column)
  if (nh > 1) {
    if (consumer) sm.tree[t] = acc;
    for (int sgap = nh >> 1; sgap >= 1; sgap >>= 1) {
      __syncthreads();
      if (h < sgap) {
        Acc mine = sm.tree[t];
        Op::combine(mine, sm.tree[t + sgap * tc]);
        sm.tree[t] = mine;
      }
    }
    __syncthreads();
    if (consumer) acc = sm.tree[t];
  }
  if (h == 0 && col_ok) j.part[chunk * j.W + col0 + c] = pack_acc(acc);
}

// Shape-only eligibility and geometry (pointer alignment is checked at launch).
inline bool col_tma_shape_ok(long long outer, long long C, long long inner) {
  const long long W = C * inner;
  return C > 1 && inner < kRowMinInner && (W % 4 == 0) && W >= 32 && outer >= 512 && outer * W >= (1ll << 20) &&
         W < (1ll << 31) && outer < (1ll << 31);
}
inline Geom make_col_tma_geom(long long outer, long long C, long long inner) {
  Geom g = make_geom(outer, C, inner, false);
  const int tc = g.W >= 128 ? 128 : (g.W >= 64 ? 64 : 32);
  const long long ncolblk = cdiv(g.W, tc);
  const int tile_rows = kTmaRowsPerThread * (kThreads / tc);
  long long rpc = cdiv(cdiv(outer * ncolblk, 148ll * LQ_TMA_BPS), tile_rows) * tile_rows;   // ~LQ_TMA_BPS blocks per SM
  if (rpc < 4ll * tile_rows) rpc = 4ll * tile_rows;
  if (rpc > 8192) rpc = 8192 / tile_rows * tile_rows;
  g.bx = tc; g.by = kThreads / tc;
  g.rpc = (int)rpc;
  g.nchunks = cdiv(outer, rpc);
  g.ncolblk = ncolblk;
  g.nblocks = g.nchunks * ncolblk;
  g.npart1 = g.nchunks * g.W;
  g.kper = g.nchunks * inner;
  g.nlin = g.kper > kFinSplit ? cdiv(g.kper, kFinChunk) : 0;
  g.npartials = g.npart1 + C * g.nlin;
  return g;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// fp32 matrix [rows, W] row-major, box [box_rows, box_cols], no swizzle (dense tile in shared memory).
inline int make_map_2d(CUtensorMap* m, const float* base, long long rows, long long W, int box_rows, int box_cols) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return LQ_E_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return LQ_E_CUDA; }
  return LQ_OK;
}

template <class Op>
int col_tma_launch(const Geom& g, const float* dy, const float* P, const float* s, float* dX, float4* part,
                   const Op& op, cudaStream_t st) {
  ColTmaJob j{};
  j.s = s; j.part = part;
  j.outer = g.outer; j.W = g.W; j.inner = g.inner; j.nchunks = g.nchunks; j.ncolblk = g.ncolblk;
  j.rpc = g.rpc; j.tc = g.bx; j.write_dx = (dX != nullptr && dX != dy) ? 1 : 0;
  const int tile_rows = kTmaRowsPerThread * (kThreads / g.bx);
  CUtensorMap mP, mD, mX;
  int rc;
  if ((rc = make_map_2d(&mP, P, g.outer, g.W, tile_rows, g.bx))) return rc;
  if ((rc = make_map_2d(&mD, dy, g.outer, g.W, tile_rows, g.bx))) return rc;
  if ((rc = make_map_2d(&mX, j.write_dx ? dX : dy, g.outer, g.W, tile_rows, g.bx))) return rc;
  // per device and per function; a host-side table update, so simply done on every launch
  cudaError_t e = cudaFuncSetAttribute(col_tma_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ColTmaSmem));
  if (e != cudaSuccess) return cuda_fail(e, "col_tma smem attribute");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)g.nblocks); cfg.blockDim = dim3(kTmaBlockThreads); cfg.dynamicSmemBytes = sizeof(ColTmaSmem); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cuda_fail(cudaLaunchKernelEx(&cfg, col_tma_kernel<Op>, j, op, mP, mD, mX), "col_tma launch");
}

}  // namespace lq
