// lq_host.cu -- host-buffer entry point: forward + threshold backward on HOST arrays.
//
// This is what a caller without device residency binds to (and what bench.py's end-to-end
// number times).  The tensor is cut into chunks that keep the [outer, C, inner] group
// structure intact; each chunk is staged H2D -> K1 -> K2 -> D2H on one of kSlots internal
// streams so the PCIe copies of neighbouring chunks overlap each other (full duplex) and
// the kernels.  All chunks write their block partials into ONE partial array laid out for
// the whole tensor, so a single K4 second stage at the end gives the same deterministic
// result as the device-resident path.
//
//   outer > 1         : chunks of whole outer-slices (multiples of the COL row-chunk)
//   outer == 1, C > 1 : chunks of whole groups (each chunk is an independent problem)
//   outer == 1, C == 1: chunks of whole ROW segments
#include <math.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include "lq_traverse.cuh"

namespace lq {

#ifndef LQ_HOST_SLOTS
#define LQ_HOST_SLOTS 3
#endif
#ifndef LQ_HOST_CHUNK_ELEMS
#define LQ_HOST_CHUNK_ELEMS (4ll << 20)          // 16 MiB per staged array
#endif
constexpr int kSlots = LQ_HOST_SLOTS;
constexpr long long kChunkElems = LQ_HOST_CHUNK_ELEMS;

struct HostPlan {
  int mode;                 // 0 outer, 1 groups, 2 segments
  long long step;           // chunk length in units of the split axis
  long long chunk_elems;    // largest chunk in elements
  long long npart;          // partial slots
  size_t off_s, off_ds, off_part, off_stage, total;
};

static HostPlan make_plan(long long outer, long long C, long long inner) {
  HostPlan p{};
  const Geom G = make_geom(outer, C, inner, inner >= kRowMinInner);
  const long long n = outer * C * inner;
  if (outer > 1) {
    p.mode = 0;
    const long long slice = C * inner;
    long long step = std::max(1ll, kChunkElems / slice);
    if (G.regime == REG_COL) {
      const long long R = G.rpc;
      step = std::max(R, step / R * R);
    }
    p.step = std::min(step, outer);
    p.chunk_elems = p.step * slice;
    p.npart = G.npartials;
  } else if (C > 1) {
    p.mode = 1;
    p.step = std::min(C, std::max(1ll, kChunkElems / inner));
    p.chunk_elems = p.step * inner;
    p.npart = kSlots * make_geom(1, p.step, inner, inner >= kRowMinInner).npartials;   // per-slot private partials
  } else {
    p.mode = 2;
    long long step = std::max(kChunkElems / kSeg * kSeg, (long long)kSegMax + kSeg);
    p.step = (n <= step + kSegMax + kSeg) ? n : step;
    p.chunk_elems = (p.step == n) ? n : p.step + kSegMax + kSeg;   // the last chunk absorbs a short tail
    p.npart = G.npartials;
  }
  size_t off = 0;
  p.off_s = off; off += align256((size_t)C * 4);
  p.off_ds = off; off += align256((size_t)C * 4);
  p.off_part = off; off += align256((size_t)p.npart * sizeof(float4));
  p.off_stage = off; off += (size_t)kSlots * 3 * align256((size_t)p.chunk_elems * 4);
  p.total = off;
  return p;
}

}  // namespace lq

using namespace lq;

extern "C" int lq_host_workspace_bytes(int64_t outer, int64_t C, int64_t inner, size_t* out_bytes) {
  if (!out_bytes) { set_error("lq_host_workspace_bytes: null out_bytes"); return LQ_E_NULLPTR; }
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  if (outer * C * inner == 0) { *out_bytes = 0; return LQ_OK; }
  if (C == 1) { inner *= outer; outer = 1; }
  *out_bytes = make_plan(outer, C, inner).total;
  return LQ_OK;
}

// Streams and events of the staged pipeline, created once (per handle) instead of per call.
struct lq_host {
  int device;
  cudaStream_t st[kSlots];
  cudaEvent_t ev_s, ev_in, ev_done[kSlots];
  std::mutex mu;               // one call at a time per handle: the slots are shared state
};

static int host_handle_init(lq_host* h) {
  int rc = cuda_fail(cudaGetDevice(&h->device), "lq_host: cudaGetDevice");
  for (int i = 0; i < kSlots && !rc; ++i) {
    rc = cuda_fail(cudaStreamCreateWithFlags(&h->st[i], cudaStreamNonBlocking), "lq_host: stream create");
    if (!rc) rc = cuda_fail(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming), "lq_host: event create");
  }
  if (!rc) rc = cuda_fail(cudaEventCreateWithFlags(&h->ev_s, cudaEventDisableTiming), "lq_host: event create");
  if (!rc) rc = cuda_fail(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming), "lq_host: event create");
  return rc;
}

extern "C" int lq_host_create(lq_host_t** out) {
  if (!out) { set_error("lq_host_create: null out"); return LQ_E_NULLPTR; }
  lq_host* h = new lq_host();
  for (int i = 0; i < kSlots; ++i) { h->st[i] = nullptr; h->ev_done[i] = nullptr; }
  h->ev_s = h->ev_in = nullptr;
  const int rc = host_handle_init(h);
  if (rc) { lq_host_destroy(h); return rc; }
  *out = h;
  return LQ_OK;
}

extern "C" int lq_host_destroy(lq_host_t* h) {
  if (!h) return LQ_OK;
  for (int i = 0; i < kSlots; ++i) {
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    if (h->st[i]) { cudaStreamSynchronize(h->st[i]); cudaStreamDestroy(h->st[i]); }
  }
  if (h->ev_s) cudaEventDestroy(h->ev_s);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  delete h;
  return LQ_OK;
}

// the handle behind the handle-less entry point: one per device, created on first use, kept for the process
static lq_host* default_handle() {
  static std::mutex mu;
  static lq_host* per_device[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!per_device[dev]) {
    lq_host* h = nullptr;
    if (lq_host_create(&h) != LQ_OK) return nullptr;
    per_device[dev] = h;
  }
  return per_device[dev];
}

extern "C" int lq_host_fq_fwd_bwd(const float* hP, const float* hs, const float* hdy, float penalty_threshold,
                                  float* hy, float* hdX, float* hds,
                                  int64_t outer, int64_t C, int64_t inner,
                                  void* dev_ws, size_t dev_ws_bytes) {
  lq_host* h = default_handle();
  if (!h) { set_error("lq_host_fq_fwd_bwd: no CUDA device / streams"); return LQ_E_CUDA; }
  return lq_host_fq_fwd_bwd_on(h, hP, hs, hdy, penalty_threshold, hy, hdX, hds, outer, C, inner, dev_ws, dev_ws_bytes,
                               nullptr, 0);
}

extern "C" int lq_host_fq_fwd_bwd_on(lq_host_t* h, const float* hP, const float* hs, const float* hdy, float penalty_threshold,
                                     float* hy, float* hdX, float* hds,
                                     int64_t outer, int64_t C, int64_t inner,
                                     void* dev_ws, size_t dev_ws_bytes, lq_stream_t after, int order_after) {
  if (!h) { set_error("lq_host_fq_fwd_bwd_on: null handle"); return LQ_E_NULLPTR; }
  std::lock_guard<std::mutex> lock(h->mu);
  int rc = check_dims(outer, C, inner);
  if (rc) return rc;
  const long long n = outer * C * inner;
  if (n == 0) return LQ_OK;
  if (!hP || !hs || !hdy || !hy || !hds) { set_error("lq_host_fq_fwd_bwd: null host array"); return LQ_E_NULLPTR; }
  if (C == 1) { inner *= outer; outer = 1; }            // one group: a flat run (as make_geom does)
  const HostPlan p = make_plan(outer, C, inner);
  if (!dev_ws || dev_ws_bytes < p.total || !aligned16(dev_ws)) {
    set_error("lq_host_fq_fwd_bwd: device workspace %zu B needed, got %zu", p.total, dev_ws_bytes);
    return LQ_E_WORKSPACE;
  }
  QMode qm = make_qmode(LQ_ROUND_FLOOR, -INFINITY, INFINITY, &rc);
  char* base = static_cast<char*>(dev_ws);
  float* d_s = reinterpret_cast<float*>(base + p.off_s);
  float* d_ds = reinterpret_cast<float*>(base + p.off_ds);
  float4* d_part = reinterpret_cast<float4*>(base + p.off_part);
  const size_t stage = align256((size_t)p.chunk_elems * 4);
  auto stage_ptr = [&](int slot, int which) {
    return reinterpret_cast<float*>(base + p.off_stage + ((size_t)slot * 3 + which) * stage);
  };

  cudaStream_t* st = h->st;
  cudaEvent_t ev_s = h->ev_s;
  cudaEvent_t* ev_done = h->ev_done;
  auto cleanup = [&]() {};
#define LQ_TRY(expr, what) do { if ((rc = cuda_fail((expr), what))) { cudaDeviceSynchronize(); cleanup(); return rc; } } while (0)
#define LQ_TRY_RC(expr) do { if ((rc = (expr))) { cudaDeviceSynchronize(); cleanup(); return rc; } } while (0)
  // `after` (order_after != 0): work the caller has enqueued on that stream -- whatever last touched dev_ws, e.g. the
  // previous tenant of a caching allocator's block -- completes before the first staged copy
  if (order_after) {
    LQ_TRY(cudaEventRecord(h->ev_in, (cudaStream_t)after), "event record");
    for (int i = 0; i < kSlots; ++i) LQ_TRY(cudaStreamWaitEvent(st[i], h->ev_in, 0), "stream wait");
  }

  LQ_TRY(cudaMemcpyAsync(d_s, hs, (size_t)C * 4, cudaMemcpyHostToDevice, st[0]), "H2D s");
  // Zero partial = {sum 0, n_below 0, max 0}: neutral, so a tail chunk that covers fewer
  // units than the whole-tensor layout reserves leaves nothing undefined.
  LQ_TRY(cudaMemsetAsync(d_part, 0, (size_t)p.npart * sizeof(float4), st[0]), "zero partials");
  LQ_TRY(cudaEventRecord(ev_s, st[0]), "event record");
  for (int i = 1; i < kSlots; ++i) LQ_TRY(cudaStreamWaitEvent(st[i], ev_s, 0), "stream wait");

  const Geom G = make_geom(outer, C, inner, inner >= kRowMinInner);
  const long long axis_len = (p.mode == 0) ? outer : (p.mode == 1 ? C : n);
  int k = 0;
  for (long long a0 = 0, a_next = 0; a0 < axis_len; a0 = a_next, ++k) {
    const int slot = k % kSlots;
    long long a1 = std::min(a0 + p.step, axis_len);
    if (p.mode == 2 && axis_len - a1 < kSegMax + kSeg) a1 = axis_len;   // keep every chunk in the ROW regime
    long long e0, ne, co, cc, ci;
    const float* s_ptr = d_s;
    float4* part_ptr = d_part;
    if (p.mode == 0) {
      e0 = a0 * C * inner; ne = (a1 - a0) * C * inner; co = a1 - a0; cc = C; ci = inner;
      part_ptr = d_part + ((G.regime == REG_ROW) ? a0 * C * G.nseg
                                                 : (a0 / (long long)G.rpc) * G.W);
    } else if (p.mode == 1) {
      e0 = a0 * inner; ne = (a1 - a0) * inner; co = 1; cc = a1 - a0; ci = inner;
      s_ptr = d_s + a0;
      part_ptr = d_part + (size_t)slot * (p.npart / kSlots);
    } else {
      e0 = a0; ne = a1 - a0; co = 1; cc = 1; ci = ne;
      part_ptr = d_part + a0 / G.seg;
    }
    float* dP = stage_ptr(slot, 0);
    float* ddy = stage_ptr(slot, 1);
    float* dY = stage_ptr(slot, 2);
    LQ_TRY(cudaMemcpyAsync(dP, hP + e0, (size_t)ne * 4, cudaMemcpyHostToDevice, st[slot]), "H2D P");
    LQ_TRY(cudaMemcpyAsync(ddy, hdy + e0, (size_t)ne * 4, cudaMemcpyHostToDevice, st[slot]), "H2D dy");
    // every slice keeps one segmentation: the whole tensor's (modes 0, 2) or that of a full group chunk (mode 1)
    const int seg_of_slices = (G.regime != REG_ROW) ? 0
        : (p.mode == 1 ? make_geom(1, p.step, inner, inner >= kRowMinInner).seg : G.seg);
    const Geom g = make_geom(co, cc, ci, ci >= kRowMinInner, (p.mode == 0 && G.regime == REG_COL) ? G.rpc : 0, seg_of_slices);
    LQ_TRY_RC(fq_fwd_launch(g, dP, s_ptr, dY, nullptr, 0, qm, st[slot]));
    LQ_TRY_RC(fq_bwd_elements_launch(g, ddy, dP, s_ptr, penalty_threshold, nullptr, part_ptr, qm, st[slot]));
    LQ_TRY(cudaMemcpyAsync(hy + e0, dY, (size_t)ne * 4, cudaMemcpyDeviceToHost, st[slot]), "D2H y");
    if (p.mode == 1) LQ_TRY_RC(fq_finalize_launch(g, part_ptr, d_ds + a0, s_ptr, penalty_threshold, st[slot]));
    a_next = a1;
  }
  for (int i = 0; i < kSlots; ++i) LQ_TRY(cudaEventRecord(ev_done[i], st[i]), "event record");
  for (int i = 1; i < kSlots; ++i) LQ_TRY(cudaStreamWaitEvent(st[0], ev_done[i], 0), "stream wait");
  if (p.mode != 1) LQ_TRY_RC(fq_finalize_launch(G, d_part, d_ds, d_s, penalty_threshold, st[0]));
  LQ_TRY(cudaMemcpyAsync(hds, d_ds, (size_t)C * 4, cudaMemcpyDeviceToHost, st[0]), "D2H ds");
  if (hdX && hdX != hdy) memcpy(hdX, hdy, (size_t)n * 4);   // NQ:118: dP is dy itself; no device work
  LQ_TRY(cudaStreamSynchronize(st[0]), "sync");
  for (int i = 1; i < kSlots; ++i) LQ_TRY(cudaStreamSynchronize(st[i]), "sync");
#undef LQ_TRY
#undef LQ_TRY_RC
  cleanup();
  return LQ_OK;
}
