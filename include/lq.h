/* lq.h -- C ABI of liblq.so: the B200 (sm_100a) learned-quantization hot path.
 *
 * Drop-in boundary for the operator the reference implements in TensorFlow as
 *   @tf.custom_gradient my_custom_gradient(parameter, scale[, penalty_threshold])
 * (IMAGENETTE/nested_quantization_layer/custom_components/custom_layers.py:49-120,
 *  IMAGENETTE/custom_loss_terms/custom_components/custom_layers.py:49-64,
 *  CIFAR-10/paper_implementation/custom_components/custom_layers.py:49-66)
 * and for the penalty reductions of
 *  IMAGENETTE/custom_loss_terms/custom_components/custom_loss_functions.py:75-116,161-195,240-275.
 * Paths are relative to the reference root; "NQ", "CL", "PI", "LF" below abbreviate them.
 *
 * Conventions
 *  - Every tensor is contiguous fp32, viewed as [outer, C, inner] row-major with
 *    group(idx) = (idx / inner) % C; `s` and `ds` hold C floats.  The four reference
 *    orientations map onto this view (NQ:147-197; table in DESIGN.md).
 *  - All pointers are DEVICE pointers on the current device unless the name says host.
 *    The caller owns every buffer including the workspace; the library never allocates
 *    device memory, never synchronises (except the *_host entry points, which say so)
 *    and enqueues all work on the caller's stream.
 *  - Every function returns LQ_OK or a negative lq_status; lq_last_error() gives the
 *    thread-local message.  Nothing throws.  There is no CPU fallback: without a CUDA
 *    device every compute entry point returns LQ_E_CUDA.
 */
#ifndef LQ_H_
#define LQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQ_VERSION 200 /* 0.2.0 */

typedef void* lq_stream_t; /* a cudaStream_t (NULL = legacy default stream) */

typedef enum lq_status {
  LQ_OK = 0,
  LQ_E_NULLPTR = -1,
  LQ_E_BADSHAPE = -2,
  LQ_E_BADMODE = -3,
  LQ_E_WORKSPACE = -4,
  LQ_E_CUDA = -5,
  LQ_E_NCCL = -6
} lq_status;

/* NQ:59 uses tf.floor; RNE (+ clamp) is the extension BASELINE.json's wording asks for. */
enum { LQ_ROUND_FLOOR = 0, LQ_ROUND_RNE = 1 };
/* optional integer-code output of the forward (utils/log_scripts.py:74-79 exports int8) */
enum { LQ_CODES_NONE = 0, LQ_CODES_F32 = 1, LQ_CODES_I32 = 2, LQ_CODES_I8 = 3 };
/* which custom loss term (LF:75, LF:161, LF:240) */
enum { LQ_PEN_MAXBIN = 0, LQ_PEN_DIFFERENCE = 1, LQ_PEN_INVERSE = 2 };
/* scale-gradient rule of the backward: NQ:62-118 (threshold) or CL:61-62 / PI:62-64 (zeros) */
enum { LQ_SGRAD_NONE = 0, LQ_SGRAD_THRESHOLD = 1 };
/* phases for lq_workspace_bytes */
enum { LQ_PHASE_FWD = 0, LQ_PHASE_BWD = 1, LQ_PHASE_PENALTY_FWD = 2, LQ_PHASE_PENALTY_BWD = 3 };
/* lq_tensor_desc.flags */
enum { LQ_F_ACCUM_DX = 1u, /* penalty bwd: dX += grad instead of dX = grad */
       LQ_F_ACCUM_DS = 2u  /* penalty bwd: ds += grad instead of ds = grad */ };

/* One quantised parameter tensor (a kernel, a weight matrix or a bias) with its scale. */
typedef struct lq_tensor_desc {
  const float* P;  /* parameter                       [outer*C*inner] */
  const float* s;  /* learned scale                   [C]             */
  float* y;        /* fake-quantised parameter (fwd)  [outer*C*inner] */
  const float* dy; /* upstream gradient (bwd)         [outer*C*inner] */
  float* dX;       /* parameter gradient (bwd); may equal dy or be NULL (dP is dy itself, NQ:118) */
  float* ds;       /* scale gradient (bwd)            [C]             */
  int64_t outer, C, inner;
  uint32_t flags;
  uint32_t reserved;
} lq_tensor_desc;

int lq_version(void);
const char* lq_last_error(void);
/* number of SMs of the current device (0 and LQ_E_CUDA through lq_last_error when there is none) */
int lq_device_sm_count(void);

/* Bytes of device workspace the given phase needs for these tensors (sum over descs,
 * each 256-B aligned).  Pure host arithmetic: works without a GPU. */
int lq_workspace_bytes(const lq_tensor_desc* descs, int n, int phase, size_t* out_bytes);

/* a1 forward, NQ:55-60:  q = floor(P / s) (IEEE division), y = q * s.
 * codes (nullable) receives q as f32 / i32 / i8 (i8 wraps like numpy .astype(int8)).
 * round_mode/qmin/qmax: LQ_ROUND_FLOOR with -inf/+inf reproduces the reference. */
int lq_fq_fwd(const float* P, const float* s, float* y, void* codes, int codes_dtype,
              int64_t outer, int64_t C, int64_t inner, int round_mode, float qmin, float qmax,
              lq_stream_t stream);

/* a1 with the per-group statistics the loss terms and the callbacks are functions of, as by-products of the same
 * HBM pass (no second read of P):  stats[3g .. 3g+2] = { min P, max P, sum |P - P / s| } over group g.
 *   max |q| of the group (custom_callbacks.py:84-129) = max(|floor(min P / s)|, |floor(max P / s)|);
 *   MaxBin's per-group term (LF:90-94) = max(|min P|, |max P|) / s   (division by s > 0 is monotone);
 *   Difference's per-group sum (LF:175-178) = stats[3g+2].
 * NaN elements drop out of the extremes and poison the sum.  ws >= lq_workspace_bytes(.., LQ_PHASE_BWD). */
int lq_fq_fwd_stats(const float* P, const float* s, float* y, void* codes, int codes_dtype,
                    int64_t outer, int64_t C, int64_t inner, int round_mode, float qmin, float qmax,
                    float* stats, void* ws, size_t ws_bytes, lq_stream_t stream);

/* a2 backward, NQ:62-118:  dX = dy (skipped when dX is NULL or == dy) and
 * ds[g] = mean_g(e) * max_g|q|, e as in NQ:103-111, recomputing q,y from P,s in registers.
 * One HBM pass + a deterministic second-stage reduce over block partials in `ws`. */
int lq_fq_bwd_threshold(const float* dy, const float* P, const float* s, float penalty_threshold,
                        float* dX, float* ds, int64_t outer, int64_t C, int64_t inner,
                        int round_mode, float qmin, float qmax,
                        void* ws, size_t ws_bytes, lq_stream_t stream);

/* a2 for a tensor SHARDED along `outer` over several devices (microbench sizes):
 * each shard runs lq_fq_bwd_threshold_stats on its slice -- same single HBM pass, but the second
 * stage stops at the raw per-group statistics stats[3*g .. 3*g+2] = {sum, n_below, max|q|}
 * (doubles) -- the shards' statistics are gathered (lq_dp_allgather_f64: 24*C bytes per rank, the
 * path's only exchange), and lq_fq_bwd_threshold_finish folds them in shard order and applies
 * NQ:103-116 with n_per_group = elements per group over ALL shards.  Every rank obtains the same
 * bits, equal to the unsharded lq_fq_bwd_threshold up to the fp64 summation order. */
int lq_fq_bwd_threshold_stats(const float* dy, const float* P, const float* s, float penalty_threshold,
                              float* dX, double* stats, int64_t outer, int64_t C, int64_t inner,
                              int round_mode, float qmin, float qmax,
                              void* ws, size_t ws_bytes, lq_stream_t stream);
int lq_fq_bwd_threshold_finish(const double* stats, int nshards, int64_t C, float penalty_threshold,
                               double n_per_group, float* ds, lq_stream_t stream);

/* a3 backward, CL:61-62 / PI:62-64:  dX = dy (copy skipped when NULL or aliased), ds = 0. */
int lq_fq_bwd_ste(const float* dy, float* dX, float* ds, int64_t n_elems, int64_t C,
                  lq_stream_t stream);

/* a5 MinValueConstraint, NQ:35-46:  s = max(s, min_value) in place. */
int lq_scale_constraint(float* s, int64_t n, float min_value, lq_stream_t stream);

/* a7-a9 loss terms over a list of tensors (LF iterates layers; a layer contributes its
 * kernel and its bias as two descs).  out_penalty[0] = compute_*_penalty() (un-scaled by
 * the penalty rate, LF:116 / LF:195 / LF:275). Uses desc fields P, s, outer, C, inner. */
int lq_penalty_fwd(int kind, const lq_tensor_desc* descs, int n, float* out_penalty,
                   void* ws, size_t ws_bytes, lq_stream_t stream);

/* TF-autodiff gradients of `u * penalty` w.r.t. every P (into desc.dX) and every s (into
 * desc.ds), u = upstream * (upstream_dev ? *upstream_dev : 1): the device scalar lets an
 * autograd engine pass its incoming gradient without a host sync.  LQ_F_ACCUM_* flags
 * select += .  Must be given the workspace that the matching lq_penalty_fwd call filled
 * (it holds the per-group maxima / tie counts).  Inverse never touches dX. */
int lq_penalty_bwd(int kind, const lq_tensor_desc* descs, int n, float upstream,
                   const float* upstream_dev, void* ws, size_t ws_bytes, lq_stream_t stream);

/* The forward of every tensor AND the MaxBin / Difference loss term over them in ONE pass (the penalty
 * forward "fused into K1: 0 extra bytes"; reference: the forward at NQ:55-60 and compute_*_penalty at LF:75-116 /
 * 161-195 read the same kernels and biases in the same step).  desc.y = floor(P / s) * s is written exactly as
 * lq_mt_fq_fwd writes it; out_penalty and the workspace are left exactly as lq_penalty_fwd(kind, ...) leaves them, so
 * the matching lq_penalty_bwd call works unchanged.  ws: lq_workspace_bytes(descs, n, LQ_PHASE_PENALTY_FWD).
 * LQ_PEN_INVERSE (which reads no tensor) is refused with LQ_E_BADMODE. */
int lq_mt_fq_fwd_penalty(int kind, const lq_tensor_desc* descs, int n, int round_mode, float qmin, float qmax,
                         float* out_penalty, void* ws, size_t ws_bytes, lq_stream_t stream);

/* K5 multi-tensor forms: ONE launch per phase over all descs (the models quantise 4-42
 * small tensors per step; per-tensor launches dominate there). */
int lq_mt_fq_fwd(const lq_tensor_desc* descs, int n, int round_mode, float qmin, float qmax,
                 lq_stream_t stream);
int lq_mt_fq_bwd(const lq_tensor_desc* descs, int n, float penalty_threshold, int scale_grad_mode,
                 int round_mode, float qmin, float qmax,
                 void* ws, size_t ws_bytes, lq_stream_t stream);

/* the step after the hot path.  One launch updates every tensor in descs with
 * Keras 2.11 Adam / SGD semantics (keras/optimizers/optimizer_experimental/{adam,sgd}.py, the
 * optimizers experiment.py:763-771 and CIFAR-10/paper_implementation/experiment.py:130-141
 * construct), folding in: g *= grad_scale * (grad_scale_dev ? *grad_scale_dev : 1);
 * g += 2 * l2 * p (keras.regularizers.l2, IMAGENETTE/.../experiment.py:638); decoupled
 * weight decay p -= p * weight_decay * lr; and MinValueConstraint p = max(p, min_value)
 * (NQ:35-46; pass -INFINITY for unconstrained tensors).  m / v are the optimizer slots
 * (v unused by SGD; m may be NULL for SGD with momentum 0).  step is 1-based; when step_dev
 * (a device int64) is given it is read instead, so that a CUDA-graph replay of the step sees
 * the current iteration. */
typedef struct lq_opt_desc {
  float* p;        /* parameter or learned scale, updated in place */
  const float* g;  /* its gradient                                 */
  float* m;        /* Adam m / SGD momentum buffer                 */
  float* v;        /* Adam v                                       */
  int64_t n;
  float l2, min_value, weight_decay;
  uint32_t reserved;
} lq_opt_desc;
int lq_opt_adam(const lq_opt_desc* descs, int n, float lr, float beta1, float beta2, float eps,
                int64_t step, const int64_t* step_dev, float grad_scale, const float* grad_scale_dev,
                lq_stream_t stream);
int lq_opt_sgd(const lq_opt_desc* descs, int n, float lr, float momentum, int nesterov,
               float grad_scale, const float* grad_scale_dev, lq_stream_t stream);

/* integer-code statistics on the device, replacing the per-epoch host round trip of
 * custom_callbacks.py:84-129 (floor(kernel / scale) -> np.unique(return_counts=True), max|q|) and
 * the unchecked .astype(np.int8) of utils/log_scripts.py:74-79.
 * lq_code_range: out_min_max[0..1] (device) = smallest / largest code (NaN, NaN for an empty or
 *   all-NaN tensor); ws: 8 bytes of device scratch.
 * lq_code_histogram: hist[b] += #elements with code == code_lo + b (hist and *outside are device
 *   buffers the CALLER zeroes; codes outside the window and NaNs are counted in *outside).
 *   np.unique's values/counts are the non-zero bins; the int8 saturation count is the mass
 *   outside [-128, 127]. */
int lq_code_range(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner,
                  int round_mode, float qmin, float qmax, float* out_min_max,
                  void* ws, size_t ws_bytes, lq_stream_t stream);
int lq_code_histogram(const float* P, const float* s, int64_t outer, int64_t C, int64_t inner,
                      int round_mode, float qmin, float qmax, int64_t code_lo, int64_t nbins,
                      uint32_t* hist, unsigned long long* outside, lq_stream_t stream);

/* SURVEY 8(f3), bias path: out[c] = sum_r x[r*cols + c] for a row-major fp32 matrix -- the gradient of
 * the quantised bias in conv2d(x, qk) + qb (NQ:343-350) is the column sum of the NHWC output gradient
 * seen as [N*H*W, Cout].  Deterministic two-stage reduction; ws >= lq_colsum_workspace_bytes. */
int lq_colsum_workspace_bytes(int64_t rows, int64_t cols, size_t* out_bytes);
int lq_colsum(const float* x, int64_t rows, int64_t cols, float* out, void* ws, size_t ws_bytes, lq_stream_t stream);

/* Host-buffer entry point (the end-to-end path bench.py times): pinned or pageable HOST
 * arrays in, HOST arrays out; chunks are staged through the caller's device workspace
 * (>= lq_host_workspace_bytes) on internal streams, copies overlapped with the kernels.
 * SYNCHRONISES before returning.  hdX may be NULL (dP is dy itself). */
int lq_host_workspace_bytes(int64_t outer, int64_t C, int64_t inner, size_t* out_bytes);
int lq_host_fq_fwd_bwd(const float* hP, const float* hs, const float* hdy, float penalty_threshold,
                       float* hy, float* hdX, float* hds,
                       int64_t outer, int64_t C, int64_t inner,
                       void* dev_ws, size_t dev_ws_bytes);
/* The same with an explicit handle (the pipeline's streams and events, created once; the handle-less form keeps one
 * per device for the life of the process) and ordered after the caller's stream: with order_after != 0 every staged
 * copy waits for the work enqueued on `after` at the time of the call (what last used dev_ws).  One call at a time
 * per handle. */
typedef struct lq_host lq_host_t;
int lq_host_create(lq_host_t** out);
int lq_host_destroy(lq_host_t* h);
int lq_host_fq_fwd_bwd_on(lq_host_t* h, const float* hP, const float* hs, const float* hdy, float penalty_threshold,
                          float* hy, float* hdX, float* hds,
                          int64_t outer, int64_t C, int64_t inner,
                          void* dev_ws, size_t dev_ws_bytes, lq_stream_t after, int order_after);

/* Host-only (no device work, no pointers dereferenced): how lq_mt_fq_fwd (phase LQ_PHASE_FWD) / lq_mt_fq_bwd
 * (LQ_PHASE_BWD) would cut every tensor of a table: out[4 i .. 4 i + 3] = {regime (0 = rows of one group, 1 = interleaved),
 * rows per chunk (interleaved) or segment length (rows), blocks, partial slots}.  For tests of the geometry rules. */
int lq_debug_mt_geometry(const lq_tensor_desc* descs, int n, int phase, int64_t* out);

/* Diagnostics: checks on the device that the kernels' hoisted-reciprocal division equals
 * IEEE division (the bit-exactness the forward is graded on, NQ:56-58 tf.truediv) for every
 * 23-bit mantissa of the dividend a = bits(a_sign_exp_bits | mantissa) against each divisor
 * s_bits[i] (n_s <= 2^24).  out3 (device, zeroed by the caller) accumulates
 * {mismatches, pairs checked inside the fast window, first mismatching (a<<32 | s)}. */
int lq_debug_divcheck(const uint32_t* s_bits, int64_t n_s, uint32_t a_sign_exp_bits,
                      unsigned long long* out3, lq_stream_t stream);

/* Data parallelism (new; the reference is single-device, experiment.py:774-788):
 * one NCCL communicator per process, sum-allreduce of a flat fp32 gradient bucket scaled
 * by `scale` (1/world) in the same call. */
typedef struct lq_dp lq_dp_t;
int lq_dp_unique_id(void* out_id128);                       /* 128-byte ncclUniqueId */
int lq_dp_init(const void* id128, int rank, int world, lq_dp_t** out);
int lq_dp_allreduce(lq_dp_t* dp, float* flat, int64_t n, float scale, lq_stream_t stream);
int lq_dp_allreduce_max(lq_dp_t* dp, float* flat, int64_t n, lq_stream_t stream);
int lq_dp_allgather_f64(lq_dp_t* dp, const double* send, double* recv /* [world*n] */, int64_t n, lq_stream_t stream);
int lq_dp_destroy(lq_dp_t* dp);

/* The same exchange + finish as lq_dp_allgather_f64 + lq_fq_bwd_threshold_finish, fused into ONE
 * kernel over NVLink peer memory: each rank stores its per-group statistics into every peer's
 * receive buffer, publishes / awaits a per-step flag, and folds the world's statistics in rank order.
 * lq_p2p_create allocates this rank's receive buffer (the one device allocation the library makes;
 * cudaMalloc, so that it is IPC-exportable) and returns its 64-byte cudaIpcMemHandle; the host program
 * gathers the handles of all ranks (world * 64 bytes, rank order) and passes them to lq_p2p_connect.
 * One process per GPU.  lq_p2p_error reports (and synchronises for) the time-out flag a rank raises
 * when a peer does not show up within ~2 s. */
typedef struct lq_p2p lq_p2p_t;
int lq_p2p_create(int rank, int world, int64_t max_groups, lq_p2p_t** out, void* out_handle64);
int lq_p2p_connect(lq_p2p_t* p, const void* handles64_all);
/* same-process peers instead of IPC handles (one process driving several devices, or several "ranks" on one
 * device on different streams -- how the single-GPU tests exercise the exchange): bufs[r] = lq_p2p_buffer of rank r */
int lq_p2p_buffer(lq_p2p_t* p, void** out);
int lq_p2p_connect_ptrs(lq_p2p_t* p, const void* const* bufs);
int lq_fq_bwd_threshold_finish_p2p(lq_p2p_t* p, const double* local_stats, int64_t C, float penalty_threshold,
                                   double n_per_group, float* ds, lq_stream_t stream);
int lq_p2p_error(lq_p2p_t* p, int* out_flag);
int lq_p2p_destroy(lq_p2p_t* p);

/* The data-parallel gradient exchange FUSED with the threshold backward, over NVLink peer memory
 * (exact mode of SURVEY 8e: all-reduce dy, then NQ:62-118 on the averaged dy on every rank).
 * lq_dpx_create allocates this rank's arena (cudaMalloc, IPC-exportable; *out_arena is the user part,
 * arena_bytes long, 256-byte aligned, zero-filled) and returns its 64-byte cudaIpcMemHandle; the host
 * program gathers the handles of all ranks (world * 64 bytes, rank order) and passes them to
 * lq_dpx_connect.  The caller lays its flat gradient bucket and the partial workspace out INSIDE the
 * arena, identically on every rank.
 * lq_dpx_exchange: ONE persistent kernel per rank.  The work items of the multi-tensor backward over
 * `descs` (plus the flat range plain[0, n_plain) of unquantised gradients) are dealt round-robin to the
 * ranks; the owner of an item loads dy from every rank's arena, averages in rank order, applies the
 * threshold-backward operator with its local P and s, and stores the averaged dy (in place: desc.dX must
 * be NULL or dy) and the item's partial to every rank.  The ordinary second stage then produces ds,
 * bit-identical on all ranks.  scale_grad_mode LQ_SGRAD_NONE: everything is only averaged, ds = 0.
 * desc.dy, plain and ws (>= lq_workspace_bytes(descs, n, LQ_PHASE_BWD)) must lie in the arena; P, s, ds are
 * ordinary device pointers.  Reference rounding only (floor, no clamp).  Every rank must issue the same
 * sequence of calls; barrier epochs live on the device, so a call may be captured in a CUDA graph and
 * replayed.  channel (0..3): exchanges that may run at the same time on different streams (a step that
 * starts exchanging its late layers while the backward is still producing the early ones) must use
 * different channels, and the same channel for the same call on every rank; max_blocks (0 = default, two
 * persistent blocks per SM) bounds the grid so that concurrent exchanges fit on the device together.
 * lq_dpx_error reports (and synchronises for) the time-out flag a rank raises when a peer
 * does not show up within ~2 s. */
typedef struct lq_dpx lq_dpx_t;
int lq_dpx_create(int rank, int world, size_t arena_bytes, lq_dpx_t** out, void* out_handle64, void** out_arena);
int lq_dpx_connect(lq_dpx_t* x, const void* handles64_all);
/* same-process peers instead of IPC handles: arenas[r] = the *out_arena of rank r's lq_dpx_create in THIS
 * process (one process driving several devices with peer access enabled, or -- for tests -- several arenas
 * on one device, launched on different streams with few enough blocks to be resident together) */
int lq_dpx_connect_ptrs(lq_dpx_t* x, const void* const* arenas);
int lq_dpx_exchange(lq_dpx_t* x, const lq_tensor_desc* descs, int n, float* plain, int64_t n_plain,
                    float penalty_threshold, int scale_grad_mode, int channel, int max_blocks,
                    void* ws, size_t ws_bytes, lq_stream_t stream);
/* Flat in-place mean all-reduce of flat[0, n) (inside the arena, 16-byte aligned, n a multiple of 4) on the copy
 * engine: bulk asynchronous copies (TMA) stage one tile from every rank's arena in shared memory, a few warps add
 * them in rank order, bulk stores send the mean to every rank.  The bytes in flight live in shared memory, so a
 * few dozen blocks (max_blocks, 0 = 32) fill the NVLinks: this is the exchange a training step starts on a side
 * stream while its backward is still computing.  Channels and epochs as for lq_dpx_exchange. */
int lq_dpx_allreduce(lq_dpx_t* x, float* flat, int64_t n, int channel, int max_blocks, lq_stream_t stream);
int lq_dpx_error(lq_dpx_t* x, int* out_flag);
int lq_dpx_destroy(lq_dpx_t* x);

#ifdef __cplusplus
}
#endif
#endif /* LQ_H_ */
