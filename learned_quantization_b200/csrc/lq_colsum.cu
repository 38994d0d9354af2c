// lq_colsum.cu -- SURVEY 8(f3), bias path: the gradient of the quantised bias qb in
// conv2d(x, qk) + qb (NQ:343-350) is the column sum of the NHWC output gradient viewed as a
// [N*H*W, Cout] matrix.  The framework's generic reduction spends ~43 us per layer on it (13 % of the
// IMAGENETTE step); this is the same interleaved traversal as the COL regime: a thread owns four
// adjacent columns, walks down a chunk of rows with 128-bit loads, threads sharing columns combine
// in shared memory, and a second stage adds the chunks in order.  Deterministic, fp32 accumulation
// per chunk and fp64 across chunks.
#include "lq_traverse.cuh"

namespace lq {

// Rows per block: the grid is held to ~4 blocks per SM -- the second stage walks the chunks, and with a fixed 256 rows per
// block the IMAGENETTE step's first layer ([802816, 64]) left it 3136 partial rows to add with two blocks (41 us), while its
// last layers ([3136, 512]) ran the first stage on 13 blocks (8.5 us for 6.4 MB).  Never fewer than 16 rows per block.
constexpr long long kCsTargetBlocks = 148 * 4;
constexpr long long kCsMinRows = 16;
inline long long colsum_rows_per_block(long long rows, long long ncolblk) {
  long long target = kCsTargetBlocks / (ncolblk > 0 ? ncolblk : 1);
  if (target < 1) target = 1;
  long long rpb = (rows + target - 1) / target;
  return rpb < kCsMinRows ? kCsMinRows : rpb;
}
inline long long colsum_ncolblk(long long cols, bool vec) {
  const long long ncv = vec ? cols / 4 : cols;
  const long long bx = ncv < 256 ? ncv : 256;
  return bx > 0 ? (ncv + bx - 1) / bx : 1;
}

// grid (nchunks, ncolblk); block 256 = bx column-vectors x by rows
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, long long rows, long long cols,
                                                             int bx, int by, int vec, long long rpb, float* __restrict__ part) {
  cudaGridDependencySynchronize();
  __shared__ float4 sm[256];
  const int t = threadIdx.x, ty = t / bx, tx = t - ty * bx;
  const long long cv = (long long)blockIdx.y * bx + tx;
  const long long ncv = vec ? cols / 4 : cols;
  const bool active = ty < by && cv < ncv;
  const long long r0 = (long long)blockIdx.x * rpb;
  const long long r1 = (r0 + rpb < rows) ? r0 + rpb : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    if (vec) {
      const float4* p = reinterpret_cast<const float4*>(x) + cv;
      const long long stride = cols / 4;
      long long r = r0 + ty;
      for (; r + 3 * by < r1; r += 4 * by) {
        const float4 a = __ldcs(p + r * stride), b = __ldcs(p + (r + by) * stride);
        const float4 c = __ldcs(p + (r + 2 * by) * stride), d = __ldcs(p + (r + 3 * by) * stride);
        acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
        acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
      }
      for (; r < r1; r += by) {
        const float4 a = __ldcs(p + r * stride);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    } else {
      for (long long r = r0 + ty; r < r1; r += by) acc.x += x[r * cols + cv];
    }
  }
  sm[t] = acc;
  int span = 1;
  while (span < by) span <<= 1;
  for (int s = span >> 1; s >= 1; s >>= 1) {
    __syncthreads();
    if (ty < s && ty + s < by) {
      float4 a = sm[t];
      const float4 b = sm[t + s * bx];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      sm[t] = a;
    }
  }
  __syncthreads();
  if (ty == 0 && cv < ncv) {
    const float4 a = sm[t];
    float* dst = part + (long long)blockIdx.x * cols + (vec ? cv * 4 : cv);
    if (vec) *reinterpret_cast<float4*>(dst) = a; else *dst = a.x;
  }
}

// Second stage: lanes are adjacent columns (coalesced), the block's 32 warps split the chunks, four loads in flight each;
// warp w adds chunks w, w + 32, ... in order, the warps' sums are added in warp order: deterministic, fp64 throughout.
constexpr int kCsFinalWarps = 32;
__global__ void __launch_bounds__(32 * kCsFinalWarps) colsum_final_kernel(const float* __restrict__ part, long long nchunks, long long cols,
                                                                          float* __restrict__ out) {
  cudaGridDependencySynchronize();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long c = (long long)blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (c < cols) {
    long long k = w;
    for (; k + 3 * kCsFinalWarps < nchunks; k += 4 * kCsFinalWarps) {
      const float a0 = part[k * cols + c], a1 = part[(k + kCsFinalWarps) * cols + c];
      const float a2 = part[(k + 2 * kCsFinalWarps) * cols + c], a3 = part[(k + 3 * kCsFinalWarps) * cols + c];
      acc += (double)a0; acc += (double)a1; acc += (double)a2; acc += (double)a3;
    }
    for (; k < nchunks; k += kCsFinalWarps) acc += (double)part[k * cols + c];
  }
  __shared__ double sm[kCsFinalWarps][33];
  sm[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c < cols) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kCsFinalWarps; ++i) s += sm[i][lane];
    out[c] = (float)s;
  }
}

}  // namespace lq

using namespace lq;

extern "C" int lq_colsum_workspace_bytes(int64_t rows, int64_t cols, size_t* out_bytes) {
  if (!out_bytes) { set_error("lq_colsum_workspace_bytes: null out"); return LQ_E_NULLPTR; }
  if (rows < 0 || cols < 0) { set_error("lq_colsum_workspace_bytes: negative size"); return LQ_E_BADSHAPE; }
  // whichever form runs (the vector form needs an aligned pointer, unknown here): the larger chunk count of the two
  long long nchunks = 0;
  for (int vec = 0; vec < 2; ++vec) {
    if (vec && cols % 4 != 0) continue;
    const long long rpb = colsum_rows_per_block(rows, colsum_ncolblk(cols, vec != 0));
    const long long n = (rows + rpb - 1) / rpb;
    if (n > nchunks) nchunks = n;
  }
  *out_bytes = align256((size_t)nchunks * (size_t)cols * sizeof(float));
  return LQ_OK;
}

extern "C" int lq_colsum(const float* x, int64_t rows, int64_t cols, float* out, void* ws, size_t ws_bytes, lq_stream_t stream) {
  if (rows < 0 || cols < 0) { set_error("lq_colsum: negative size"); return LQ_E_BADSHAPE; }
  if (cols == 0) return LQ_OK;
  if (!out) { set_error("lq_colsum: null out"); return LQ_E_NULLPTR; }
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) return cuda_fail(cudaMemsetAsync(out, 0, (size_t)cols * 4, st), "lq_colsum: zero");
  if (!x) { set_error("lq_colsum: null x"); return LQ_E_NULLPTR; }
  const int vec = (aligned16(x) && cols % 4 == 0) ? 1 : 0;
  const long long ncv = vec ? cols / 4 : cols;
  int bx = ncv < 256 ? (int)ncv : 256;
  int by = 256 / bx;
  const long long ncolblk = (ncv + bx - 1) / bx;
  if (ncolblk > 65535) { set_error("lq_colsum: too many columns"); return LQ_E_BADSHAPE; }
  const long long rpb = colsum_rows_per_block(rows, ncolblk);
  const long long nchunks = (rows + rpb - 1) / rpb;
  const size_t need = align256((size_t)nchunks * (size_t)cols * sizeof(float));
  if (!ws || ws_bytes < need || !aligned16(ws)) { set_error("lq_colsum: workspace %zu B needed, got %zu", need, ws_bytes); return LQ_E_WORKSPACE; }
  if (nchunks > 0x7fffffffll) { set_error("lq_colsum: too many rows"); return LQ_E_BADSHAPE; }
  float* part = static_cast<float*>(ws);
  int rc = launch_after(colsum_partial_kernel, dim3((unsigned)nchunks, (unsigned)ncolblk), dim3(256), st, "colsum launch",
                        x, (long long)rows, (long long)cols, bx, by, vec, rpb, part);
  if (rc) return rc;
  return launch_after(colsum_final_kernel, dim3((unsigned)((cols + 31) / 32)), dim3(32 * kCsFinalWarps), st, "colsum final launch",
                      (const float*)part, nchunks, (long long)cols, out);
}
