// lq_colsum.cu -- SURVEY 8(f3), bias path: the gradient of the quantised bias qb in
// conv2d(x, qk) + qb (NQ:343-350) is the column sum of the NHWC output gradient viewed as a
// [N*H*W, Cout] matrix.  The framework's generic reduction spends ~43 us per layer on it (13 % of the
// IMAGENETTE step); this is the same interleaved traversal as the COL regime: a thread owns four
// adjacent columns, walks down a chunk of rows with 128-bit loads, threads sharing columns combine
// in shared memory, and a second stage adds the chunks in order.  Deterministic, fp32 accumulation
// per chunk and fp64 across chunks.
#include "lq_traverse.cuh"

namespace lq {

constexpr int kCsRowsPerBlock = 256;

// grid (nchunks, ncolblk); block 256 = bx column-vectors x by rows
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, long long rows, long long cols,
                                                             int bx, int by, int vec, float* __restrict__ part) {
  cudaGridDependencySynchronize();
  __shared__ float4 sm[256];
  const int t = threadIdx.x, ty = t / bx, tx = t - ty * bx;
  const long long cv = (long long)blockIdx.y * bx + tx;
  const long long ncv = vec ? cols / 4 : cols;
  const bool active = ty < by && cv < ncv;
  const long long r0 = (long long)blockIdx.x * kCsRowsPerBlock;
  const long long r1 = (r0 + kCsRowsPerBlock < rows) ? r0 + kCsRowsPerBlock : rows;
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

__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, long long nchunks, long long cols,
                                                           float* __restrict__ out) {
  cudaGridDependencySynchronize();
  const long long c = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;          // 8 warps split the chunks; lanes are adjacent columns (coalesced)
  double acc = 0.0;
  if (c < cols)
    for (long long k = w; k < nchunks; k += 8) acc += (double)part[k * cols + c];
  __shared__ double sm[8][33];
  sm[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && c < cols) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][threadIdx.x];
    out[c] = (float)s;
  }
}

}  // namespace lq

using namespace lq;

extern "C" int lq_colsum_workspace_bytes(int64_t rows, int64_t cols, size_t* out_bytes) {
  if (!out_bytes) { set_error("lq_colsum_workspace_bytes: null out"); return LQ_E_NULLPTR; }
  if (rows < 0 || cols < 0) { set_error("lq_colsum_workspace_bytes: negative size"); return LQ_E_BADSHAPE; }
  *out_bytes = align256((size_t)((rows + kCsRowsPerBlock - 1) / kCsRowsPerBlock) * (size_t)cols * sizeof(float));
  return LQ_OK;
}

extern "C" int lq_colsum(const float* x, int64_t rows, int64_t cols, float* out, void* ws, size_t ws_bytes, lq_stream_t stream) {
  if (rows < 0 || cols < 0) { set_error("lq_colsum: negative size"); return LQ_E_BADSHAPE; }
  if (cols == 0) return LQ_OK;
  if (!out) { set_error("lq_colsum: null out"); return LQ_E_NULLPTR; }
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) return cuda_fail(cudaMemsetAsync(out, 0, (size_t)cols * 4, st), "lq_colsum: zero");
  if (!x) { set_error("lq_colsum: null x"); return LQ_E_NULLPTR; }
  const long long nchunks = (rows + kCsRowsPerBlock - 1) / kCsRowsPerBlock;
  const size_t need = align256((size_t)nchunks * (size_t)cols * sizeof(float));
  if (!ws || ws_bytes < need || !aligned16(ws)) { set_error("lq_colsum: workspace %zu B needed, got %zu", need, ws_bytes); return LQ_E_WORKSPACE; }
  if (nchunks > 0x7fffffffll) { set_error("lq_colsum: too many rows"); return LQ_E_BADSHAPE; }
  const int vec = (aligned16(x) && cols % 4 == 0) ? 1 : 0;
  const long long ncv = vec ? cols / 4 : cols;
  int bx = ncv < 256 ? (int)ncv : 256;
  int by = 256 / bx;
  const long long ncolblk = (ncv + bx - 1) / bx;
  if (ncolblk > 65535) { set_error("lq_colsum: too many columns"); return LQ_E_BADSHAPE; }
  float* part = static_cast<float*>(ws);
  int rc = launch_after(colsum_partial_kernel, dim3((unsigned)nchunks, (unsigned)ncolblk), dim3(256), st, "colsum launch",
                        x, (long long)rows, (long long)cols, bx, by, vec, part);
  if (rc) return rc;
  return launch_after(colsum_final_kernel, dim3((unsigned)((cols + 31) / 32)), dim3(256), st, "colsum final launch",
                      (const float*)part, nchunks, (long long)cols, out);
}
