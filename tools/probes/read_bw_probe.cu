// How fast can a read-only max|x| pass over 64 Mi floats go on this GPU, and with which structure?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/read_bw_probe tools/probes/read_bw_probe.cu && /tmp/read_bw_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ float fold4(float m, float4 q) {
  return fmaxf(fmaxf(m, fabsf(q.x)), fmaxf(fabsf(q.y), fmaxf(fabsf(q.z), fabsf(q.w))));
}

// one unit (1024 floats) per warp, 8 loads per lane, then exit
__global__ void __launch_bounds__(256, 4) oneshot(const float4* p, float* out, long long units) {
  const long long unit = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (unit >= units) return;
  const int lane = threadIdx.x & 31;
  const float4* b = p + unit * 256;
  float4 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = __ldcs(b + u * 32 + lane);
  float m = -1.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) m = fold4(m, v[u]);
  m = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(m)));
  if (lane == 0) out[unit] = m;
}

// persistent: warp walks units w, w+nw, ...; register refilled right after it is folded
template <int MINB>
__global__ void __launch_bounds__(256, MINB) rolling(const float4* p, float* out, long long units) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= units) return;
  float4 v[8];
  const float4* b = p + w * 256 + lane;
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = __ldcs(b + u * 32);
  for (long long unit = w; unit < units; unit += nw) {
    const long long nxt = unit + nw;
    const bool more = nxt < units;
    const float4* nb = p + nxt * 256 + lane;
    float m = -1.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      m = fold4(m, v[u]);
      if (more) v[u] = __ldcs(nb + u * 32);
    }
    m = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(m)));
    if (lane == 0) out[unit] = m;
  }
}

// persistent, one accumulator per warp for the whole walk (one partial per warp)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) rolling_acc(const float4* p, float* out, long long units) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= units) return;
  float4 v[8];
  const float4* b = p + w * 256 + lane;
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = __ldcs(b + u * 32);
  float m = -1.f;
  for (long long unit = w; unit < units; unit += nw) {
    const long long nxt = unit + nw;
    const bool more = nxt < units;
    const float4* nb = p + nxt * 256 + lane;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      m = fold4(m, v[u]);
      if (more) v[u] = __ldcs(nb + u * 32);
    }
  }
  m = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(m)));
  if (lane == 0) out[w] = m;
}

// classic grid-stride, U loads per thread per trip
template <int U, int T>
__global__ void __launch_bounds__(T) gridstride(const float4* p, float* out, long long n4) {
  const long long stride = (long long)gridDim.x * T;
  long long i = (long long)blockIdx.x * T + threadIdx.x;
  float m = -1.f;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) m = fold4(m, v[u]);
  }
  for (; i < n4; i += stride) m = fold4(m, __ldcs(p + i));
  m = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(m)));
  if ((threadIdx.x & 31) == 0) out[((long long)blockIdx.x * T + threadIdx.x) >> 5] = m;
}

template <class F>
float timeit(F launch, int reps = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) launch(i);
  cudaDeviceSynchronize();
  std::vector<float> ts;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch(i);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    ts.push_back(ms / reps * 1e3f);
  }
  std::sort(ts.begin(), ts.end());
  return ts[2];
}

int main() {
  const long long n = 1ll << 26, n4 = n / 4, units = n / 1024;
  float* buf[3]; float* out;
  for (int i = 0; i < 3; ++i) { cudaMalloc(&buf[i], n * 4); cudaMemset(buf[i], 0x3c, n * 4); }
  cudaMalloc(&out, units * 4 * 8);
  auto rep = [&](const char* name, float us) { printf("%-44s %7.1f us  %6.0f GB/s\n", name, us, n * 4 / (us * 1e-6) / 1e9); };
  rep("oneshot (unit per warp, 8192 blocks)", timeit([&](int i) { oneshot<<<(unsigned)(units / 8), 256>>>((const float4*)buf[i % 3], out, units); }));
  for (int bps : {2, 3, 4}) {
    char nm[64]; snprintf(nm, 64, "rolling, %d blocks/SM", bps);
    if (bps == 4) rep(nm, timeit([&](int i) { rolling<4><<<148 * 4, 256>>>((const float4*)buf[i % 3], out, units); }));
    if (bps == 3) rep(nm, timeit([&](int i) { rolling<3><<<148 * 3, 256>>>((const float4*)buf[i % 3], out, units); }));
    if (bps == 2) rep(nm, timeit([&](int i) { rolling<2><<<148 * 2, 256>>>((const float4*)buf[i % 3], out, units); }));
  }
  rep("rolling, 8 blocks/SM (2 waves of 4)", timeit([&](int i) { rolling<4><<<148 * 8, 256>>>((const float4*)buf[i % 3], out, units); }));
  rep("rolling_acc, 4 blocks/SM", timeit([&](int i) { rolling_acc<4><<<148 * 4, 256>>>((const float4*)buf[i % 3], out, units); }));
  rep("rolling_acc, 6 blocks/SM", timeit([&](int i) { rolling_acc<6><<<148 * 6, 256>>>((const float4*)buf[i % 3], out, units); }));
  rep("gridstride U=8 T=256 x4/SM", timeit([&](int i) { gridstride<8, 256><<<148 * 4, 256>>>((const float4*)buf[i % 3], out, n4); }));
  rep("gridstride U=8 T=512 x2/SM", timeit([&](int i) { gridstride<8, 512><<<148 * 2, 512>>>((const float4*)buf[i % 3], out, n4); }));
  rep("gridstride U=4 T=256 x8/SM", timeit([&](int i) { gridstride<4, 256><<<148 * 8, 256>>>((const float4*)buf[i % 3], out, n4); }));
  rep("gridstride U=16 T=256 x2/SM", timeit([&](int i) { gridstride<16, 256><<<148 * 2, 256>>>((const float4*)buf[i % 3], out, n4); }));
  rep("gridstride U=8 T=1024 x1/SM", timeit([&](int i) { gridstride<8, 1024><<<148, 1024>>>((const float4*)buf[i % 3], out, n4); }));
  rep("gridstride U=8 T=1024 x2/SM", timeit([&](int i) { gridstride<8, 1024><<<148 * 2, 1024>>>((const float4*)buf[i % 3], out, n4); }));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
