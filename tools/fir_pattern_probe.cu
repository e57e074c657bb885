// Which access pattern lets a 2-D op with misaligned rows ([P,1025,1025] -> [P,1024,1024], fp32) approach the copy bandwidth?
//   strip : a warp walks down a 32-column strip, one 128-byte request per row (8 rows in flight)
//   flat  : a thread per 4 consecutive outputs, linear in the output (reads 4 scalars)
//   band  : a CTA stages a band of R+3 whole input rows (one contiguous chunk) in shared memory with 16-byte loads, then writes
//           the R output rows from there
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fir_pattern_probe fir_pattern_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
constexpr int P = 256, IW = 1025, IH = 1025, OW = 1024, OH = 1024;

__global__ void __launch_bounds__(128) k_strip(const float* __restrict__ in, float* __restrict__ out, int rh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 4 + warp) * 32;
  const int r0 = blockIdx.y * rh;
  const float* src = in + (size_t)blockIdx.z * IW * IH + (size_t)r0 * IW + c0 + lane;
  float* dst = out + (size_t)blockIdx.z * OW * OH + (size_t)r0 * OW + c0 + lane;
  for (int j = 0; j < rh; j += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (j + u) * IW);
#pragma unroll
    for (int u = 0; u < 8; ++u) dst[(j + u) * OW] = v[u];
  }
}
__global__ void __launch_bounds__(256) k_flat(const float* __restrict__ in, float* __restrict__ out) {
  const size_t total4 = (size_t)P * OH * OW / 4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += (size_t)gridDim.x * 256) {
    const size_t o = i * 4;
    const int x = (int)(o % OW);
    const size_t t = o / OW;
    const int y = (int)(t % OH);
    const size_t p = t / OH;
    const float* s = in + (p * IH + y) * IW + x;
    float4 v = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), __ldg(s + 3));
    *reinterpret_cast<float4*>(out + o) = v;
  }
}
template <int R>
__global__ void __launch_bounds__(256) k_band(const float* __restrict__ in, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int band = blockIdx.x, plane = blockIdx.y;
  const int r0 = band * R;
  const int rows_in = min(R + 3, IH - r0);
  const size_t first = ((size_t)plane * IH + r0) * IW;       // first float of the chunk
  const size_t first_al = first & ~(size_t)3;                 // 16-byte aligned start
  const int skew = (int)(first - first_al);
  const int nvec = (skew + rows_in * IW + 3) / 4;
  const float4* src = reinterpret_cast<const float4*>(in + first_al);
  const size_t lim = ((size_t)P * IH * IW + 3) / 4;          // do not read past the allocation (padded by the host)
  for (int i = threadIdx.x; i < nvec; i += 256) reinterpret_cast<float4*>(sm)[i] = (first_al / 4 + i < lim) ? __ldg(src + i) : make_float4(0, 0, 0, 0);
  __syncthreads();
  const int rows_out = min(R, OH - r0);
  for (int i = threadIdx.x; i < rows_out * (OW / 4); i += 256) {
    const int y = i / (OW / 4), x = (i % (OW / 4)) * 4;
    const float* s = sm + skew + y * IW + x;
    *reinterpret_cast<float4*>(out + ((size_t)plane * OH + r0 + y) * OW + x) = make_float4(s[0], s[1], s[2], s[3]);
  }
}
template <typename F>
static void timeit(const char* name, F launch, void* flush, size_t flush_bytes) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 8; ++it) {
    cudaMemsetAsync(flush, it, flush_bytes);
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 2 && ms < best) best = ms;
  }
  const double bytes = 4.0 * P * ((double)IW * IH + (double)OW * OH);
  printf("%-22s %8.1f us %7.0f GB/s  (%s)\n", name, best * 1e3, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *in, *out; void* flush; const size_t fb = 256u << 20;
  cudaMalloc(&in, (size_t)P * IW * IH * 4 + 64); cudaMalloc(&out, (size_t)P * OW * OH * 4); cudaMalloc(&flush, fb);
  cudaMemset(in, 0, (size_t)P * IW * IH * 4 + 64);
  for (int rh : {128, 64, 32}) {
    char nm[64]; snprintf(nm, 64, "strip rh=%d", rh);
    timeit(nm, [&] { k_strip<<<dim3(OW / 128, OH / rh, P), 128>>>(in, out, rh); }, flush, fb);
  }
  timeit("flat", [&] { k_flat<<<148 * 16, 256>>>(in, out); }, flush, fb);
  { constexpr int R = 13; const int smem = ((R + 3) * IW + 8) * 4; cudaFuncSetAttribute(k_band<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    timeit("band R=13", [&] { k_band<R><<<dim3((OH + R - 1) / R, P), 256, smem>>>(in, out); }, flush, fb); }
  { constexpr int R = 5; const int smem = ((R + 3) * IW + 8) * 4; cudaFuncSetAttribute(k_band<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    timeit("band R=5", [&] { k_band<R><<<dim3((OH + R - 1) / R, P), 256, smem>>>(in, out); }, flush, fb); }
  { constexpr int R = 29; const int smem = ((R + 3) * IW + 8) * 4; cudaFuncSetAttribute(k_band<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    timeit("band R=29", [&] { k_band<R><<<dim3((OH + R - 1) / R, P), 256, smem>>>(in, out); }, flush, fb); }
  return 0;
}
