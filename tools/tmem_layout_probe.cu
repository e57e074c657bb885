// Probe (sm_100a): register <-> (lane, column) mapping of tcgen05.ld.16x256b.x4, used to design an epilogue
// in which a thread owns 4 pixels x 8 channels instead of 1 pixel x 32 channels.
// TMEM is filled with value = lane * 1000 + column through tcgen05.st.32x32b (thread t of warp w <-> lane 32 w + t).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_layout_probe tmem_layout_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) probe(float* out) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 32; ++c) {
    const uint32_t v = __float_as_uint((float)((warp * 32 + lane) * 1000 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(row_addr + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    const uint32_t a = tmem + ((uint32_t)(warp * 32 + half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(tid * 2 + half) * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}
int main() {
  float* d; cudaMalloc(&d, 128 * 32 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  static float h[128 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int tid = 0; tid < 128; ++tid)
    for (int half = 0; half < 2; ++half)
      for (int i = 0; i < 16; ++i) {
        const int v = (int)h[(tid * 2 + half) * 16 + i];
        const int lane = v / 1000, col = v % 1000;
        const int t = tid & 31, w = tid >> 5, k = i >> 2, j = i & 3;
        const int exp_lane = w * 32 + half * 16 + (t >> 2) + (j >= 2 ? 8 : 0);
        const int exp_col = 8 * k + 2 * (t & 3) + (j & 1);
        if (lane != exp_lane || col != exp_col) { if (bad < 12) printf("tid %d half %d reg %d: got lane %d col %d, expected lane %d col %d\n", tid, half, i, lane, col, exp_lane, exp_col); ++bad; }
      }
  printf(bad ? "MAPPING DIFFERS (%d mismatches)\n" : "mapping confirmed: reg 4k+j of thread t <-> lane base + t/4 + 8*(j/2), column 8k + 2*(t%%4) + j%%2 (%d mismatches)\n", bad);
  for (int i = 0; i < 16; ++i) printf("thread 5 half 0 reg %d -> %d\n", i, (int)h[(5 * 2 + 0) * 16 + i]);
  return 0;
}
