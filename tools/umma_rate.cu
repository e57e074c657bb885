// Issue-rate probe for tcgen05.mma kind::tf32 (M = 128, K = 8, SS mode) on sm_100a: how many SM cycles one MMA costs as a
// function of N and of how the A operand is addressed, with nothing else running on the SM.  Used to tell whether the
// N <= 64 layers of conv_tc.cu sit on a tensor-pipe / operand-fetch floor or lose their time elsewhere.
//   mode 0: the conv pattern - 9 shifted windows (taps) of a 10 x 18 haloed image, SBO = 10 rows, 4 K-steps per tap
//   mode 1: plain GEMM pattern - one 128-row tile, SBO = 8 rows (1024 B), 4 K-steps, repeated
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate umma_rate.cu ; run: ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128) rate_kernel(int N, int mode, int iters, int nacc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  // zero the operands (values do not matter for the rate; NaN payloads should not either, but keep it clean)
  for (uint32_t i = tid * 16; i < 2 * 23552 + 4 * 256 * 128; i += 128 * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem0 + i), "r"(0));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sbo = mode == 0 ? 10 * 128 : 1024;
    const uint32_t a_hi = (uint32_t)(sbo >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (smem0 >> 4) | 0x10000u, b_lo0 = ((smem0 + 2 * 23552) >> 4) | 0x10000u;
    const uint32_t slice16 = (uint32_t)(N * 128) >> 4;
    uint32_t tap_a[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) tap_a[t] = mode == 0 ? (uint32_t)(((t / 3) * 10 + (t % 3)) * 8) : 0u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t tacc = tmem + (uint32_t)((it % nacc) * N);
      const uint32_t a_stage = a_lo0 + (uint32_t)(it & 1) * (23552 >> 4);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint32_t a_lo = a_stage + tap_a[t], b_lo = b_lo0 + (uint32_t)(t & 3) * slice16;   // four weight-slice slots
        umma_tf32_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, t > 0);
        umma_tf32_lohi(tacc, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
        umma_tf32_lohi(tacc, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
        umma_tf32_lohi(tacc, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 2 * 23552 + 4 * 256 * 128 + 2048;   // two activation stages, four weight-slice slots of up to 256 rows
  if (smem > 232448 - 1024) { printf("smem plan too large\n"); return 1; }
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  printf("tcgen05.mma kind::tf32 M=128 K=8, %d x 36 MMAs per CTA, 148 CTAs\n", iters);
  printf("%5s %5s %5s %12s %12s %10s\n", "N", "mode", "nacc", "cycles/MMA", "ideal N/2", "TFLOP/s");
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {32, 64, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * N > 512) continue;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        rate_kernel<<<148, 128, smem>>>(N, mode, 10, nacc, d);   // warm-up
        cudaEventRecord(e0);
        rate_kernel<<<148, 128, smem>>>(N, mode, iters, nacc, d);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        const double mmas = (double)iters * 36;
        printf("%5d %5d %5d %12.1f %12.1f %10.1f\n", N, mode, nacc, cyc / mmas, N / 2.0, 148.0 * mmas * 2.0 * 128 * N * 8 / (ms * 1e-3) / 1e12);
      }
  return 0;
}
