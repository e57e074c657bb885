// Probe of tcgen05 (UMMA) shared-memory descriptor semantics on sm_100a, used to decide how the
// conv kernel may address a haloed activation tile:
//   * K-major, SWIZZLE_128B, kind::tf32, M=128, N=32, K=32 (4 MMAs of K=8), accumulators in TMEM
//   * A rows taken from a 512-row smem image at row  shift + (m/8)*pitch + (m%8)
//     i.e. descriptor start = base + shift*128 B, SBO = pitch*128 B
//   * every 128-byte row of the image is stored with the 16-byte chunk index XOR-ed with
//     (absolute row index & 7) - what TMA SWIZZLE_128B writes into a 1024-byte aligned buffer.
// For each (shift, pitch) the MMA is issued with base_offset = 0 and base_offset = shift & 7 and the
// result compared with the exact CPU product (inputs are small integers, exact in tf32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int ROWS = 512, N = 32, K = 32, M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                               // LBO (ignored for swizzled K-major), 16 B
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ a_img, const float* __restrict__ b_img,
                                                    float* __restrict__ out, int shift, int pitch, int base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sa = (float*)smem;                          // ROWS x 32 floats
  float* sb = (float*)(smem + ROWS * 128);           // N x 32 floats
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // fill smem images with the swizzle of the absolute row index
  for (int i = tid; i < ROWS * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(a_img + row * 32 + ch * 4);
    *reinterpret_cast<float4*>(smem + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(b_img + row * 32 + ch * 4);
    *reinterpret_cast<float4*>((uint8_t*)sb + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // make generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    const uint32_t b_addr = smem_u32(sb);
    for (int k = 0; k < K / 8; ++k) {
      const uint64_t da = make_desc(a_addr + k * 32, pitch * 128, base_offset);
      const uint64_t db = make_desc(b_addr + k * 32, 1024, 0);
      const uint32_t acc = k > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int m = warp * 32 + lane;
  for (int j = 0; j < 32; ++j) out[m * N + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
  std::vector<float> a(ROWS * 32), b(N * 32), ref(M * N), got(M * N);
  for (int i = 0; i < ROWS * 32; ++i) a[i] = (float)((i * 7 + (i / 32) * 3) % 17 - 8);
  for (int i = 0; i < N * 32; ++i) b[i] = (float)((i * 5 + (i / 32)) % 13 - 6);
  float *da, *db, *dout;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, ref.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = ROWS * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shifts[] = {0, 8, 1, 3, 10, 11, 21};
  const int pitches[] = {8, 16, 10, 18};
  int fails_plain = 0;
  for (int pitch : pitches)
    for (int shift : shifts) {
      for (int mode = 0; mode < 2; ++mode) {
        const int bo = mode == 0 ? 0 : (shift & 7);
        if (mode == 1 && bo == 0) continue;
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < N; ++n) {
            const int row = shift + (m / 8) * pitch + (m % 8);
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += a[row * 32 + k] * b[n * 32 + k];
            ref[m * N + n] = s;
          }
        cudaMemset(dout, 0xff, ref.size() * 4);
        probe_kernel<<<1, 128, smem>>>(da, db, dout, shift, pitch, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s (pitch %d shift %d bo %d)\n", cudaGetErrorString(e), pitch, shift, bo); return 1; }
        cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; float maxerr = 0.f;
        for (int i = 0; i < M * N; ++i) { float d = fabsf(got[i] - ref[i]); if (d > 1e-3f) ++bad; if (d > maxerr) maxerr = d; }
        printf("pitch %2d shift %2d base_offset %d : %s (%d/%d wrong, max err %.1f)\n", pitch, shift, bo, bad ? "MISMATCH" : "ok", bad, M * N, maxerr);
        if (pitch == 8 && shift == 0 && bad) ++fails_plain;
      }
    }
  printf(fails_plain ? "BASELINE (aligned, pitch 8) FAILED\n" : "baseline ok\n");
  return 0;
}
