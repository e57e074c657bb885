// Fingerprint embed (w0 = U^T alpha + mu, wx = w0 + sd V^T diag(sigma) sigmoid(key)) and the MSE
// loss glue of the attribution loop, batched over trajectories (C-ABI group 4).
// Reference: src/main.py:60-61, src/generator.py:148-161, src/utils.py:46-47.
#include "common.cuh"

namespace lfp {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// one thread per (b, j): coalesced over j in U[i][j] and V[k][j]
__global__ void __launch_bounds__(128) embed_fwd_kernel(const float* __restrict__ alpha,
                                                        const float* __restrict__ key,
                                                        const float* __restrict__ U,
                                                        const float* __restrict__ V,
                                                        const float* __restrict__ sigma,
                                                        const float* __restrict__ mu, float sd,
                                                        int n_main, int key_len, int dim,
                                                        float* __restrict__ w0, float* __restrict__ wx) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= dim) return;
  // four interleaved accumulators (rows i, i+1, i+2, i+3 of U): one dependent chain of 448 loads made the launch
  // latency-bound (80 us per step)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* al = alpha + (int64_t)b * n_main;
  int i = 0;
  for (; i + 4 <= n_main; i += 4) {
    a0 = fmaf(__ldg(U + (int64_t)i * dim + j), __ldg(al + i), a0);
    a1 = fmaf(__ldg(U + (int64_t)(i + 1) * dim + j), __ldg(al + i + 1), a1);
    a2 = fmaf(__ldg(U + (int64_t)(i + 2) * dim + j), __ldg(al + i + 2), a2);
    a3 = fmaf(__ldg(U + (int64_t)(i + 3) * dim + j), __ldg(al + i + 3), a3);
  }
  for (; i < n_main; ++i) a0 = fmaf(__ldg(U + (int64_t)i * dim + j), __ldg(al + i), a0);
  const float acc = (a0 + a1) + (a2 + a3);
  const float base = acc + __ldg(mu + j);
  float e = 0.f;
  for (int k = 0; k < key_len; ++k)
    e = fmaf(__ldg(V + (int64_t)k * dim + j) * __ldg(sigma + k), sigmoidf_(__ldg(key + (int64_t)b * key_len + k)), e);
  w0[(int64_t)b * dim + j] = base;
  wx[(int64_t)b * dim + j] = base + sd * e;
}

// one warp per output element: dot over dim
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ d_wx,
                                                        const float* __restrict__ key,
                                                        const float* __restrict__ U,
                                                        const float* __restrict__ V,
                                                        const float* __restrict__ sigma, float sd,
                                                        int n_main, int key_len, int dim,
                                                        float* __restrict__ d_alpha,
                                                        float* __restrict__ d_key) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n_main + key_len) return;
  const float* m = row < n_main ? U + (int64_t)row * dim : V + (int64_t)(row - n_main) * dim;
  float acc = 0.f;
  for (int j = lane; j < dim; j += 32) acc = fmaf(__ldg(m + j), __ldg(d_wx + (int64_t)b * dim + j), acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane != 0) return;
  if (row < n_main) {
    d_alpha[(int64_t)b * n_main + row] = acc;
  } else {
    const int k = row - n_main;
    const float sg = sigmoidf_(__ldg(key + (int64_t)b * key_len + k));
    d_key[(int64_t)b * key_len + k] = sd * __ldg(sigma + k) * acc * sg * (1.f - sg);
  }
}

// loss_total[b] = mse[b] + weight * sum_i (relu(alpha - max) + relu(min - alpha))      (src/main.py:65, src/utils.py:53-58)
// one CTA per trajectory, fixed summation order
__global__ void __launch_bounds__(256) bound_loss_kernel(const float* __restrict__ alpha, const float* __restrict__ amax,
                                                         const float* __restrict__ amin, const float* __restrict__ mse,
                                                         int n_main, float weight, float* __restrict__ loss_total) {
  __shared__ float part[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_main; i += 256) {
    const float a = __ldg(alpha + (int64_t)b * n_main + i);
    acc += fmaxf(a - __ldg(amax + i), 0.f) + fmaxf(__ldg(amin + i) - a, 0.f);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) part[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k];
    loss_total[b] = __ldg(mse + b) + weight * t;
  }
}

// Backward of the embed (as embed_bwd_kernel) + gradient of the alpha bound + one Adam step, in place
// (src/main.py:66-70; torch.optim.Adam defaults).  One warp per (parameter element, trajectory).
__global__ void __launch_bounds__(256) adam_update_kernel(const float* __restrict__ d_wx, float* __restrict__ alpha,
                                                          float* __restrict__ key, const float* __restrict__ U,
                                                          const float* __restrict__ V, const float* __restrict__ sigma,
                                                          const float* __restrict__ amax, const float* __restrict__ amin,
                                                          float sd, float bound_weight, float* __restrict__ m_a,
                                                          float* __restrict__ v_a, float* __restrict__ m_k,
                                                          float* __restrict__ v_k, int n_main, int key_len, int dim,
                                                          float step_size, float sqrt_bc2, float beta1, float beta2,
                                                          float omb1, float omb2, float eps, int optimise_alpha,
                                                          const float2* __restrict__ hyper, const int* __restrict__ step_ptr) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n_main + key_len) return;
  if (row < n_main && !optimise_alpha) return;
  if (hyper != nullptr) {   // graph-replayed step: the schedule scalars of step *step_ptr come from a device table
    const float2 hp = hyper[*step_ptr];
    step_size = hp.x; sqrt_bc2 = hp.y;
  }
  const float* mrow = row < n_main ? U + (int64_t)row * dim : V + (int64_t)(row - n_main) * dim;
  float acc = 0.f;
  for (int j = lane; j < dim; j += 32) acc = fmaf(__ldg(mrow + j), __ldg(d_wx + (int64_t)b * dim + j), acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane != 0) return;
  float g, *p, *m, *v;
  if (row < n_main) {
    const int64_t i = (int64_t)b * n_main + row;
    p = alpha + i; m = m_a + i; v = v_a + i;
    const float a = *p;
    g = acc + bound_weight * ((a - __ldg(amax + row) > 0.f ? 1.f : 0.f) - (__ldg(amin + row) - a > 0.f ? 1.f : 0.f));
  } else {
    const int k = row - n_main;
    const int64_t i = (int64_t)b * key_len + k;
    p = key + i; m = m_k + i; v = v_k + i;
    const float sg = sigmoidf_(*p);
    g = sd * __ldg(sigma + k) * acc * sg * (1.f - sg);
  }
  // same operation order as torch.optim.Adam: m.mul_(b1).add_(g, alpha=1-b1); v.mul_(b2).addcmul_(g, g, value=1-b2);
  // denom = v.sqrt() / sqrt(bc2) + eps; p.addcdiv_(m, denom, value=-lr/bc1)
  const float mn = __fadd_rn(__fmul_rn(*m, beta1), __fmul_rn(omb1, g));
  const float vn = __fadd_rn(__fmul_rn(*v, beta2), __fmul_rn(omb2, __fmul_rn(g, g)));
  *m = mn; *v = vn;
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vn), sqrt_bc2), eps);
  *p = __fadd_rn(*p, __fmul_rn(-step_size, __fdiv_rn(mn, denom)));
}

// MSE: pass 1 writes d_est and per-CTA partial sums of squared error; pass 2 reduces in order
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ est,
                                                          const float* __restrict__ target,
                                                          int64_t target_bstride, int64_t numel_per,
                                                          float* __restrict__ d_est,
                                                          float* __restrict__ partial, int chunks) {
  __shared__ float sm[256];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = ceil_div(ceil_div(numel_per, 4), chunks) * 4;
  const int64_t lo = per * chunk, hi = lo + per < numel_per ? lo + per : numel_per;
  const float scale = 2.f / (float)numel_per;
  const float* e = est + (int64_t)b * numel_per;
  const float* t = target + (int64_t)b * target_bstride;
  float* g = d_est ? d_est + (int64_t)b * numel_per : nullptr;
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const float d = e[i] - t[i];
    acc = fmaf(d, d, acc);
    if (g) g[i] = d * scale;
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(int64_t)b * chunks + chunk] = sm[0];
}
__global__ void mse_final_kernel(const float* __restrict__ partial, float* __restrict__ loss, int batch,
                                 int chunks, float inv_n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  float acc = 0.f;
  for (int k = 0; k < chunks; ++k) acc += partial[(int64_t)b * chunks + k];
  loss[b] = acc * inv_n;
}

// Chunk count of the two-pass reduction.  It depends on numel_per ONLY (fixed chunk length of 16384 elements, at most
// 2048 chunks): the summation order of a trajectory's loss must not change with the batch it runs in, otherwise near-tied
// guesses could flip the per-image arg-min between differently sharded runs.  grid.y = batch supplies the parallelism.
static int mse_chunks(int batch, int64_t numel_per) {
  (void)batch;
  int64_t c = ceil_div(numel_per, 16384);
  if (c < 1) c = 1;
  if (c > 2048) c = 2048;
  return (int)c;
}

}  // namespace lfp

using namespace lfp;

extern "C" int lfp_embed_forward(const float* alpha, const float* key_logits, const float* U,
                                 const float* V, const float* sigma_key, const float* mu, float sd,
                                 int batch, int n_main, int key_len, int dim, float* w0, float* wx,
                                 void* stream) {
  LFP_CHECK_ARG(alpha && key_logits && U && V && sigma_key && mu && w0 && wx, "embed_forward: null argument");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535 && n_main >= 0 && key_len >= 0 && dim >= 1, "embed_forward: bad extent");
  dim3 grid((unsigned)ceil_div(dim, 128), (unsigned)batch);
  embed_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(alpha, key_logits, U, V, sigma_key, mu, sd, n_main, key_len, dim, w0, wx);
  LFP_LAUNCH_CHECK();
  return 0;
}

extern "C" int lfp_embed_backward(const float* d_wx, const float* key_logits, const float* U,
                                  const float* V, const float* sigma_key, float sd, int batch,
                                  int n_main, int key_len, int dim, float* d_alpha, float* d_key,
                                  void* stream) {
  LFP_CHECK_ARG(d_wx && key_logits && U && V && sigma_key && d_alpha && d_key, "embed_backward: null argument");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535, "embed_backward: bad batch");
  dim3 grid((unsigned)ceil_div(n_main + key_len, 8), (unsigned)batch);
  embed_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_wx, key_logits, U, V, sigma_key, sd, n_main, key_len, dim, d_alpha, d_key);
  LFP_LAUNCH_CHECK();
  return 0;
}

extern "C" int lfp_attrib_bound_loss(const float* alpha, const float* max_alpha, const float* min_alpha, const float* mse,
                                     int batch, int n_main, float weight, float* loss_total, void* stream) {
  LFP_CHECK_ARG(alpha && max_alpha && min_alpha && mse && loss_total && batch >= 1 && n_main >= 1, "attrib_bound_loss: bad argument");
  bound_loss_kernel<<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>(alpha, max_alpha, min_alpha, mse, n_main, weight, loss_total);
  LFP_LAUNCH_CHECK();
  return 0;
}

extern "C" int lfp_attrib_adam_update(const float* d_wx, float* alpha, float* key_logits, const float* U, const float* V,
                                      const float* sigma_key, const float* max_alpha, const float* min_alpha, float sd,
                                      float bound_weight, float* m_alpha, float* v_alpha, float* m_key, float* v_key,
                                      int batch, int n_main, int key_len, int dim, float step_size, float sqrt_bc2,
                                      float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                                      int optimise_alpha, void* stream) {
  LFP_CHECK_ARG(d_wx && alpha && key_logits && U && V && sigma_key && max_alpha && min_alpha && m_alpha && v_alpha && m_key && v_key,
                "attrib_adam_update: null argument");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535 && n_main >= 0 && key_len >= 1 && dim >= 1, "attrib_adam_update: bad shape");
  dim3 grid((unsigned)ceil_div(n_main + key_len, 8), (unsigned)batch);
  adam_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_wx, alpha, key_logits, U, V, sigma_key, max_alpha, min_alpha, sd,
                                                             bound_weight, m_alpha, v_alpha, m_key, v_key, n_main, key_len, dim,
                                                             step_size, sqrt_bc2, beta1, beta2, one_minus_beta1, one_minus_beta2, eps,
                                                             optimise_alpha, nullptr, nullptr);
  LFP_LAUNCH_CHECK();
  return 0;
}

namespace lfp {
// same kernel with step_size / sqrt_bc2 read from hyper[*step_ptr] (device memory): what a captured CUDA graph replays
int launch_adam_update_dev(const float* d_wx, float* alpha, float* key_logits, const float* U, const float* V, const float* sigma_key,
                           const float* max_alpha, const float* min_alpha, float sd, float bound_weight, float* m_alpha, float* v_alpha,
                           float* m_key, float* v_key, int batch, int n_main, int key_len, int dim, const float2* hyper,
                           const int* step_ptr, float beta1, float beta2, float omb1, float omb2, float eps, int optimise_alpha,
                           cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(n_main + key_len, 8), (unsigned)batch);
  adam_update_kernel<<<grid, 256, 0, s>>>(d_wx, alpha, key_logits, U, V, sigma_key, max_alpha, min_alpha, sd, bound_weight, m_alpha,
                                          v_alpha, m_key, v_key, n_main, key_len, dim, 0.f, 1.f, beta1, beta2, omb1, omb2, eps,
                                          optimise_alpha, hyper, step_ptr);
  LFP_LAUNCH_CHECK();
  return 0;
}
}  // namespace lfp

extern "C" size_t lfp_mse_scratch_bytes(int batch, int64_t numel_per) {
  if (batch <= 0 || numel_per <= 0) return 0;
  return (size_t)batch * mse_chunks(batch, numel_per) * sizeof(float);
}

extern "C" int lfp_mse_loss_grad(const float* est, const float* target, int target_batch, int batch,
                                 int64_t numel_per, float* loss, float* d_est, void* scratch,
                                 size_t scratch_bytes, void* stream) {
  LFP_CHECK_ARG(est && target && loss && scratch, "mse: null argument");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535 && numel_per >= 1, "mse: bad extent");
  LFP_CHECK_ARG(target_batch == 1 || target_batch == batch, "mse: target batch must be 1 or %d", batch);
  const int chunks = mse_chunks(batch, numel_per);
  if (scratch_bytes < (size_t)batch * chunks * sizeof(float)) { set_error("mse: scratch too small"); return LFP_ENOMEM; }
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)chunks, (unsigned)batch);
  mse_partial_kernel<<<grid, 256, 0, s>>>(est, target, target_batch == 1 ? 0 : numel_per, numel_per, d_est, (float*)scratch, chunks);
  LFP_LAUNCH_CHECK();
  mse_final_kernel<<<(unsigned)ceil_div(batch, 128), 128, 0, s>>>((const float*)scratch, loss, batch, chunks, 1.f / (float)numel_per);
  LFP_LAUNCH_CHECK();
  return 0;
}
