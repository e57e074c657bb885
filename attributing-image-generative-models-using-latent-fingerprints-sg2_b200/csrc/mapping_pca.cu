// One-off set-up of the fingerprint basis on the GPU (additive C-ABI group 8 of include/lfp_sg2.h):
//   * the mapping network `style` = PixelNorm + n_mlp x (EqualLinear + fused bias-lrelu) on [n, dim] latents
//     (src/model.py:407-416, 132-161, 14-19; called on 10 000 samples by GetPCA.perform_pca, src/PCA.py:68-70),
//   * mean and covariance of the mapped latents in fp64 with a fixed summation order - the input of the PCA
//     (src/PCA.py:72-74: sklearn PCA().fit; explained_variance_ = eigenvalues of this covariance, components_ = its
//     eigenvectors).  The 512 x 512 symmetric eigendecomposition itself stays a library call (torch.linalg.eigh).
// Every linear layer runs as a one-tap "convolution" over n pixels on the same CUDA-core gather kernel as the fp32
// generator path, with the bias + leaky-ReLU * sqrt(2) epilogue (demod = 1, noise weight = 0).
#include <math.h>
#include <string>
#include <vector>

#include "synth_kernels.cuh"

namespace lfp {

// x * rsqrt(mean(x^2) + 1e-8) per row (src/model.py:14-19), one warp per row
__global__ void __launch_bounds__(256) pixel_norm_kernel(const float* __restrict__ z, float* __restrict__ out, int dim, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* p = z + row * dim;
  float ss = 0.f;
  for (int j = lane; j < dim; j += 32) ss = fmaf(p[j], p[j], ss);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float r = rsqrtf(ss / (float)dim + 1e-8f);
  for (int j = lane; j < dim; j += 32) out[row * dim + j] = p[j] * r;
}

// wt[k][n] = W[n][k] * scale  (EqualLinear: F.linear(x, W * scale), src/model.py:153-159)
__global__ void linear_prep_kernel(const float* __restrict__ W, float* __restrict__ wt, float scale, int dim) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)dim * dim) return;
  const int n = (int)(i / dim), k = (int)(i - (int64_t)n * dim);
  wt[(int64_t)k * dim + n] = W[i] * scale;
}

// mean[j] = sum_i w[i, j] / n in fp64, rows in ascending order within 64 fixed row-chunks, chunks summed in order
__global__ void __launch_bounds__(256) col_mean_kernel(const float* __restrict__ w, double* __restrict__ mean, int dim, int64_t n) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= dim) return;
  const int64_t per = (n + 63) / 64;
  double tot = 0.0;
  for (int c = 0; c < 64; ++c) {
    double acc = 0.0;
    const int64_t i1 = (c + 1) * per < n ? (c + 1) * per : n;
    for (int64_t i = c * per; i < i1; ++i) acc += (double)w[i * dim + j];
    tot += acc;
  }
  mean[j] = tot / (double)n;
}

// cov[a, b] = sum_i (w[i,a] - mean[a]) (w[i,b] - mean[b]) / (n - 1), fp64, 16 x 16 output tile per CTA, rows staged in
// shared memory 64 at a time; each output element is summed by one thread in ascending row order
__global__ void __launch_bounds__(256) covariance_kernel(const float* __restrict__ w, const double* __restrict__ mean,
                                                         double* __restrict__ cov, int dim, int64_t n) {
  __shared__ double sa[64][17], sb[64][17];
  const int ta = threadIdx.x >> 4, tb = threadIdx.x & 15;
  const int a0 = blockIdx.y * 16, b0 = blockIdx.x * 16;
  double acc = 0.0;
  for (int64_t r0 = 0; r0 < n; r0 += 64) {
    for (int t = threadIdx.x; t < 64 * 16; t += 256) {
      const int rr = t >> 4, cc = t & 15;
      const int64_t r = r0 + rr;
      sa[rr][cc] = (r < n && a0 + cc < dim) ? (double)w[r * dim + a0 + cc] - mean[a0 + cc] : 0.0;
      sb[rr][cc] = (r < n && b0 + cc < dim) ? (double)w[r * dim + b0 + cc] - mean[b0 + cc] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) acc = fma(sa[rr][ta], sb[rr][tb], acc);
    __syncthreads();
  }
  if (a0 + ta < dim && b0 + tb < dim) cov[(int64_t)(a0 + ta) * dim + b0 + tb] = acc / (double)(n - 1);
}

}  // namespace lfp

using namespace lfp;

struct lfp_mapping {
  int dim = 0, n_mlp = 0; float lr_mul = 0.01f;
  std::vector<float*> W, bias, wt, bs;   // raw [dim, dim] / [dim]; prepared transposed-scaled / bias * lr_mul
  float *ones = nullptr, *zero = nullptr;
  bool finalized = false;
  std::vector<void*> owned;
  ~lfp_mapping() { for (void* p : owned) cudaFree(p); }
  int alloc(float** p, size_t n) { LFP_CUDA(cudaMalloc((void**)p, (n ? n : 1) * sizeof(float))); owned.push_back(*p); return 0; }
};

extern "C" int lfp_mapping_create(lfp_mapping** out, int dim, int n_mlp, float lr_mul) {
  LFP_CHECK_ARG(out && dim >= 16 && dim % 16 == 0 && n_mlp >= 1 && n_mlp <= 64, "mapping_create: dim must be a multiple of 16, 1 <= n_mlp <= 64");
  lfp_mapping* h = new lfp_mapping();
  h->dim = dim; h->n_mlp = n_mlp; h->lr_mul = lr_mul;
  int rc = 0;
  h->W.resize(n_mlp); h->bias.resize(n_mlp); h->wt.resize(n_mlp); h->bs.resize(n_mlp);
  for (int i = 0; i < n_mlp; ++i) {
    rc |= h->alloc(&h->W[i], (size_t)dim * dim); rc |= h->alloc(&h->bias[i], dim);
    rc |= h->alloc(&h->wt[i], (size_t)dim * dim); rc |= h->alloc(&h->bs[i], dim);
  }
  rc |= h->alloc(&h->ones, dim); rc |= h->alloc(&h->zero, 1);
  if (rc != 0) { set_error("mapping_create: device allocation failed"); delete h; return LFP_ENOMEM; }
  std::vector<float> one((size_t)dim, 1.f);
  float z = 0.f;
  if (cudaMemcpy(h->ones, one.data(), dim * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->zero, &z, sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("mapping_create: copy failed"); delete h; return LFP_ENOMEM; }
  *out = h;
  return 0;
}

extern "C" void lfp_mapping_destroy(lfp_mapping* h) { delete h; }

extern "C" int lfp_mapping_set_param(lfp_mapping* h, const char* name, const float* data, int64_t numel, void* stream) {
  LFP_CHECK_ARG(h && name && data, "mapping_set_param: null argument");
  const std::string n(name);
  float* dst = nullptr; int64_t want = -1;
  for (int i = 0; i < h->n_mlp; ++i) {   // Generator.style = Sequential(PixelNorm, EqualLinear x n_mlp): names style.1 .. style.n_mlp
    if (n == "style." + std::to_string(i + 1) + ".weight") { dst = h->W[i]; want = (int64_t)h->dim * h->dim; }
    if (n == "style." + std::to_string(i + 1) + ".bias") { dst = h->bias[i]; want = h->dim; }
  }
  LFP_CHECK_ARG(dst != nullptr, "mapping_set_param: unknown parameter '%s' (style.<1..n_mlp>.weight|bias)", name);
  LFP_CHECK_ARG(want == numel, "mapping_set_param: '%s' expects %lld elements, got %lld", name, (long long)want, (long long)numel);
  LFP_CUDA(cudaMemcpyAsync(dst, data, numel * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  h->finalized = false;
  return 0;
}

extern "C" int lfp_mapping_finalize(lfp_mapping* h, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "mapping_finalize: null handle");
  cudaStream_t s = (cudaStream_t)stream;
  const float scale = (1.f / sqrtf((float)h->dim)) * h->lr_mul;   // src/model.py:148
  for (int i = 0; i < h->n_mlp; ++i) {
    linear_prep_kernel<<<(unsigned)ceil_div((int64_t)h->dim * h->dim, 256), 256, 0, s>>>(h->W[i], h->wt[i], scale, h->dim);
    LFP_LAUNCH_CHECK();
    LFP_TRY(launch_scale_copy(h->bias[i], h->bs[i], h->lr_mul, h->dim, s));   // fused_leaky_relu(out, bias * lr_mul), :154
  }
  h->finalized = true;
  return 0;
}

// w_out [n, dim] = style(z [n, dim]); scratch: 2 * n * dim floats
extern "C" int lfp_mapping_forward(lfp_mapping* h, const float* z, int64_t n, float* w_out, void* scratch, size_t scratch_bytes, void* stream) {
  LFP_CHECK_ARG(h && z && w_out && scratch && n >= 1 && n < (1ll << 31), "mapping_forward: bad argument");
  if (!h->finalized) { set_error("mapping_forward: lfp_mapping_finalize has not been called since the last set_param"); return LFP_ESTATE; }
  if (scratch_bytes < 2 * (size_t)n * h->dim * sizeof(float)) { set_error("mapping_forward: scratch too small"); return LFP_ENOMEM; }
  cudaStream_t s = (cudaStream_t)stream;
  float* buf[2] = {(float*)scratch, (float*)scratch + (size_t)n * h->dim};
  pixel_norm_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, s>>>(z, buf[0], h->dim, n);
  LFP_LAUNCH_CHECK();
  ConvGeom g{};
  g.batch = 1; g.gh = (int)n; g.gw = 1; g.in_h = (int)n; g.in_w = 1; g.in_bstride = 0; g.in_stride = 1;
  g.out_h = (int)n; g.out_w = 1; g.out_stride = 1; g.K = h->dim; g.N = h->dim; g.ntaps = 1; g.dy[0] = g.dx[0] = 0; g.widx[0] = 0;
  int cur = 0;
  for (int i = 0; i < h->n_mlp; ++i) {
    ConvEpiArgs e;
    e.demod = h->ones; e.noise = z; e.noise_bstride = 0; e.noise_w = h->zero; e.bias = h->bs[i];   // lrelu(acc + bias) * sqrt 2
    float* o = i + 1 == h->n_mlp ? w_out : buf[cur ^ 1];
    LFP_TRY(launch_conv_simt(buf[cur], nullptr, h->wt[i], o, g, EPI_ACT, e, s));
    cur ^= 1;
  }
  return 0;
}

// mean [dim] and covariance [dim, dim] (unbiased, n - 1) of w [n, dim], fp64 outputs
extern "C" int lfp_pca_covariance(const float* w, int64_t n, int dim, double* mean, double* cov, void* stream) {
  LFP_CHECK_ARG(w && mean && cov && n >= 2 && dim >= 1, "pca_covariance: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  col_mean_kernel<<<(unsigned)ceil_div(dim, 256), 256, 0, s>>>(w, mean, dim, n);
  LFP_LAUNCH_CHECK();
  dim3 grid((unsigned)ceil_div(dim, 16), (unsigned)ceil_div(dim, 16));
  covariance_kernel<<<grid, 256, 0, s>>>(w, mean, cov, dim, n);
  LFP_LAUNCH_CHECK();
  return 0;
}
