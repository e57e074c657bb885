// The whole Adam step of the attribution loop (src/main.py:57-72) as one native call, replayable as a CUDA graph
// (additive C-ABI group 6 of include/lfp_sg2.h):
//   w0 = U^T alpha + mu ; wx = w0 + sd V^T diag(sigma) sigmoid(key)          (src/main.py:60-61, src/generator.py:148-161)
//   est = G([wx], noise)                                                      (:62, lfp_synth_forward)
//   loss = MSE(target, est) + 0.1 alpha_bound(alpha)                          (:63-65, src/utils.py:46-58)
//   lr_i = lr0 exp(-0.001 (i + 1))                                            (:67, :42-43)
//   backward to (alpha, key) + Adam                                           (:69-70)
// Nothing in the step depends on a host value that changes from step to step: the schedule scalars of every step live
// in a device table indexed by a device step counter, all buffers are bound once.  So one step (~95 kernel launches) is
// captured once and replayed with cudaGraphLaunch: the host issues one call per step instead of ~95 launches + torch
// allocations, and consecutive steps queue back to back.
#include <math.h>
#include <vector>

#include "synth_kernels.cuh"

namespace lfp {
int launch_adam_update_dev(const float* d_wx, float* alpha, float* key_logits, const float* U, const float* V, const float* sigma_key,
                           const float* max_alpha, const float* min_alpha, float sd, float bound_weight, float* m_alpha, float* v_alpha,
                           float* m_key, float* v_key, int batch, int n_main, int key_len, int dim, const float2* hyper,
                           const int* step_ptr, float beta1, float beta2, float omb1, float omb2, float eps, int optimise_alpha,
                           cudaStream_t s);

// latent[b, slot, :] = wx[b, :]   (src/model.py:531-535: one latent for every layer)
__global__ void __launch_bounds__(256) broadcast_latent_kernel(const float* __restrict__ wx, float* __restrict__ latent, int n_latent, int dim, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int j = (int)(i % dim);
  const int64_t b = i / ((int64_t)dim * n_latent);
  latent[i] = wx[b * dim + j];
}
// d_wx[b, :] = sum over slots of d_latent[b, slot, :], ascending slot order (what d_latent.sum(1) of the repeat's backward does)
__global__ void __launch_bounds__(256) slot_sum_kernel(const float* __restrict__ d_latent, float* __restrict__ d_wx, int n_latent, int dim, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int j = (int)(i % dim);
  const int64_t b = i / dim;
  float acc = 0.f;
  for (int s = 0; s < n_latent; ++s) acc += d_latent[(b * n_latent + s) * dim + j];
  d_wx[i] = acc;
}
__global__ void step_inc_kernel(int* step) { *step += 1; }
}  // namespace lfp

using namespace lfp;

struct lfp_attrib {
  lfp_synth* plan = nullptr;
  int batch = 0, n_main = 0, key_len = 0, dim = 0, n_latent = 0, size = 0, precision = 0, max_steps = 0;
  const float *U = nullptr, *V = nullptr, *sigma_key = nullptr, *mu = nullptr, *max_alpha = nullptr, *min_alpha = nullptr;
  float sd = 1.f, bound_weight = 0.1f, lr0 = 0.2f;
  // bound buffers
  std::vector<const float*> noise; std::vector<int> noise_batch;
  const float* target = nullptr; int target_batch = 1;
  float *alpha = nullptr, *key = nullptr, *m_a = nullptr, *v_a = nullptr, *m_k = nullptr, *v_k = nullptr, *loss_total = nullptr;
  int optimise_alpha = 1;
  bool bound = false;
  // owned
  float2* hyper = nullptr; int* step = nullptr;
  // workspace carve-up (floats from the caller's workspace)
  float *w0 = nullptr, *wx = nullptr, *latent = nullptr, *image = nullptr, *d_image = nullptr, *mse = nullptr, *d_latent = nullptr, *d_wx = nullptr;
  void* mse_scratch = nullptr; size_t mse_scratch_bytes = 0;
  void* synth_ws = nullptr; size_t synth_ws_bytes = 0;
  lfp_lpips* lpips = nullptr; void* lpips_ws = nullptr; size_t lpips_ws_bytes = 0;   // optional perceptual loss instead of MSE
  cudaGraphExec_t exec = nullptr; cudaGraph_t graph = nullptr; bool warmed = false;
  ~lfp_attrib() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (hyper) cudaFree(hyper);
    if (step) cudaFree(step);
  }
};

namespace {
size_t up256(size_t v) { return (v + 255) / 256 * 256; }
struct StepLayout { size_t w0, wx, latent, image, d_image, mse, d_latent, d_wx, scratch, synth, total; };
StepLayout step_layout(const lfp_synth* plan, int B, int dim, int n_latent, int size) {
  StepLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = up256(off + bytes); return o; };
  const size_t img = (size_t)B * 3 * size * size * 4;
  L.w0 = take((size_t)B * dim * 4); L.wx = take((size_t)B * dim * 4);
  L.latent = take((size_t)B * n_latent * dim * 4);
  L.image = take(img); L.d_image = take(img);
  L.mse = take((size_t)B * 4);
  L.d_latent = take((size_t)B * n_latent * dim * 4); L.d_wx = take((size_t)B * dim * 4);
  L.scratch = take(lfp_mse_scratch_bytes(B, (int64_t)3 * size * size));
  L.synth = take(lfp_synth_workspace_bytes(plan, B));
  L.total = off;
  return L;
}

int enqueue_step(lfp_attrib* h, cudaStream_t s) {
  const int B = h->batch;
  LFP_TRY(lfp_embed_forward(h->alpha, h->key, h->U, h->V, h->sigma_key, h->mu, h->sd, B, h->n_main, h->key_len, h->dim, h->w0, h->wx, s));
  const int64_t nl = (int64_t)B * h->n_latent * h->dim;
  broadcast_latent_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, s>>>(h->wx, h->latent, h->n_latent, h->dim, nl);
  LFP_LAUNCH_CHECK();
  LFP_TRY(lfp_synth_forward(h->plan, B, h->latent, h->noise.data(), h->noise_batch.data(), h->image, h->synth_ws, h->synth_ws_bytes, h->precision, s));
  if (h->lpips != nullptr)   // src/main.py:63 with the reference's default loss; target features were cached by lfp_lpips_set_target
    LFP_TRY(lfp_lpips_loss_grad(h->lpips, B, h->image, h->mse, h->d_image, h->lpips_ws, h->lpips_ws_bytes, h->precision, s));
  else
    LFP_TRY(lfp_mse_loss_grad(h->image, h->target, h->target_batch, B, (int64_t)3 * h->size * h->size, h->mse, h->d_image, h->mse_scratch,
                              h->mse_scratch_bytes, s));
  LFP_TRY(lfp_synth_backward(h->plan, B, h->d_image, h->d_latent, h->synth_ws, h->synth_ws_bytes, h->precision, s));
  const int64_t nw = (int64_t)B * h->dim;
  slot_sum_kernel<<<(unsigned)ceil_div(nw, 256), 256, 0, s>>>(h->d_latent, h->d_wx, h->n_latent, h->dim, nw);
  LFP_LAUNCH_CHECK();
  LFP_TRY(lfp_attrib_bound_loss(h->alpha, h->max_alpha, h->min_alpha, h->mse, B, h->n_main, h->bound_weight, h->loss_total, s));
  const float b1 = 0.9f, b2 = 0.999f;
  LFP_TRY(launch_adam_update_dev(h->d_wx, h->alpha, h->key, h->U, h->V, h->sigma_key, h->max_alpha, h->min_alpha, h->sd, h->bound_weight,
                                 h->m_a, h->v_a, h->m_k, h->v_k, B, h->n_main, h->key_len, h->dim, h->hyper, h->step, b1, b2,
                                 (float)(1 - 0.9), (float)(1 - 0.999), 1e-8f, h->optimise_alpha, s));
  step_inc_kernel<<<1, 1, 0, s>>>(h->step);
  LFP_LAUNCH_CHECK();
  return 0;
}
}  // namespace

extern "C" int lfp_attrib_create(lfp_attrib** out, lfp_synth* plan, int size, int batch, int n_main, int key_len, int dim,
                                 const float* U, const float* V, const float* sigma_key, const float* mu, const float* max_alpha,
                                 const float* min_alpha, float sd, float bound_weight, double lr0, int max_steps, int precision) {
  LFP_CHECK_ARG(out && plan && U && V && sigma_key && mu && max_alpha && min_alpha, "attrib_create: null argument");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535 && n_main >= 0 && key_len >= 1 && dim >= 1 && max_steps >= 1, "attrib_create: bad extent");
  LFP_CHECK_ARG(precision == LFP_PREC_FP32 || precision == LFP_PREC_TF32, "attrib_create: unknown precision mode %d", precision);
  lfp_attrib* h = new lfp_attrib();
  h->plan = plan; h->size = size; h->batch = batch; h->n_main = n_main; h->key_len = key_len; h->dim = dim;
  h->n_latent = lfp_synth_n_latent(plan); h->precision = precision; h->max_steps = max_steps;
  h->U = U; h->V = V; h->sigma_key = sigma_key; h->mu = mu; h->max_alpha = max_alpha; h->min_alpha = min_alpha;
  h->sd = sd; h->bound_weight = bound_weight; h->lr0 = (float)lr0;
  // schedule table, computed as the Python driver does (double arithmetic, then one rounding to float):
  //   step_size = lr_i / (1 - beta1^t), sqrt_bc2 = sqrt(1 - beta2^t), t = i + 1   (torch.optim.Adam, src/main.py:42-43, 67)
  std::vector<float2> tab((size_t)max_steps);
  for (int i = 0; i < max_steps; ++i) {
    const double lr = lr0 * exp(-0.001 * (double)(i + 1));   // lr0 as a double: 0.2f widened is not the Python driver's 0.2
    const int t = i + 1;
    tab[i].x = (float)(lr / (1.0 - pow(0.9, (double)t)));
    tab[i].y = (float)sqrt(1.0 - pow(0.999, (double)t));
  }
  cudaError_t e = cudaMalloc((void**)&h->hyper, tab.size() * sizeof(float2));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->step, sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(h->hyper, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->step, 0, sizeof(int));
  if (e != cudaSuccess) { set_error("attrib_create: %s", cudaGetErrorString(e)); delete h; return (int)e; }
  *out = h;
  return 0;
}

extern "C" void lfp_attrib_destroy(lfp_attrib* h) { delete h; }

extern "C" size_t lfp_attrib_workspace_bytes(const lfp_attrib* h) {
  if (!h) return 0;
  return step_layout(h->plan, h->batch, h->dim, h->n_latent, h->size).total;
}

extern "C" int lfp_attrib_bind(lfp_attrib* h, const float* const* noise, const int* noise_batch, const float* target, int target_batch,
                               float* alpha, float* key_logits, float* m_alpha, float* v_alpha, float* m_key, float* v_key,
                               float* loss_total, int optimise_alpha, void* workspace, size_t workspace_bytes) {
  LFP_CHECK_ARG(h && noise && noise_batch && target && alpha && key_logits && m_alpha && v_alpha && m_key && v_key && loss_total && workspace,
                "attrib_bind: null argument");
  LFP_CHECK_ARG(target_batch == 1 || target_batch == h->batch, "attrib_bind: target batch must be 1 or %d", h->batch);
  LFP_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "attrib_bind: workspace must be 256-byte aligned");
  const StepLayout L = step_layout(h->plan, h->batch, h->dim, h->n_latent, h->size);
  if (workspace_bytes < L.total) { set_error("attrib_bind: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total); return LFP_ENOMEM; }
  const int nn = lfp_synth_num_noise(h->plan);
  h->noise.assign(noise, noise + nn); h->noise_batch.assign(noise_batch, noise_batch + nn);
  h->target = target; h->target_batch = target_batch;
  h->alpha = alpha; h->key = key_logits; h->m_a = m_alpha; h->v_a = v_alpha; h->m_k = m_key; h->v_k = v_key; h->loss_total = loss_total;
  h->optimise_alpha = optimise_alpha;
  unsigned char* w = (unsigned char*)workspace;
  h->w0 = (float*)(w + L.w0); h->wx = (float*)(w + L.wx); h->latent = (float*)(w + L.latent); h->image = (float*)(w + L.image);
  h->d_image = (float*)(w + L.d_image); h->mse = (float*)(w + L.mse); h->d_latent = (float*)(w + L.d_latent); h->d_wx = (float*)(w + L.d_wx);
  h->mse_scratch = w + L.scratch; h->mse_scratch_bytes = lfp_mse_scratch_bytes(h->batch, (int64_t)3 * h->size * h->size);
  h->synth_ws = w + L.synth; h->synth_ws_bytes = lfp_synth_workspace_bytes(h->plan, h->batch);
  if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }   // pointers are baked into a captured graph
  if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
  h->bound = true; h->warmed = false;
  return 0;
}

extern "C" int lfp_attrib_set_lpips(lfp_attrib* h, lfp_lpips* lpips, void* workspace, size_t workspace_bytes) {
  LFP_CHECK_ARG(h != nullptr, "attrib_set_lpips: null handle");
  LFP_CHECK_ARG(lpips == nullptr || (workspace != nullptr && workspace_bytes >= lfp_lpips_workspace_bytes(lpips, h->batch)),
                "attrib_set_lpips: workspace missing or smaller than lfp_lpips_workspace_bytes(batch)");
  h->lpips = lpips; h->lpips_ws = workspace; h->lpips_ws_bytes = workspace_bytes;
  if (h->exec) { cudaGraphExecDestroy(h->exec); h->exec = nullptr; }
  if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
  h->warmed = false;
  return 0;
}

extern "C" int lfp_attrib_set_step(lfp_attrib* h, int step, void* stream) {
  LFP_CHECK_ARG(h != nullptr && step >= 0 && step < h->max_steps, "attrib_set_step: step out of range");
  LFP_CUDA(cudaMemcpyAsync(h->step, &step, sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  LFP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // `step` is a stack variable
  return 0;
}

extern "C" int lfp_attrib_get_w0(const lfp_attrib* h, const float** w0, const float** wx) {
  LFP_CHECK_ARG(h && h->bound, "attrib_get_w0: not bound");
  if (w0) *w0 = h->w0;
  if (wx) *wx = h->wx;
  return 0;
}

extern "C" int lfp_attrib_run(lfp_attrib* h, int steps, int use_graph, void* stream) {
  LFP_CHECK_ARG(h != nullptr && steps >= 0, "attrib_run: bad argument");
  if (!h->bound) { set_error("attrib_run: lfp_attrib_bind has not been called"); return LFP_ESTATE; }
  cudaStream_t s = (cudaStream_t)stream;
  int done = 0;
  if (!use_graph) {
    for (; done < steps; ++done) LFP_TRY(enqueue_step(h, s));
    h->warmed = true;
    return 0;
  }
  LFP_CHECK_ARG(s != nullptr, "attrib_run: graph replay needs an explicit (non-default) stream");
  if (!h->warmed && steps > 0) {
    // first step eagerly: one-off work that must not happen inside a capture (kernel attributes, table uploads)
    LFP_TRY(enqueue_step(h, s));
    h->warmed = true; ++done;
  }
  if (h->exec == nullptr && done < steps) {
    LFP_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_step(h, s);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
    if (ce != cudaSuccess) { set_error("attrib_run: stream capture failed: %s", cudaGetErrorString(ce)); return (int)ce; }
    h->graph = g;
    LFP_CUDA(cudaGraphInstantiate(&h->exec, g, 0));
  }
  for (; done < steps; ++done) LFP_CUDA(cudaGraphLaunch(h->exec, s));
  return 0;
}
