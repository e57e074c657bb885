/*
 * lfp_sg2.h - C ABI of the B200-native latent-fingerprint (StyleGAN2 synthesis) hot path.
 *
 * One shared library, liblfp_sg2.so, built by nvcc for sm_100a with no torch / ATen
 * dependency.  Every entry point takes plain pointers and sizes; device pointers are
 * raw CUDA device addresses, `stream` is a cudaStream_t passed as void* (NULL = legacy
 * default stream).  All functions return 0 on success, a positive cudaError_t value
 * when a CUDA call failed, or a negative LFP_E* code for argument errors;
 * lfp_last_error() returns a thread-local description of the last failure.
 * Launches are asynchronous on `stream` unless the name ends in `_host`.
 *
 * The first two groups are exactly what the reference's FFI for this path binds
 * (its two pybind11 modules); the third group is additive: the fused whole-synthesis
 * forward / backward that replaces the Python module chain in src/model.py.
 * Citations are relative to /root/reference/.
 */
#ifndef LFP_SG2_H_
#define LFP_SG2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFP_OK 0
#define LFP_EINVAL (-1)    /* bad argument (shape, dtype, null pointer)            */
#define LFP_ESTATE (-2)    /* call sequence error (e.g. backward without forward) */
#define LFP_ENOMEM (-3)    /* workspace too small                                 */
#define LFP_EUNSUPPORTED (-4)

/* element types accepted by the two legacy ops (the reference dispatches
 * AT_DISPATCH_FLOATING_TYPES_AND_HALF, src/op/upfirdn2d_kernel.cu:310,
 * src/op/fused_bias_act_kernel.cu:96) */
#define LFP_F32 0
#define LFP_F64 1
#define LFP_F16 2

const char* lfp_last_error(void);
int lfp_version(void);
/* number of kernels this library has launched since load (all entry points) */
uint64_t lfp_launch_count(void);

/* ---------------------------------------------------------------------------------
 * 1. upfirdn2d  -  replaces `upfirdn2d(input, kernel, up_x, up_y, down_x, down_y,
 *    pad_x0, pad_x1, pad_y0, pad_y1)` of module "upfirdn2d"
 *    (src/op/upfirdn2d.cpp:17-31 -> upfirdn2d_op, src/op/upfirdn2d_kernel.cu:209-369).
 *    input  [major, in_h, in_w, minor] contiguous, kernel [kernel_h, kernel_w] same dtype,
 *    out    [major, out_h, out_w, minor] with
 *           out_h = (in_h*up_y + pad_y0 + pad_y1 - kernel_h + down_y) / down_y (likewise w).
 *    The FIR is applied flipped (true convolution); negative pads crop.  64-bit indexing
 *    throughout (the reference overflows int at 2^31 elements).
 * --------------------------------------------------------------------------------- */
int lfp_upfirdn2d_out_size(int in_h, int in_w, int kernel_h, int kernel_w, int up_x, int up_y,
                           int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                           int* out_h, int* out_w);

int lfp_upfirdn2d(const void* input, const void* kernel, void* out, int dtype, int64_t major,
                  int in_h, int in_w, int64_t minor, int kernel_h, int kernel_w, int up_x, int up_y,
                  int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                  void* stream);

/* same op on HOST buffers: allocates device scratch, copies in, runs, copies out, syncs */
int lfp_upfirdn2d_host(const void* input, const void* kernel, void* out, int dtype, int64_t major,
                       int in_h, int in_w, int64_t minor, int kernel_h, int kernel_w, int up_x,
                       int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                       int pad_y1);

/* ---------------------------------------------------------------------------------
 * 2. fused_bias_act  -  replaces `fused_bias_act(input, bias, refer, act, grad, alpha,
 *    scale)` of module "fused" (src/op/fused_bias_act.cpp:18-32 -> fused_bias_act_op,
 *    src/op/fused_bias_act_kernel.cu:67-104).
 *    x[size_x] contiguous; bias[size_b] or NULL ("empty tensor"); ref[size_x] or NULL;
 *    bias index of element i is (i / step_b) % size_b, step_b = prod(dims >= 2).
 *    act*10+grad: 10,11 -> y=x; 30 -> x>0?x:x*alpha; 31 -> ref>0?x:x*alpha; 12,32 -> 0;
 *    out = y*scale.
 * --------------------------------------------------------------------------------- */
int lfp_fused_bias_act(const void* x, const void* bias, const void* ref, void* out, int dtype,
                       int64_t size_x, int64_t step_b, int64_t size_b, int act, int grad,
                       float alpha, float scale, void* stream);

int lfp_fused_bias_act_host(const void* x, const void* bias, const void* ref, void* out, int dtype,
                            int64_t size_x, int64_t step_b, int64_t size_b, int act, int grad,
                            float alpha, float scale);

/* grad_bias = sum over all dims but 1 of grad_input (src/op/fused_act.py:34-40), fp32 only,
 * deterministic (fixed summation order).  g [outer, size_b, step_b]; out [size_b]. */
int lfp_bias_grad_reduce(const float* g, float* out, int64_t outer, int64_t size_b, int64_t step_b,
                         void* scratch, size_t scratch_bytes, void* stream);
size_t lfp_bias_grad_reduce_scratch(int64_t outer, int64_t size_b, int64_t step_b);

/* ---------------------------------------------------------------------------------
 * 3. Whole-synthesis plan (additive).  Replaces the module chain executed by
 *    Generator.forward with input_is_latent=True (src/model.py:551-566): ConstantInput,
 *    StyledConv (ModulatedConv2d :258-302 + NoiseInjection :311-316 + FusedLeakyReLU),
 *    ToRGB (:379-388) incl. Blur / Upsample, and its autograd backward to the latent.
 *    Generator parameters are frozen (the attribution loop discards their gradients,
 *    src/main.py:58); only d(image)/d(latent) is produced.
 *
 *    Activations are kept NHWC fp32 inside the workspace; the public tensors keep the
 *    reference layouts: latent [B, n_latent, style_dim], noise_i [nb, 1, h, h] with nb in
 *    {1, B}, image / d_image [B, 3, size, size] (NCHW), d_latent [B, n_latent, style_dim].
 * --------------------------------------------------------------------------------- */
typedef struct lfp_synth lfp_synth;

/* precision modes for the 3x3 convolutions */
#define LFP_PREC_FP32 0 /* CUDA-core FFMA, fp32 operands and accumulation                  */
#define LFP_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 storage, fp32 accumulation in TMEM      */

int lfp_synth_create(lfp_synth** out, int size, int style_dim, int channel_multiplier,
                     const float* blur_kernel_1d, int blur_taps);
void lfp_synth_destroy(lfp_synth* h);

int lfp_synth_n_latent(const lfp_synth* h);   /* 2*log2(size)-2  (src/model.py:474)       */
int lfp_synth_num_noise(const lfp_synth* h);  /* 2*(log2(size)-2)+1 (src/model.py:437-438) */

/* Upload one tensor by its reference state_dict name, e.g.
 * "conv1.conv.weight", "convs.3.conv.modulation.bias", "to_rgbs.2.bias", "input.input".
 * `data` is a DEVICE pointer to fp32, `numel` must match the expected shape.  Unknown
 * names (mapping network, stored noises, FIR buffers) return LFP_EINVAL. */
int lfp_synth_set_param(lfp_synth* h, const char* name, const float* data, int64_t numel,
                        void* stream);
/* Build derived tables (scaled / re-laid-out weights, sum-of-squares, stacked modulation
 * matrix).  Must be called after the last set_param and before forward. */
int lfp_synth_finalize(lfp_synth* h, void* stream);

size_t lfp_synth_workspace_bytes(const lfp_synth* h, int batch);

/* noise: host array of lfp_synth_num_noise() device pointers; noise_batch[i] in {1, batch}.
 * The workspace must stay untouched between forward and the matching backward. */
int lfp_synth_forward(lfp_synth* h, int batch, const float* latent, const float* const* noise,
                      const int* noise_batch, float* image, void* workspace,
                      size_t workspace_bytes, int precision, void* stream);
/* The backward is validated against the forward that last ran on THIS workspace (batch, precision, noise pointers are
 * recorded per workspace), so two workspaces of one plan can be interleaved freely. */
int lfp_synth_backward(lfp_synth* h, int batch, const float* d_image, float* d_latent,
                       void* workspace, size_t workspace_bytes, int precision, void* stream);

/* Forward only (fingerprinted generation, src/generator.py:185-198: no backward follows): same computation and bits as
 * lfp_synth_forward, but activations and skip images ping-pong between two buffers instead of being kept, so the
 * workspace is ~3 activation maps instead of all of them (27 GB instead of 93 GB for batch 64 at 1024 px).
 * lfp_synth_backward / lfp_synth_read_activation refuse a workspace whose last forward was this call. */
size_t lfp_synth_generate_workspace_bytes(const lfp_synth* h, int batch);
int lfp_synth_generate(lfp_synth* h, int batch, const float* latent, const float* const* noise,
                       const int* noise_batch, float* image, void* workspace, size_t workspace_bytes,
                       int precision, void* stream);

/* Inspection: copy the saved output of StyledConv `conv_index` (0 = conv1, 1 + i = convs.i,
 * src/model.py:552-563) of the last forward on `workspace` to `out` as [B, C, res, res] (NCHW,
 * device pointer).  Returns the channel count and resolution through `channels` / `res` (either
 * may be NULL; `out` may be NULL to query only).  Parity tests use it to compare leaky-ReLU
 * branch patterns with the oracle. */
int lfp_synth_num_convs(const lfp_synth* h);
int lfp_synth_read_activation(lfp_synth* h, int batch, int conv_index, const void* workspace,
                              float* out, int* channels, int* res, void* stream);

/* Host-buffer convenience (what a non-torch caller binds): latent / noise / image / d_image /
 * d_latent are HOST pointers; copies are inside the call; synchronous.  d_image may be NULL
 * (forward only: runs lfp_synth_generate).  The device staging buffers and the workspace belong to the
 * plan and only grow, so calling this in a loop allocates nothing after the first iteration. */
int lfp_synth_forward_backward_host(lfp_synth* h, int batch, const float* latent,
                                    const float* const* noise, const int* noise_batch,
                                    float* image, const float* d_image, float* d_latent,
                                    int precision);

/* Per-kernel-class timing with CUDA events recorded on the launching stream (bench.py's
 * roofline leg).  profile_begin(mask) starts recording every launch whose class bit is set;
 * profile_end synchronises the recorded events and returns, per class, the summed device time
 * in ms, the launch count and the algorithmic flops / bytes of those launches (arrays of
 * LFP_KIND_COUNT entries). */
#define LFP_KIND_CONV_FWD 0
#define LFP_KIND_CONV_DGRAD 1
#define LFP_KIND_FIR 2
#define LFP_KIND_TORGB 3
#define LFP_KIND_ACTBWD 4
#define LFP_KIND_COUNT 5
int lfp_synth_profile_begin(lfp_synth* h, int kind_mask);
int lfp_synth_profile_end(lfp_synth* h, double* ms, int64_t* launches, double* flops, double* bytes);
/* After profile_end: the recorded launches one by one, in launch order (class, device ms, algorithmic
 * flops and bytes); returns how many were recorded, fills at most max_launches entries. */
int lfp_synth_profile_launches(lfp_synth* h, int max_launches, int* kinds, float* ms, double* flops, double* bytes);

/* ---------------------------------------------------------------------------------
 * 4. Fingerprint embed + loss glue used by the attribution loop (additive).
 *    w0 = U^T alpha + mu (src/main.py:60); wx = w0 + sd * V^T diag(sigma) sigmoid(key)
 *    (src/generator.py:148-161, src/main.py:61), batched over B trajectories, row-major:
 *    alpha [B, n_main], key_logits [B, key_len], U [n_main, dim], V [key_len, dim],
 *    sigma_key [key_len], mu [dim]; outputs w0, wx [B, dim].
 *    Backward: given d_wx [B, dim] -> d_alpha [B, n_main], d_key [B, key_len].
 * --------------------------------------------------------------------------------- */
int lfp_embed_forward(const float* alpha, const float* key_logits, const float* U, const float* V,
                      const float* sigma_key, const float* mu, float sd, int batch, int n_main,
                      int key_len, int dim, float* w0, float* wx, void* stream);
int lfp_embed_backward(const float* d_wx, const float* key_logits, const float* U, const float* V,
                       const float* sigma_key, float sd, int batch, int n_main, int key_len,
                       int dim, float* d_alpha, float* d_key, void* stream);

/* mean-squared-error per trajectory against a shared or per-trajectory target, and its
 * gradient (src/utils.py:46-47): loss[b] = mean((est[b]-target[b or 0])^2),
 * d_est = 2*(est-target)/numel_per.  Deterministic two-pass reduction. */
int lfp_mse_loss_grad(const float* est, const float* target, int target_batch, int batch,
                      int64_t numel_per, float* loss, float* d_est, void* scratch,
                      size_t scratch_bytes, void* stream);
size_t lfp_mse_scratch_bytes(int batch, int64_t numel_per);

/* Step glue of the hot loop (src/main.py:65-70), batched over trajectories:
 *   loss_total[b] = mse[b] + weight * alpha_bound(alpha[b], max_alpha, min_alpha)            (src/utils.py:53-58)
 *   one torch.optim.Adam step (defaults beta 0.9 / 0.999, eps 1e-8) of alpha and the key logits, in place, from
 *   d(loss)/d(wx): the embed backward (as lfp_embed_backward), the bound's sub-gradient and the Adam update in one
 *   kernel, in torch's operation order.  step_size = lr_t / (1 - beta1^t), sqrt_bc2 = sqrt(1 - beta2^t), t = 1-based
 *   step; one_minus_beta* are passed in (1 - 0.9 evaluated in float is not the float nearest to 0.1).
 *   optimise_alpha = 0 freezes alpha (the key-only fixture). */
int lfp_attrib_bound_loss(const float* alpha, const float* max_alpha, const float* min_alpha, const float* mse,
                          int batch, int n_main, float weight, float* loss_total, void* stream);
int lfp_attrib_adam_update(const float* d_wx, float* alpha, float* key_logits, const float* U, const float* V,
                           const float* sigma_key, const float* max_alpha, const float* min_alpha, float sd,
                           float bound_weight, float* m_alpha, float* v_alpha, float* m_key, float* v_key,
                           int batch, int n_main, int key_len, int dim, float step_size, float sqrt_bc2,
                           float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                           int optimise_alpha, void* stream);

/* ---------------------------------------------------------------------------------
 * 6. The whole Adam step of the attribution loop (src/main.py:57-72) as one native call, replayable as a CUDA graph
 *    (additive): embed -> lfp_synth_forward -> MSE -> lfp_synth_backward -> bound loss -> Adam, for `batch` trajectories.
 *    All buffers are bound once (lfp_attrib_bind); the per-step schedule scalars lr_i / (1 - beta1^t), sqrt(1 - beta2^t)
 *    (src/main.py:42-43, 67; torch.optim.Adam defaults) come from a device table indexed by a device step counter, so a
 *    captured step can be replayed unchanged.  lfp_attrib_run(h, n, use_graph = 1, stream) runs the first step eagerly,
 *    captures the second and replays it for the rest: one host call per step instead of ~95 launches.  Results are
 *    bit-identical to driving the group 3/4 entry points one by one.
 *    U [n_main, dim], V [key_len, dim], sigma_key [key_len], mu [dim], max_alpha / min_alpha [n_main] (device, borrowed
 *    for the handle's lifetime); state alpha [B, n_main], key_logits [B, key_len] and the Adam moments are updated in
 *    place; loss_total [B] holds the loss of the last step run.
 * --------------------------------------------------------------------------------- */
typedef struct lfp_attrib lfp_attrib;
int lfp_attrib_create(lfp_attrib** out, lfp_synth* plan, int size, int batch, int n_main, int key_len, int dim,
                      const float* U, const float* V, const float* sigma_key, const float* mu, const float* max_alpha,
                      const float* min_alpha, float sd, float bound_weight, double lr0, int max_steps, int precision);
void lfp_attrib_destroy(lfp_attrib* h);
size_t lfp_attrib_workspace_bytes(const lfp_attrib* h);
int lfp_attrib_bind(lfp_attrib* h, const float* const* noise, const int* noise_batch, const float* target, int target_batch,
                    float* alpha, float* key_logits, float* m_alpha, float* v_alpha, float* m_key, float* v_key,
                    float* loss_total, int optimise_alpha, void* workspace, size_t workspace_bytes);
/* Use the perceptual loss (group 7) instead of the MSE in the step: `lpips` must have had lfp_lpips_set_target called with
 * this step's target; `workspace` >= lfp_lpips_workspace_bytes(lpips, batch).  NULL switches back to the MSE. */
struct lfp_lpips;
int lfp_attrib_set_lpips(lfp_attrib* h, struct lfp_lpips* lpips, void* workspace, size_t workspace_bytes);
int lfp_attrib_set_step(lfp_attrib* h, int step, void* stream);   /* index of the next step (0-based) */
int lfp_attrib_get_w0(const lfp_attrib* h, const float** w0, const float** wx);   /* latents of the last step, [B, dim] */
int lfp_attrib_run(lfp_attrib* h, int steps, int use_graph, void* stream);

/* ---------------------------------------------------------------------------------
 * 5. Stand-alone modulated convolution (additive).  Replaces ModulatedConv2d.forward(input, style)
 *    (src/model.py:169-302; fused branch :258-302, algebra of the unfused branch :229-256) for ONE layer and its
 *    autograd backward to the input and the style: plain k x k (k = 1: ToRGB's conv, no demodulation; k = 3), or the
 *    3x3 stride-2 transposed conv followed by Blur (upsample = 1, :269-282).  What model.ModulatedConv2d / StyledConv /
 *    ToRGB run when they are used outside Generator.forward.
 *    input [B, Cin, H, W], style [B, style_dim], out / d_out [B, Cout, H', W'] (H' = 2H when upsample), d_input like
 *    input, d_style like style; all NCHW fp32 device pointers.  Parameters by name: "weight" [1, Cout, Cin, k, k],
 *    "modulation.weight" [Cin, style_dim], "modulation.bias" [Cin]; they are frozen constants (no weight gradients).
 *    Any channel count works (padded inside); shapes the tcgen05 kernel tiles run on it when precision = LFP_PREC_TF32.
 *    The workspace keeps what the backward needs; one forward in flight per handle.
 * --------------------------------------------------------------------------------- */
typedef struct lfp_modconv lfp_modconv;
int lfp_modconv_create(lfp_modconv** out, int in_channel, int out_channel, int kernel_size, int style_dim,
                       int demodulate, int upsample, const float* blur_kernel_1d, int blur_taps);
void lfp_modconv_destroy(lfp_modconv* h);
int lfp_modconv_set_param(lfp_modconv* h, const char* name, const float* data, int64_t numel, void* stream);
int lfp_modconv_finalize(lfp_modconv* h, void* stream);
int lfp_modconv_out_size(const lfp_modconv* h, int in_h, int in_w, int* out_h, int* out_w);
size_t lfp_modconv_workspace_bytes(const lfp_modconv* h, int batch, int in_h, int in_w);
int lfp_modconv_forward(lfp_modconv* h, int batch, int in_h, int in_w, const float* input, const float* style,
                        float* out, void* workspace, size_t workspace_bytes, int precision, void* stream);
/* d_input or d_style may be NULL when that gradient is not needed */
int lfp_modconv_backward(lfp_modconv* h, int batch, int in_h, int in_w, const float* d_out, float* d_input,
                         float* d_style, void* workspace, size_t workspace_bytes, int precision, void* stream);

/* ---------------------------------------------------------------------------------
 * 7. Perceptual loss (additive).  Replaces `percept(target, est)` (src/utils.py:16, 44-50) =
 *    PNetLin.forward, LPIPS v0.1 with the VGG16 backbone and linear heads (src/custom_lpips/networks_basic.py:27-91,
 *    pretrained_networks.py:97-135), forward and backward to the estimated image.  The target's features are computed
 *    once (lfp_lpips_set_target) instead of every step as the reference does (networks_basic.py:66).
 *    Parameters by PNetLin state_dict name: "net.slice<S>.<I>.weight" [Cout, Cin, 3, 3] / ".bias" [Cout] for the 13
 *    convolutions (I = torchvision vgg16.features index 0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28 in slices 1..5) and
 *    "lin<K>.model.1.weight" [1, C_K, 1, 1], K = 0..4.  Images are [B, 3, H, W] NCHW in [-1, 1], H and W multiples of 16.
 *    loss[b] = LPIPS(target[b or 0], est[b]); d_est = d loss[b] / d est[b] (may be NULL: value only).
 * --------------------------------------------------------------------------------- */
typedef struct lfp_lpips lfp_lpips;
int lfp_lpips_create(lfp_lpips** out, int height, int width);
void lfp_lpips_destroy(lfp_lpips* h);
int lfp_lpips_set_param(lfp_lpips* h, const char* name, const float* data, int64_t numel, void* stream);
int lfp_lpips_finalize(lfp_lpips* h, void* stream);
size_t lfp_lpips_workspace_bytes(const lfp_lpips* h, int batch);
int lfp_lpips_set_target(lfp_lpips* h, int target_batch, const float* target, void* workspace, size_t workspace_bytes,
                         int precision, void* stream);
int lfp_lpips_loss_grad(lfp_lpips* h, int batch, const float* est, float* loss, float* d_est, void* workspace,
                        size_t workspace_bytes, int precision, void* stream);

/* ---------------------------------------------------------------------------------
 * 8. Set-up of the fingerprint basis (additive): the mapping network Generator.style = PixelNorm +
 *    n_mlp x (EqualLinear(lr_mul) + fused bias-lrelu) (src/model.py:407-416) on [n, dim] latents - what
 *    GetPCA.perform_pca runs on 10 000 samples (src/PCA.py:68-70) - and the fp64 mean / covariance of the mapped latents
 *    whose eigen-decomposition is the PCA (src/PCA.py:72-74).  Parameters by state_dict name "style.<i>.weight" [dim, dim]
 *    / "style.<i>.bias" [dim], i = 1..n_mlp.  lfp_mapping_forward needs 2 * n * dim floats of scratch.
 * --------------------------------------------------------------------------------- */
typedef struct lfp_mapping lfp_mapping;
int lfp_mapping_create(lfp_mapping** out, int dim, int n_mlp, float lr_mul);
void lfp_mapping_destroy(lfp_mapping* h);
int lfp_mapping_set_param(lfp_mapping* h, const char* name, const float* data, int64_t numel, void* stream);
int lfp_mapping_finalize(lfp_mapping* h, void* stream);
int lfp_mapping_forward(lfp_mapping* h, const float* z, int64_t n, float* w_out, void* scratch, size_t scratch_bytes, void* stream);
int lfp_pca_covariance(const float* w, int64_t n, int dim, double* mean, double* cov, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LFP_SG2_H_ */
