// Kernel launchers of the fused synthesis path (declarations).  Activations are NHWC fp32.
#pragma once
#include "common.cuh"

namespace lfp {

constexpr float kLreluSlope = 0.2f;
constexpr float kLreluGain = 1.4142135623730951f;  // 2 ** 0.5 (src/op/fused_act.py:93)

// Gather convolution over a flattened grid of output pixels:
//   acc[b, gy, gx, n] = sum_t sum_k in[b, gy*in_stride + dy[t], gx*in_stride + dx[t], k]
//                                   * (mod ? mod[b, k] : 1) * wtab[widx[t]][k][n]
// (out-of-range input reads are zero), written at output pixel
// (gy*out_stride + out_oy, gx*out_stride + out_ox) of an [B, out_h, out_w, N] tensor.
// This one shape covers the plain 3x3 conv, the four sub-pixel phases of the stride-2
// transposed conv, the plain data-gradient and the stride-2 data-gradient.
struct ConvGeom {
  int batch;
  int gh, gw;            // grid pixels per sample
  int in_h, in_w;        // input extent
  int64_t in_bstride;    // elements between samples of `in` (0: broadcast, e.g. the constant input)
  int in_stride;         // 1 or 2
  int out_h, out_w, out_stride, out_oy, out_ox;
  int K, N;              // reduction / output channels
  int ntaps;
  signed char dy[9], dx[9], widx[9];
};

enum ConvEpilogue {
  EPI_STORE = 0,  // raw accumulator
  EPI_ACT = 1,    // *demod[b,n] + noise_w*noise[b?,gy,gx] + bias[n] -> lrelu*sqrt2   (src/model.py:360-366)
  EPI_DGRAD = 2,  // dx = acc*mod_out[b,n];  partial[seg,n] = sum_pix xsave*acc       (style gradient)
  // EPI_DGRAD followed, in the same epilogue, by the backward through noise/bias/lrelu (+ ToRGB branch) of the layer
  // that produced xsave (what act_bwd_kernel does as a separate pass): out = gpre*demod, partial_T, partial_R.
  // Tensor-core kernel only.
  EPI_DGRAD_ACT = 3,
  // plain (unmodulated) convolution layers of the perceptual loss's VGG16 backbone (src/custom_lpips/pretrained_networks.py:97-135):
  EPI_RELU = 4,        // max(acc + bias[n], 0)
  EPI_DGRAD_RELU = 5,  // acc * (xsave > 0): data gradient through the ReLU that produced this layer's input
};

struct ConvEpiArgs {
  // EPI_ACT
  const float* demod = nullptr;      // [B, N]
  const float* noise = nullptr;      // [nb, gh*gw]
  int64_t noise_bstride = 0;         // 0 when nb == 1
  const float* noise_w = nullptr;    // [1] device scalar
  const float* bias = nullptr;       // [N]
  // EPI_DGRAD
  const float* mod_out = nullptr;    // [B, N]  style of the layer whose input gradient this is
  const float* xsave = nullptr;      // [B or 1, gh, gw, N] forward input of that layer
  int64_t xsave_bstride = 0;
  float* partial = nullptr;          // [ceil(B*gh*gw / seglen), N]
  // EPI_DGRAD_ACT (fields demod / noise / noise_w / bias above then describe the layer that produced xsave)
  const float* drgb = nullptr;       // [B, 3, gh*gw] gradient of the skip image at this resolution, or null
  const float* s_rgb = nullptr;      // [B, N] ToRGB style
  const float* wrgb = nullptr;       // [3, N]
  float* partial_T = nullptr;        // [B * tiles, N]
  float* partial_R = nullptr;        // [B * tiles, N]
  // EPI_ACT on the tensor-core kernel, layers feeding a ToRGB (s_rgb / wrgb above): rgb_out[b,o,pix] =
  // sum_n act[b,pix,n]*s_rgb[b,n]*wrgb[o,n] + rgb_bias[o]; the FIR-upsampled skip is added by launch_skip_add
  float* rgb_out = nullptr;          // [B, 3, gh*gw] or null
  const float* rgb_bias = nullptr;   // [3]
};

// `out` may be null for EPI_DGRAD (gradient wrt the constant input is not needed)
int launch_conv_simt(const float* in, const float* mod, const float* wtab, float* out,
                     const ConvGeom& g, int epi, const ConvEpiArgs& e, cudaStream_t s);
int conv_dgrad_seglen(const ConvGeom& g);   // pixels per partial row (power of two <= 128)

// ---- tensor-core (tcgen05, tf32) version of the same gather convolution, csrc/conv_tc.cu ----
// Taps are organised in up to four groups; every group reads one "plane" of the input tensor
// [B, in_planes, in_h, in_w, K] (planes = the four sub-pixel phases of the stride-2 transposed
// conv's [2H+1, 2W+1] intermediate stored phase-major; 1 plane for ordinary activations).
// Tap t of group g contributes in[b, plane_g, gy + dy[t], gx + dx[t], :] * wtab[widx[t]], dy,dx in [-1,1].
struct TcTaps {
  int ngroups;
  int group_plane[4];
  int group_tap0[5];  // taps of group g are [group_tap0[g], group_tap0[g+1])
  signed char dy[9], dx[9], widx[9];
  // fused sub-pixel phases (EPI_STORE only): tap t accumulates into accumulator acc[t] in [0, nphase); accumulator p is
  // written to output plane p over the grid (gh - (p >> 1), gw - (p & 1)).  nphase = 1: ordinary convolution.
  signed char acc[9];
  int nphase;
};
struct TcConv {
  const float* in; int in_planes, in_h, in_w; bool in_bcast;
  const float* mod;   // [B, K] style (forward) or null
  const void* wmap;   // four CUtensorMaps (box rows 256/128/64/32) of the K-major weight table [ntaps * N, K] (tc_make_weight_maps)
  float* out; int out_planes, out_plane, out_h, out_w;   // out tensor [B, out_planes, out_h, out_w, N]
  int out_stride = 1, out_oy = 0, out_ox = 0;            // grid pixel (gy,gx) is written at (gy*stride+oy, gx*stride+ox)
  int batch, gh, gw, K, N;
  TcTaps taps;
  int epi; ConvEpiArgs e;   // EPI_DGRAD: partial rows are [b * tc_tiles_per_sample + tile]
};
int launch_conv_tc(const TcConv& c, cudaStream_t s);
bool tc_supported(int K, int N, int gh, int gw);
int tc_tiles_per_sample(int gh, int gw);
int tc_make_weight_maps(void* maps_out_4x128B, const float* table, int rows, int K, int N);

// 4x4 FIR on NHWC: out[b,oy,ox,c] = sum_{ty,tx} in[b, oy+ty-pad, ox+tx-pad, c] * coef[ty*4+tx]
// optional fused epilogue (same as EPI_ACT).  coef is a device pointer to 16 floats.
struct FirArgs {
  int batch, in_h, in_w, out_h, out_w, C, pad;
  const float* coef;
  // phase-major storage of the odd-sized side ([B, 4, (n+1)/2, (n+1)/2, C], plane = (y&1)*2 + (x&1))
  bool in_planar = false, out_planar = false;
  // separable form of the same taps, coef[ty*4+tx] = ky[ty]*kx[tx] (device pointers to 4 floats each); when set the
  // wide layers use the row/column two-pass kernel (half the FMAs, 1.75 loads per output instead of 2.5)
  const float* kx = nullptr; const float* ky = nullptr;
  bool act = false;
  const float* demod = nullptr; const float* noise = nullptr; int64_t noise_bstride = 0;
  const float* noise_w = nullptr; const float* bias = nullptr;
};
int launch_fir4x4_nhwc(const float* in, float* out, const FirArgs& a, cudaStream_t s);

// ToRGB: rgb[b,o,y,x] = sum_c act[b,y,x,c]*s[b,c]*wrgb[o,c] + bias[o] + up2(skip)[b,o,y,x]
// (src/model.py:379-388); skip [B,3,h/2,w/2] may be null; kup = 16 taps of the Upsample FIR.
int launch_torgb_fwd(const float* act, const float* s, const float* wrgb, const float* bias,
                     const float* skip, const float* kup, float* rgb, int batch, int h, int w,
                     int C, cudaStream_t s_);

// rgb[b,o,y,x] += up2(skip)[b,o,y,x]   (the Upsample branch of ToRGB, src/model.py:384-386)
int launch_skip_add(float* rgb, const float* skip, const float* kup, int batch, int h, int w, cudaStream_t st);

// Backward through noise/bias/lrelu (+ the ToRGB branch) of one StyledConv, in place on g:
//   gtot = g (or 0) + sum_o drgb[b,o,pix]*s_rgb[b,c]*wrgb[o,c]
//   gpre = gtot * (act>0 ? 1 : slope) * gain
//   g   <- gpre * demod[b,c]
//   pT[seg,c] = sum_pix gpre * (pre - noise_w*noise - bias[c])     pre = lrelu^-1(act)
//   pR[seg,c] = sum_pix act * sum_o drgb[b,o,pix]*wrgb[o,c]
struct ActBwdArgs {
  int batch, hw, C;
  const float* act; float* g; bool g_has_input;
  const float* demod; const float* noise; int64_t noise_bstride; const float* noise_w;
  const float* bias;
  const float* drgb = nullptr;   // [B,3,hw] or null
  const float* s_rgb = nullptr;  // [B,C]
  const float* wrgb = nullptr;   // [3,C]
  float* pT; float* pR;          // [B*hw/seglen, C]
};
int actbwd_seglen(int hw, int C);
int launch_act_bwd(const ActBwdArgs& a, cudaStream_t s);

// out[b, c] = sum_{q<Q} partial[b*Q+q, c]   (fixed order -> deterministic)
int launch_partial_reduce(const float* partial, float* out, int batch, int Q, int C, int64_t out_bstride,
                          cudaStream_t s);

// Batched forms (one launch per pass): items address the workspace `ws` by float offsets.
struct ReduceItem { int64_t src_off, dst_off; int Q, C; };                     // dst[b, c] = sum_q src[b*Q + q, c]
struct GradItem { int64_t r1_off, s_off, T_off, d_off, ds_off; const float* wsq; int cin, cout; };   // launch_style_grad
struct DemodItem { int64_t s_off, d_off; const float* wsq; int cin, cout; };   // launch_demod
// blocks[i] = (item, 32-channel block (reduce, style-grad) or 8-channel block (demod)); grid = (nblocks, batch)
int launch_batched_partial_reduce(const ReduceItem* items, const int2* blocks, int nblocks, float* ws, int batch, cudaStream_t s);
int launch_batched_style_grad(const GradItem* items, const int2* blocks, int nblocks, float* ws, int batch, cudaStream_t s);
int launch_batched_demod(const DemodItem* items, const int2* blocks, int nblocks, float* ws, int batch, cudaStream_t s);

// s[b, r] = sum_j latent[b, slot[r], j] * A[r, j] + bias[r]      (src/model.py:151-161, 258)
// s is stored per modulation unit as compact [B, cin] blocks: element (b, r) lives at
// batch*row_base[r] + b*row_cin[r] + (r - row_base[r]).
int launch_style_affine(const float* latent, const float* A, const float* bias, const int* row_slot,
                        const int* row_base, const int* row_cin, float* s, int batch, int rows,
                        int n_latent, int dim, cudaStream_t st);
// d_latent[b, slot, j] = sum_{r: slot[r]==slot} ds[b, r] * A[r, j]
int launch_style_affine_bwd(const float* ds, const float* A, const int* slot_row_begin,
                            const int* slot_row_end, const int* row_base, const int* row_cin,
                            float* d_latent, int batch, int rows, int n_latent, int dim,
                            cudaStream_t st);
// d[b, co] = rsqrt(sum_ci s[b,ci]^2 * wsq[co,ci] + 1e-8)       (src/model.py:261-263)
int launch_demod(const float* s, int64_t s_bstride, const float* wsq, float* d, int64_t d_bstride,
                 int batch, int cin, int cout, cudaStream_t st);
// ds[b,ci] = r1[b,ci] - s[b,ci] * sum_co T[b,co]*d[b,co]^2*wsq[co,ci]
int launch_style_grad(const float* r1, const float* s, int64_t s_bstride, const float* T,
                      const float* d, int64_t d_bstride, const float* wsq, float* ds, int batch,
                      int cin, int cout, cudaStream_t st);

// weight preparation (finalize): from W[Cout,Cin,3,3] (unscaled) build
//   wf[t][ci][co] = scale*W[co][ci][t], wg[t][co][ci] = scale*W[co][ci][t], wsq[co][ci] = sum_t (scale*W)^2
int launch_prep_conv3x3(const float* W, float scale, float* wf, float* wg, float* wsq, int cin,
                        int cout, cudaStream_t st);
int launch_round_tf32(const float* src, float* dst, int64_t n, cudaStream_t st);
int launch_scale_copy(const float* src, float* dst, float scale, int64_t n, cudaStream_t st);

// NHWC [B,H,W,C] <-> NCHW [B,C,H,W] (layer-level entry points / tests only)
int launch_nchw_to_nhwc(const float* in, float* out, int batch, int C, int hw, cudaStream_t st);
int launch_nhwc_to_nchw(const float* in, float* out, int batch, int C, int hw, cudaStream_t st);

int upfirdn2d_dispatch(const void* input, const void* kernel, void* out, int dtype, int64_t major,
                       int in_h, int in_w, int64_t minor, int kh, int kw, int up_x, int up_y,
                       int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                       cudaStream_t s, bool allow_tiled, bool allow_planes = true);
// allow_planes = false keeps the kernel choice independent of the plane count (the synthesis plan's skip-gradient FIRs: a
// trajectory's bits must not depend on what shares its batch)

}  // namespace lfp
