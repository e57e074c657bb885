// Perceptual loss of the attribution loop (additive C-ABI group 7 of include/lfp_sg2.h): LPIPS v0.1, VGG16 backbone +
// linear heads, forward AND backward to the estimated image, with the target's features computed once per image.
//
// Replaces `percept(target, est)` of src/utils.py:16,44-50 -> PNetLin.forward (src/custom_lpips/networks_basic.py:63-91):
//   scaling layer (:93-100) -> VGG16 slices relu1_2 .. relu5_3 (src/custom_lpips/pretrained_networks.py:97-135) for BOTH images
//   -> unit-normalise each tap over channels (custom_lpips/__init__.py:42-44) -> squared difference -> 1x1 "lin" conv ->
//   spatial mean -> sum over the five taps.
// The reference recomputes the constant target's 13 convolutions every step (networks_basic.py:66); here
// lfp_lpips_set_target runs them once and keeps the normalised tap features.
//
// Layout NHWC fp32; the 13 convolutions run on the same gather kernels as the generator (tcgen05 kind::tf32 for
// LFP_PREC_TF32 - the arithmetic the reference's cuDNN convs use by default - or CUDA-core fp32), with the bias + ReLU and the
// ReLU-mask of the data gradient in the conv epilogues (EPI_RELU / EPI_DGRAD_RELU).  The loss, its gradient through the
// normalisation, the un-pooling of the gradient that arrives from the next slice and the ReLU mask of a tap are one kernel
// per tap (tap_backward_kernel).  Every reduction has a fixed order: a trajectory's loss does not depend on its batch.
#include <math.h>
#include <string.h>
#include <string>
#include <vector>

#include "synth_kernels.cuh"

namespace lfp {

static const int kVggCin[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kVggCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const int kVggLevel[13] = {0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4};        // resolution = size >> level
static const int kVggFeat[13] = {0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28};   // torchvision features index
static const int kTapConv[5] = {1, 3, 6, 9, 12};                                 // relu1_2, relu2_2, relu3_3, relu4_3, relu5_3

// NCHW [B, 3, hw] image -> NHWC [B, hw, 32]: (x - shift) / scale in channels 0..2 (networks_basic.py:93-100), zeros above.
// One warp per 32 consecutive pixels (hw is a multiple of 256, so a group never straddles two samples): three coalesced loads,
// then eight 512-byte stores whose first-channel values come from the owning lane by shuffle.  The first version gave every
// 16-byte piece its own thread with the loads on every eighth lane: 1.18 ms for 2.7 GB written at 1024 px, B = 20.
__global__ void __launch_bounds__(256) lpips_prep_kernel(const float* __restrict__ img, float* __restrict__ out, int64_t hw, int64_t npix) {
  const int lane = threadIdx.x & 31;
  const int64_t p0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;   // over B * hw
  if (p0 >= npix) return;
  const int64_t b = p0 / hw, p = p0 - b * hw + lane;
  const float* s = img + b * 3 * hw + p;
  const float x = (s[0] - (-.030f)) / .458f;
  const float y = (s[hw] - (-.088f)) / .448f;
  const float z = (s[2 * hw] - (-.188f)) / .450f;
  float4* o = reinterpret_cast<float4*>(out) + p0 * 8 + lane;
  const int q = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int src = 4 * j + (lane >> 3);
    const float vx = __shfl_sync(0xffffffffu, x, src), vy = __shfl_sync(0xffffffffu, y, src), vz = __shfl_sync(0xffffffffu, z, src);
    o[j * 32] = q == 0 ? make_float4(vx, vy, vz, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// gradient back through the scaling layer: d_img[b, c, p] = D[b, p, c] / scale_c
__global__ void __launch_bounds__(256) lpips_unprep_kernel(const float* __restrict__ D, float* __restrict__ d_img, int64_t hw, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;   // over B * hw
  if (i >= n) return;
  const int64_t b = i / hw, p = i - b * hw;
  const float4 v = reinterpret_cast<const float4*>(D)[i * 8];
  float* o = d_img + b * 3 * hw + p;
  o[0] = v.x / .458f; o[hw] = v.y / .448f; o[2 * hw] = v.z / .450f;
}

// 2x2 max-pool, NHWC (torchvision features[4], [9], [16], [23])
__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int C4, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;   // over B * (H/2) * (W/2) * C4
  if (i >= n) return;
  const int c = (int)(i % C4);
  int64_t r = i / C4;
  const int ox = (int)(r % (W / 2)); r /= (W / 2);
  const int oy = (int)(r % (H / 2));
  const int64_t b = r / (H / 2);
  const float4* p = reinterpret_cast<const float4*>(in) + ((b * H + 2 * oy) * W + 2 * ox) * C4 + c;
  const float4 a = p[0], bb = p[C4], cc = p[(int64_t)W * C4], d = p[(int64_t)W * C4 + C4];
  float4 m;
  m.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)); m.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y));
  m.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)); m.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w));
  reinterpret_cast<float4*>(out)[i] = m;
}

// t = f / (sqrt(sum_c f^2) + 1e-10), one warp per pixel (custom_lpips/__init__.py:42-44): the cached target features
__global__ void __launch_bounds__(256) normalize_nhwc_kernel(const float* __restrict__ f, float* __restrict__ t, int C, int64_t npix) {
  const int lane = threadIdx.x & 31;
  const int64_t pix = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= npix) return;
  const float4* p = reinterpret_cast<const float4*>(f + pix * C);
  float ss = 0.f;
  for (int j = lane; j < C / 4; j += 32) { const float4 v = p[j]; ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss)))); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float inv = 1.f / (sqrtf(ss) + 1e-10f);
  float4* o = reinterpret_cast<float4*>(t + pix * C);
  for (int j = lane; j < C / 4; j += 32) { float4 v = p[j]; v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; o[j] = v; }
}

// One LPIPS tap, forward value + backward, for a 2x2 window of pixels per warp (the pool window of the slice above):
//   n = f / (|f| + eps) ; loss += sum_c lin_c (n_c - t_c)^2 / HW                    (networks_basic.py:69-76)
//   d n_c = 2 lin_c (n_c - t_c) / HW ; d f_c = d n_c / (|f| + eps) - f_c (sum_c' d n_c' f_c') / (|f| (|f| + eps)^2)
//   + the gradient arriving from the next slice through its 2x2 max-pool: up[b, y/2, x/2, c] goes to the FIRST maximal
//     element of the window in row-major order (what torch's max_pool2d backward does)
//   G = (d f + routed) * (f > 0): gradient w.r.t. the pre-ReLU output of the tap's convolution
// partial[(b * nwarps + window) ] = the window's loss contribution (summed in a fixed order afterwards).
// LPW lanes share a window (16 for C = 64: two windows per warp, 32 otherwise) and NJ = C / (4 LPW) float4 per lane and pixel.
// The window's features stay in registers across the three passes (norms; loss and sum_c g_c f_c; gradient), so f is read
// once; the first version re-read it per pass with half of the lanes idle at C = 64 and ran at 1.15 TB/s on the 1024 px tap
// (10.8 ms of the 89 ms LPIPS step at 1024 px, B = 20).  Lane li holds channels 4 (li + LPW jj) ..+3: the per-lane fmaf chains
// and the xor trees add the same numbers in the same order as before (the idle upper half-warp contributed exact zeros).
template <bool UP, int LPW, int NJ>
__global__ void __launch_bounds__(256) tap_backward_kernel(const float* __restrict__ f, const float* __restrict__ t, int64_t t_bstride,
                                                           const float* __restrict__ lin, const float* __restrict__ up,
                                                           float* __restrict__ G, float* __restrict__ partial, int H, int W, int C,
                                                           float inv_hw, int64_t nwin_per, int batch) {
  constexpr int WPW = 32 / LPW;   // windows per warp
  const int lane = threadIdx.x & 31, li = lane % LPW;
  // The sample is the fastest-varying part of the block index: the CTAs that work on the same windows of the batch's samples
  // run together, so a target shared by the batch (t_bstride = 0, the attribution loop) comes from HBM once and from L2 for the
  // other samples.  With the sample as grid.y every sample re-read the whole cached tap from HBM (268 MB per sample at the
  // 1024 px tap: 5.4 GB of the 17.5 GB that launch moved at B = 20, ncu launch list of profiles/r02_lpips_launches.md).
  const int b = (int)(blockIdx.x % (unsigned)batch);
  const int64_t win = ((int64_t)(blockIdx.x / (unsigned)batch) * 8 + (threadIdx.x >> 5)) * WPW + lane / LPW;
  const bool live = win < nwin_per;            // (a dead half-warp still takes part in the shuffles)
  const int W2 = W / 2;
  const int64_t wc = live ? win : 0;
  const int wy = (int)(wc / W2), wx = (int)(wc - (int64_t)wy * W2);
  const float4* fp[4];
  const float4* tp[4];
  float4* gp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t pix = (int64_t)(2 * wy + (k >> 1)) * W + 2 * wx + (k & 1);
    fp[k] = reinterpret_cast<const float4*>(f + ((int64_t)b * H * W + pix) * C) + li;
    tp[k] = reinterpret_cast<const float4*>(t + (int64_t)b * t_bstride + pix * C) + li;
    gp[k] = reinterpret_cast<float4*>(G + ((int64_t)b * H * W + pix) * C) + li;
  }
  // pass 1: the four pixels' squared norms
  float4 v[4][NJ];
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k][jj] = live ? fp[k][jj * LPW] : make_float4(0.f, 0.f, 0.f, 0.f);
  float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj)
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float4 q = v[k][jj]; ss[k] = fmaf(q.x, q.x, fmaf(q.y, q.y, fmaf(q.z, q.z, fmaf(q.w, q.w, ss[k])))); }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int off = LPW / 2; off > 0; off >>= 1) ss[k] += __shfl_xor_sync(0xffffffffu, ss[k], off);
  float nrm[4], inv[4], dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) { nrm[k] = sqrtf(ss[k]); inv[k] = 1.f / (nrm[k] + 1e-10f); }
  // pass 2: loss and sum_c lin_c (n_c - t_c) f_c
  const float4* lp = reinterpret_cast<const float4*>(lin) + li;
  float loss = 0.f;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const float4 l4 = __ldg(lp + jj * LPW);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 q = v[k][jj];
      const float4 tt = live ? tp[k][jj * LPW] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float dx = q.x * inv[k] - tt.x, dy = q.y * inv[k] - tt.y, dz = q.z * inv[k] - tt.z, dw = q.w * inv[k] - tt.w;
      loss = fmaf(l4.x * dx, dx, fmaf(l4.y * dy, dy, fmaf(l4.z * dz, dz, fmaf(l4.w * dw, dw, loss))));
      dot[k] = fmaf(l4.x * dx, q.x, fmaf(l4.y * dy, q.y, fmaf(l4.z * dz, q.z, fmaf(l4.w * dw, q.w, dot[k]))));
    }
  }
#pragma unroll
  for (int off = LPW / 2; off > 0; off >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, off);
#pragma unroll
    for (int k = 0; k < 4; ++k) dot[k] += __shfl_xor_sync(0xffffffffu, dot[k], off);
  }
  if (!live) return;
  // d f_c = 2/HW * ( lin_c (n_c - t_c) inv - f_c * dot * inv^2 / nrm )      (second term 0 where the pixel is all zero)
  float c2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) c2[k] = nrm[k] > 0.f ? dot[k] * inv[k] * inv[k] / nrm[k] : 0.f;
  const float s2 = 2.f * inv_hw;
  const float4* upp = UP ? reinterpret_cast<const float4*>(up + (((int64_t)b * (H / 2) + wy) * W2 + wx) * C) + li : nullptr;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const float4 l4 = __ldg(lp + jj * LPW);
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    int ax = 0, ay = 0, az = 0, aw = 0;   // index of the first maximum of the window, per channel
    if (UP) {
      u = upp[jj * LPW];
#define LFP_ARGMAX(comp, a)                                                                 \
  {                                                                                         \
    float m = v[0][jj].comp; a = 0;                                                         \
    if (v[1][jj].comp > m) { m = v[1][jj].comp; a = 1; }                                    \
    if (v[2][jj].comp > m) { m = v[2][jj].comp; a = 2; }                                    \
    if (v[3][jj].comp > m) { m = v[3][jj].comp; a = 3; }                                    \
  }
      LFP_ARGMAX(x, ax) LFP_ARGMAX(y, ay) LFP_ARGMAX(z, az) LFP_ARGMAX(w, aw)
#undef LFP_ARGMAX
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 q = v[k][jj];
      const float4 tt = tp[k][jj * LPW];
      float4 g;
      g.x = s2 * (l4.x * (q.x * inv[k] - tt.x) * inv[k] - q.x * c2[k]);
      g.y = s2 * (l4.y * (q.y * inv[k] - tt.y) * inv[k] - q.y * c2[k]);
      g.z = s2 * (l4.z * (q.z * inv[k] - tt.z) * inv[k] - q.z * c2[k]);
      g.w = s2 * (l4.w * (q.w * inv[k] - tt.w) * inv[k] - q.w * c2[k]);
      if (UP) { if (ax == k) g.x += u.x; if (ay == k) g.y += u.y; if (az == k) g.z += u.z; if (aw == k) g.w += u.w; }
      g.x = q.x > 0.f ? g.x : 0.f; g.y = q.y > 0.f ? g.y : 0.f; g.z = q.z > 0.f ? g.z : 0.f; g.w = q.w > 0.f ? g.w : 0.f;
      gp[k][jj * LPW] = g;
    }
  }
  if (li == 0) partial[(int64_t)b * nwin_per + win] = loss * inv_hw;
}

template <bool UP>
static int launch_tap_backward(const float* f, const float* t, int64_t tstride, const float* lin, const float* up, float* G, float* partial,
                               int H, int W, int C, int batch, int64_t nwin, cudaStream_t s) {
  const float inv_hw = 1.f / (float)((int64_t)H * W);
#define LFP_TAP(LPW, NJ)                                                                                          \
  {                                                                                                               \
    const int64_t nblk = ceil_div(nwin, (int64_t)8 * (32 / LPW)) * batch;                                          \
    LFP_CHECK_ARG(nblk < (1ll << 31), "lpips: too many windows");                                                 \
    tap_backward_kernel<UP, LPW, NJ><<<(unsigned)nblk, 256, 0, s>>>(f, t, tstride, lin, up, G, partial, H, W, C, inv_hw, nwin, batch); \
  }
  switch (C) {
    case 64: LFP_TAP(16, 1) break;
    case 128: LFP_TAP(32, 1) break;
    case 256: LFP_TAP(32, 2) break;
    case 512: LFP_TAP(32, 4) break;
    default: set_error("lpips: unsupported tap width %d", C); return LFP_EUNSUPPORTED;
  }
#undef LFP_TAP
  LFP_LAUNCH_CHECK();
  return 0;
}

// loss[b] (+)= sum over windows of partial[b, :] in a fixed order: 256 interleaved accumulators, then a tree
__global__ void __launch_bounds__(256) lpips_loss_reduce_kernel(const float* __restrict__ partial, int64_t n, float* __restrict__ loss, int accumulate) {
  __shared__ float sm[256];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) acc += partial[(int64_t)b * n + i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) loss[b] = accumulate ? loss[b] + sm[0] : sm[0];
}

// W [cout, cin, 3, 3] (torch layout, unscaled) -> wf[t][cin_p][cout], wg[t][cout][cin_p], cross-correlation taps as stored
__global__ void vgg_prep_kernel(const float* __restrict__ W, float* __restrict__ wf, float* __restrict__ wg, int cin, int cin_p, int cout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cin_p * cout) return;
  const int co = (int)(i / cin_p), ci = (int)(i - (int64_t)co * cin_p);
  for (int t = 0; t < 9; ++t) {
    const float v = ci < cin ? W[((int64_t)co * cin + ci) * 9 + t] : 0.f;
    wf[((int64_t)t * cin_p + ci) * cout + co] = v;
    wg[((int64_t)t * cout + co) * cin_p + ci] = v;
  }
}

}  // namespace lfp

using namespace lfp;

struct lfp_lpips {
  int H = 0, W = 0;
  struct Conv { int cin, cin_p, cout, level; float *W, *bias, *wf, *wg, *wf_t, *wg_t; alignas(64) unsigned char map_fwd[512]; alignas(64) unsigned char map_bwd[512]; };
  Conv conv[13];
  float* lin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* tfeat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // cached normalised target features [Bt, h, w, C]
  int target_batch = 0, target_prec = -1;
  bool finalized = false;
  std::vector<void*> owned;
  ~lfp_lpips() { for (void* p : owned) cudaFree(p); for (float* p : tfeat) if (p) cudaFree(p); }
  int alloc(float** p, size_t n) { LFP_CUDA(cudaMalloc((void**)p, (n ? n : 1) * sizeof(float))); owned.push_back(*p); return 0; }
  int res_h(int level) const { return H >> level; }
  int res_w(int level) const { return W >> level; }
};

namespace {
struct LpLayout { size_t x0, act[13], pool, gA, gB, partial, total; };
LpLayout lp_layout(const lfp_lpips* h, int B) {
  LpLayout L{};
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (off + n + 63) / 64 * 64; return o; };
  const size_t hw = (size_t)h->H * h->W;
  L.x0 = take((size_t)B * hw * 32);
  size_t maxg = (size_t)B * hw * 64, maxpool = 0;
  for (int c = 0; c < 13; ++c) {
    const size_t n = (size_t)B * h->res_h(h->conv[c].level) * h->res_w(h->conv[c].level) * h->conv[c].cout;
    L.act[c] = take(n);
    if (c + 1 < 13 && h->conv[c + 1].level != h->conv[c].level && n / 4 > maxpool) maxpool = n / 4;
  }
  L.pool = take(maxpool);
  L.gA = take(maxg);
  L.gB = take(maxg);
  L.partial = take((size_t)B * (hw / 4));
  L.total = off;
  return L;
}

int vgg_conv(const lfp_lpips* h, int c, const float* in, float* out, int B, int epi, const float* xsave, int precision, bool fwd, cudaStream_t s) {
  const lfp_lpips::Conv& L = h->conv[c];
  const int R = h->res_h(L.level), Wd = h->res_w(L.level);
  const int K = fwd ? L.cin_p : L.cout, N = fwd ? L.cout : L.cin_p;
  ConvGeom g{};
  g.batch = B; g.gh = R; g.gw = Wd; g.in_h = R; g.in_w = Wd; g.in_bstride = (int64_t)R * Wd * K; g.in_stride = 1;
  g.out_h = R; g.out_w = Wd; g.out_stride = 1; g.K = K; g.N = N; g.ntaps = 9;
  for (int t = 0; t < 9; ++t) {
    // forward: cross-correlation, out[y, x] += in[y + ky - 1, x + kx - 1] W[ky, kx]; data gradient: the adjoint, flipped offsets
    g.dy[t] = (signed char)(fwd ? t / 3 - 1 : 1 - t / 3); g.dx[t] = (signed char)(fwd ? t % 3 - 1 : 1 - t % 3); g.widx[t] = (signed char)t;
  }
  ConvEpiArgs e;
  e.bias = L.bias; e.xsave = xsave; e.xsave_bstride = (int64_t)R * Wd * N;
  const bool use_tc = precision == LFP_PREC_TF32 && tc_supported(K, N, R, Wd);
  if (use_tc) {
    TcConv t{};
    t.in = in; t.in_planes = 1; t.in_h = R; t.in_w = Wd; t.in_bcast = false; t.mod = nullptr; t.wmap = fwd ? L.map_fwd : L.map_bwd;
    t.out = out; t.out_planes = 1; t.out_plane = 0; t.out_h = R; t.out_w = Wd; t.batch = B; t.gh = R; t.gw = Wd; t.K = K; t.N = N;
    t.taps.ngroups = 1; t.taps.group_plane[0] = 0; t.taps.group_tap0[0] = 0; t.taps.group_tap0[1] = 9;
    for (int i = 0; i < 9; ++i) { t.taps.dy[i] = g.dy[i]; t.taps.dx[i] = g.dx[i]; t.taps.widx[i] = g.widx[i]; }
    t.epi = epi; t.e = e;
    return launch_conv_tc(t, s);
  }
  return launch_conv_simt(in, nullptr, fwd ? L.wf : L.wg, out, g, epi, e, s);
}

// VGG16 forward: fills L.act[0..12] from the NCHW image
int vgg_forward(const lfp_lpips* h, int B, const float* img, float* ws, const LpLayout& L, int precision, cudaStream_t s) {
  const int64_t hw = (int64_t)h->H * h->W;
  const int64_t npix = (int64_t)B * hw;   // a CTA converts 256 pixels
  lpips_prep_kernel<<<(unsigned)ceil_div(npix, 256), 256, 0, s>>>(img, ws + L.x0, hw, npix);
  LFP_LAUNCH_CHECK();
  const float* x = ws + L.x0;
  for (int c = 0; c < 13; ++c) {
    const lfp_lpips::Conv& cv = h->conv[c];
    if (c > 0 && cv.level != h->conv[c - 1].level) {
      const int Hp = h->res_h(h->conv[c - 1].level), Wp = h->res_w(h->conv[c - 1].level), C4 = h->conv[c - 1].cout / 4;
      const int64_t n = (int64_t)B * (Hp / 2) * (Wp / 2) * C4;
      maxpool_nhwc_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(x, ws + L.pool, Hp, Wp, C4, n);
      LFP_LAUNCH_CHECK();
      x = ws + L.pool;
    }
    LFP_TRY(vgg_conv(h, c, x, ws + L.act[c], B, EPI_RELU, nullptr, precision, true, s));
    x = ws + L.act[c];
  }
  return 0;
}
}  // namespace

extern "C" int lfp_lpips_create(lfp_lpips** out, int height, int width) {
  LFP_CHECK_ARG(out != nullptr, "lpips_create: null out");
  LFP_CHECK_ARG(height >= 32 && width >= 32 && height % 16 == 0 && width % 16 == 0 && height <= 4096 && width <= 4096,
                "lpips_create: image %dx%d must be a multiple of 16 in [32, 4096] (four 2x2 max-pools)", height, width);
  lfp_lpips* h = new lfp_lpips();
  h->H = height; h->W = width;
  int rc = 0;
  for (int c = 0; c < 13; ++c) {
    lfp_lpips::Conv& L = h->conv[c];
    L.cin = kVggCin[c]; L.cin_p = c == 0 ? 32 : kVggCin[c]; L.cout = kVggCout[c]; L.level = kVggLevel[c];
    const size_t wn = (size_t)9 * L.cin_p * L.cout;
    rc |= h->alloc(&L.W, (size_t)L.cout * L.cin * 9); rc |= h->alloc(&L.bias, L.cout);
    rc |= h->alloc(&L.wf, wn); rc |= h->alloc(&L.wg, wn); rc |= h->alloc(&L.wf_t, wn); rc |= h->alloc(&L.wg_t, wn);
  }
  for (int k = 0; k < 5; ++k) rc |= h->alloc(&h->lin[k], kVggCout[kTapConv[k]]);
  if (rc != 0) { set_error("lpips_create: device allocation failed"); delete h; return LFP_ENOMEM; }
  *out = h;
  return 0;
}

extern "C" void lfp_lpips_destroy(lfp_lpips* h) { delete h; }

extern "C" int lfp_lpips_set_param(lfp_lpips* h, const char* name, const float* data, int64_t numel, void* stream) {
  LFP_CHECK_ARG(h && name && data, "lpips_set_param: null argument");
  const std::string n(name);
  float* dst = nullptr; int64_t want = -1;
  for (int c = 0; c < 13; ++c) {
    const std::string base = "net.slice" + std::to_string(kVggLevel[c] + 1) + "." + std::to_string(kVggFeat[c]);
    if (n == base + ".weight") { dst = h->conv[c].W; want = (int64_t)h->conv[c].cout * h->conv[c].cin * 9; }
    if (n == base + ".bias") { dst = h->conv[c].bias; want = h->conv[c].cout; }
  }
  for (int k = 0; k < 5; ++k)
    if (n == "lin" + std::to_string(k) + ".model.1.weight") { dst = h->lin[k]; want = kVggCout[kTapConv[k]]; }
  LFP_CHECK_ARG(dst != nullptr, "lpips_set_param: unknown parameter '%s' (net.slice<S>.<I>.weight|bias, lin<K>.model.1.weight)", name);
  LFP_CHECK_ARG(want == numel, "lpips_set_param: '%s' expects %lld elements, got %lld", name, (long long)want, (long long)numel);
  LFP_CUDA(cudaMemcpyAsync(dst, data, numel * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  h->finalized = false; h->target_batch = 0;
  return 0;
}

extern "C" int lfp_lpips_finalize(lfp_lpips* h, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "lpips_finalize: null handle");
  cudaStream_t s = (cudaStream_t)stream;
  for (int c = 0; c < 13; ++c) {
    lfp_lpips::Conv& L = h->conv[c];
    vgg_prep_kernel<<<(unsigned)ceil_div((int64_t)L.cin_p * L.cout, 256), 256, 0, s>>>(L.W, L.wf, L.wg, L.cin, L.cin_p, L.cout);
    LFP_LAUNCH_CHECK();
    const int64_t wn = (int64_t)9 * L.cin_p * L.cout;
    LFP_TRY(launch_round_tf32(L.wf, L.wf_t, wn, s));
    LFP_TRY(launch_round_tf32(L.wg, L.wg_t, wn, s));
    if (tc_supported(L.cin_p, L.cout, 4, 4)) LFP_TRY(tc_make_weight_maps(L.map_fwd, L.wg_t, 9 * L.cout, L.cin_p, L.cout));
    if (tc_supported(L.cout, L.cin_p, 4, 4)) LFP_TRY(tc_make_weight_maps(L.map_bwd, L.wf_t, 9 * L.cin_p, L.cout, L.cin_p));
  }
  h->finalized = true;
  return 0;
}

extern "C" size_t lfp_lpips_workspace_bytes(const lfp_lpips* h, int batch) {
  if (!h || batch <= 0) return 0;
  return lp_layout(h, batch).total * sizeof(float);
}

static int lp_check(const lfp_lpips* h, int batch, const void* ws, size_t ws_bytes, int precision, const LpLayout& L) {
  LFP_CHECK_ARG(h != nullptr && ws != nullptr, "lpips: null handle or workspace");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535, "lpips: batch %d out of range", batch);
  if (!h->finalized) { set_error("lpips: lfp_lpips_finalize has not been called since the last set_param"); return LFP_ESTATE; }
  if (ws_bytes < L.total * sizeof(float)) { set_error("lpips: workspace too small (%zu < %zu bytes)", ws_bytes, L.total * sizeof(float)); return LFP_ENOMEM; }
  LFP_CHECK_ARG(((uintptr_t)ws & 255) == 0, "lpips: workspace must be 256-byte aligned");
  LFP_CHECK_ARG(precision == LFP_PREC_FP32 || precision == LFP_PREC_TF32, "lpips: unknown precision mode %d", precision);
  return 0;
}

extern "C" int lfp_lpips_set_target(lfp_lpips* h, int target_batch, const float* target, void* workspace, size_t workspace_bytes,
                                    int precision, void* stream) {
  LFP_CHECK_ARG(h != nullptr && target_batch >= 1, "lpips_set_target: bad argument");
  const LpLayout L = lp_layout(h, target_batch);
  LFP_TRY(lp_check(h, target_batch, workspace, workspace_bytes, precision, L));
  LFP_CHECK_ARG(target != nullptr, "lpips_set_target: null target");
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  LFP_TRY(vgg_forward(h, target_batch, target, ws, L, precision, s));
  for (int k = 0; k < 5; ++k) {
    const lfp_lpips::Conv& cv = h->conv[kTapConv[k]];
    const int64_t npix = (int64_t)target_batch * h->res_h(cv.level) * h->res_w(cv.level);
    if (h->target_batch != target_batch) {
      if (h->tfeat[k]) { LFP_CUDA(cudaStreamSynchronize(s)); cudaFree(h->tfeat[k]); h->tfeat[k] = nullptr; }
      LFP_CUDA(cudaMalloc((void**)&h->tfeat[k], (size_t)npix * cv.cout * sizeof(float)));
    }
    normalize_nhwc_kernel<<<(unsigned)ceil_div(npix, 8), 256, 0, s>>>(ws + L.act[kTapConv[k]], h->tfeat[k], cv.cout, npix);
    LFP_LAUNCH_CHECK();
  }
  h->target_batch = target_batch; h->target_prec = precision;
  return 0;
}

extern "C" int lfp_lpips_loss_grad(lfp_lpips* h, int batch, const float* est, float* loss, float* d_est, void* workspace,
                                   size_t workspace_bytes, int precision, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "lpips_loss_grad: null handle");
  const LpLayout L = lp_layout(h, batch);
  LFP_TRY(lp_check(h, batch, workspace, workspace_bytes, precision, L));
  LFP_CHECK_ARG(est && loss, "lpips_loss_grad: null argument");
  if (h->target_batch != 1 && h->target_batch != batch) { set_error("lpips_loss_grad: target batch %d (lfp_lpips_set_target) must be 1 or %d", h->target_batch, batch); return LFP_ESTATE; }
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  LFP_TRY(vgg_forward(h, batch, est, ws, L, precision, s));
  float* gbuf[2] = {ws + L.gA, ws + L.gB};
  int cur = 0;
  const float* up = nullptr;   // gradient w.r.t. the pooled input of the slice above (null at the top tap)
  for (int k = 4; k >= 0; --k) {
    const int c_tap = kTapConv[k];
    const lfp_lpips::Conv& cv = h->conv[c_tap];
    const int R = h->res_h(cv.level), Wd = h->res_w(cv.level), C = cv.cout;
    const int64_t nwin = (int64_t)(R / 2) * (Wd / 2);
    const int64_t tstride = h->target_batch == 1 ? 0 : (int64_t)R * Wd * C;
    float* G = gbuf[cur];
    if (up) LFP_TRY(launch_tap_backward<true>(ws + L.act[c_tap], h->tfeat[k], tstride, h->lin[k], up, G, ws + L.partial, R, Wd, C, batch, nwin, s));
    else LFP_TRY(launch_tap_backward<false>(ws + L.act[c_tap], h->tfeat[k], tstride, h->lin[k], nullptr, G, ws + L.partial, R, Wd, C, batch, nwin, s));
    lpips_loss_reduce_kernel<<<(unsigned)batch, 256, 0, s>>>(ws + L.partial, nwin, loss, k == 4 ? 0 : 1);
    LFP_LAUNCH_CHECK();
    if (d_est == nullptr) { up = nullptr; continue; }   // value only: no gradient chain (the taps' G buffers are scratch)
    // data gradients down to the first convolution of this slice
    const int c_first = k == 0 ? 0 : kTapConv[k - 1] + 1;
    const float* g = G;
    for (int c = c_tap; c >= c_first; --c) {
      float* o = gbuf[cur ^ 1];
      if (c > c_first) {
        LFP_TRY(vgg_conv(h, c, g, o, batch, EPI_DGRAD_RELU, ws + L.act[c - 1], precision, false, s));   // through the ReLU of conv c - 1
      } else {
        // first conv of a slice: its input is the pooled tap below (masked and un-pooled by that tap's kernel) or the image
        LFP_TRY(vgg_conv(h, c, g, o, batch, EPI_STORE, nullptr, precision, false, s));
      }
      g = o; cur ^= 1;
    }
    up = g;
    cur ^= 1;   // keep `up` intact while the next tap writes its G
  }
  if (d_est != nullptr) {
    const int64_t hw = (int64_t)h->H * h->W, n = (int64_t)batch * hw;
    lpips_unprep_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(up, d_est, hw, n);
    LFP_LAUNCH_CHECK();
  }
  return 0;
}
