// Stand-alone modulated convolution (additive C-ABI group 5 of include/lfp_sg2.h): the reference's
// ModulatedConv2d.forward(input, style) (src/model.py:169-302) for one layer - plain k x k (k = 1 or 3), or 3x3
// stride-2 transposed + Blur - and its backward to the input and the style.  Same algebra and the same kernels as the
// whole-synthesis plan (activation-modulated form, src/model.py:229-256), without the
// noise / bias / leaky-ReLU epilogue: this is what model.ModulatedConv2d / StyledConv / ToRGB call when used on their own,
// and what the per-layer parity tests and the op microbenchmark (BASELINE.json configs[4]) exercise.
//
//   s = A style + b ; demod = rsqrt(sum_ci s^2 Wsq + 1e-8) (or 1) ; raw = conv(x s, scale W) ; out = raw demod
//   backward: dRaw = dOut demod ; dXm = dgrad(dRaw) ; dx = dXm s ; ds = sum_pix x dXm - s ((sum_pix dOut out) demod^2) Wsq
//             d_style = ds A
// Public tensors are NCHW / [B, style_dim] like the module's; inside, activations are NHWC with the channel counts padded
// (zeros) to what the gather kernels need (K multiple of 16, N multiple of 4), so any channel count works.  Layers whose
// shape the tcgen05 kernel supports run on it when precision = LFP_PREC_TF32.  Parameters are frozen constants (as in the
// plan): no weight gradients.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "synth_kernels.cuh"

namespace lfp {

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// W [cout, cin, k, k] (unscaled) -> wf[t][cin_p][cout_p], wg[t][cout_p][cin_p], wsq[cout_p][cin_p]; padding = 0
__global__ void prep_conv_generic_kernel(const float* __restrict__ W, float scale, float* __restrict__ wf, float* __restrict__ wg,
                                         float* __restrict__ wsq, int cin, int cout, int cin_p, int cout_p, int taps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cin_p * cout_p) return;
  const int co = (int)(i / cin_p), ci = (int)(i - (int64_t)co * cin_p);
  const bool real = co < cout && ci < cin;
  float sq = 0.f;
  for (int t = 0; t < taps; ++t) {
    const float v = real ? W[((int64_t)co * cin + ci) * taps + t] * scale : 0.f;
    wf[((int64_t)t * cin_p + ci) * cout_p + co] = v;
    wg[((int64_t)t * cout_p + co) * cin_p + ci] = v;
    sq = fmaf(v, v, sq);
  }
  wsq[(int64_t)co * cin_p + ci] = sq;
}

// NCHW [B, C, hw] -> NHWC [B, hw, Cp] with out = in * scale[b, c] (scale may be null); channels >= C are written as 0
__global__ void __launch_bounds__(256) nchw_to_nhwc_pad_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               const float* __restrict__ scale, int C, int Cp, int64_t hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t p = p0 + tx;
    float v = 0.f;
    if (c < C && p < hw) {
      v = in[((int64_t)b * C + c) * hw + p];
      if (scale) v *= scale[(int64_t)b * Cp + c];
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j;
    const int c = c0 + tx;
    if (p < hw && c < Cp) out[((int64_t)b * hw + p) * Cp + c] = tile[tx][j];
  }
}

// NHWC [B, hw, Cp] -> NCHW [B, C, hw] with out = in * scale[b, c] (scale may be null); padded channels are dropped
__global__ void __launch_bounds__(256) nhwc_to_nchw_scale_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                 const float* __restrict__ scale, int C, int Cp, int64_t hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (p < hw && c < Cp) ? in[((int64_t)b * hw + p) * Cp + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t p = p0 + tx;
    if (c < C && p < hw) {
      float v = tile[tx][j];
      if (scale) v *= scale[(int64_t)b * Cp + c];
      out[((int64_t)b * C + c) * hw + p] = v;
    }
  }
}

// Vector forms of the two layout kernels for the common case (hw and C multiples of 4, 16-byte aligned bases): a thread
// moves a 4 channel x 4 pixel block with four 128-bit loads, transposes it in registers and stores four 128-bit words; a warp
// covers CQ channel quads x 32 / CQ pixel quads, so the NHWC side is written / read as whole 16 CQ-byte runs per pixel (the
// full 128-byte line for 32 channels) and the NCHW side as 64-byte (CQ = 8) or 128-byte (CQ = 4) runs per channel row.  No
// shared memory, no barrier, 64 bytes in flight per thread.  Same arithmetic as the scalar kernels (one multiply), so the
// results are bit-identical; the scalar kernels remain for odd extents and channel counts.
template <int CQ>
__global__ void __launch_bounds__(256) nchw_to_nhwc_vec_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               const float* __restrict__ scale, int C, int Cp, int64_t hw) {
  constexpr int PG = 32 / CQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cq = lane % CQ, pg = lane / CQ;
  const int b = blockIdx.y;
  const int64_t pix = ((int64_t)blockIdx.x * 8 + warp) * (4 * PG) + 4 * pg;
  if (pix >= hw) return;
  for (int c4 = 4 * cq; c4 < Cp; c4 += 4 * CQ) {
    float4 v[4];
    if (c4 < C) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ldg_stream(reinterpret_cast<const float4*>(in + ((int64_t)b * C + c4 + j) * hw + pix));
      if (scale) {
        const float4 sc = *reinterpret_cast<const float4*>(scale + (int64_t)b * Cp + c4);
        v[0].x *= sc.x; v[0].y *= sc.x; v[0].z *= sc.x; v[0].w *= sc.x;
        v[1].x *= sc.y; v[1].y *= sc.y; v[1].z *= sc.y; v[1].w *= sc.y;
        v[2].x *= sc.z; v[2].y *= sc.z; v[2].z *= sc.z; v[2].w *= sc.z;
        v[3].x *= sc.w; v[3].y *= sc.w; v[3].z *= sc.w; v[3].w *= sc.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* o = out + ((int64_t)b * hw + pix) * Cp + c4;
    stg_stream(reinterpret_cast<float4*>(o), make_float4(v[0].x, v[1].x, v[2].x, v[3].x));
    stg_stream(reinterpret_cast<float4*>(o + Cp), make_float4(v[0].y, v[1].y, v[2].y, v[3].y));
    stg_stream(reinterpret_cast<float4*>(o + 2 * (int64_t)Cp), make_float4(v[0].z, v[1].z, v[2].z, v[3].z));
    stg_stream(reinterpret_cast<float4*>(o + 3 * (int64_t)Cp), make_float4(v[0].w, v[1].w, v[2].w, v[3].w));
  }
}

template <int CQ>
__global__ void __launch_bounds__(256) nhwc_to_nchw_vec_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               const float* __restrict__ scale, int C, int Cp, int64_t hw) {
  constexpr int PG = 32 / CQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cq = lane % CQ, pg = lane / CQ;
  const int b = blockIdx.y;
  const int64_t pix = ((int64_t)blockIdx.x * 8 + warp) * (4 * PG) + 4 * pg;
  if (pix >= hw) return;
  for (int c4 = 4 * cq; c4 < C; c4 += 4 * CQ) {   // padded channels are dropped (C is a multiple of 4 here)
    const float* i0 = in + ((int64_t)b * hw + pix) * Cp + c4;
    const float4 p0 = ldg_stream(reinterpret_cast<const float4*>(i0));
    const float4 p1 = ldg_stream(reinterpret_cast<const float4*>(i0 + Cp));
    const float4 p2 = ldg_stream(reinterpret_cast<const float4*>(i0 + 2 * (int64_t)Cp));
    const float4 p3 = ldg_stream(reinterpret_cast<const float4*>(i0 + 3 * (int64_t)Cp));
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (scale) sc = *reinterpret_cast<const float4*>(scale + (int64_t)b * Cp + c4);
    float4 r0 = make_float4(p0.x, p1.x, p2.x, p3.x), r1 = make_float4(p0.y, p1.y, p2.y, p3.y);
    float4 r2 = make_float4(p0.z, p1.z, p2.z, p3.z), r3 = make_float4(p0.w, p1.w, p2.w, p3.w);
    if (scale) {
      r0.x *= sc.x; r0.y *= sc.x; r0.z *= sc.x; r0.w *= sc.x;
      r1.x *= sc.y; r1.y *= sc.y; r1.z *= sc.y; r1.w *= sc.y;
      r2.x *= sc.z; r2.y *= sc.z; r2.z *= sc.z; r2.w *= sc.z;
      r3.x *= sc.w; r3.y *= sc.w; r3.z *= sc.w; r3.w *= sc.w;
    }
    float* o = out + ((int64_t)b * C + c4) * hw + pix;
    stg_stream(reinterpret_cast<float4*>(o), r0);
    stg_stream(reinterpret_cast<float4*>(o + hw), r1);
    stg_stream(reinterpret_cast<float4*>(o + 2 * hw), r2);
    stg_stream(reinterpret_cast<float4*>(o + 3 * hw), r3);
  }
}

// partial[(b * Q + q) * C + c] = sum over the pixels of segment q of a[b, pix, c] * bb[b, pix, c]   (NHWC, fixed order)
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ bb,
                                                          float* __restrict__ partial, int C, int64_t hw, int seglen, int Q) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int q = blockIdx.y, b = blockIdx.z;
  const int sub = threadIdx.x >> 5;   // eight pixel lanes per segment
  __shared__ float sm[8][32];
  float acc = 0.f;
  if (c < C) {
    const int64_t p0 = (int64_t)q * seglen, p1 = p0 + seglen < hw ? p0 + seglen : hw;
    for (int64_t p = p0 + sub; p < p1; p += 8) {
      const int64_t i = ((int64_t)b * hw + p) * C + c;
      acc = fmaf(a[i], bb[i], acc);
    }
  }
  sm[sub][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sub == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x & 31];
    partial[((int64_t)b * Q + q) * C + c] = t;
  }
}

// 128-bit form of dot_partial_kernel: 256 threads = 4 segments x 8 pixel lanes x 8 channel quads.  Every channel keeps the
// scalar kernel's summation order (pixel lane `sub` runs p0 + sub, p0 + sub + 8, ... with one fmaf chain; the eight lanes are
// added in index order), so the partial sums are bit-identical.
__global__ void __launch_bounds__(256) dot_partial_vec_kernel(const float* __restrict__ a, const float* __restrict__ bb,
                                                              float* __restrict__ partial, int C, int64_t hw, int seglen, int Q) {
  const int cq = threadIdx.x & 7, sub = (threadIdx.x >> 3) & 7, qs = threadIdx.x >> 6;
  const int c4 = (blockIdx.x * 8 + cq) * 4;
  const int q = blockIdx.y * 4 + qs, b = blockIdx.z;
  __shared__ float4 sm[4][8][8];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool live = c4 < C && q < Q;
  if (live) {
    const int64_t p0 = (int64_t)q * seglen, p1 = p0 + seglen < hw ? p0 + seglen : hw;
#pragma unroll 4
    for (int64_t p = p0 + sub; p < p1; p += 8) {
      const int64_t i = ((int64_t)b * hw + p) * C + c4;
      const float4 va = ldg_stream(reinterpret_cast<const float4*>(a + i));
      const float4 vb = ldg_stream(reinterpret_cast<const float4*>(bb + i));
      acc.x = fmaf(va.x, vb.x, acc.x); acc.y = fmaf(va.y, vb.y, acc.y);
      acc.z = fmaf(va.z, vb.z, acc.z); acc.w = fmaf(va.w, vb.w, acc.w);
    }
  }
  sm[qs][sub][cq] = acc;
  __syncthreads();
  if (sub == 0 && live) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float4 v = sm[qs][k][cq]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    *reinterpret_cast<float4*>(partial + ((int64_t)b * Q + q) * C + c4) = t;
  }
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// d_style[b, j] = sum_ci ds[b, ci] * A[ci, j]   (A already carries the EqualLinear scale)
__global__ void __launch_bounds__(128) dstyle_kernel(const float* __restrict__ ds, const float* __restrict__ A, float* __restrict__ out,
                                                     int cin_p, int dim) {
  const int j = blockIdx.x * 128 + threadIdx.x, b = blockIdx.y;
  if (j >= dim) return;
  float acc = 0.f;
  for (int ci = 0; ci < cin_p; ++ci) acc = fmaf(ds[(int64_t)b * cin_p + ci], __ldg(A + (int64_t)ci * dim + j), acc);
  out[(int64_t)b * dim + j] = acc;
}

}  // namespace lfp

using namespace lfp;

struct lfp_modconv {
  int cin = 0, cout = 0, k = 3, style_dim = 0, cin_p = 0, cout_p = 0, taps = 9;
  bool demodulate = true, upsample = false, finalized = false, tc_fwd = false, tc_bwd = false;
  float blur1d[4] = {1, 3, 3, 1};
  float *W = nullptr, *modw = nullptr, *modb = nullptr;                 // raw parameters
  float *wf = nullptr, *wg = nullptr, *wsq = nullptr, *wf_t = nullptr, *wg_t = nullptr, *A = nullptr, *bvec = nullptr, *fir = nullptr;
  int *row_slot = nullptr, *row_base = nullptr, *row_cin = nullptr;
  alignas(64) unsigned char map_fwd[4 * 128];
  alignas(64) unsigned char map_bwd[4 * 128];
  std::vector<void*> owned;
  // last forward per plan (one workspace in flight per handle, like the module it backs)
  const void* fwd_ws = nullptr; int fwd_batch = -1, fwd_h = 0, fwd_w = 0, fwd_prec = -1;
  ~lfp_modconv() { for (void* p : owned) cudaFree(p); }
  int alloc(float** p, size_t n) { LFP_CUDA(cudaMalloc((void**)p, (n ? n : 1) * sizeof(float))); owned.push_back(*p); return 0; }
};

namespace {
struct McLayout { size_t x, s, d, raw, T, g, dxm, part, R1, Tv, ds, total; int oh, ow, seglen_o, Qo, seglen_i, Qi; };

McLayout mc_layout(const lfp_modconv* h, int B, int H, int W) {
  McLayout L{};
  L.oh = h->upsample ? 2 * H : H; L.ow = h->upsample ? 2 * W : W;
  const size_t hw = (size_t)H * W, ohw = (size_t)L.oh * L.ow;
  L.seglen_o = 1024; L.Qo = (int)((ohw + 1023) / 1024);
  L.seglen_i = 1024; L.Qi = (int)((hw + 1023) / 1024);
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (off + n + 63) / 64 * 64; return o; };
  L.x = take((size_t)B * hw * h->cin_p);
  L.s = take((size_t)B * h->cin_p);
  L.d = take((size_t)B * h->cout_p);
  L.raw = take((size_t)B * ohw * h->cout_p);
  L.T = h->upsample ? take((size_t)B * (L.oh + 2) * (L.ow + 2) * h->cout_p) : 0;
  L.g = take((size_t)B * ohw * h->cout_p);
  L.dxm = take((size_t)B * hw * h->cin_p);
  const size_t po = (size_t)B * L.Qo * h->cout_p, pi = (size_t)B * L.Qi * h->cin_p;
  L.part = take(po > pi ? po : pi);
  L.R1 = take((size_t)B * h->cin_p);
  L.Tv = take((size_t)B * h->cout_p);
  L.ds = take((size_t)B * h->cin_p);
  L.total = off;
  return L;
}

// the vector layout kernels need whole 4 x 4 blocks and 128-bit accesses on both sides
bool scalar_layout_forced() {   // A/B switch, read once
  static const bool v = getenv("LFP_MC_SCALAR_LAYOUT") != nullptr;
  return v;
}
bool layout_vec_ok(const void* a, const void* b, const void* scale, int C, int Cp, int64_t hw) {
  return C % 4 == 0 && Cp % 16 == 0 && hw % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)scale) & 15) == 0 && !scalar_layout_forced();
}
int launch_layout_in(const float* in, float* out, const float* scale, int B, int C, int Cp, int64_t hw, cudaStream_t s) {
  if (layout_vec_ok(in, out, scale, C, Cp, hw)) {
    if (Cp % 32 == 0) {
      nchw_to_nhwc_vec_kernel<8><<<dim3((unsigned)ceil_div(hw, 8 * 16), (unsigned)B), 256, 0, s>>>(in, out, scale, C, Cp, hw);
    } else {
      nchw_to_nhwc_vec_kernel<4><<<dim3((unsigned)ceil_div(hw, 8 * 32), (unsigned)B), 256, 0, s>>>(in, out, scale, C, Cp, hw);
    }
    LFP_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)ceil_div(hw, 32), (unsigned)ceil_div(Cp, 32), (unsigned)B);
  nchw_to_nhwc_pad_kernel<<<grid, 256, 0, s>>>(in, out, scale, C, Cp, hw);
  LFP_LAUNCH_CHECK();
  return 0;
}
int launch_layout_out(const float* in, float* out, const float* scale, int B, int C, int Cp, int64_t hw, cudaStream_t s) {
  if (layout_vec_ok(in, out, scale, C, Cp, hw)) {
    if (Cp % 32 == 0) {
      nhwc_to_nchw_vec_kernel<8><<<dim3((unsigned)ceil_div(hw, 8 * 16), (unsigned)B), 256, 0, s>>>(in, out, scale, C, Cp, hw);
    } else {
      nhwc_to_nchw_vec_kernel<4><<<dim3((unsigned)ceil_div(hw, 8 * 32), (unsigned)B), 256, 0, s>>>(in, out, scale, C, Cp, hw);
    }
    LFP_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((unsigned)ceil_div(hw, 32), (unsigned)ceil_div(Cp, 32), (unsigned)B);
  nhwc_to_nchw_scale_kernel<<<grid, 256, 0, s>>>(in, out, scale, C, Cp, hw);
  LFP_LAUNCH_CHECK();
  return 0;
}
int launch_dot(const float* a, const float* b, float* partial, float* out, int B, int C, int64_t hw, int seglen, int Q, cudaStream_t s) {
  if (C % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)partial) & 15) == 0 && !scalar_layout_forced()) {
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)ceil_div(Q, 4), (unsigned)B);
    dot_partial_vec_kernel<<<grid, 256, 0, s>>>(a, b, partial, C, hw, seglen, Q);
  } else {
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)Q, (unsigned)B);
    dot_partial_kernel<<<grid, 256, 0, s>>>(a, b, partial, C, hw, seglen, Q);
  }
  LFP_LAUNCH_CHECK();
  return launch_partial_reduce(partial, out, B, Q, C, C, s);
}
}  // namespace

extern "C" int lfp_modconv_create(lfp_modconv** out, int in_channel, int out_channel, int kernel_size, int style_dim,
                                  int demodulate, int upsample, const float* blur_kernel_1d, int blur_taps) {
  LFP_CHECK_ARG(out != nullptr, "modconv_create: null out");
  LFP_CHECK_ARG(in_channel >= 1 && out_channel >= 1 && style_dim >= 4 && style_dim % 4 == 0, "modconv_create: bad channel counts / style_dim");
  LFP_CHECK_ARG(kernel_size == 1 || kernel_size == 3, "modconv_create: kernel size %d (the generator uses 1 and 3)", kernel_size);
  if (upsample && kernel_size != 3) { set_error("modconv_create: upsample needs a 3x3 kernel"); return LFP_EUNSUPPORTED; }
  if (blur_kernel_1d != nullptr && blur_taps != 4) { set_error("modconv_create: only 4-tap blur kernels are supported (got %d)", blur_taps); return LFP_EUNSUPPORTED; }
  lfp_modconv* h = new lfp_modconv();
  h->cin = in_channel; h->cout = out_channel; h->k = kernel_size; h->taps = kernel_size * kernel_size; h->style_dim = style_dim;
  h->demodulate = demodulate != 0; h->upsample = upsample != 0;
  // pad to what the gather kernels take; when that is also a tensor-core shape the tcgen05 kernel can run the layer
  h->cin_p = round_up(in_channel, in_channel >= 32 ? 32 : 16);
  h->cout_p = round_up(out_channel, out_channel >= 32 ? 32 : 16);   // also the reduction width of the data gradient
  if (blur_kernel_1d) memcpy(h->blur1d, blur_kernel_1d, 4 * sizeof(float));
  int rc = 0;
  const size_t wn = (size_t)h->taps * h->cin_p * h->cout_p;
  rc |= h->alloc(&h->W, (size_t)h->cout * h->cin * h->taps);
  rc |= h->alloc(&h->modw, (size_t)h->cin * style_dim);
  rc |= h->alloc(&h->modb, h->cin);
  rc |= h->alloc(&h->wf, wn); rc |= h->alloc(&h->wg, wn); rc |= h->alloc(&h->wf_t, wn); rc |= h->alloc(&h->wg_t, wn);
  rc |= h->alloc(&h->wsq, (size_t)h->cin_p * h->cout_p);
  rc |= h->alloc(&h->A, (size_t)h->cin_p * style_dim);
  rc |= h->alloc(&h->bvec, h->cin_p);
  rc |= h->alloc(&h->fir, 80);
  int* ip = nullptr;
  if (rc == 0 && cudaMalloc((void**)&ip, 3 * (size_t)h->cin_p * sizeof(int)) == cudaSuccess) {
    h->owned.push_back(ip);
    h->row_slot = ip; h->row_base = ip + h->cin_p; h->row_cin = ip + 2 * h->cin_p;
    std::vector<int> t(3 * (size_t)h->cin_p, 0);
    for (int i = 0; i < h->cin_p; ++i) t[2 * (size_t)h->cin_p + i] = h->cin_p;
    if (cudaMemcpy(ip, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) rc = LFP_ENOMEM;
  } else if (rc == 0) rc = LFP_ENOMEM;
  // FIR tables, as in the synthesis plan: blur forward (flipped taps, pad 1), adjoint (pad 2), separable factors
  float k2[16], tab[80] = {0}, sum = 0.f, s1 = 0.f;
  for (int i = 0; i < 4; ++i) { s1 += h->blur1d[i]; for (int j = 0; j < 4; ++j) { k2[i * 4 + j] = h->blur1d[i] * h->blur1d[j]; sum += k2[i * 4 + j]; } }
  for (int i = 0; i < 16; ++i) k2[i] = k2[i] / sum * 4.f;
  for (int ty = 0; ty < 4; ++ty)
    for (int tx = 0; tx < 4; ++tx) { tab[ty * 4 + tx] = k2[(3 - ty) * 4 + (3 - tx)]; tab[16 + ty * 4 + tx] = k2[ty * 4 + tx]; }
  for (int i = 0; i < 4; ++i) { tab[64 + i] = 2.f * h->blur1d[3 - i] / s1; tab[68 + i] = 2.f * h->blur1d[i] / s1; }
  if (rc == 0 && cudaMemcpy(h->fir, tab, sizeof(tab), cudaMemcpyHostToDevice) != cudaSuccess) rc = LFP_ENOMEM;
  if (rc != 0) { set_error("modconv_create: device allocation failed"); delete h; return rc; }
  *out = h;
  return 0;
}

extern "C" void lfp_modconv_destroy(lfp_modconv* h) { delete h; }

extern "C" int lfp_modconv_set_param(lfp_modconv* h, const char* name, const float* data, int64_t numel, void* stream) {
  LFP_CHECK_ARG(h && name && data, "modconv_set_param: null argument");
  const std::string n(name);
  float* dst = nullptr; int64_t want = -1;
  if (n == "weight") { dst = h->W; want = (int64_t)h->cout * h->cin * h->taps; }
  else if (n == "modulation.weight") { dst = h->modw; want = (int64_t)h->cin * h->style_dim; }
  else if (n == "modulation.bias") { dst = h->modb; want = h->cin; }
  LFP_CHECK_ARG(dst != nullptr, "modconv_set_param: unknown parameter '%s' (weight, modulation.weight, modulation.bias)", name);
  LFP_CHECK_ARG(want == numel, "modconv_set_param: '%s' expects %lld elements, got %lld", name, (long long)want, (long long)numel);
  LFP_CUDA(cudaMemcpyAsync(dst, data, numel * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  h->finalized = false;
  return 0;
}

extern "C" int lfp_modconv_finalize(lfp_modconv* h, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "modconv_finalize: null handle");
  cudaStream_t s = (cudaStream_t)stream;
  const float wscale = 1.f / sqrtf((float)(h->cin * h->taps));   // src/model.py:208-209
  const float mscale = 1.f / sqrtf((float)h->style_dim);         // EqualLinear, lr_mul = 1 (src/model.py:148)
  prep_conv_generic_kernel<<<(unsigned)ceil_div((int64_t)h->cin_p * h->cout_p, 256), 256, 0, s>>>(h->W, wscale, h->wf, h->wg, h->wsq, h->cin, h->cout,
                                                                                                  h->cin_p, h->cout_p, h->taps);
  LFP_LAUNCH_CHECK();
  const int64_t wn = (int64_t)h->taps * h->cin_p * h->cout_p;
  LFP_TRY(launch_round_tf32(h->wf, h->wf_t, wn, s));
  LFP_TRY(launch_round_tf32(h->wg, h->wg_t, wn, s));
  LFP_CUDA(cudaMemsetAsync(h->A, 0, (size_t)h->cin_p * h->style_dim * sizeof(float), s));
  LFP_CUDA(cudaMemsetAsync(h->bvec, 0, (size_t)h->cin_p * sizeof(float), s));
  LFP_TRY(launch_scale_copy(h->modw, h->A, mscale, (int64_t)h->cin * h->style_dim, s));
  LFP_TRY(launch_scale_copy(h->modb, h->bvec, 1.f, h->cin, s));
  // tensor-core eligibility: 3x3, channel counts the tcgen05 kernel tiles
  h->tc_fwd = h->k == 3 && tc_supported(h->cin_p, h->cout_p, 4, 4);
  h->tc_bwd = h->k == 3 && tc_supported(h->cout_p, h->cin_p, 4, 4);
  if (h->tc_fwd) LFP_TRY(tc_make_weight_maps(h->map_fwd, h->wg_t, 9 * h->cout_p, h->cin_p, h->cout_p));
  if (h->tc_bwd) LFP_TRY(tc_make_weight_maps(h->map_bwd, h->wf_t, 9 * h->cin_p, h->cout_p, h->cin_p));
  h->finalized = true;
  return 0;
}

extern "C" int lfp_modconv_out_size(const lfp_modconv* h, int in_h, int in_w, int* out_h, int* out_w) {
  LFP_CHECK_ARG(h && out_h && out_w && in_h >= 1 && in_w >= 1, "modconv_out_size: bad argument");
  *out_h = h->upsample ? 2 * in_h : in_h; *out_w = h->upsample ? 2 * in_w : in_w;
  return 0;
}

extern "C" size_t lfp_modconv_workspace_bytes(const lfp_modconv* h, int batch, int in_h, int in_w) {
  if (!h || batch <= 0 || in_h <= 0 || in_w <= 0) return 0;
  return mc_layout(h, batch, in_h, in_w).total * sizeof(float);
}

static int mc_check(const lfp_modconv* h, int batch, int H, int W, const void* ws, size_t ws_bytes, int precision, const McLayout& L) {
  LFP_CHECK_ARG(h != nullptr && ws != nullptr, "modconv: null handle or workspace");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535 && H >= 1 && W >= 1, "modconv: bad batch / extent");
  if (!h->finalized) { set_error("modconv: lfp_modconv_finalize has not been called since the last set_param"); return LFP_ESTATE; }
  if (ws_bytes < L.total * sizeof(float)) { set_error("modconv: workspace too small (%zu < %zu bytes)", ws_bytes, L.total * sizeof(float)); return LFP_ENOMEM; }
  LFP_CHECK_ARG(((uintptr_t)ws & 255) == 0, "modconv: workspace must be 256-byte aligned");
  LFP_CHECK_ARG(precision == LFP_PREC_FP32 || precision == LFP_PREC_TF32, "modconv: unknown precision mode %d", precision);
  return 0;
}

extern "C" int lfp_modconv_forward(lfp_modconv* h, int batch, int in_h, int in_w, const float* input, const float* style,
                                   float* out, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "modconv_forward: null handle");
  const McLayout L = mc_layout(h, batch, in_h, in_w);
  LFP_TRY(mc_check(h, batch, in_h, in_w, workspace, workspace_bytes, precision, L));
  LFP_CHECK_ARG(input && style && out, "modconv_forward: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const int B = batch, H = in_h, W = in_w, Kp = h->cin_p, Np = h->cout_p;
  float *x = ws + L.x, *sm = ws + L.s, *d = ws + L.d, *raw = ws + L.raw;
  LFP_TRY(launch_layout_in(input, x, nullptr, B, h->cin, Kp, (int64_t)H * W, s));
  LFP_TRY(launch_style_affine(style, h->A, h->bvec, h->row_slot, h->row_base, h->row_cin, sm, B, Kp, 1, h->style_dim, s));
  if (h->demodulate) LFP_TRY(launch_demod(sm, Kp, h->wsq, d, Np, B, Kp, Np, s));
  else { fill_kernel<<<(unsigned)ceil_div((int64_t)B * Np, 256), 256, 0, s>>>(d, 1.f, (int64_t)B * Np); LFP_LAUNCH_CHECK(); }
  const bool use_tc = precision == LFP_PREC_TF32 && h->tc_fwd && H >= 4 && W >= 4;
  if (!h->upsample) {
    ConvGeom g{};
    g.batch = B; g.gh = H; g.gw = W; g.in_h = H; g.in_w = W; g.in_bstride = (int64_t)H * W * Kp; g.in_stride = 1;
    g.out_h = H; g.out_w = W; g.out_stride = 1; g.K = Kp; g.N = Np; g.ntaps = h->taps;
    for (int t = 0; t < h->taps; ++t) {
      g.dy[t] = (signed char)(h->k == 3 ? t / 3 - 1 : 0); g.dx[t] = (signed char)(h->k == 3 ? t % 3 - 1 : 0); g.widx[t] = (signed char)t;
    }
    if (use_tc) {
      TcConv t{};
      t.in = x; t.in_planes = 1; t.in_h = H; t.in_w = W; t.in_bcast = false; t.mod = sm; t.wmap = h->map_fwd;
      t.out = raw; t.out_planes = 1; t.out_plane = 0; t.out_h = H; t.out_w = W; t.batch = B; t.gh = H; t.gw = W; t.K = Kp; t.N = Np;
      t.taps.ngroups = 1; t.taps.group_plane[0] = 0; t.taps.group_tap0[0] = 0; t.taps.group_tap0[1] = 9;
      for (int i = 0; i < 9; ++i) { t.taps.dy[i] = g.dy[i]; t.taps.dx[i] = g.dx[i]; t.taps.widx[i] = g.widx[i]; }
      t.epi = EPI_STORE;
      LFP_TRY(launch_conv_tc(t, s));
    } else {
      LFP_TRY(launch_conv_simt(x, sm, h->wf, raw, g, EPI_STORE, ConvEpiArgs{}, s));
    }
  } else {
    // stride-2 transposed conv as four sub-pixel phases into [B, 2H+1, 2W+1, Np] (src/model.py:269-279), then the Blur (:280-282)
    float* T = ws + L.T;
    // up to 128 output channels (or an 8 px input): all four phases in one launch, as the synthesis plan does - the input tile
    // is loaded and modulated once and every tap accumulates into its phase's TMEM accumulator (LFP_FUSE_PHASES=0: four launches)
    static const bool fuse_off = getenv("LFP_FUSE_PHASES") != nullptr && atoi(getenv("LFP_FUSE_PHASES")) == 0;
    const bool fuse_phases = use_tc && !fuse_off && H == W && (Np <= 128 || H <= 8);   // square: the configuration the plan runs
    if (fuse_phases) {
      TcConv q{};
      q.in = x; q.in_planes = 1; q.in_h = H; q.in_w = W; q.in_bcast = false; q.mod = sm; q.wmap = h->map_fwd;
      q.out = T; q.out_planes = 4; q.out_plane = 0; q.out_h = H + 1; q.out_w = W + 1; q.out_stride = 2;   // interleaved [2H+1, 2W+1]
      q.batch = B; q.gh = H + 1; q.gw = W + 1; q.K = Kp; q.N = Np;
      q.taps.ngroups = 1; q.taps.group_plane[0] = 0; q.taps.group_tap0[0] = 0; q.taps.nphase = 4;
      int t = 0;
      for (int a = 0; a < 2; ++a)
        for (int bb = 0; bb < 2; ++bb)
          for (int ky = a == 0 ? 0 : 1; ky < 3; ky += 2)
            for (int kx = bb == 0 ? 0 : 1; kx < 3; kx += 2) {
              q.taps.dy[t] = (signed char)(ky == 2 ? -1 : 0); q.taps.dx[t] = (signed char)(kx == 2 ? -1 : 0);
              q.taps.widx[t] = (signed char)(ky * 3 + kx); q.taps.acc[t] = (signed char)(a * 2 + bb); ++t;
            }
      q.taps.group_tap0[1] = t;
      q.epi = EPI_STORE;
      LFP_TRY(launch_conv_tc(q, s));
    } else
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb) {
        ConvGeom g{};
        g.batch = B; g.gh = a == 0 ? H + 1 : H; g.gw = bb == 0 ? W + 1 : W;
        g.in_h = H; g.in_w = W; g.in_bstride = (int64_t)H * W * Kp; g.in_stride = 1;
        g.out_h = 2 * H + 1; g.out_w = 2 * W + 1; g.out_stride = 2; g.out_oy = a; g.out_ox = bb; g.K = Kp; g.N = Np;
        int t = 0;
        for (int ky = a == 0 ? 0 : 1; ky < 3; ky += 2)
          for (int kx = bb == 0 ? 0 : 1; kx < 3; kx += 2) {
            g.dy[t] = (signed char)(ky == 2 ? -1 : 0); g.dx[t] = (signed char)(kx == 2 ? -1 : 0); g.widx[t] = (signed char)(ky * 3 + kx); ++t;
          }
        g.ntaps = t;
        if (use_tc) {
          TcConv q{};
          q.in = x; q.in_planes = 1; q.in_h = H; q.in_w = W; q.in_bcast = false; q.mod = sm; q.wmap = h->map_fwd;
          q.out = T; q.out_planes = 1; q.out_plane = 0; q.out_h = 2 * H + 1; q.out_w = 2 * W + 1; q.out_stride = 2; q.out_oy = a; q.out_ox = bb;
          q.batch = B; q.gh = g.gh; q.gw = g.gw; q.K = Kp; q.N = Np;
          q.taps.ngroups = 1; q.taps.group_plane[0] = 0; q.taps.group_tap0[0] = 0; q.taps.group_tap0[1] = t;
          for (int i = 0; i < t; ++i) { q.taps.dy[i] = g.dy[i]; q.taps.dx[i] = g.dx[i]; q.taps.widx[i] = g.widx[i]; }
          q.epi = EPI_STORE;
          LFP_TRY(launch_conv_tc(q, s));
        } else {
          LFP_TRY(launch_conv_simt(x, sm, h->wf, T, g, EPI_STORE, ConvEpiArgs{}, s));
        }
      }
    FirArgs f{};
    f.kx = f.ky = h->fir + 64;
    f.batch = B; f.in_h = 2 * H + 1; f.in_w = 2 * W + 1; f.out_h = 2 * H; f.out_w = 2 * W; f.C = Np; f.pad = 1; f.coef = h->fir + 0;
    LFP_TRY(launch_fir4x4_nhwc(T, raw, f, s));
  }
  LFP_TRY(launch_layout_out(raw, out, d, B, h->cout, Np, (int64_t)L.oh * L.ow, s));
  h->fwd_ws = workspace; h->fwd_batch = batch; h->fwd_h = in_h; h->fwd_w = in_w; h->fwd_prec = precision;
  return 0;
}

extern "C" int lfp_modconv_backward(lfp_modconv* h, int batch, int in_h, int in_w, const float* d_out, float* d_input,
                                    float* d_style, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "modconv_backward: null handle");
  const McLayout L = mc_layout(h, batch, in_h, in_w);
  LFP_TRY(mc_check(h, batch, in_h, in_w, workspace, workspace_bytes, precision, L));
  LFP_CHECK_ARG(d_out && (d_input || d_style), "modconv_backward: null argument");
  if (h->fwd_ws != workspace || h->fwd_batch != batch || h->fwd_h != in_h || h->fwd_w != in_w || h->fwd_prec != precision) {
    set_error("modconv_backward: no matching forward on this workspace");
    return LFP_ESTATE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const int B = batch, H = in_h, W = in_w, Kp = h->cin_p, Np = h->cout_p;
  const int64_t hw = (int64_t)H * W, ohw = (int64_t)L.oh * L.ow;
  float *x = ws + L.x, *sm = ws + L.s, *d = ws + L.d, *raw = ws + L.raw, *g = ws + L.g, *dxm = ws + L.dxm;
  // dRaw = dOut * demod (NHWC), T[b, co] = sum_pix dOut out = sum_pix dRaw raw
  LFP_TRY(launch_layout_in(d_out, g, d, B, h->cout, Np, ohw, s));
  if (h->demodulate) LFP_TRY(launch_dot(g, raw, ws + L.part, ws + L.Tv, B, Np, ohw, L.seglen_o, L.Qo, s));
  const bool use_tc = precision == LFP_PREC_TF32 && h->tc_bwd && H >= 4 && W >= 4;
  ConvGeom gg{};
  gg.batch = B; gg.gh = H; gg.gw = W; gg.K = Np; gg.N = Kp; gg.out_h = H; gg.out_w = W; gg.out_stride = 1;
  TcConv tq{};
  tq.in_bcast = false; tq.mod = nullptr; tq.wmap = h->map_bwd; tq.out = dxm; tq.out_planes = 1; tq.out_plane = 0; tq.out_h = H; tq.out_w = W;
  tq.batch = B; tq.gh = H; tq.gw = W; tq.K = Np; tq.N = Kp; tq.epi = EPI_STORE;
  const float* gin = g;
  if (!h->upsample) {
    gg.in_h = H; gg.in_w = W; gg.in_stride = 1; gg.in_bstride = ohw * Np; gg.ntaps = h->taps;
    for (int t = 0; t < h->taps; ++t) {
      gg.dy[t] = (signed char)(h->k == 3 ? 1 - t / 3 : 0); gg.dx[t] = (signed char)(h->k == 3 ? 1 - t % 3 : 0); gg.widx[t] = (signed char)t;
    }
    tq.in_planes = 1; tq.in_h = H; tq.in_w = W;
    tq.taps.ngroups = 1; tq.taps.group_plane[0] = 0; tq.taps.group_tap0[0] = 0; tq.taps.group_tap0[1] = 9;
    for (int i = 0; i < 9; ++i) { tq.taps.dy[i] = (signed char)(1 - i / 3); tq.taps.dx[i] = (signed char)(1 - i % 3); tq.taps.widx[i] = (signed char)i; }
  } else {
    // adjoint of the Blur: [2H, 2W] -> [2H+1, 2W+1], un-flipped taps, pad 2 (src/op/upfirdn2d.py:112-115); phase-major for the tensor-core kernel
    float* T = ws + L.T;
    FirArgs f{};
    f.batch = B; f.in_h = 2 * H; f.in_w = 2 * W; f.out_h = 2 * H + 1; f.out_w = 2 * W + 1; f.C = Np; f.pad = 2; f.coef = h->fir + 16;
    f.out_planar = use_tc && H == W;   // the phase-major store assumes square planes
    f.kx = f.ky = h->fir + 68;
    const bool planar = f.out_planar;
    LFP_TRY(launch_fir4x4_nhwc(g, T, f, s));
    gin = T;
    gg.in_h = 2 * H + 1; gg.in_w = 2 * W + 1; gg.in_stride = 2; gg.in_bstride = (int64_t)(2 * H + 1) * (2 * W + 1) * Np; gg.ntaps = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) { const int t = ky * 3 + kx; gg.dy[t] = (signed char)ky; gg.dx[t] = (signed char)kx; gg.widx[t] = (signed char)t; }
    if (planar) {
      tq.in_planes = 4; tq.in_h = H + 1; tq.in_w = W + 1; tq.taps.ngroups = 4;
      int t = 0;
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          const int gi = py * 2 + px;
          tq.taps.group_plane[gi] = gi; tq.taps.group_tap0[gi] = t;
          for (int ky = py; ky < 3; ky += 2)
            for (int kx = px; kx < 3; kx += 2) { tq.taps.dy[t] = (signed char)(ky >> 1); tq.taps.dx[t] = (signed char)(kx >> 1); tq.taps.widx[t] = (signed char)(ky * 3 + kx); ++t; }
        }
      tq.taps.group_tap0[4] = t;
    }
    if (use_tc && planar) { tq.in = gin; LFP_TRY(launch_conv_tc(tq, s)); }
    else LFP_TRY(launch_conv_simt(gin, nullptr, h->wg, dxm, gg, EPI_STORE, ConvEpiArgs{}, s));
  }
  if (!h->upsample) {
    if (use_tc) { tq.in = gin; LFP_TRY(launch_conv_tc(tq, s)); }
    else LFP_TRY(launch_conv_simt(gin, nullptr, h->wg, dxm, gg, EPI_STORE, ConvEpiArgs{}, s));
  }
  if (d_input) LFP_TRY(launch_layout_out(dxm, d_input, sm, B, h->cin, Kp, hw, s));
  if (d_style) {
    float* R1 = ws + L.R1; float* ds = ws + L.ds;
    LFP_TRY(launch_dot(x, dxm, ws + L.part, R1, B, Kp, hw, L.seglen_i, L.Qi, s));
    if (h->demodulate) LFP_TRY(launch_style_grad(R1, sm, Kp, ws + L.Tv, d, Np, h->wsq, ds, B, Kp, Np, s));
    else ds = R1;
    dstyle_kernel<<<dim3((unsigned)ceil_div(h->style_dim, 128), (unsigned)B), 128, 0, s>>>(ds, h->A, d_style, Kp, h->style_dim);
    LFP_LAUNCH_CHECK();
  }
  return 0;
}
