// Kernels of the fused synthesis path: fp32 CUDA-core gather convolution with fused epilogues,
// NHWC FIR, ToRGB, activation backward and the small style / demodulation kernels.
// Algebra: the reference's activation-modulated ("unfused") form, src/model.py:229-256, which
// keeps the weights shared across the batch and makes the style gradient two reductions
// (SURVEY.md 7.3 "Style gradient without wgrad").  Reductions use fixed-order partials, so a
// trajectory's result does not depend on which other trajectories share its batch or GPU.
#include "synth_kernels.cuh"

namespace lfp {

__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 c) { return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w)); }
__device__ __forceinline__ float4 f4_fma4(float4 a, float4 b, float4 c) { return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w)); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ float lrelu_fwd(float v) { return (v > 0.f ? v : v * kLreluSlope) * kLreluGain; }

// =============================================================================================
// Gather convolution, CUDA cores (fp32 path)
// =============================================================================================
struct ConvKArgs {
  const float* in; const float* mod; const float* wtab; float* out;
  ConvGeom g; ConvEpiArgs e; int seglen;
};

template <int TN, int EPI, bool MOD>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvKArgs a) {
  constexpr int KC = 16, TM = 128, LDA = TM + 4;
  constexpr int NT = TN / 4;    // threads along n
  constexpr int NG = 256 / NT;  // row groups
  constexpr int TMR = TM / NG;  // rows per thread (8 for TN=64, 4 for TN=32)
  __shared__ __align__(16) float As[2][KC][LDA];
  __shared__ __align__(16) float Bs[2][KC][TN];

  const ConvGeom& g = a.g;
  const int tid = threadIdx.x;
  const int pix_per = g.gh * g.gw;
  const int64_t total_pix = (int64_t)g.batch * pix_per;
  const int64_t tile0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int K = g.K, N = g.N;

  // ---- loader roles ----
  const int a_row = tid >> 2;
  const int a_kq = (tid & 3) * 4;
  int lb[2], lgy[2], lgx[2];
  bool lvalid[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int64_t p = tile0 + a_row + 64 * i;
    lvalid[i] = p < total_pix;
    const int64_t pp = lvalid[i] ? p : 0;
    lb[i] = (int)(pp / pix_per);
    const int rem = (int)(pp - (int64_t)lb[i] * pix_per);
    lgy[i] = rem / g.gw;
    lgx[i] = rem - lgy[i] * g.gw;
  }
  const int b_row = tid / NT;
  const int b_c4 = tid % NT;
  const bool b_active = b_row < KC && (n0 + b_c4 * 4) < N;

  const int kchunks = K / KC;
  const int nit = g.ntaps * kchunks;

  float4 ar[2], br;
  auto load = [&](int it) {
    const int tap = it / kchunks;
    const int k0 = (it - tap * kchunks) * KC;
    const int dy = g.dy[tap], dx = g.dx[tap], wi = g.widx[tap];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int iy = lgy[i] * g.in_stride + dy, ix = lgx[i] * g.in_stride + dx;
      const bool ok = lvalid[i] && iy >= 0 && iy < g.in_h && ix >= 0 && ix < g.in_w;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        v = __ldg(reinterpret_cast<const float4*>(a.in + (int64_t)lb[i] * g.in_bstride +
                                                  ((int64_t)iy * g.in_w + ix) * K + k0 + a_kq));
        if (MOD) v = f4_mul(v, __ldg(reinterpret_cast<const float4*>(a.mod + (int64_t)lb[i] * K + k0 + a_kq)));
      }
      ar[i] = v;
    }
    br = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_active)
      br = __ldg(reinterpret_cast<const float4*>(a.wtab + ((int64_t)wi * K + k0 + b_row) * N + n0 + b_c4 * 4));
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][a_kq + 0][a_row + 64 * i] = ar[i].x;
      As[buf][a_kq + 1][a_row + 64 * i] = ar[i].y;
      As[buf][a_kq + 2][a_row + 64 * i] = ar[i].z;
      As[buf][a_kq + 3][a_row + 64 * i] = ar[i].w;
    }
    if (b_row < KC) *reinterpret_cast<float4*>(&Bs[buf][b_row][b_c4 * 4]) = br;
  };

  const int tm = tid / NT, tn = tid % NT;
  float acc[TMR][4];
#pragma unroll
  for (int r = 0; r < TMR; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  load(0);
  store(0);
  __syncthreads();
  for (int it = 0; it < nit; ++it) {
    const int buf = it & 1;
    if (it + 1 < nit) load(it + 1);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float av[TMR];
#pragma unroll
      for (int q = 0; q < TMR / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][tm * TMR + q * 4]);
        av[q * 4 + 0] = t.x; av[q * 4 + 1] = t.y; av[q * 4 + 2] = t.z; av[q * 4 + 3] = t.w;
      }
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn * 4]);
#pragma unroll
      for (int r = 0; r < TMR; ++r) {
        acc[r][0] = fmaf(av[r], bv.x, acc[r][0]);
        acc[r][1] = fmaf(av[r], bv.y, acc[r][1]);
        acc[r][2] = fmaf(av[r], bv.z, acc[r][2]);
        acc[r][3] = fmaf(av[r], bv.w, acc[r][3]);
      }
    }
    if (it + 1 < nit) store(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const int n = n0 + tn * 4;
  const bool n_ok = n < N;
  float4 red = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float nw = 0.f;
  if (EPI == EPI_ACT && n_ok) {
    bias4 = __ldg(reinterpret_cast<const float4*>(a.e.bias + n));
    nw = __ldg(a.e.noise_w);
  }
  if (EPI == EPI_RELU && n_ok) bias4 = __ldg(reinterpret_cast<const float4*>(a.e.bias + n));
#pragma unroll
  for (int r = 0; r < TMR; ++r) {
    const int64_t p = tile0 + tm * TMR + r;
    if (p >= total_pix || !n_ok) continue;
    const int b = (int)(p / pix_per);
    const int rem = (int)(p - (int64_t)b * pix_per);
    const int gy = rem / g.gw, gx = rem - gy * g.gw;
    float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    if (EPI == EPI_ACT) {
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.e.demod + (int64_t)b * N + n));
      const float nz = nw * __ldg(a.e.noise + (int64_t)b * a.e.noise_bstride + rem);
      v.x = lrelu_fwd(fmaf(v.x, d4.x, nz) + bias4.x);
      v.y = lrelu_fwd(fmaf(v.y, d4.y, nz) + bias4.y);
      v.z = lrelu_fwd(fmaf(v.z, d4.z, nz) + bias4.z);
      v.w = lrelu_fwd(fmaf(v.w, d4.w, nz) + bias4.w);
    } else if (EPI == EPI_DGRAD) {
      const float4 xs = __ldg(reinterpret_cast<const float4*>(a.e.xsave + (int64_t)b * a.e.xsave_bstride + (int64_t)rem * N + n));
      red = f4_fma4(xs, v, red);
      v = f4_mul(v, __ldg(reinterpret_cast<const float4*>(a.e.mod_out + (int64_t)b * N + n)));
    } else if (EPI == EPI_RELU) {
      v.x = fmaxf(v.x + bias4.x, 0.f); v.y = fmaxf(v.y + bias4.y, 0.f); v.z = fmaxf(v.z + bias4.z, 0.f); v.w = fmaxf(v.w + bias4.w, 0.f);
    } else if (EPI == EPI_DGRAD_RELU) {
      const float4 xs = __ldg(reinterpret_cast<const float4*>(a.e.xsave + (int64_t)b * a.e.xsave_bstride + (int64_t)rem * N + n));
      v.x = xs.x > 0.f ? v.x : 0.f; v.y = xs.y > 0.f ? v.y : 0.f; v.z = xs.z > 0.f ? v.z : 0.f; v.w = xs.w > 0.f ? v.w : 0.f;
    }
    if (a.out != nullptr) {
      const int oy = gy * g.out_stride + g.out_oy, ox = gx * g.out_stride + g.out_ox;
      *reinterpret_cast<float4*>(a.out + (((int64_t)b * g.out_h + oy) * g.out_w + ox) * N + n) = v;
    }
  }
  if (EPI == EPI_DGRAD) {
    // fixed-order reduction over the rows of each segment (seglen pixels of one sample)
    float* redbuf = &As[0][0][0];  // NG x TN floats
    *reinterpret_cast<float4*>(&redbuf[tm * TN + tn * 4]) = red;
    __syncthreads();
    const int gps = a.seglen / TMR;        // row groups per segment
    const int nseg = TM / a.seglen;        // segments per tile
    for (int t = tid; t < nseg * TN; t += 256) {
      const int seg = t / TN, nn = t - seg * TN;
      const int64_t p0 = tile0 + (int64_t)seg * a.seglen;
      if (p0 >= total_pix || n0 + nn >= N) continue;
      float sacc = 0.f;
      for (int q = 0; q < gps; ++q) sacc += redbuf[(seg * gps + q) * TN + nn];
      a.e.partial[(p0 / a.seglen) * N + n0 + nn] = sacc;
    }
  }
}

int conv_dgrad_seglen(const ConvGeom& g) {
  const int hw = g.gh * g.gw;
  return hw < 128 ? hw : 128;
}

template <int TN, int EPI>
static int conv_simt_launch2(const ConvKArgs& ka, bool mod, dim3 grid, cudaStream_t s) {
  if (mod) conv_simt_kernel<TN, EPI, true><<<grid, 256, 0, s>>>(ka);
  else conv_simt_kernel<TN, EPI, false><<<grid, 256, 0, s>>>(ka);
  LFP_LAUNCH_CHECK();
  return 0;
}

int launch_conv_simt(const float* in, const float* mod, const float* wtab, float* out,
                     const ConvGeom& g, int epi, const ConvEpiArgs& e, cudaStream_t s) {
  LFP_CHECK_ARG(g.K % 16 == 0 && g.N % 4 == 0, "conv: K=%d must be a multiple of 16, N=%d of 4", g.K, g.N);
  LFP_CHECK_ARG(g.ntaps >= 1 && g.ntaps <= 9, "conv: bad tap count");
  ConvKArgs ka{in, mod, wtab, out, g, e, 0};
  const int64_t total_pix = (int64_t)g.batch * g.gh * g.gw;
  if (total_pix == 0) return 0;
  if (epi == EPI_DGRAD) {
    const int hw = g.gh * g.gw;
    LFP_CHECK_ARG((hw & (hw - 1)) == 0 && hw >= 8, "dgrad: grid %dx%d must be a power of two >= 8 pixels", g.gh, g.gw);
    ka.seglen = conv_dgrad_seglen(g);
  }
  const int TN = g.N >= 64 ? 64 : 32;
  dim3 grid((unsigned)ceil_div(total_pix, 128), (unsigned)ceil_div(g.N, TN));
  const bool m = mod != nullptr;
  LFP_CHECK_ARG(epi == EPI_STORE || epi == EPI_ACT || epi == EPI_DGRAD || epi == EPI_RELU || epi == EPI_DGRAD_RELU,
                "conv: epilogue %d is not available on the CUDA-core kernel", epi);
  if (TN == 64) {
    if (epi == EPI_STORE) return conv_simt_launch2<64, EPI_STORE>(ka, m, grid, s);
    if (epi == EPI_ACT) return conv_simt_launch2<64, EPI_ACT>(ka, m, grid, s);
    if (epi == EPI_RELU) return conv_simt_launch2<64, EPI_RELU>(ka, m, grid, s);
    if (epi == EPI_DGRAD_RELU) return conv_simt_launch2<64, EPI_DGRAD_RELU>(ka, m, grid, s);
    return conv_simt_launch2<64, EPI_DGRAD>(ka, m, grid, s);
  }
  if (epi == EPI_STORE) return conv_simt_launch2<32, EPI_STORE>(ka, m, grid, s);
  if (epi == EPI_ACT) return conv_simt_launch2<32, EPI_ACT>(ka, m, grid, s);
  if (epi == EPI_RELU) return conv_simt_launch2<32, EPI_RELU>(ka, m, grid, s);
  if (epi == EPI_DGRAD_RELU) return conv_simt_launch2<32, EPI_DGRAD_RELU>(ka, m, grid, s);
  return conv_simt_launch2<32, EPI_DGRAD>(ka, m, grid, s);
}

// =============================================================================================
// 4x4 FIR on NHWC with a register sliding window (blur after the transposed conv, and its adjoint)
// =============================================================================================
// Each thread owns a channel quad of XO adjacent output columns and slides a 4 x (XO+3) window of
// input pixels down RY output rows: (XO+3) 128-bit loads per XO outputs per row.
template <bool ACT, int XO>
__global__ void __launch_bounds__(256) fir4x4_nhwc_kernel(const float* __restrict__ in,
                                                          float* __restrict__ out, const FirArgs a,
                                                          int C4, int PX, int RY) {
  constexpr int WC = XO + 3;
  const int c4 = threadIdx.x % C4;
  const int px = threadIdx.x / C4;
  const int ox0 = (blockIdx.x * PX + px) * XO;
  const int oy0 = blockIdx.y * RY;
  const int b = blockIdx.z;
  if (px >= PX || ox0 >= a.out_w) return;
  float coef[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) coef[i] = __ldg(a.coef + i);
  const int C = a.C;
  const int iph = (a.in_h + 1) >> 1, ipw = (a.in_w + 1) >> 1, oph = (a.out_h + 1) >> 1, opw = (a.out_w + 1) >> 1;
  const int64_t in_per = a.in_planar ? (int64_t)4 * iph * ipw : (int64_t)a.in_h * a.in_w;
  const int64_t out_per = a.out_planar ? (int64_t)4 * oph * opw : (int64_t)a.out_h * a.out_w;
  const float* src = in + (int64_t)b * in_per * C + c4 * 4;
  float4 win[4][WC];
  auto load_row = [&](int iy, float4(&row)[WC]) {
#pragma unroll
    for (int tx = 0; tx < WC; ++tx) {
      const int ix = ox0 + tx - a.pad;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < a.in_h && ix >= 0 && ix < a.in_w) {
        int64_t e = (int64_t)iy * a.in_w + ix;
        if (a.in_planar) e = ((int64_t)((iy & 1) * 2 + (ix & 1)) * iph + (iy >> 1)) * ipw + (ix >> 1);
        v = __ldg(reinterpret_cast<const float4*>(src + e * C));
      }
      row[tx] = v;
    }
  };
  load_row(oy0 - a.pad + 0, win[0]);
  load_row(oy0 - a.pad + 1, win[1]);
  load_row(oy0 - a.pad + 2, win[2]);
  float4 d4 = make_float4(1.f, 1.f, 1.f, 1.f), bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float nw = 0.f;
  if (ACT) {
    d4 = __ldg(reinterpret_cast<const float4*>(a.demod + (int64_t)b * C + c4 * 4));
    bias4 = __ldg(reinterpret_cast<const float4*>(a.bias + c4 * 4));
    nw = __ldg(a.noise_w);
  }
  for (int o = 0; o < RY; o += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int oy = oy0 + o + u;
      if (oy >= a.out_h) return;
      load_row(oy + 3 - a.pad, win[(u + 3) & 3]);
#pragma unroll
      for (int xo = 0; xo < XO; ++xo) {
        const int ox = ox0 + xo;
        if (ox >= a.out_w) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) acc = f4_fma(coef[r * 4 + tx], win[(u + r) & 3][xo + tx], acc);
        if (ACT) {
          const float nz = nw * __ldg(a.noise + (int64_t)b * a.noise_bstride + (int64_t)oy * a.out_w + ox);
          acc.x = lrelu_fwd(fmaf(acc.x, d4.x, nz) + bias4.x);
          acc.y = lrelu_fwd(fmaf(acc.y, d4.y, nz) + bias4.y);
          acc.z = lrelu_fwd(fmaf(acc.z, d4.z, nz) + bias4.z);
          acc.w = lrelu_fwd(fmaf(acc.w, d4.w, nz) + bias4.w);
        }
        int64_t oe = (int64_t)oy * a.out_w + ox;
        if (a.out_planar) oe = ((int64_t)((oy & 1) * 2 + (ox & 1)) * oph + (oy >> 1)) * opw + (ox >> 1);
        *reinterpret_cast<float4*>(out + ((int64_t)b * out_per + oe) * C + c4 * 4) = acc;
      }
    }
  }
}

// Separable variant: horizontal 4-tap pass on every input row as it is loaded (XO results per row), vertical 4-tap pass
// over a sliding window of the last four horizontal rows.
template <bool ACT, int XO>
__global__ void __launch_bounds__(256) fir4x4_sep_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                              const FirArgs a, int C4, int PX, int RY) {
  constexpr int WC = XO + 3;
  const int c4 = threadIdx.x % C4;
  const int px = threadIdx.x / C4;
  const int ox0 = (blockIdx.x * PX + px) * XO;
  const int oy0 = blockIdx.y * RY;
  const int b = blockIdx.z;
  if (px >= PX || ox0 >= a.out_w) return;
  float kx[4], ky[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { kx[i] = __ldg(a.kx + i); ky[i] = __ldg(a.ky + i); }
  const int C = a.C;
  const int iph = (a.in_h + 1) >> 1, ipw = (a.in_w + 1) >> 1, oph = (a.out_h + 1) >> 1, opw = (a.out_w + 1) >> 1;
  const int64_t in_per = a.in_planar ? (int64_t)4 * iph * ipw : (int64_t)a.in_h * a.in_w;
  const int64_t out_per = a.out_planar ? (int64_t)4 * oph * opw : (int64_t)a.out_h * a.out_w;
  const float* src = in + (int64_t)b * in_per * C + c4 * 4;
  float4 hwin[4][XO];
  auto hrow = [&](int iy, float4(&h)[XO]) {
    float4 v[WC];
#pragma unroll
    for (int tx = 0; tx < WC; ++tx) {
      const int ix = ox0 + tx - a.pad;
      v[tx] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < a.in_h && ix >= 0 && ix < a.in_w) {
        int64_t e = (int64_t)iy * a.in_w + ix;
        if (a.in_planar) e = ((int64_t)((iy & 1) * 2 + (ix & 1)) * iph + (iy >> 1)) * ipw + (ix >> 1);
        v[tx] = __ldg(reinterpret_cast<const float4*>(src + e * C));
      }
    }
#pragma unroll
    for (int xo = 0; xo < XO; ++xo) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int tx = 0; tx < 4; ++tx) acc = f4_fma(kx[tx], v[xo + tx], acc);
      h[xo] = acc;
    }
  };
  hrow(oy0 - a.pad + 0, hwin[0]);
  hrow(oy0 - a.pad + 1, hwin[1]);
  hrow(oy0 - a.pad + 2, hwin[2]);
  float4 d4 = make_float4(1.f, 1.f, 1.f, 1.f), bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float nw = 0.f;
  if (ACT) {
    d4 = __ldg(reinterpret_cast<const float4*>(a.demod + (int64_t)b * C + c4 * 4));
    bias4 = __ldg(reinterpret_cast<const float4*>(a.bias + c4 * 4));
    nw = __ldg(a.noise_w);
  }
  for (int o = 0; o < RY; o += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int oy = oy0 + o + u;
      if (oy >= a.out_h) return;
      hrow(oy + 3 - a.pad, hwin[(u + 3) & 3]);
#pragma unroll
      for (int xo = 0; xo < XO; ++xo) {
        const int ox = ox0 + xo;
        if (ox >= a.out_w) break;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r) acc = f4_fma(ky[r], hwin[(u + r) & 3][xo], acc);
        if (ACT) {
          const float nz = nw * __ldg(a.noise + (int64_t)b * a.noise_bstride + (int64_t)oy * a.out_w + ox);
          acc.x = lrelu_fwd(fmaf(acc.x, d4.x, nz) + bias4.x);
          acc.y = lrelu_fwd(fmaf(acc.y, d4.y, nz) + bias4.y);
          acc.z = lrelu_fwd(fmaf(acc.z, d4.z, nz) + bias4.z);
          acc.w = lrelu_fwd(fmaf(acc.w, d4.w, nz) + bias4.w);
        }
        int64_t oe = (int64_t)oy * a.out_w + ox;
        if (a.out_planar) oe = ((int64_t)((oy & 1) * 2 + (ox & 1)) * oph + (oy >> 1)) * opw + (ox >> 1);
        *reinterpret_cast<float4*>(out + ((int64_t)b * out_per + oe) * C + c4 * 4) = acc;
      }
    }
  }
}

// ---- row-ring variant for the wide layers (dense NHWC input, C <= 128) -----------------------------------------------------
// The register-streaming kernel above keeps the bytes it has in flight in registers, which at HBM latency under load is not
// enough to saturate the memory system.  Here one producer thread streams the input rows of a pixel segment into a
// shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier transaction bytes; an NHWC row piece is contiguous and
// 16-byte aligned) while eight consumer warps run the same separable passes out of shared memory: horizontal 4-tap pass per
// input row (LDS.128, one float4 of channels per lane), vertical pass over the last four horizontal rows held in registers,
// fused noise/bias/lrelu epilogue.  The sums are formed in the same order as in fir4x4_sep_nhwc_kernel: identical results.
namespace nring {
constexpr int NCW = 8;  // consumer warps
__host__ __device__ constexpr int segp(int C, int NV) { return NCW * 32 * 4 * NV / C; }          // output pixels per CTA row
__host__ __device__ constexpr int rowb(int C, int NV) { return (segp(C, NV) + 3) * C * 4; }      // bytes of a staged row piece
constexpr int NZB = 128;  // bytes reserved per slot for the noise of the output row the staged input row completes (<= 32 floats)
__host__ __device__ constexpr int slotb(int C, int NV) { return rowb(C, NV) + NZB; }
__host__ __device__ constexpr int depth(int C, int NV) { return 49152 / slotb(C, NV) < 12 ? 49152 / slotb(C, NV) : 12; }
__host__ __device__ constexpr int smem_bytes(int C, int NV) { return depth(C, NV) * slotb(C, NV) + 2 * depth(C, NV) * 8 + 16; }
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
}  // namespace nring

template <int C, int NV, bool ACT, bool OUT_PLANAR>
__global__ void __launch_bounds__((nring::NCW + 1) * 32, 2) fir_ring_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                            const FirArgs a, int RB) {
  using namespace nring;
  constexpr int SEGP = segp(C, NV), ROWB = rowb(C, NV), SLOTB = slotb(C, NV), D = depth(C, NV);
  extern __shared__ __align__(128) uint8_t nring_smem[];
  const uint32_t sbase = smem_u32(nring_smem);
  const uint32_t bars = sbase + D * SLOTB;  // full[D], empty[D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * SEGP, r0 = blockIdx.y * RB;
  const int rows = min(RB, a.out_h - r0);
  const int total = rows + 3;          // input rows feeding the band
  const int iy0 = r0 - a.pad;
  const int pfirst = ox0 - a.pad;      // input pixel of tap 0 of the segment's first output pixel
  const int pbase = max(pfirst, 0), pend = min(pfirst + SEGP + 3, a.in_w);
  const int npx = pend - pbase;        // staged pixels per row
  const int jlo = max(0, -iy0), jhi = npx > 0 ? min(total, a.in_h - iy0) : 0;   // staged input rows of the band
  if (threadIdx.x == 0) {
    for (int i = 0; i < D; ++i) { mbar_init(bars + i * 8, 1); mbar_init(bars + (D + i) * 8, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NCW) {
    // ---- producer ----
    if (lane == 0) {
      const float* src = in + (((int64_t)b * a.in_h + iy0) * a.in_w + pbase) * C;
      const uint32_t bytes = (uint32_t)(npx > 0 ? npx : 0) * C * 4;
      // ACT: the noise of output row j - 3 (one value per pixel of the segment) rides in the slot of input row j
      const float* nsrc = ACT ? a.noise + (int64_t)b * a.noise_bstride + (int64_t)(r0 - 3) * a.out_w + ox0 : nullptr;
      const uint32_t nbytes = ACT ? (uint32_t)min(SEGP, a.out_w - ox0) * 4 : 0u;
      uint32_t slot_addr = sbase, full_bar = bars, empty_bar = bars + D * 8, par = 0;
      int slot = 0;
      for (int j = 0; j < total; ++j) {
        if (j >= D) mbar_wait(empty_bar, par);
        const bool ok = j >= jlo && j < jhi;
        const bool nok = ACT && j >= 3;
        mbar_expect_tx(full_bar, (ok ? bytes : 0u) + (nok ? nbytes : 0u));
        if (ok) bulk_g2s(slot_addr, src + (int64_t)j * a.in_w * C, bytes, full_bar);
        if (nok) bulk_g2s(slot_addr + ROWB, nsrc + (int64_t)j * a.out_w, nbytes, full_bar);
        slot_addr += SLOTB; full_bar += 8; empty_bar += 8;
        if (++slot == D) { slot = 0; slot_addr = sbase; full_bar = bars; empty_bar = bars + D * 8; if (j >= D) par ^= 1u; }
      }
    }
    return;
  }
  // ---- consumers ----
  float kx[4], ky[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { kx[i] = __ldg(a.kx + i); ky[i] = __ldg(a.ky + i); }
  const int t = threadIdx.x;
  uint32_t loff[NV];          // byte offset of tap 0 in a staged row
  uint32_t tokmask = 0;       // bit (v*4 + tx): tap inside the image
  int stok[NV];
  float4 d4[NV], bias4[NV];
  uint32_t nzoff[NV];         // byte offset of the lane's noise value in a slot
  float* op0[NV];             // output pointer of the band's first row (dense), or of the even / odd rows (phase-major)
  float* op1[NV];
  int64_t ostep;              // floats between consecutive rows of one output pointer
  const int oph = (a.out_h + 1) >> 1, opw = (a.out_w + 1) >> 1;
  ostep = OUT_PLANAR ? (int64_t)opw * C : (int64_t)a.out_w * C;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int fo = (v * NCW * 32 + t) * 4;
    const int px = fo / C, c = fo % C;
    const int ox = ox0 + px;
    stok[v] = ox < a.out_w;
    const int oxc = min(ox, a.out_w - 1);
    loff[v] = (uint32_t)(((pfirst + px - pbase) * C + c) * 4);
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) {
      const int ip = pfirst + px + tx;
      tokmask |= (ip >= 0 && ip < a.in_w ? 1u : 0u) << (v * 4 + tx);
    }
    if (ACT) {
      d4[v] = __ldg(reinterpret_cast<const float4*>(a.demod + (int64_t)b * C + c));
      bias4[v] = __ldg(reinterpret_cast<const float4*>(a.bias + c));
      nzoff[v] = (uint32_t)(ROWB + px * 4);
    }
    if (OUT_PLANAR) {
      // r0 is even: even rows of the band live in phase plane (0, ox&1), odd rows in (1, ox&1), both from row r0/2 on
      const int64_t plane_sz = (int64_t)oph * opw;
      op0[v] = out + (((int64_t)b * 4 + (oxc & 1)) * plane_sz + (int64_t)(r0 >> 1) * opw + (oxc >> 1)) * C + c;
      op1[v] = out + (((int64_t)b * 4 + 2 + (oxc & 1)) * plane_sz + (int64_t)(r0 >> 1) * opw + (oxc >> 1)) * C + c;
    } else {
      op0[v] = out + (((int64_t)b * a.out_h + r0) * a.out_w + oxc) * C + c;
      op1[v] = op0[v];
    }
  }
  const bool edge = __any_sync(0xffffffffu, tokmask != (NV == 1 ? 0xfu : 0xffu));
  const float nw = ACT ? __ldg(a.noise_w) : 0.f;
  float4 h1[NV], h2[NV], h3[NV];   // horizontal passes of the previous three input rows (h3 oldest)
#pragma unroll
  for (int v = 0; v < NV; ++v) h1[v] = h2[v] = h3[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t slot_addr = sbase, full_bar = bars, empty_bar = bars + D * 8, par = 0;
  int slot = 0;
  for (int j0 = 0; j0 < total; j0 += 2) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + u;
      if (j < total) {
        mbar_wait(full_bar, par);
        const bool rowok = j >= jlo && j < jhi;   // warp-uniform
        float4 h0[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float4 x[4];
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) {
            x[tx] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowok && (!edge || ((tokmask >> (v * 4 + tx)) & 1u))) x[tx] = lds128(slot_addr + loff[v] + tx * C * 4);
          }
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) acc = f4_fma(kx[tx], x[tx], acc);
          h0[v] = acc;
        }
        float nz[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) nz[v] = ACT && j >= 3 && stok[v] ? lds32(slot_addr + nzoff[v]) : 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar);
        slot_addr += SLOTB; full_bar += 8; empty_bar += 8;
        if (++slot == D) { slot = 0; slot_addr = sbase; full_bar = bars; empty_bar = bars + D * 8; par ^= 1u; }
        const int tr = j - 3;    // output row of the band completed by input row j
        if (tr >= 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            acc = f4_fma(ky[0], h3[v], acc);
            acc = f4_fma(ky[1], h2[v], acc);
            acc = f4_fma(ky[2], h1[v], acc);
            acc = f4_fma(ky[3], h0[v], acc);
            if (ACT) {
              const float nzv = nw * nz[v];
              acc.x = lrelu_fwd(fmaf(acc.x, d4[v].x, nzv) + bias4[v].x);
              acc.y = lrelu_fwd(fmaf(acc.y, d4[v].y, nzv) + bias4[v].y);
              acc.z = lrelu_fwd(fmaf(acc.z, d4[v].z, nzv) + bias4[v].z);
              acc.w = lrelu_fwd(fmaf(acc.w, d4[v].w, nzv) + bias4[v].w);
            }
            // j = j0 + u with j0 even: the output row tr = j - 3 is odd for u == 0 and even for u == 1
            if (OUT_PLANAR && u == 0) {
              if (stok[v]) *reinterpret_cast<float4*>(op1[v]) = acc;
              op1[v] += ostep;
            } else {
              if (stok[v]) *reinterpret_cast<float4*>(op0[v]) = acc;
              op0[v] += ostep;
            }
          }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) { h3[v] = h2[v]; h2[v] = h1[v]; h1[v] = h0[v]; }
      }
    }
  }
}

template <int C, int NV, bool ACT, bool OUT_PLANAR>
static int launch_fir_ring(const float* in, float* out, const FirArgs& a, cudaStream_t s) {
  using namespace nring;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    LFP_CUDA(cudaFuncSetAttribute(fir_ring_nhwc_kernel<C, NV, ACT, OUT_PLANAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(C, NV)));
    attr_done[dev] = true;
  }
  // output rows per CTA (even: the phase-major output alternates planes by row parity): 128 costs 2 % of halo rows; halved
  // while the grid is shorter than ~6 waves of the 2 CTAs an SM holds
  int rb = a.out_h >= 512 ? 128 : 64;
  while (rb > 32 && ceil_div(a.out_w, segp(C, NV)) * ceil_div(a.out_h, rb) * a.batch < (int64_t)num_sms() * 12) rb >>= 1;
  dim3 grid((unsigned)ceil_div(a.out_w, segp(C, NV)), (unsigned)ceil_div(a.out_h, rb), (unsigned)a.batch);
  fir_ring_nhwc_kernel<C, NV, ACT, OUT_PLANAR><<<grid, (NCW + 1) * 32, smem_bytes(C, NV), s>>>(in, out, a, rb);
  LFP_LAUNCH_CHECK();
  return 0;
}

int launch_fir4x4_nhwc(const float* in, float* out, const FirArgs& a, cudaStream_t s) {
  static const bool use_ring = !(getenv("LFP_FIR_RING") && atoi(getenv("LFP_FIR_RING")) == 0);
  const bool noise_ok = !a.act || ((a.out_w & 3) == 0 && (a.noise_bstride & 3) == 0 && (reinterpret_cast<uintptr_t>(a.noise) & 15) == 0);
  if (use_ring && a.kx != nullptr && a.ky != nullptr && !a.in_planar && a.out_w >= 64 && (a.act ? !a.out_planar : true) && noise_ok &&
      (a.C == 32 || a.C == 64 || a.C == 128 || a.C == 256) && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
#define LFP_RING_CASE(CC, NVV)                                                                      \
    if (a.C == CC) {                                                                                 \
      if (a.act) return launch_fir_ring<CC, NVV, true, false>(in, out, a, s);                        \
      if (a.out_planar) return launch_fir_ring<CC, NVV, false, true>(in, out, a, s);                 \
      return launch_fir_ring<CC, NVV, false, false>(in, out, a, s);                                  \
    }
    LFP_RING_CASE(32, 1)
    LFP_RING_CASE(64, 1)
    LFP_RING_CASE(128, 1)
    LFP_RING_CASE(256, 1)
#undef LFP_RING_CASE
  }
  if (a.kx != nullptr && a.ky != nullptr && a.out_w >= 64 && a.C % 4 == 0 && a.C <= 1024) {
    const int C4 = a.C / 4;
    const int PX = 256 / C4 > 0 ? 256 / C4 : 1;
    const int RY = a.out_h >= 256 ? 32 : 16;
    dim3 grid((unsigned)ceil_div(a.out_w, PX * 4), (unsigned)ceil_div(a.out_h, RY), (unsigned)a.batch);
    if (a.act) fir4x4_sep_nhwc_kernel<true, 4><<<grid, 256, 0, s>>>(in, out, a, C4, PX, RY);
    else fir4x4_sep_nhwc_kernel<false, 4><<<grid, 256, 0, s>>>(in, out, a, C4, PX, RY);
    LFP_LAUNCH_CHECK();
    return 0;
  }
  LFP_CHECK_ARG(a.C % 4 == 0 && a.C <= 1024, "fir: C=%d must be a multiple of 4 and <= 1024", a.C);
  const int C4 = a.C / 4;
  const int PX = 256 / C4 > 0 ? 256 / C4 : 1;
  const bool wide = a.out_w >= 32;   // two output columns per thread once rows are long enough to fill the grid
  const int XO = wide ? 2 : 1;
  const int RY = a.out_h >= 256 ? 16 : 8;
  dim3 grid((unsigned)ceil_div(a.out_w, PX * XO), (unsigned)ceil_div(a.out_h, RY), (unsigned)a.batch);
  if (a.act) {
    if (wide) fir4x4_nhwc_kernel<true, 2><<<grid, 256, 0, s>>>(in, out, a, C4, PX, RY);
    else fir4x4_nhwc_kernel<true, 1><<<grid, 256, 0, s>>>(in, out, a, C4, PX, RY);
  } else {
    if (wide) fir4x4_nhwc_kernel<false, 2><<<grid, 256, 0, s>>>(in, out, a, C4, PX, RY);
    else fir4x4_nhwc_kernel<false, 1><<<grid, 256, 0, s>>>(in, out, a, C4, PX, RY);
  }
  LFP_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// ToRGB forward (+ bias + FIR-upsampled skip)
// =============================================================================================
__device__ __forceinline__ float up2_sample(const float* __restrict__ sp, int h2, int w2,
                                            const float* __restrict__ kup, int oy, int ox) {
  // Upsample: upfirdn2d(up=2, pad=(2,1)) (src/model.py:33-51); flipped tap (ty,tx) = kup[3-ty][3-tx].  Only the taps whose
  // zero-stuffed coordinate oy + ty - 2 is even meet a sample: ty = (oy & 1) + {0, 2}, likewise tx; ascending tap order.
  float acc = 0.f;
  const int py = oy & 1, px = ox & 1;
#pragma unroll
  for (int jy = 0; jy < 2; ++jy) {
    const int ty = py + 2 * jy;
    const int iy = (oy + ty - 2) >> 1;
    if (iy < 0 || iy >= h2) continue;
#pragma unroll
    for (int jx = 0; jx < 2; ++jx) {
      const int tx = px + 2 * jx;
      const int ix = (ox + tx - 2) >> 1;
      if (ix < 0 || ix >= w2) continue;
      acc = fmaf(__ldg(sp + (int64_t)iy * w2 + ix), __ldg(kup + (3 - ty) * 4 + (3 - tx)), acc);
    }
  }
  return acc;
}

template <int LPP, int NCH>
__global__ void __launch_bounds__(256) torgb_fwd_kernel(const float* __restrict__ act,
                                                        const float* __restrict__ s,
                                                        const float* __restrict__ wrgb,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ skip,
                                                        const float* __restrict__ kup,
                                                        float* __restrict__ rgb, int h, int w, int C) {
  constexpr int PPW = 32 / LPP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int hw = h * w;
  const int pix0 = (blockIdx.x * 8 + warp) * 32;
  if (pix0 >= hw) return;
  const int sub = lane / LPP, li = lane % LPP;
  float4 m[3][NCH];
#pragma unroll
  for (int q = 0; q < NCH; ++q) {
    const int c = (q * LPP + li) * 4;
    const float4 sv = __ldg(reinterpret_cast<const float4*>(s + (int64_t)b * C + c));
#pragma unroll
    for (int o = 0; o < 3; ++o) m[o][q] = f4_mul(sv, __ldg(reinterpret_cast<const float4*>(wrgb + o * C + c)));
  }
  float mine[3] = {0.f, 0.f, 0.f};
  // C = 4 * LPP * NCH is fixed by the instantiation: one 64-bit pointer per lane, every other address a compile-time
  // offset from it (the kernel is bound by instruction issue, not by bytes)
  constexpr int CT = 4 * LPP * NCH;
  const float* lane_base = act + ((int64_t)b * hw + pix0 + sub) * CT + li * 4;
  const bool whole = pix0 + 32 <= hw;   // warp-uniform: all 32 pixels of the group exist
  // groups of up to 8 pixel-iterations with all activation loads issued before the first use
  constexpr int GRP = LPP * NCH <= 8 ? LPP : (8 / NCH >= 1 ? 8 / NCH : 1);   // <= 8 float4 in flight per lane
#pragma unroll 1
  for (int it0 = 0; it0 < LPP; it0 += GRP) {
    float4 v[GRP][NCH];
    const float* gp = lane_base + (int64_t)it0 * PPW * CT;
    if (whole) {
#pragma unroll
      for (int g = 0; g < GRP; ++g)
#pragma unroll
        for (int q = 0; q < NCH; ++q) v[g][q] = __ldg(reinterpret_cast<const float4*>(gp + g * PPW * CT + q * LPP * 4));
    } else {
#pragma unroll
      for (int g = 0; g < GRP; ++g) {
        const int pix = pix0 + (it0 + g) * PPW + sub;
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
          v[g][q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pix < hw) v[g][q] = __ldg(reinterpret_cast<const float4*>(gp + g * PPW * CT + q * LPP * 4));
        }
      }
    }
    float r[GRP][3];
#pragma unroll
    for (int g = 0; g < GRP; ++g) {
#pragma unroll
      for (int o = 0; o < 3; ++o) r[g][o] = 0.f;
#pragma unroll
      for (int q = 0; q < NCH; ++q)
#pragma unroll
        for (int o = 0; o < 3; ++o) r[g][o] += f4_dot(v[g][q], m[o][q]);
    }
    // Reduce over the LPP lanes of a pixel by recursive halving: at each step a lane keeps half of the passes it still holds
    // and hands the other half to its partner, so GRP passes cost 3*(GRP-1) shuffles instead of 3*GRP*log2(LPP).  Afterwards
    // the lanes whose low bits are j hold pass it0 + j; the fixed exchange pattern keeps the sums independent of the batch.
#pragma unroll
    for (int bit = GRP / 2; bit >= 1; bit >>= 1) {
      const bool upper = (li & bit) != 0;
#pragma unroll
      for (int g = 0; g < bit; ++g)
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const float send = upper ? r[g][o] : r[g + bit][o];
          const float keep = upper ? r[g + bit][o] : r[g][o];
          r[g][o] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
#pragma unroll
    for (int off = GRP; off < LPP; off <<= 1)
#pragma unroll
      for (int o = 0; o < 3; ++o) r[0][o] += __shfl_xor_sync(0xffffffffu, r[0][o], off);
    // lane li keeps pass `li`: it is computed in the batch with it0 == (li rounded down to GRP)
    if ((li & ~(GRP - 1)) == it0) {
#pragma unroll
      for (int o = 0; o < 3; ++o) mine[o] = r[0][o];
    }
  }
  // lane (sub, li) owns the pixel of pass li: pix0 + li * PPW + sub
  const int pix = pix0 + li * PPW + sub;
  if (pix < hw) {
    const int y = pix / w, x = pix - y * w;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      float v = mine[o] + __ldg(bias + o);
      if (skip != nullptr) v += up2_sample(skip + ((int64_t)b * 3 + o) * (h / 2) * (w / 2), h / 2, w / 2, kup, y, x);
      rgb[((int64_t)b * 3 + o) * hw + pix] = v;
    }
  }
}

int launch_torgb_fwd(const float* act, const float* s, const float* wrgb, const float* bias,
                     const float* skip, const float* kup, float* rgb, int batch, int h, int w,
                     int C, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div((int64_t)h * w, 256), (unsigned)batch);
#define LFP_TORGB(LPP, NCH) torgb_fwd_kernel<LPP, NCH><<<grid, 256, 0, st>>>(act, s, wrgb, bias, skip, kup, rgb, h, w, C)
  switch (C) {
    case 16: LFP_TORGB(4, 1); break;
    case 32: LFP_TORGB(8, 1); break;
    case 64: LFP_TORGB(16, 1); break;
    case 128: LFP_TORGB(32, 1); break;
    case 256: LFP_TORGB(32, 2); break;
    case 512: LFP_TORGB(32, 4); break;
    default: set_error("torgb: unsupported channel count %d", C); return LFP_EUNSUPPORTED;
  }
#undef LFP_TORGB
  LFP_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) skip_add_kernel(float* __restrict__ rgb, const float* __restrict__ skip,
                                                       const float* __restrict__ kup, int h, int w) {
  const int64_t hw = (int64_t)h * w;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
  const int plane = blockIdx.y;   // b*3 + o
  const int y = (int)(i / w), x = (int)(i - (int64_t)y * w);
  rgb[plane * hw + i] += up2_sample(skip + (int64_t)plane * (h / 2) * (w / 2), h / 2, w / 2, kup, y, x);
}

int launch_skip_add(float* rgb, const float* skip, const float* kup, int batch, int h, int w, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div((int64_t)h * w, 256), (unsigned)(batch * 3));
  skip_add_kernel<<<grid, 256, 0, st>>>(rgb, skip, kup, h, w);
  LFP_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Backward through noise / bias / lrelu (+ ToRGB branch), with the two style-gradient reductions
// =============================================================================================
__global__ void __launch_bounds__(256) act_bwd_kernel(const ActBwdArgs a, int C4, int PX, int seglen) {
  __shared__ float redT[1024];
  __shared__ float redR[1024];
  const int tid = threadIdx.x;
  const int c4 = tid % C4, py = tid / C4;
  const bool active = py < PX;
  const int segs_per = a.hw / seglen;
  const int seg = blockIdx.x;
  const int b = seg / segs_per;
  const int pix0 = (seg - b * segs_per) * seglen;
  const int C = a.C;
  const bool has_rgb = a.drgb != nullptr;
  float4 T4 = make_float4(0.f, 0.f, 0.f, 0.f), R4 = T4;
  if (active) {
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.demod + (int64_t)b * C + c4 * 4));
    const float4 bias4 = __ldg(reinterpret_cast<const float4*>(a.bias + c4 * 4));
    const float nw = __ldg(a.noise_w);
    float4 s4 = T4, w0 = T4, w1 = T4, w2 = T4;
    if (has_rgb) {
      s4 = __ldg(reinterpret_cast<const float4*>(a.s_rgb + (int64_t)b * C + c4 * 4));
      w0 = __ldg(reinterpret_cast<const float4*>(a.wrgb + 0 * C + c4 * 4));
      w1 = __ldg(reinterpret_cast<const float4*>(a.wrgb + 1 * C + c4 * 4));
      w2 = __ldg(reinterpret_cast<const float4*>(a.wrgb + 2 * C + c4 * 4));
    }
    const float inv_pos = 1.f / kLreluGain, inv_neg = 1.f / (kLreluGain * kLreluSlope);
    // four pixels per trip with all loads issued first (memory-level parallelism); the accumulation order is the
    // plain i order, so results do not depend on the unrolling
    constexpr int U = 4;
    for (int i0 = py; i0 < seglen; i0 += U * PX) {
      float4 act_v[U], g_v[U];
      float r0v[U], r1v[U], r2v[U], nzv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * PX;
        const bool in = i < seglen;
        const int pix = pix0 + (in ? i : py);
        const int64_t idx = ((int64_t)b * a.hw + pix) * C + c4 * 4;
        act_v[u] = __ldg(reinterpret_cast<const float4*>(a.act + idx));
        g_v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.g_has_input) g_v[u] = *reinterpret_cast<const float4*>(a.g + idx);
        r0v[u] = r1v[u] = r2v[u] = 0.f;
        if (has_rgb) {
          r0v[u] = __ldg(a.drgb + ((int64_t)b * 3 + 0) * a.hw + pix);
          r1v[u] = __ldg(a.drgb + ((int64_t)b * 3 + 1) * a.hw + pix);
          r2v[u] = __ldg(a.drgb + ((int64_t)b * 3 + 2) * a.hw + pix);
        }
        nzv[u] = nw * __ldg(a.noise + (int64_t)b * a.noise_bstride + pix);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * PX;
        if (i >= seglen) break;
        const int pix = pix0 + i;
        const int64_t idx = ((int64_t)b * a.hw + pix) * C + c4 * 4;
        const float4 act4 = act_v[u];
        float4 g4 = g_v[u];
        if (has_rgb) {
          const float r0 = r0v[u], r1 = r1v[u], r2 = r2v[u];
          float4 q4 = make_float4(r0 * w0.x, r0 * w0.y, r0 * w0.z, r0 * w0.w);
          q4 = f4_fma(r1, w1, q4);
          q4 = f4_fma(r2, w2, q4);
          g4 = f4_fma4(q4, s4, g4);
          R4 = f4_fma4(act4, q4, R4);
        }
        const float nz = nzv[u];
        float4 o4;
#define LFP_ACTB(comp)                                                          \
  {                                                                             \
    const bool pos = act4.comp > 0.f;                                           \
    const float gpre = g4.comp * (pos ? kLreluGain : kLreluGain * kLreluSlope); \
    const float pre = act4.comp * (pos ? inv_pos : inv_neg);                    \
    T4.comp = fmaf(gpre, pre - nz - bias4.comp, T4.comp);                       \
    o4.comp = gpre * d4.comp;                                                   \
  }
        LFP_ACTB(x) LFP_ACTB(y) LFP_ACTB(z) LFP_ACTB(w)
#undef LFP_ACTB
        *reinterpret_cast<float4*>(a.g + idx) = o4;
      }
    }
    *reinterpret_cast<float4*>(&redT[py * C + c4 * 4]) = T4;
    *reinterpret_cast<float4*>(&redR[py * C + c4 * 4]) = R4;
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float t = 0.f, r = 0.f;
    for (int q = 0; q < PX; ++q) { t += redT[q * C + c]; r += redR[q * C + c]; }
    a.pT[(int64_t)seg * C + c] = t;
    if (has_rgb) a.pR[(int64_t)seg * C + c] = r;
  }
}

// ---- row-ring variant for the large maps (C <= 128): same arithmetic, inputs streamed through shared memory ------------
// The kernel above is a one-shot CTA per 256 pixels whose loads in flight live in registers; at 117 registers two CTAs fit
// an SM, which keeps ~35 KB in flight and the kernel at 4 TB/s.  Here a producer thread streams 2048-float chunks of the
// activation (and of g, the per-pixel ToRGB gradients and the noise) into a shared-memory ring with bulk copies, and 256
// consumer threads walk a 2048-pixel segment chunk by chunk.  Each thread still sums its pixels in ascending order.
namespace aring {
constexpr int CHF = 2048;            // floats of activation per chunk
constexpr int SEG = 2048;            // pixels per CTA
constexpr int AUXB = 4 * 512;        // three ToRGB-gradient planes + noise, <= 128 floats each
__host__ __device__ constexpr int slotb(bool gin) { return CHF * 4 * (gin ? 2 : 1) + AUXB; }
constexpr int D = 5;
__host__ __device__ constexpr int smem_bytes(bool gin) { return D * slotb(gin) + 2 * D * 8 + 16 + 2 * 1024 * 4; }
}  // namespace aring

template <int C, bool GIN, bool RGB>
__global__ void __launch_bounds__(288) act_bwd_ring_kernel(const ActBwdArgs a) {
  using namespace nring;
  constexpr int CHF = aring::CHF, CHP = CHF / C, SLOTB = aring::slotb(GIN), D = aring::D, NCHUNK = aring::SEG / CHP;
  constexpr int C4 = C / 4, PXT = 256 / C4;
  extern __shared__ __align__(128) uint8_t aring_smem[];
  const uint32_t sbase = smem_u32(aring_smem);
  const uint32_t bars = sbase + D * SLOTB;
  float* redT = reinterpret_cast<float*>(aring_smem + D * SLOTB + 2 * D * 8 + 16);
  float* redR = redT + 1024;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int segs_per = a.hw / aring::SEG;
  const int seg = blockIdx.x;
  const int b = seg / segs_per;
  const int pix0 = (seg - b * segs_per) * aring::SEG;
  if (threadIdx.x == 0) {
    for (int i = 0; i < D; ++i) { mbar_init(bars + i * 8, 1); mbar_init(bars + (D + i) * 8, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    if (lane == 0) {
      const float* asrc = a.act + ((int64_t)b * a.hw + pix0) * C;
      const float* gsrc = a.g + ((int64_t)b * a.hw + pix0) * C;
      const float* nsrc = a.noise + (int64_t)b * a.noise_bstride + pix0;
      const uint32_t bytes = CHF * 4 * (GIN ? 2 : 1) + CHP * 4 * (RGB ? 4 : 1);
      for (int k = 0; k < NCHUNK; ++k) {
        const int slot = k % D;
        const uint32_t sa = sbase + slot * SLOTB, fb = bars + slot * 8;
        if (k >= D) mbar_wait(bars + (D + slot) * 8, ((k / D) - 1) & 1);
        mbar_expect_tx(fb, bytes);
        bulk_g2s(sa, asrc + (int64_t)k * CHF, CHF * 4, fb);
        if (GIN) bulk_g2s(sa + CHF * 4, gsrc + (int64_t)k * CHF, CHF * 4, fb);
        const uint32_t aux = sa + CHF * 4 * (GIN ? 2 : 1);
        bulk_g2s(aux, nsrc + k * CHP, CHP * 4, fb);
        if (RGB) {
#pragma unroll
          for (int o = 0; o < 3; ++o)
            bulk_g2s(aux + (1 + o) * 512, a.drgb + ((int64_t)b * 3 + o) * a.hw + pix0 + k * CHP, CHP * 4, fb);
        }
      }
    }
  } else {
    const int t = threadIdx.x;
    const int c4 = t % C4, py = t / C4;
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.demod + (int64_t)b * C + c4 * 4));
    const float4 bias4 = __ldg(reinterpret_cast<const float4*>(a.bias + c4 * 4));
    const float nw = __ldg(a.noise_w);
    float4 T4 = make_float4(0.f, 0.f, 0.f, 0.f), R4 = T4;
    float4 s4 = T4, w0 = T4, w1 = T4, w2 = T4;
    if (RGB) {
      s4 = __ldg(reinterpret_cast<const float4*>(a.s_rgb + (int64_t)b * C + c4 * 4));
      w0 = __ldg(reinterpret_cast<const float4*>(a.wrgb + 0 * C + c4 * 4));
      w1 = __ldg(reinterpret_cast<const float4*>(a.wrgb + 1 * C + c4 * 4));
      w2 = __ldg(reinterpret_cast<const float4*>(a.wrgb + 2 * C + c4 * 4));
    }
    const float inv_pos = 1.f / kLreluGain, inv_neg = 1.f / (kLreluGain * kLreluSlope);
    float* gout = a.g + ((int64_t)b * a.hw + pix0) * C;
    constexpr int NV = CHF / 1024;   // float4 per thread and chunk
    for (int k = 0; k < NCHUNK; ++k) {
      const int slot = k % D;
      const uint32_t sa = sbase + slot * SLOTB;
      mbar_wait(bars + slot * 8, (k / D) & 1);
      float4 act_v[NV], g_v[NV];
      float r0v[NV], r1v[NV], r2v[NV], nzv[NV];
      const uint32_t aux = sa + CHF * 4 * (GIN ? 2 : 1);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * 256 + t) * 4, px = f / C;
        act_v[v] = lds128(sa + f * 4);
        g_v[v] = GIN ? lds128(sa + CHF * 4 + f * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        nzv[v] = nw * lds32(aux + px * 4);
        r0v[v] = r1v[v] = r2v[v] = 0.f;
        if (RGB) { r0v[v] = lds32(aux + 512 + px * 4); r1v[v] = lds32(aux + 1024 + px * 4); r2v[v] = lds32(aux + 1536 + px * 4); }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + (D + slot) * 8);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * 256 + t) * 4;
        const float4 act4 = act_v[v];
        float4 g4 = g_v[v];
        if (RGB) {
          const float r0 = r0v[v], r1 = r1v[v], r2 = r2v[v];
          float4 q4 = make_float4(r0 * w0.x, r0 * w0.y, r0 * w0.z, r0 * w0.w);
          q4 = f4_fma(r1, w1, q4);
          q4 = f4_fma(r2, w2, q4);
          g4 = f4_fma4(q4, s4, g4);
          R4 = f4_fma4(act4, q4, R4);
        }
        const float nz = nzv[v];
        float4 o4;
#define LFP_ACTB(comp)                                                          \
  {                                                                             \
    const bool pos = act4.comp > 0.f;                                           \
    const float gpre = g4.comp * (pos ? kLreluGain : kLreluGain * kLreluSlope); \
    const float pre = act4.comp * (pos ? inv_pos : inv_neg);                    \
    T4.comp = fmaf(gpre, pre - nz - bias4.comp, T4.comp);                       \
    o4.comp = gpre * d4.comp;                                                   \
  }
        LFP_ACTB(x) LFP_ACTB(y) LFP_ACTB(z) LFP_ACTB(w)
#undef LFP_ACTB
        *reinterpret_cast<float4*>(gout + (int64_t)k * CHF + f) = o4;
      }
    }
    *reinterpret_cast<float4*>(&redT[py * C + c4 * 4]) = T4;
    *reinterpret_cast<float4*>(&redR[py * C + c4 * 4]) = R4;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 288) {
    float tt = 0.f, r = 0.f;
    for (int q = 0; q < PXT; ++q) { tt += redT[q * C + c]; r += redR[q * C + c]; }
    a.pT[(int64_t)seg * C + c] = tt;
    if (RGB) a.pR[(int64_t)seg * C + c] = r;
  }
}

static bool actbwd_ring_ok(int hw, int C) {
  static const bool use_ring = !(getenv("LFP_ACTBWD_RING") && atoi(getenv("LFP_ACTBWD_RING")) == 0);
  return use_ring && (C == 32 || C == 64 || C == 128) && hw >= 65536 && hw % aring::SEG == 0;
}

template <int C, bool GIN, bool RGB>
static int launch_act_bwd_ring(const ActBwdArgs& a, cudaStream_t s) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    LFP_CUDA(cudaFuncSetAttribute(act_bwd_ring_kernel<C, GIN, RGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, aring::smem_bytes(GIN)));
    attr_done[dev] = true;
  }
  const int64_t nseg = (int64_t)a.batch * (a.hw / aring::SEG);
  act_bwd_ring_kernel<C, GIN, RGB><<<(unsigned)nseg, 288, aring::smem_bytes(GIN), s>>>(a);
  LFP_LAUNCH_CHECK();
  return 0;
}

int actbwd_seglen(int hw, int C) {
  if (actbwd_ring_ok(hw, C)) return aring::SEG;
  const int PX = 1024 / C;  // pixels per CTA pass
  int seg = PX * 8;
  if (seg < 256) seg = 256;
  if (seg > hw) seg = hw;
  return seg;
}

int launch_act_bwd(const ActBwdArgs& a, cudaStream_t s) {
  LFP_CHECK_ARG(a.C % 4 == 0 && a.C <= 1024 && 1024 % a.C == 0, "act_bwd: unsupported C=%d", a.C);
  const bool aligned = (reinterpret_cast<uintptr_t>(a.act) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.g) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a.noise) & 15) == 0 && (a.noise_bstride & 3) == 0 &&
                       (a.drgb == nullptr || (reinterpret_cast<uintptr_t>(a.drgb) & 15) == 0);
  if (actbwd_ring_ok(a.hw, a.C)) {
    LFP_CHECK_ARG(aligned, "act_bwd: the ring kernel needs 16-byte aligned tensors");
#define LFP_AR(CC)                                                                                         \
    if (a.C == CC) {                                                                                        \
      if (a.g_has_input) return a.drgb ? launch_act_bwd_ring<CC, true, true>(a, s) : launch_act_bwd_ring<CC, true, false>(a, s); \
      return a.drgb ? launch_act_bwd_ring<CC, false, true>(a, s) : launch_act_bwd_ring<CC, false, false>(a, s);                \
    }
    LFP_AR(32) LFP_AR(64) LFP_AR(128)
#undef LFP_AR
  }
  const int C4 = a.C / 4, PX = 256 / C4;
  const int seglen = actbwd_seglen(a.hw, a.C);
  LFP_CHECK_ARG(a.hw % seglen == 0, "act_bwd: hw=%d not divisible by segment %d", a.hw, seglen);
  const int64_t nseg = (int64_t)a.batch * (a.hw / seglen);
  act_bwd_kernel<<<(unsigned)nseg, 256, 0, s>>>(a, C4, PX, seglen);
  LFP_LAUNCH_CHECK();
  return 0;
}

// out[b, c] = sum_q partial[b*Q + q, c], fixed order: NR contiguous q-ranges (four independent accumulators each, combined
// in a fixed tree), then the NR range sums in ascending order.  NR = 32 for the long reductions of the high-resolution
// layers (Q up to 8192 tiles per sample), 8 otherwise.
template <int NR>
__global__ void __launch_bounds__(NR * 32) partial_reduce_kernel(const float* __restrict__ partial,
                                                                 float* __restrict__ out, int Q, int C,
                                                                 int64_t out_bstride) {
  __shared__ float sm[NR][32];
  const int cx = threadIdx.x & 31, qy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int b = blockIdx.y;
  const int per = (Q + NR - 1) / NR;
  const int q0 = qy * per, q1 = min(Q, q0 + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < C) {
    const float* p = partial + ((int64_t)b * Q + q0) * C + c;
    int q = q0;
    for (; q + 4 <= q1; q += 4, p += 4 * (int64_t)C) {
      a0 += __ldg(p); a1 += __ldg(p + C); a2 += __ldg(p + 2 * (int64_t)C); a3 += __ldg(p + 3 * (int64_t)C);
    }
    for (; q < q1; ++q, p += C) a0 += __ldg(p);
  }
  sm[qy][cx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (qy == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < NR; ++k) t += sm[k][cx];
    out[(int64_t)b * out_bstride + c] = t;
  }
}

int launch_partial_reduce(const float* partial, float* out, int batch, int Q, int C, int64_t out_bstride,
                          cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)batch);
  if (Q >= 256) partial_reduce_kernel<32><<<grid, 1024, 0, s>>>(partial, out, Q, C, out_bstride);
  else partial_reduce_kernel<8><<<grid, 256, 0, s>>>(partial, out, Q, C, out_bstride);
  LFP_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Style affine (all modulation linears of one forward in one launch), demodulation, style grads
// =============================================================================================
__global__ void __launch_bounds__(256) style_affine_kernel(const float* __restrict__ latent,
                                                           const float* __restrict__ A,
                                                           const float* __restrict__ bias,
                                                           const int* __restrict__ row_slot,
                                                           const int* __restrict__ row_base,
                                                           const int* __restrict__ row_cin,
                                                           float* __restrict__ s, int batch, int rows,
                                                           int n_latent, int dim) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int slot = row_slot[r];
  const float* arow = A + (int64_t)r * dim;
  const float bv = __ldg(bias + r);
  const int base = row_base[r], cin = row_cin[r];
  for (int b = 0; b < batch; ++b) {
    const float* lrow = latent + ((int64_t)b * n_latent + slot) * dim;
    float acc = 0.f;
    for (int j = lane * 4; j < dim; j += 128)
      acc += f4_dot(__ldg(reinterpret_cast<const float4*>(arow + j)), __ldg(reinterpret_cast<const float4*>(lrow + j)));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) s[(int64_t)batch * base + (int64_t)b * cin + (r - base)] = acc + bv;
  }
}

int launch_style_affine(const float* latent, const float* A, const float* bias, const int* row_slot,
                        const int* row_base, const int* row_cin, float* s, int batch, int rows,
                        int n_latent, int dim, cudaStream_t st) {
  LFP_CHECK_ARG(dim % 4 == 0, "style_affine: dim must be a multiple of 4");
  style_affine_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, st>>>(latent, A, bias, row_slot, row_base, row_cin, s, batch, rows, n_latent, dim);
  LFP_LAUNCH_CHECK();
  return 0;
}

// d_latent[b, slot, j] = sum over the rows r of the slot of ds[b, r] * A[r, j], eight samples per CTA: a row of A is read
// once for the eight.  The slot's ds values are staged in shared memory 256 rows at a time (coalesced), then warp w walks
// rows w + 8u (u < 8) + 64i of the chunk with eight independent loads of A in flight (the first version issued one
// dependent L2 round trip per row and took 210 us per step; this one is bandwidth-shaped).  Fixed summation order per
// sample: two accumulators per sample (even / odd u), warps combined in ascending order.
__global__ void __launch_bounds__(256) style_affine_bwd8_kernel(const float* __restrict__ ds, const float* __restrict__ A,
                                                                const int* __restrict__ row_begin, const int* __restrict__ row_end,
                                                                const int* __restrict__ row_base, const int* __restrict__ row_cin,
                                                                float* __restrict__ d_latent, int batch, int n_latent, int dim) {
  __shared__ float ds_s[8][256];
  __shared__ float part[8][8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int slot = blockIdx.y, b0 = blockIdx.z * 8;
  const int nb = min(8, batch - b0);
  const int r0 = row_begin[slot], r1 = row_end[slot];
  float acc[8][2];
#pragma unroll
  for (int bb = 0; bb < 8; ++bb) acc[bb][0] = acc[bb][1] = 0.f;
  for (int c0 = r0; c0 < r1; c0 += 256) {
    const int nrow = min(256, r1 - c0);
    __syncthreads();
    {   // thread t stages row c0 + t of the eight samples
      const int r = c0 + threadIdx.x;
      if (threadIdx.x < nrow) {
        const int base = row_base[r], cin = row_cin[r];
        const float* dsr = ds + (int64_t)batch * base + (int64_t)b0 * cin + (r - base);
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) ds_s[bb][threadIdx.x] = bb < nb ? __ldg(dsr + (int64_t)bb * cin) : 0.f;
      }
    }
    __syncthreads();
    if (j < dim)
      for (int rr = w; rr < nrow; rr += 64) {
        float av[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) av[u] = rr + 8 * u < nrow ? __ldg(A + (int64_t)(c0 + rr + 8 * u) * dim + j) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = rr + 8 * u < nrow ? rr + 8 * u : 0;
#pragma unroll
          for (int bb = 0; bb < 8; ++bb) acc[bb][u & 1] = fmaf(ds_s[bb][r], av[u], acc[bb][u & 1]);
        }
      }
  }
#pragma unroll
  for (int bb = 0; bb < 8; ++bb) part[bb][w][lane] = acc[bb][0] + acc[bb][1];
  __syncthreads();
  if (w < nb && j < dim) {   // warp w finishes sample b0 + w
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[w][k][lane];
    d_latent[((int64_t)(b0 + w) * n_latent + slot) * dim + j] = t;
  }
}

int launch_style_affine_bwd(const float* ds, const float* A, const int* slot_row_begin,
                            const int* slot_row_end, const int* row_base, const int* row_cin,
                            float* d_latent, int batch, int rows, int n_latent, int dim,
                            cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(dim, 32), (unsigned)n_latent, (unsigned)ceil_div(batch, 8));
  style_affine_bwd8_kernel<<<grid, 256, 0, st>>>(ds, A, slot_row_begin, slot_row_end, row_base, row_cin, d_latent, batch, n_latent, dim);
  LFP_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ s, int64_t s_bstride,
                                                    const float* __restrict__ wsq,
                                                    float* __restrict__ d, int64_t d_bstride,
                                                    int cin, int cout) {
  const int lane = threadIdx.x & 31;
  const int co = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (co >= cout) return;
  float acc = 0.f;
  for (int ci = lane; ci < cin; ci += 32) {
    const float sv = __ldg(s + (int64_t)b * s_bstride + ci);
    acc = fmaf(sv * sv, __ldg(wsq + (int64_t)co * cin + ci), acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) d[(int64_t)b * d_bstride + co] = rsqrtf(acc + 1e-8f);
}

int launch_demod(const float* s, int64_t s_bstride, const float* wsq, float* d, int64_t d_bstride,
                 int batch, int cin, int cout, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(cout, 8), (unsigned)batch);
  demod_kernel<<<grid, 256, 0, st>>>(s, s_bstride, wsq, d, d_bstride, cin, cout);
  LFP_LAUNCH_CHECK();
  return 0;
}

// ds[b,ci] = r1[b,ci] - s[b,ci] * sum_co (T[b,co] d[b,co]^2) wsq[co,ci]: one CTA per (32 input channels, sample);
// warp w walks co = w, w+8, ... (coalesced 128-byte rows of wsq), the eight partial sums are added in warp order.
__global__ void __launch_bounds__(256) style_grad_kernel(const float* __restrict__ r1,
                                                         const float* __restrict__ s, int64_t s_bstride,
                                                         const float* __restrict__ T,
                                                         const float* __restrict__ d, int64_t d_bstride,
                                                         const float* __restrict__ wsq,
                                                         float* __restrict__ ds, int cin, int cout) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ci = blockIdx.x * 32 + lane;
  const int b = blockIdx.y;
  float acc = 0.f;
  if (ci < cin)
    for (int co = w; co < cout; co += 8) {
      const float dv = __ldg(d + (int64_t)b * d_bstride + co);
      acc = fmaf(__ldg(T + (int64_t)b * d_bstride + co) * dv * dv, __ldg(wsq + (int64_t)co * cin + ci), acc);
    }
  part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && ci < cin) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    const int64_t i = (int64_t)b * s_bstride + ci;
    ds[i] = r1[i] - __ldg(s + i) * t;
  }
}

int launch_style_grad(const float* r1, const float* s, int64_t s_bstride, const float* T,
                      const float* d, int64_t d_bstride, const float* wsq, float* ds, int batch,
                      int cin, int cout, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(cin, 32), (unsigned)batch);
  style_grad_kernel<<<grid, 256, 0, st>>>(r1, s, s_bstride, T, d, d_bstride, wsq, ds, cin, cout);
  LFP_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Batched forms: every reduction / style-gradient / demodulation of a pass in ONE launch each.
// Items live in device memory and address the workspace by float offsets, so a table depends on the
// plan, the batch and the kernel choices only (built once per configuration by the plan).
// The per-(sample, channel) summation orders are exactly those of the single-layer kernels above.
// =============================================================================================
__global__ void __launch_bounds__(1024) batched_partial_reduce_kernel(const ReduceItem* __restrict__ items,
                                                                      const int2* __restrict__ blocks,
                                                                      float* __restrict__ ws) {
  __shared__ float sm[32][32];
  const int2 blk = blocks[blockIdx.x];
  const ReduceItem it = items[blk.x];
  const int NR = it.Q >= 256 ? 32 : 8;
  const int cx = threadIdx.x & 31, qy = threadIdx.x >> 5;
  const int c = blk.y * 32 + cx;
  const int b = blockIdx.y;
  const int Q = it.Q, C = it.C;
  const int per = (Q + NR - 1) / NR;
  const int q0 = qy * per, q1 = min(Q, q0 + per);
  // eight independent accumulators: the loads of a range are L2 round trips, and four in flight per thread left the
  // launch latency-bound (210 us per step for ~100 MB of partials)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
  if (c < C && qy < NR) {
    const float* p = ws + it.src_off + ((int64_t)b * Q + q0) * C + c;
    const int64_t Cs = C;
    int q = q0;
    for (; q + 8 <= q1; q += 8, p += 8 * Cs) {
      a0 += __ldg(p); a1 += __ldg(p + Cs); a2 += __ldg(p + 2 * Cs); a3 += __ldg(p + 3 * Cs);
      a4 += __ldg(p + 4 * Cs); a5 += __ldg(p + 5 * Cs); a6 += __ldg(p + 6 * Cs); a7 += __ldg(p + 7 * Cs);
    }
    for (; q < q1; ++q, p += Cs) a0 += __ldg(p);
  }
  sm[qy][cx] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  __syncthreads();
  if (qy == 0 && c < C) {
    float t = 0.f;
    for (int k = 0; k < NR; ++k) t += sm[k][cx];
    ws[it.dst_off + (int64_t)b * C + c] = t;
  }
}

int launch_batched_partial_reduce(const ReduceItem* items, const int2* blocks, int nblocks, float* ws, int batch, cudaStream_t s) {
  if (nblocks == 0) return 0;
  batched_partial_reduce_kernel<<<dim3((unsigned)nblocks, (unsigned)batch), 1024, 0, s>>>(items, blocks, ws);
  LFP_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) batched_style_grad_kernel(const GradItem* __restrict__ items, const int2* __restrict__ blocks,
                                                                 float* __restrict__ ws) {
  __shared__ float part[8][32];
  const int2 blk = blocks[blockIdx.x];
  const GradItem it = items[blk.x];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ci = blk.y * 32 + lane;
  const int b = blockIdx.y;
  const float* T = ws + it.T_off + (int64_t)b * it.cout;
  const float* d = ws + it.d_off + (int64_t)b * it.cout;
  float acc = 0.f;
  if (ci < it.cin)
    for (int co = w; co < it.cout; co += 8) {
      const float dv = d[co];
      acc = fmaf(T[co] * dv * dv, __ldg(it.wsq + (int64_t)co * it.cin + ci), acc);
    }
  part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && ci < it.cin) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    const int64_t i = (int64_t)b * it.cin + ci;
    ws[it.ds_off + i] = ws[it.r1_off + i] - ws[it.s_off + i] * t;
  }
}

int launch_batched_style_grad(const GradItem* items, const int2* blocks, int nblocks, float* ws, int batch, cudaStream_t s) {
  if (nblocks == 0) return 0;
  batched_style_grad_kernel<<<dim3((unsigned)nblocks, (unsigned)batch), 256, 0, s>>>(items, blocks, ws);
  LFP_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) batched_demod_kernel(const DemodItem* __restrict__ items, const int2* __restrict__ blocks,
                                                            float* __restrict__ ws) {
  const int2 blk = blocks[blockIdx.x];
  const DemodItem it = items[blk.x];
  const int lane = threadIdx.x & 31;
  const int co = blk.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (co >= it.cout) return;
  const float* s = ws + it.s_off + (int64_t)b * it.cin;
  float acc = 0.f;
  for (int ci = lane; ci < it.cin; ci += 32) {
    const float sv = s[ci];
    acc = fmaf(sv * sv, __ldg(it.wsq + (int64_t)co * it.cin + ci), acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) ws[it.d_off + (int64_t)b * it.cout + co] = rsqrtf(acc + 1e-8f);
}

int launch_batched_demod(const DemodItem* items, const int2* blocks, int nblocks, float* ws, int batch, cudaStream_t s) {
  if (nblocks == 0) return 0;
  batched_demod_kernel<<<dim3((unsigned)nblocks, (unsigned)batch), 256, 0, s>>>(items, blocks, ws);
  LFP_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// One-off weight preparation and layout helpers
// =============================================================================================
__global__ void prep_conv3x3_kernel(const float* __restrict__ W, float scale, float* __restrict__ wf,
                                    float* __restrict__ wg, float* __restrict__ wsq, int cin, int cout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cin * cout) return;
  const int co = (int)(i / cin), ci = (int)(i - (int64_t)co * cin);
  float sq = 0.f;
  for (int t = 0; t < 9; ++t) {
    const float v = W[((int64_t)co * cin + ci) * 9 + t] * scale;
    wf[((int64_t)t * cin + ci) * cout + co] = v;
    wg[((int64_t)t * cout + co) * cin + ci] = v;
    sq = fmaf(v, v, sq);
  }
  wsq[(int64_t)co * cin + ci] = sq;
}

int launch_prep_conv3x3(const float* W, float scale, float* wf, float* wg, float* wsq, int cin,
                        int cout, cudaStream_t st) {
  prep_conv3x3_kernel<<<(unsigned)ceil_div((int64_t)cin * cout, 256), 256, 0, st>>>(W, scale, wf, wg, wsq, cin, cout);
  LFP_LAUNCH_CHECK();
  return 0;
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(src[i]));
    dst[i] = __uint_as_float(r);
  }
}
int launch_round_tf32(const float* src, float* dst, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 4096) blocks = 4096;
  round_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n);
  LFP_LAUNCH_CHECK();
  return 0;
}

__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, float scale, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * scale;
}
int launch_scale_copy(const float* src, float* dst, float scale, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 4096) blocks = 4096;
  scale_copy_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, scale, n);
  LFP_LAUNCH_CHECK();
  return 0;
}

// tiled transpose of the [C, hw] <-> [hw, C] planes of each sample
template <bool TO_NHWC>
__global__ void __launch_bounds__(256) layout_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int hw) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = in + (int64_t)b * C * hw;
  float* dst = out + (int64_t)b * C * hw;
  for (int k = ty; k < 32; k += 8) {
    if (TO_NHWC) { const int c = c0 + k, p = p0 + tx; if (c < C && p < hw) t[k][tx] = src[(int64_t)c * hw + p]; }
    else { const int p = p0 + k, c = c0 + tx; if (c < C && p < hw) t[k][tx] = src[(int64_t)p * C + c]; }
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    if (TO_NHWC) { const int p = p0 + k, c = c0 + tx; if (c < C && p < hw) dst[(int64_t)p * C + c] = t[tx][k]; }
    else { const int c = c0 + k, p = p0 + tx; if (c < C && p < hw) dst[(int64_t)c * hw + p] = t[tx][k]; }
  }
}
int launch_nchw_to_nhwc(const float* in, float* out, int batch, int C, int hw, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(hw, 32), (unsigned)ceil_div(C, 32), (unsigned)batch);
  layout_kernel<true><<<grid, 256, 0, st>>>(in, out, C, hw);
  LFP_LAUNCH_CHECK();
  return 0;
}
int launch_nhwc_to_nchw(const float* in, float* out, int batch, int C, int hw, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(hw, 32), (unsigned)ceil_div(C, 32), (unsigned)batch);
  layout_kernel<false><<<grid, 256, 0, st>>>(in, out, C, hw);
  LFP_LAUNCH_CHECK();
  return 0;
}

}  // namespace lfp
