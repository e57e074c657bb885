// upfirdn2d: zero-insert upsample -> pad/crop -> FIR (true convolution) -> decimate.
//
// Semantics restated from src/op/upfirdn2d_kernel.cu:49-207 (reference) on the layout
// [major, in_h, in_w, minor]:
//   out[m, oy, ox, c] = sum_{ty,tx} S[oy*down_y + ty - pad_y0, ox*down_x + tx - pad_x0] * kflip[ty][tx]
// with S the zero-stuffed input (S[j] = in[j/up] when up | j and 0 <= j/up < in, else 0) and
// kflip[ty][tx] = kernel[kh-1-ty][kw-1-tx].
//
// B200 design: the op is HBM-bound (4 B in + 4 B out per sample at up=down=1).  The hot
// configurations in the synthesis path all have minor == 1 and a <=4x4 kernel, so they go to a
// shared-memory tiled kernel: one CTA stages a zero-padded input tile once (coalesced 4-byte
// loads; rows of 2H+1 floats rule out 16-byte global alignment) and every thread produces a
// 2x4 micro-tile from 128-bit shared loads, i.e. 10 LDS.128 per 128 FMAs at up=down=1.
// Everything else (minor > 1, big kernels, odd up/down, fp16/fp64) takes the direct kernel.
// All indexing is 64-bit (the reference overflows at 2^31 elements).
#include "common.cuh"

namespace lfp {

struct UpfirdnParams {
  int64_t major, minor;
  int in_h, in_w, out_h, out_w;
  int kh, kw;
  int up_x, up_y, down_x, down_y;
  int pad_x0, pad_y0;
};

template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };
template <typename T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v) { return (typename Acc<T>::type)v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_acc(typename Acc<T>::type v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }

// ---- direct kernel: one thread per output sample, any configuration --------------------------
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_direct_kernel(const T* __restrict__ in,
                                                               const T* __restrict__ kernel,
                                                               T* __restrict__ out,
                                                               UpfirdnParams p) {
  using A = typename Acc<T>::type;
  const int64_t total = p.major * p.out_h * p.out_w * p.minor;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = idx;
    const int64_t c = t % p.minor; t /= p.minor;
    const int ox = (int)(t % p.out_w); t /= p.out_w;
    const int oy = (int)(t % p.out_h);
    const int64_t m = t / p.out_h;
    const int base_y = oy * p.down_y - p.pad_y0;
    const int base_x = ox * p.down_x - p.pad_x0;
    A acc = 0;
    for (int ty = 0; ty < p.kh; ++ty) {
      const int sy = base_y + ty;
      if (sy < 0) continue;
      const int iy = sy / p.up_y;
      if (iy * p.up_y != sy || iy >= p.in_h) continue;
      for (int tx = 0; tx < p.kw; ++tx) {
        const int sx = base_x + tx;
        if (sx < 0) continue;
        const int ix = sx / p.up_x;
        if (ix * p.up_x != sx || ix >= p.in_w) continue;
        const A kv = to_acc<T>(kernel[(p.kh - 1 - ty) * p.kw + (p.kw - 1 - tx)]);
        acc += to_acc<T>(in[((m * p.in_h + iy) * p.in_w + ix) * p.minor + c]) * kv;
      }
    }
    out[idx] = from_acc<T>(acc);
  }
}

// ---- tiled kernel: minor == 1, fp32, kernel <= 4x4, (UP,DOWN) in {(1,1),(1,2),(2,1)} ----------
constexpr int TILE_OH = 32;
constexpr int TILE_OW = 64;

template <int UP, int DOWN>
struct TileGeom {
  // extent on the zero-stuffed grid covered by one output tile, then in input samples
  static constexpr int OH = DOWN == 2 ? 16 : TILE_OH;   // output rows per tile (halved for down-2: its input tile is 4x larger)
  static constexpr int RPT = OH / 16;                   // output rows per thread
  static constexpr int SPAN_H = (OH - 1) * DOWN + 4;
  static constexpr int SPAN_W = (TILE_OW - 1) * DOWN + 4;
  static constexpr int IN_H = UP == 1 ? SPAN_H : SPAN_H / 2 + 1;
  static constexpr int IN_W_RAW = UP == 1 ? SPAN_W : SPAN_W / 2 + 1;
  static constexpr int IN_W = (IN_W_RAW + 3) / 4 * 4;  // 16-byte aligned rows for LDS.128
};

// One CTA owns one output tile position and walks the planes (N*C) with a register-staged pipeline: the zero-padded
// input tile of plane p+1 is loaded (coalesced 4-byte loads - rows of 2H+1 floats rule out 16-byte alignment and TMA)
// while plane p is filtered out of shared memory.  Per-thread source offsets, shared
// offsets and bounds flags do not depend on the plane and are computed once.
template <int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_tiled_kernel(const float* __restrict__ in,
                                                              const float* __restrict__ kernel,
                                                              float* __restrict__ out,
                                                              UpfirdnParams p) {
  using G = TileGeom<UP, DOWN>;
  constexpr int TOT = G::IN_H * G::IN_W;
  constexpr int NL = (TOT + 255) / 256;
  __shared__ __align__(16) float tile[1][G::IN_H][G::IN_W];
  __shared__ float kf[4][4];  // flipped taps, zero-extended to 4x4

  const int tid = threadIdx.x;
  if (tid < 16) {
    const int ty = tid >> 2, tx = tid & 3;
    float v = 0.f;
    if (ty < p.kh && tx < p.kw) v = kernel[(p.kh - 1 - ty) * p.kw + (p.kw - 1 - tx)];
    kf[ty][tx] = v;
  }
  const int tile_ox = blockIdx.x * TILE_OW;
  const int tile_oy = blockIdx.y * G::OH;
  // stuffed-grid origin of the tile and the first input sample at/after it
  const int sy0 = tile_oy * DOWN - p.pad_y0;
  const int sx0 = tile_ox * DOWN - p.pad_x0;
  const int iy0 = UP == 1 ? sy0 : floor_div_i(sy0 + 1, 2);
  const int ix0 = UP == 1 ? sx0 : floor_div_i(sx0 + 1, 2);
  const int64_t plane_in = (int64_t)p.in_h * p.in_w;
  const int64_t plane_out = (int64_t)p.out_h * p.out_w;

  int goff[NL];
  uint32_t okmask = 0;
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const int i = tid + j * 256;
    const int r = i / G::IN_W, c = i - r * G::IN_W;
    const int iy = iy0 + r, ix = ix0 + c;
    const bool ok = i < TOT && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
    goff[j] = ok ? iy * p.in_w + ix : 0;
    okmask |= (ok ? 1u : 0u) << j;
  }
  // register-staged software pipeline: the next plane's tile is in flight (plain coalesced 4-byte loads: LDG issues ~4x
  // faster than 4-byte cp.async) while the current plane is filtered out of shared memory
  float stage[NL];
  auto fetch = [&](int64_t plane) {
    const float* src = in + plane * plane_in;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      stage[j] = 0.f;
      if ((okmask >> j) & 1u) stage[j] = __ldg(src + goff[j]);
    }
  };
  float* const tile_flat = &tile[0][0][0];

  constexpr int RPT = G::RPT;
  const int lx = (tid & 15) * 4;    // 4 output columns per thread
  const int ly = (tid >> 4) * RPT;  // RPT output rows per thread
  constexpr int buf = 0;
  if ((int64_t)blockIdx.z < p.major) fetch(blockIdx.z);
  for (int64_t plane = blockIdx.z; plane < p.major; plane += gridDim.z) {
    __syncthreads();  // the previous plane's readers are done (also orders the kf writes)
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int i = tid + j * 256;
      if (i < TOT) tile_flat[i] = stage[j];
    }
    __syncthreads();
    const int64_t next = plane + gridDim.z;
    if (next < p.major) fetch(next);
    float acc[RPT][4] = {};
    if (UP == 1) {
      constexpr int WR = (RPT - 1) * DOWN + 4;   // window rows
      constexpr int WC = 3 * DOWN + 4;  // window cols:   (4-1)*DOWN + 4
      constexpr int WCV = (WC + 3) / 4;
      float win[WR][WCV * 4];
#pragma unroll
      for (int r = 0; r < WR; ++r)
#pragma unroll
        for (int v = 0; v < WCV; ++v) {
          const float4 q = *reinterpret_cast<const float4*>(&tile[buf][ly * DOWN + r][lx * DOWN + v * 4]);
          win[r][v * 4 + 0] = q.x; win[r][v * 4 + 1] = q.y; win[r][v * 4 + 2] = q.z; win[r][v * 4 + 3] = q.w;
        }
#pragma unroll
      for (int ty = 0; ty < 4; ++ty)
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) {
          const float k = kf[ty][tx];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(win[r * DOWN + ty][c * DOWN + tx], k, acc[r][c]);
        }
    } else {
      // UP == 2 (DOWN == 1): only the taps that land on an even stuffed coordinate see a sample - two per axis, chosen
      // by the parity of the output position (ascending tap order, as in the 16-tap form)
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int py = (sy0 + ly + r) & 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int px = (sx0 + lx + c) & 1;
          float a = 0.f;
#pragma unroll
          for (int jy = 0; jy < 2; ++jy) {
            const int ty = py + 2 * jy;
            const int rr = ((sy0 + ly + r + ty) >> 1) - iy0;
#pragma unroll
            for (int jx = 0; jx < 2; ++jx) {
              const int tx = px + 2 * jx;
              a = fmaf(tile[buf][rr][((sx0 + lx + c + tx) >> 1) - ix0], kf[ty][tx], a);
            }
          }
          acc[r][c] = a;
        }
      }
    }
    float* dst = out + plane * plane_out;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int oy = tile_oy + ly + r;
      if (oy >= p.out_h) continue;
      const int ox = tile_ox + lx;
      float* row = dst + (int64_t)oy * p.out_w + ox;
      if (ox + 3 < p.out_w && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
        *reinterpret_cast<float4*>(row) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (ox + c < p.out_w) row[c] = acc[r][c];
      }
    }
  }
}

template <typename T>
static int upfirdn_direct_launch(const void* in, const void* kernel, void* out, const UpfirdnParams& p,
                                 cudaStream_t s) {
  const int64_t total = p.major * p.out_h * p.out_w * p.minor;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  upfirdn2d_direct_kernel<T><<<(unsigned)blocks, 256, 0, s>>>((const T*)in, (const T*)kernel, (T*)out, p);
  LFP_LAUNCH_CHECK();
  return 0;
}

template <int UP, int DOWN>
static int upfirdn_tiled_launch(const float* in, const float* kernel, float* out, const UpfirdnParams& p,
                                cudaStream_t s) {
  // enough CTAs for ~8 per SM; each walks several planes so that the two-stage pipeline has something to overlap
  constexpr int OH = TileGeom<UP, DOWN>::OH;
  const int64_t tiles = ceil_div(p.out_w, TILE_OW) * ceil_div(p.out_h, OH);
  int64_t gz = ceil_div((int64_t)num_sms() * 8, tiles);
  if (gz < 1) gz = 1;
  if (gz > p.major) gz = p.major;
  if (gz > 65535) gz = 65535;
  dim3 grid((unsigned)ceil_div(p.out_w, TILE_OW), (unsigned)ceil_div(p.out_h, OH), (unsigned)gz);
  upfirdn2d_tiled_kernel<UP, DOWN><<<grid, 256, 0, s>>>(in, kernel, out, p);
  LFP_LAUNCH_CHECK();
  return 0;
}

int upfirdn2d_dispatch(const void* input, const void* kernel, void* out, int dtype, int64_t major,
                       int in_h, int in_w, int64_t minor, int kh, int kw, int up_x, int up_y,
                       int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                       cudaStream_t s, bool allow_tiled) {
  UpfirdnParams p;
  p.major = major; p.minor = minor; p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  LFP_TRY(lfp_upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1,
                                 pad_y0, pad_y1, &p.out_h, &p.out_w));
  if (major == 0 || minor == 0 || p.out_h <= 0 || p.out_w <= 0) return 0;
  const bool small_fir = kh <= 4 && kw <= 4 && up_x == up_y && down_x == down_y;
  const bool tiny = p.out_h * (int64_t)p.out_w < 64;  // 4x4 / 8x8 maps: a tile would be mostly halo
  if (allow_tiled && dtype == LFP_F32 && minor == 1 && small_fir && !tiny) {
    if (up_x == 1 && down_x == 1) return upfirdn_tiled_launch<1, 1>((const float*)input, (const float*)kernel, (float*)out, p, s);
    if (up_x == 1 && down_x == 2) return upfirdn_tiled_launch<1, 2>((const float*)input, (const float*)kernel, (float*)out, p, s);
    if (up_x == 2 && down_x == 1) return upfirdn_tiled_launch<2, 1>((const float*)input, (const float*)kernel, (float*)out, p, s);
  }
  switch (dtype) {
    case LFP_F32: return upfirdn_direct_launch<float>(input, kernel, out, p, s);
    case LFP_F64: return upfirdn_direct_launch<double>(input, kernel, out, p, s);
    case LFP_F16: return upfirdn_direct_launch<__half>(input, kernel, out, p, s);
  }
  set_error("upfirdn2d: unsupported dtype %d", dtype);
  return LFP_EINVAL;
}

}  // namespace lfp

using namespace lfp;

extern "C" int lfp_upfirdn2d_out_size(int in_h, int in_w, int kh, int kw, int up_x, int up_y,
                                      int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                                      int pad_y1, int* out_h, int* out_w) {
  LFP_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  LFP_CHECK_ARG(kh >= 1 && kw >= 1 && in_h >= 0 && in_w >= 0, "upfirdn2d: bad kernel/input extent");
  // same integer expression as src/op/upfirdn2d_kernel.cu:236-239 (C division truncates)
  const int oh = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;
  const int ow = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  if (out_h) *out_h = oh;
  if (out_w) *out_w = ow;
  return 0;
}

extern "C" int lfp_upfirdn2d(const void* input, const void* kernel, void* out, int dtype,
                             int64_t major, int in_h, int in_w, int64_t minor, int kh, int kw,
                             int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                             int pad_y0, int pad_y1, void* stream) {
  LFP_CHECK_ARG(dtype >= 0 && dtype <= 2, "upfirdn2d: unsupported dtype %d", dtype);
  LFP_CHECK_ARG(major >= 0 && minor >= 0, "upfirdn2d: negative extent");
  if (major * minor * in_h * in_w != 0) LFP_CHECK_ARG(input && kernel && out, "upfirdn2d: null pointer");
  return upfirdn2d_dispatch(input, kernel, out, dtype, major, in_h, in_w, minor, kh, kw, up_x, up_y,
                            down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, (cudaStream_t)stream, true);
}

extern "C" int lfp_upfirdn2d_host(const void* input, const void* kernel, void* out, int dtype,
                                  int64_t major, int in_h, int in_w, int64_t minor, int kh, int kw,
                                  int up_x, int up_y, int down_x, int down_y, int pad_x0,
                                  int pad_x1, int pad_y0, int pad_y1) {
  LFP_CHECK_ARG(dtype >= 0 && dtype <= 2, "upfirdn2d_host: unsupported dtype %d", dtype);
  int oh = 0, ow = 0;
  LFP_TRY(lfp_upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, &oh, &ow));
  const size_t es = dtype == LFP_F64 ? 8 : dtype == LFP_F16 ? 2 : 4;
  const size_t nin = (size_t)major * in_h * in_w * minor, nout = (size_t)major * (oh > 0 ? oh : 0) * (ow > 0 ? ow : 0) * minor;
  if (nin == 0 || nout == 0) return 0;
  void *din = nullptr, *dk = nullptr, *dout = nullptr;
  LFP_CUDA(cudaMalloc(&din, nin * es));
  LFP_CUDA(cudaMalloc(&dk, (size_t)kh * kw * es));
  LFP_CUDA(cudaMalloc(&dout, nout * es));
  LFP_CUDA(cudaMemcpy(din, input, nin * es, cudaMemcpyHostToDevice));
  LFP_CUDA(cudaMemcpy(dk, kernel, (size_t)kh * kw * es, cudaMemcpyHostToDevice));
  int rc = lfp_upfirdn2d(din, dk, dout, dtype, major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, nullptr);
  if (rc == 0) {
    cudaError_t e = cudaMemcpy(out, dout, nout * es, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("upfirdn2d_host: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  cudaFree(din); cudaFree(dk); cudaFree(dout);
  return rc;
}
