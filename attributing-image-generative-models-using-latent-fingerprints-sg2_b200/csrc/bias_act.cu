// bias + leaky-ReLU + gain, forward and backward (the "fused" native module).
//
// Semantics restated from src/op/fused_bias_act_kernel.cu:19-65 (reference): per element
//   x += b[(i / step_b) % size_b];  y = act(x | ref);  out = y * scale
// Design for B200: this is a pure streaming op (8 B/elem forward, 12 B/elem backward in fp32),
// so the kernel is built around 128-bit loads/stores with the channel index hoisted out of the
// element loop (one plane = one (n, c) pair of step_b contiguous elements), 64-bit offsets, and
// a grid sized in multiples of the SM count.  A scalar kernel handles ragged cases (rank-2
// inputs where step_b == 1, misaligned pointers, step_b % vec != 0).
#include <stdlib.h>

#include "common.cuh"

namespace lfp {

template <typename T>
struct Arith;  // per-op arithmetic in the reference's scalar type

template <>
struct Arith<float> {
  using C = float;
  static __device__ __forceinline__ C ld(float v) { return v; }
  static __device__ __forceinline__ float st(C v) { return v; }
  static __device__ __forceinline__ C rnd(C v) { return v; }
};
template <>
struct Arith<double> {
  using C = double;
  static __device__ __forceinline__ C ld(double v) { return v; }
  static __device__ __forceinline__ double st(C v) { return v; }
  static __device__ __forceinline__ C rnd(C v) { return v; }
};
// c10::Half arithmetic = convert to float, operate, round back to half after every op
template <>
struct Arith<__half> {
  using C = float;
  static __device__ __forceinline__ C ld(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half st(C v) { return __float2half_rn(v); }
  static __device__ __forceinline__ C rnd(C v) { return __half2float(__float2half_rn(v)); }
};

template <typename T>
__device__ __forceinline__ typename Arith<T>::C bias_act_one(typename Arith<T>::C x,
                                                              typename Arith<T>::C b,
                                                              typename Arith<T>::C ref, int code,
                                                              typename Arith<T>::C alpha,
                                                              typename Arith<T>::C scale,
                                                              bool use_bias) {
  using A = Arith<T>;
  typename A::C y;
  if (use_bias) x = A::rnd(x + b);
  switch (code) {
    case 30: y = (x > 0) ? x : A::rnd(x * alpha); break;
    case 31: y = (ref > 0) ? x : A::rnd(x * alpha); break;
    case 12:
    case 32: y = 0; break;
    default: y = x; break;
  }
  return A::rnd(y * scale);
}

// ---- scalar fallback: any layout, 64-bit indices -------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(
    const T* __restrict__ x, const T* __restrict__ bias, const T* __restrict__ ref,
    T* __restrict__ out, int64_t size_x, int64_t step_b, int64_t size_b, int code, float alpha_f,
    float scale_f) {
  using A = Arith<T>;
  typename A::C alpha = A::ld((T)alpha_f), scale = A::ld((T)scale_f);
  const bool use_bias = bias != nullptr, use_ref = ref != nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size_x;
       i += (int64_t)gridDim.x * blockDim.x) {
    typename A::C b = 0, r = 0;
    if (use_bias) b = A::ld(bias[(i / step_b) % size_b]);
    if (use_ref) r = A::ld(ref[i]);
    out[i] = A::st(bias_act_one<T>(A::ld(x[i]), b, r, code, alpha, scale, use_bias));
  }
}

// ---- vector kernel: one plane (step_b elements sharing a bias value) per blockIdx.y ---------
template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) bias_act_plane_kernel(
    const T* __restrict__ x, const T* __restrict__ bias, const T* __restrict__ ref,
    T* __restrict__ out, int64_t planes, int64_t step_b, int64_t size_b, int code, float alpha_f,
    float scale_f) {
  using A = Arith<T>;
  constexpr int N = Vec16<T>::N;
  typename A::C alpha = A::ld((T)alpha_f), scale = A::ld((T)scale_f);
  const bool use_bias = bias != nullptr, use_ref = ref != nullptr;
  const int64_t vec_per_plane = step_b / N;
  for (int64_t plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    typename A::C b = 0;
    if (use_bias) b = A::ld(bias[plane % size_b]);
    const uint4* xp = reinterpret_cast<const uint4*>(x + plane * step_b);
    const uint4* rp = use_ref ? reinterpret_cast<const uint4*>(ref + plane * step_b) : nullptr;
    uint4* op = reinterpret_cast<uint4*>(out + plane * step_b);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < vec_per_plane;
         i0 += stride * UNROLL) {
      uint4 xv[UNROLL], rv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        int64_t i = i0 + u * stride;
        if (i < vec_per_plane) {
          xv[u] = __ldcs(xp + i);
          if (use_ref) rv[u] = __ldcs(rp + i);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        int64_t i = i0 + u * stride;
        if (i < vec_per_plane) {
          Vec16<T> a, r, o;
          *reinterpret_cast<uint4*>(&a) = xv[u];
          if (use_ref) *reinterpret_cast<uint4*>(&r) = rv[u];
#pragma unroll
          for (int k = 0; k < N; ++k) {
            typename A::C rr = use_ref ? A::ld(r.v[k]) : (typename A::C)0;
            o.v[k] = A::st(bias_act_one<T>(A::ld(a.v[k]), b, rr, code, alpha, scale, use_bias));
          }
          __stcs(op + i, *reinterpret_cast<uint4*>(&o));
        }
      }
    }
  }
}

// ---- flat vector kernel: small planes (a plane is shorter than one CTA pass) ------------------
// The plane kernel gives every plane its own CTA row; at 16 x 16 (64 vectors) three quarters of each CTA idle and the launch is
// 32768 one-kilobyte CTAs (0.2 of the HBM rate at [64, 512, 16, 16]).  Here the tensor is one run of 16-byte vectors and the
// bias index is recovered per vector (plane = i / vec_per_plane, one 32-bit division per 16 bytes where the reference divides
// per element, src/op/fused_bias_act_kernel.cu:33-37).  Same per-element arithmetic as the other two kernels.
template <typename T, int UNROLL>
__global__ void __launch_bounds__(256) bias_act_flat_kernel(
    const T* __restrict__ x, const T* __restrict__ bias, const T* __restrict__ ref,
    T* __restrict__ out, int64_t total_vec, int64_t vec_per_plane, int64_t size_b, int code, float alpha_f,
    float scale_f) {
  using A = Arith<T>;
  constexpr int N = Vec16<T>::N;
  typename A::C alpha = A::ld((T)alpha_f), scale = A::ld((T)scale_f);
  const bool use_bias = bias != nullptr, use_ref = ref != nullptr;
  const bool small = total_vec < (1ll << 31);   // 32-bit divisions
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  const uint4* rp = reinterpret_cast<const uint4*>(ref);
  uint4* op = reinterpret_cast<uint4*>(out);
  for (int64_t base = (int64_t)blockIdx.x * (256 * UNROLL); base < total_vec; base += (int64_t)gridDim.x * (256 * UNROLL)) {
    uint4 xv[UNROLL], rv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + u * 256 + threadIdx.x;
      if (i < total_vec) {
        xv[u] = __ldcs(xp + i);
        if (use_ref) rv[u] = __ldcs(rp + i);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = base + u * 256 + threadIdx.x;
      if (i < total_vec) {
        typename A::C b = 0;
        if (use_bias) {
          const int64_t plane = small ? (int64_t)((uint32_t)i / (uint32_t)vec_per_plane) : i / vec_per_plane;
          const int64_t c = small ? (int64_t)((uint32_t)plane % (uint32_t)size_b) : plane % size_b;
          b = A::ld(bias[c]);
        }
        Vec16<T> a, r, o;
        *reinterpret_cast<uint4*>(&a) = xv[u];
        if (use_ref) *reinterpret_cast<uint4*>(&r) = rv[u];
#pragma unroll
        for (int k = 0; k < N; ++k) {
          typename A::C rr = use_ref ? A::ld(r.v[k]) : (typename A::C)0;
          o.v[k] = A::st(bias_act_one<T>(A::ld(a.v[k]), b, rr, code, alpha, scale, use_bias));
        }
        __stcs(op + i, *reinterpret_cast<uint4*>(&o));
      }
    }
  }
}

template <typename T>
static int bias_act_launch(const void* x_, const void* bias_, const void* ref_, void* out_,
                           int64_t size_x, int64_t step_b, int64_t size_b, int code, float alpha,
                           float scale, cudaStream_t stream) {
  const T* x = (const T*)x_;
  const T* bias = (const T*)bias_;
  const T* ref = (const T*)ref_;
  T* out = (T*)out_;
  constexpr int N = Vec16<T>::N;
  auto aligned = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  constexpr int UNROLL_F = 4;
  const bool vec_ok = step_b >= N && step_b % N == 0 && size_x % step_b == 0 && aligned(x) && aligned(ref) && aligned(out);
  static const bool no_flat = getenv("LFP_BIAS_ACT_NO_FLAT") != nullptr;   // A/B switch: the first version's kernel choice
  const bool flat = vec_ok && !no_flat && step_b / N < 256 * UNROLL_F && (size_b < (1ll << 31));
  const bool planar = vec_ok && !flat && step_b >= 64 * N;
  if (flat) {
    const int64_t total_vec = size_x / N;
    int64_t blocks = ceil_div(total_vec, 256 * UNROLL_F);
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    bias_act_flat_kernel<T, UNROLL_F><<<(unsigned)blocks, 256, 0, stream>>>(x, bias, ref, out, total_vec, step_b / N,
                                                                          bias ? size_b : 1, code, alpha, scale);
  } else if (planar) {
    const int64_t planes = size_x / step_b;
    const int64_t vec_per_plane = step_b / N;
    constexpr int UNROLL = 4;
    int64_t bx = ceil_div(vec_per_plane, 256 * UNROLL);
    // keep the grid around 16 CTAs per SM: split planes over y, shrink x for tiny planes
    int64_t want = (int64_t)num_sms() * 16;
    int64_t by = planes < 65535 ? planes : 65535;
    if (bx * by > want * 4) {
      bx = ceil_div(want * 4, by);
      if (bx < 1) bx = 1;
    }
    dim3 grid((unsigned)bx, (unsigned)by);
    bias_act_plane_kernel<T, UNROLL><<<grid, 256, 0, stream>>>(x, bias, ref, out, planes, step_b,
                                                               bias ? size_b : 1, code, alpha,
                                                               scale);
  } else {
    int64_t blocks = ceil_div(size_x, 256);
    int64_t cap = (int64_t)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    bias_act_scalar_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(
        x, bias, ref, out, size_x, step_b > 0 ? step_b : 1, bias ? size_b : 1, code, alpha, scale);
  }
  LFP_LAUNCH_CHECK();
  return 0;
}

// ---- deterministic grad_bias = sum over (outer, inner) of g[outer, c, inner] ------------------
// pass 1: one CTA per (c, chunk of outer*inner) -> partial; pass 2: one warp per c sums partials.
__global__ void __launch_bounds__(256) bias_grad_partial_kernel(const float* __restrict__ g,
                                                                float* __restrict__ partial,
                                                                int64_t outer, int64_t size_b,
                                                                int64_t step_b, int chunks) {
  const int c = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t total = outer * step_b;  // elements of channel c
  const int64_t per = ceil_div(total, chunks);
  const int64_t lo = per * chunk, hi = (lo + per < total) ? lo + per : total;
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    int64_t o = i / step_b, r = i - o * step_b;
    acc += g[(o * size_b + c) * step_b + r];
  }
  __shared__ float sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(int64_t)c * chunks + chunk] = sm[0];
}
__global__ void bias_grad_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                       int64_t size_b, int chunks) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= size_b) return;
  float acc = 0.f;
  for (int k = 0; k < chunks; ++k) acc += partial[c * chunks + k];
  out[c] = acc;
}

static int bias_grad_chunks(int64_t outer, int64_t size_b, int64_t step_b) {
  int64_t total = outer * step_b;
  int64_t want = ceil_div((int64_t)num_sms() * 8, size_b > 0 ? size_b : 1);
  int64_t maxc = ceil_div(total, 1024);
  if (want > maxc) want = maxc;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

}  // namespace lfp

using namespace lfp;

extern "C" int lfp_fused_bias_act(const void* x, const void* bias, const void* ref, void* out,
                                  int dtype, int64_t size_x, int64_t step_b, int64_t size_b,
                                  int act, int grad, float alpha, float scale, void* stream) {
  LFP_CHECK_ARG(size_x >= 0, "fused_bias_act: negative size_x");
  if (size_x == 0) return 0;
  LFP_CHECK_ARG(x && out, "fused_bias_act: null input/output");
  LFP_CHECK_ARG(bias == nullptr || (size_b > 0 && step_b > 0),
                "fused_bias_act: bias given but size_b=%lld step_b=%lld", (long long)size_b,
                (long long)step_b);
  const int code = act * 10 + grad;
  cudaStream_t s = (cudaStream_t)stream;
  if (step_b <= 0) step_b = 1;
  switch (dtype) {
    case LFP_F32: return bias_act_launch<float>(x, bias, ref, out, size_x, step_b, size_b, code, alpha, scale, s);
    case LFP_F64: return bias_act_launch<double>(x, bias, ref, out, size_x, step_b, size_b, code, alpha, scale, s);
    case LFP_F16: return bias_act_launch<__half>(x, bias, ref, out, size_x, step_b, size_b, code, alpha, scale, s);
    default: set_error("fused_bias_act: unsupported dtype %d", dtype); return LFP_EINVAL;
  }
}

static size_t dtype_size(int dtype) { return dtype == LFP_F64 ? 8 : dtype == LFP_F16 ? 2 : 4; }

extern "C" int lfp_fused_bias_act_host(const void* x, const void* bias, const void* ref, void* out,
                                       int dtype, int64_t size_x, int64_t step_b, int64_t size_b,
                                       int act, int grad, float alpha, float scale) {
  LFP_CHECK_ARG(dtype >= 0 && dtype <= 2, "fused_bias_act_host: bad dtype");
  if (size_x == 0) return 0;
  const size_t es = dtype_size(dtype);
  HostStaging& hs = host_staging();
  void* dx = hs.get(0, size_x * es);
  void* dout = hs.get(1, size_x * es);
  void* db = bias ? hs.get(2, size_b * es) : nullptr;
  void* dr = ref ? hs.get(3, size_x * es) : nullptr;
  if (!dx || !dout || (bias && !db) || (ref && !dr)) { set_error("fused_bias_act_host: device allocation failed"); return LFP_ENOMEM; }
  cudaStream_t s = 0;
  cudaError_t e = cudaMemcpyAsync(dx, x, size_x * es, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && bias) e = cudaMemcpyAsync(db, bias, size_b * es, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && ref) e = cudaMemcpyAsync(dr, ref, size_x * es, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) { set_error("fused_bias_act_host: %s", cudaGetErrorString(e)); return (int)e; }
  int rc = lfp_fused_bias_act(dx, db, dr, dout, dtype, size_x, step_b, size_b, act, grad, alpha, scale, s);
  if (rc == 0) {
    e = cudaMemcpyAsync(out, dout, size_x * es, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("fused_bias_act_host: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  return rc;
}

extern "C" size_t lfp_bias_grad_reduce_scratch(int64_t outer, int64_t size_b, int64_t step_b) {
  return (size_t)size_b * bias_grad_chunks(outer, size_b, step_b) * sizeof(float);
}

extern "C" int lfp_bias_grad_reduce(const float* g, float* out, int64_t outer, int64_t size_b,
                                    int64_t step_b, void* scratch, size_t scratch_bytes,
                                    void* stream) {
  LFP_CHECK_ARG(g && out && size_b > 0 && outer > 0 && step_b > 0, "bias_grad_reduce: bad args");
  LFP_CHECK_ARG(size_b <= 65535, "bias_grad_reduce: size_b too large");
  const int chunks = bias_grad_chunks(outer, size_b, step_b);
  if (scratch_bytes < (size_t)size_b * chunks * sizeof(float) || !scratch) {
    set_error("bias_grad_reduce: scratch too small");
    return LFP_ENOMEM;
  }
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(chunks, (unsigned)size_b);
  bias_grad_partial_kernel<<<grid, 256, 0, s>>>(g, (float*)scratch, outer, size_b, step_b, chunks);
  LFP_LAUNCH_CHECK();
  bias_grad_final_kernel<<<(unsigned)ceil_div(size_b, 128), 128, 0, s>>>((const float*)scratch, out, size_b, chunks);
  LFP_LAUNCH_CHECK();
  return 0;
}
