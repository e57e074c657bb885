// upfirdn2d: zero-insert upsample -> pad/crop -> FIR (true convolution) -> decimate.
//
// Semantics restated from src/op/upfirdn2d_kernel.cu:49-207 (reference) on the layout
// [major, in_h, in_w, minor]:
//   out[m, oy, ox, c] = sum_{ty,tx} S[oy*down_y + ty - pad_y0, ox*down_x + tx - pad_x0] * kflip[ty][tx]
// with S the zero-stuffed input (S[j] = in[j/up] when up | j and 0 <= j/up < in, else 0) and
// kflip[ty][tx] = kernel[kh-1-ty][kw-1-tx].
//
// B200 design: the op is HBM-bound (4 B in + 4 B out per sample at up=down=1), and at the memory latency a loaded HBM system
// has (~2 us) the bound is only reached with >= 50 KB of reads in flight per SM.  The hot configurations of the synthesis
// path (minor == 1, taps <= 4x4, fp32) therefore go to
//   * up = down = 1 (Blur and its adjoint): a shared-memory row ring filled by 1-D bulk copies (cp.async.bulk + mbarrier),
//     consumed by warps that keep the partially summed output rows in registers - upfirdn2d_ring11_kernel;
//   * up = 2 / down = 2: warp-streaming kernels without shared memory (a lane owns an output column, neighbours by shuffle);
//   * many small planes (<= 34 x 34, e.g. the 4 ... 32 px layers at a large batch): runs of planes staged zero-haloed in
//     shared memory, vertical register strips - upfirdn2d_planes_kernel;
// everything else (minor > 1, big kernels, odd up/down, fp16/fp64) takes the direct kernel.
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

// ---- warp-streaming kernels: minor == 1, fp32, kernel <= 4x4 ----------------------------------------------------------
// One warp owns a strip of output columns and walks down the rows of a plane.  A lane loads one input column per row
// (one coalesced 128-byte request per warp and row, no alignment requirement - rows of 2H+1 floats are fine), gets its
// right-hand neighbours by warp shuffle, and carries the partially summed output rows in registers, so every input sample
// is loaded once per strip and there is no shared memory and no CTA barrier.  U rows are in flight per lane (register
// double buffer).  Taps are summed in ascending (ty, tx) order, as the 16-tap loop of the reference kernel does.
constexpr int STREAM_U = 8;

__device__ __forceinline__ void load_flipped_taps(const float* __restrict__ kernel, const UpfirdnParams& p, float (&k)[4][4]) {
#pragma unroll
  for (int ty = 0; ty < 4; ++ty)
#pragma unroll
    for (int tx = 0; tx < 4; ++tx)
      k[ty][tx] = (ty < p.kh && tx < p.kw) ? __ldg(kernel + (p.kh - 1 - ty) * p.kw + (p.kw - 1 - tx)) : 0.f;
}

// up = down = 1: 29 output columns per warp (32 lanes minus the 3-sample halo).
// Rows are handled in batches of U; a batch whose rows all lie inside the image (and whose outputs all exist) takes the
// FAST instantiation, which has no per-row predicates - the kernel is otherwise bound by instruction issue, not by HBM.
constexpr int S11_W = 29;
#ifndef LFP_S11_MINB
#define LFP_S11_MINB 5
#endif
// Two planes ride in one lane: the samples of planes (2z, 2z+1) at the same position are the halves of a 64-bit register pair
// and every tap is one packed FFMA2 (fma.rn.f32x2), which halves the FMA issue slots per output.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
struct S11Ctx {
  float k[4][4];     // flipped taps (broadcast to both halves of an FFMA2 by the instruction's scalar operand form)
  int in_w, in_h, out_w, iy0, total;
  int64_t in_b, out_b;  // offset of the lane's second plane (0 / unused when the pair has only one)
  bool colok, stok, stok_b;
};
template <bool FAST>
__device__ __forceinline__ void s11_load(float (&da)[STREAM_U], float (&db)[STREAM_U], const S11Ctx& c, const float* __restrict__ col,
                                         int jb) {
  // col: first plane's base + row iy0 + clamped input column (32-bit row offsets from there); rows jb .. jb + U-1 of the block
#pragma unroll
  for (int u = 0; u < STREAM_U; ++u) {
    const int iy = c.iy0 + jb + u;
    const float* ptr = col + (jb + u) * c.in_w;
    if (FAST) {
      da[u] = __ldg(ptr);
      db[u] = __ldg(ptr + c.in_b);
    } else {
      const bool ok = c.colok && jb + u < c.total && iy >= 0 && iy < c.in_h;
      da[u] = ok ? __ldg(ptr) : 0.f;
      db[u] = ok ? __ldg(ptr + c.in_b) : 0.f;
    }
  }
}
template <int MODE>  // 0: every row and store predicated, 1: no predicates, 2: first batch of a block (rows 0..2 complete nothing)
__device__ __forceinline__ void s11_rows(const float (&va)[STREAM_U], const float (&vb)[STREAM_U], const S11Ctx& c, float* __restrict__ dst,
                                         int jb, uint64_t& a1, uint64_t& a2, uint64_t& a3) {
  // dst: first plane's base + r0 * out_w + ox; output row of input row j is j - 3
#pragma unroll
  for (int u = 0; u < STREAM_U; ++u) {
    const int j = jb + u;
    // FAST loads keep raw values: out-of-image columns are zeroed here
    const float xa = c.colok ? va[u] : 0.f, xb = c.colok ? vb[u] : 0.f;
    const uint64_t v0 = pack2(xa, xb);
    const uint64_t v1 = pack2(__shfl_down_sync(0xffffffffu, xa, 1), __shfl_down_sync(0xffffffffu, xb, 1));
    const uint64_t v2 = pack2(__shfl_down_sync(0xffffffffu, xa, 2), __shfl_down_sync(0xffffffffu, xb, 2));
    const uint64_t v3 = pack2(__shfl_down_sync(0xffffffffu, xa, 3), __shfl_down_sync(0xffffffffu, xb, 3));
#define LFP_K2(ty, tx) pack2(c.k[ty][tx], c.k[ty][tx])
    uint64_t a0 = ffma2(v0, LFP_K2(0, 0), 0ull);
    a0 = ffma2(v1, LFP_K2(0, 1), a0); a0 = ffma2(v2, LFP_K2(0, 2), a0); a0 = ffma2(v3, LFP_K2(0, 3), a0);
    a1 = ffma2(v0, LFP_K2(1, 0), a1); a1 = ffma2(v1, LFP_K2(1, 1), a1); a1 = ffma2(v2, LFP_K2(1, 2), a1); a1 = ffma2(v3, LFP_K2(1, 3), a1);
    a2 = ffma2(v0, LFP_K2(2, 0), a2); a2 = ffma2(v1, LFP_K2(2, 1), a2); a2 = ffma2(v2, LFP_K2(2, 2), a2); a2 = ffma2(v3, LFP_K2(2, 3), a2);
    a3 = ffma2(v0, LFP_K2(3, 0), a3); a3 = ffma2(v1, LFP_K2(3, 1), a3); a3 = ffma2(v2, LFP_K2(3, 2), a3); a3 = ffma2(v3, LFP_K2(3, 3), a3);
#undef LFP_K2
    float* ptr = dst + (j - 3) * c.out_w;
    float ra, rb;
    unpack2(a3, ra, rb);
    const bool row_ok = MODE == 1 || (MODE == 2 && u >= 3) || (MODE == 0 && j >= 3 && j < c.total);
    if (row_ok && c.stok) *ptr = ra;
    if (row_ok && c.stok_b) ptr[c.out_b] = rb;
    a3 = a2; a2 = a1; a1 = a0;
  }
}
__global__ void __launch_bounds__(128, LFP_S11_MINB) upfirdn2d_stream11_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                                               float* __restrict__ out, UpfirdnParams p, int rh) {
  constexpr int U = STREAM_U;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 4 + warp) * S11_W;
  if (c0 >= p.out_w) return;
  S11Ctx c;
  load_flipped_taps(kernel, p, c.k);
  const int r0 = blockIdx.y * rh;
  const int rows = min(rh, p.out_h - r0);
  c.total = rows + 3;  // input rows feeding this row block
  c.in_w = p.in_w; c.in_h = p.in_h; c.out_w = p.out_w;
  const int ix = c0 + lane - p.pad_x0;
  c.colok = ix >= 0 && ix < p.in_w;
  const int ixc = min(max(ix, 0), p.in_w - 1);
  c.iy0 = r0 - p.pad_y0;
  const int ox = c0 + lane;
  c.stok = lane < S11_W && ox < p.out_w;
  const int64_t plane_in = (int64_t)p.in_h * p.in_w, plane_out = (int64_t)p.out_h * p.out_w;
  auto load_fast = [&](int jb) { return c.iy0 + jb >= 0 && c.iy0 + jb + U <= c.in_h && jb + U <= c.total; };
  for (int64_t plane = 2 * (int64_t)blockIdx.z; plane < p.major; plane += 2 * (int64_t)gridDim.z) {
    const bool has_b = plane + 1 < p.major;
    c.in_b = has_b ? plane_in : 0; c.out_b = has_b ? plane_out : 0; c.stok_b = c.stok && has_b;
    const float* col = in + plane * plane_in + (int64_t)c.iy0 * p.in_w + ixc;
    float* dst = out + plane * plane_out + (int64_t)r0 * p.out_w + min(ox, p.out_w - 1);
    float va[U], vb[U], na[U], nb[U];
    uint64_t a1 = 0, a2 = 0, a3 = 0;
    if (load_fast(0)) s11_load<true>(va, vb, c, col, 0); else s11_load<false>(va, vb, c, col, 0);
    for (int jb = 0; jb < c.total; jb += U) {
      if (load_fast(jb + U)) s11_load<true>(na, nb, c, col, jb + U); else s11_load<false>(na, nb, c, col, jb + U);
      if (jb + U <= c.total) {
        if (jb == 0) s11_rows<2>(va, vb, c, dst, jb, a1, a2, a3); else s11_rows<1>(va, vb, c, dst, jb, a1, a2, a3);
      } else {
        s11_rows<0>(va, vb, c, dst, jb, a1, a2, a3);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { va[u] = na[u]; vb[u] = nb[u]; }
    }
  }
}

// ---- row-ring kernel: up = down = 1, wide maps ---------------------------------------------------------------------------
// The streaming kernel above holds the bytes it has in flight in registers, and at HBM latency under load (~2 us) a 4x4 FIR
// has not enough of them to keep ~50 KB per SM in flight.  Here a producer thread streams the input rows of a 256-column
// segment into a shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier transaction bytes): a row of 2H+1 floats
// is not 16-byte aligned, so each copy takes the 16-byte aligned superset of the row piece and the consumers add the row's
// skew (0..3 floats) to their shared-memory index.  Eight consumer warps walk down the rows, 32 output columns each (full,
// aligned 128-byte stores), two planes per lane (FFMA2), partial sums in registers as in the streaming kernel.
namespace ring {
constexpr int G = 4;                           // input rows per ring slot (per plane)
constexpr int D = 8;                           // ring slots
// NCW consumer warps own NCW * 32 output columns; one staged row piece = columns + 3 halo + up to 3 skew floats, 16-byte multiple
__host__ __device__ constexpr int rowb(int ncw) { return ((ncw * 32 + 3 + 3) * 4 + 15) / 16 * 16; }
__host__ __device__ constexpr int slotb(int ncw) { return G * 2 * rowb(ncw); }
__host__ __device__ constexpr int smem_bytes(int ncw) { return D * slotb(ncw) + 2 * D * 8 + 16; }

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
__device__ __forceinline__ float lds(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
}  // namespace ring

struct RingCtx {
  float k[4][4];
  uint32_t sbase, bars, slotb;      // ring base, barrier base (full[D], empty[D]), bytes per slot
  uint32_t roff[ring::G][2];        // byte offset of the lane's tap 0 inside a slot, per row of a group and plane: row r of
                                    // every group has the same skew, since G * in_w is a multiple of 4 floats
  int ngroups, total, jlo, jhi;     // input rows [jlo, jhi) of the band lie inside the image
  int out_w;
  uint32_t tokmask;                 // EDGE: bit tx set when the lane's tap tx is inside the image
  int stok, stok_b;
  bool any;
};
// One group of G input rows.  EDGE: some tap of some lane of the warp falls outside the image (per-tap predicates).
// FULL: all G rows lie inside the image and each completes an output row of the band (no per-row predicates).
template <bool EDGE, bool FULL>
__device__ __forceinline__ void ring_group(const RingCtx& c, uint32_t slot_addr, int g, float*& pa, float*& pb, uint64_t& a1, uint64_t& a2,
                                           uint64_t& a3) {
  using namespace ring;
#pragma unroll
  for (int r = 0; r < G; ++r) {
    const int j = g * G + r;
    float xa[4], xb[4];
    const bool rowok = FULL || (unsigned)(j - c.jlo) < (unsigned)(c.jhi - c.jlo);   // warp-uniform
    const uint32_t ra = slot_addr + c.roff[r][0], rb = slot_addr + c.roff[r][1];
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) {
      xa[tx] = 0.f; xb[tx] = 0.f;
      if (rowok && (!EDGE || ((c.tokmask >> tx) & 1u))) { xa[tx] = lds(ra + tx * 4); xb[tx] = lds(rb + tx * 4); }
    }
    const uint64_t v0 = pack2(xa[0], xb[0]), v1 = pack2(xa[1], xb[1]), v2 = pack2(xa[2], xb[2]), v3 = pack2(xa[3], xb[3]);
#define LFP_K2(ty, tx) pack2(c.k[ty][tx], c.k[ty][tx])
    uint64_t a0 = ffma2(v0, LFP_K2(0, 0), 0ull);
    a0 = ffma2(v1, LFP_K2(0, 1), a0); a0 = ffma2(v2, LFP_K2(0, 2), a0); a0 = ffma2(v3, LFP_K2(0, 3), a0);
    a1 = ffma2(v0, LFP_K2(1, 0), a1); a1 = ffma2(v1, LFP_K2(1, 1), a1); a1 = ffma2(v2, LFP_K2(1, 2), a1); a1 = ffma2(v3, LFP_K2(1, 3), a1);
    a2 = ffma2(v0, LFP_K2(2, 0), a2); a2 = ffma2(v1, LFP_K2(2, 1), a2); a2 = ffma2(v2, LFP_K2(2, 2), a2); a2 = ffma2(v3, LFP_K2(2, 3), a2);
    a3 = ffma2(v0, LFP_K2(3, 0), a3); a3 = ffma2(v1, LFP_K2(3, 1), a3); a3 = ffma2(v2, LFP_K2(3, 2), a3); a3 = ffma2(v3, LFP_K2(3, 3), a3);
#undef LFP_K2
    float oa, ob;
    unpack2(a3, oa, ob);
    const bool row_st = FULL || (unsigned)(j - 3) < (unsigned)(c.total - 3);
    if (row_st && c.stok) *pa = oa;
    if (row_st && c.stok_b) *pb = ob;
    pa += c.out_w; pb += c.out_w;
    a3 = a2; a2 = a1; a1 = a0;
  }
}
template <bool EDGE>
__device__ __forceinline__ void ring_consume(const RingCtx& c, float* pa, float* pb) {
  using namespace ring;
  // pa / pb: output pointers of the row that input row 0 would complete (3 rows above the band), advanced by one row per input row
  uint64_t a1 = 0, a2 = 0, a3 = 0;
  uint32_t slot_addr = c.sbase, full_bar = c.bars, empty_bar = c.bars + D * 8, par = 0;
  int slot = 0;
  for (int g = 0; g < c.ngroups; ++g) {
    if (c.any) mbar_wait(full_bar, par);
    const int j0 = g * G;
    if (j0 >= c.jlo && j0 + G <= c.jhi && j0 >= 3 && j0 + G <= c.total) ring_group<EDGE, true>(c, slot_addr, g, pa, pb, a1, a2, a3);
    else ring_group<EDGE, false>(c, slot_addr, g, pa, pb, a1, a2, a3);
    __syncwarp();
    if ((threadIdx.x & 31) == 0 && c.any) mbar_arrive(empty_bar);
    slot_addr += c.slotb; full_bar += 8; empty_bar += 8;
    if (++slot == D) { slot = 0; slot_addr = c.sbase; full_bar = c.bars; empty_bar = c.bars + D * 8; par ^= 1u; }
  }
}

template <int NCW>
__global__ void __launch_bounds__((NCW + 1) * 32) upfirdn2d_ring11_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                                        float* __restrict__ out, UpfirdnParams p, int RB) {
  using namespace ring;
  constexpr int CW = NCW * 32, ROWB = rowb(NCW), SLOTB = slotb(NCW);
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const uint32_t sbase = smem_u32(ring_smem);
  const uint32_t bars = sbase + D * SLOTB;  // full[D], empty[D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cx0 = blockIdx.x * CW;                         // first output column of the segment
  const int r0 = blockIdx.y * RB;
  const int rows = min(RB, p.out_h - r0);
  const int total = rows + 3;                              // input rows feeding the band
  const int ngroups = (total + G - 1) / G;
  const int iy0 = r0 - p.pad_y0;
  // staged input columns [cbase, cend)
  const int cfirst = cx0 - p.pad_x0;
  const int cbase = max(cfirst, 0), cend = min(cfirst + CW + 3, p.in_w);
  const int ncols = cend - cbase;
  const int64_t plane_in = (int64_t)p.in_h * p.in_w, plane_out = (int64_t)p.out_h * p.out_w;
  const int64_t plane = 2 * (int64_t)blockIdx.z;
  const bool has_b = plane + 1 < p.major;
  // consumer warps that own at least one output column (the last segment of a row may be narrow); the others leave at once
  const int nwork = min(NCW, (p.out_w - cx0 + 31) >> 5);
  if (threadIdx.x == 0) {
    for (int i = 0; i < D; ++i) { mbar_init(bars + i * 8, 1); mbar_init(bars + (D + i) * 8, nwork); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int jlo = max(0, -iy0), jhi = min(total, p.in_h - iy0);
  if (warp == NCW) {
    // ---- producer: one thread streams the rows of both planes into the ring ----
    if (lane == 0 && ncols > 0) {
      const float* rowa = in + plane * plane_in + (int64_t)iy0 * p.in_w + cbase;   // first staged sample of row 0
      int ska = (int)((plane * plane_in + (int64_t)iy0 * p.in_w + cbase) & 3), skb = (int)((ska + plane_in) & 3);
      const int iw4 = p.in_w & 3;
      uint32_t slot_addr = sbase, full_bar = bars, empty_bar = bars + D * 8, par = 0;   // parity of the first reuse wait
      int slot = 0;
      for (int g = 0; g < ngroups; ++g) {
        if (g >= D) mbar_wait(empty_bar, par);
        // transaction bytes of the group, then the copies
        uint32_t ba[G], bb[G], sum = 0;
        int sa = ska, sb = skb;
#pragma unroll
        for (int r = 0; r < G; ++r) {
          const int j = g * G + r;
          const bool ok = j >= jlo && j < jhi;
          ba[r] = ok ? (uint32_t)(((sa + ncols) * 4 + 15) & ~15) : 0u;
          bb[r] = ok && has_b ? (uint32_t)(((sb + ncols) * 4 + 15) & ~15) : 0u;
          sum += ba[r] + bb[r];
          sa = (sa + iw4) & 3; sb = (sb + iw4) & 3;
        }
        mbar_expect_tx(full_bar, sum);
#pragma unroll
        for (int r = 0; r < G; ++r) {
          const float* src = rowa + (int64_t)(g * G + r) * p.in_w;
          if (ba[r]) bulk_g2s(slot_addr + (r * 2) * ROWB, src - ska, ba[r], full_bar);
          if (bb[r]) bulk_g2s(slot_addr + (r * 2 + 1) * ROWB, src + plane_in - skb, bb[r], full_bar);
          ska = (ska + iw4) & 3; skb = (skb + iw4) & 3;
        }
        slot_addr += SLOTB; full_bar += 8; empty_bar += 8;
        if (++slot == D) { slot = 0; slot_addr = sbase; full_bar = bars; empty_bar = bars + D * 8; if (g >= D) par ^= 1u; }
      }
    }
    return;
  }
  // ---- consumers ----
  if (warp >= nwork) return;
  RingCtx c;
  load_flipped_taps(kernel, p, c.k);
  const int ox = cx0 + warp * 32 + lane;
  c.stok = ox < p.out_w; c.stok_b = c.stok && has_b;
  // opaque to the compiler, which would otherwise recompute the flags from the thread index at every store
  asm volatile("" : "+r"(c.stok), "+r"(c.stok_b));
  const int ix0 = cfirst + warp * 32 + lane;               // input column of the lane's tap 0
  c.tokmask = 0;
#pragma unroll
  for (int tx = 0; tx < 4; ++tx) c.tokmask |= (ix0 + tx >= 0 && ix0 + tx < p.in_w ? 1u : 0u) << tx;
  c.sbase = sbase; c.bars = bars; c.slotb = SLOTB;
  c.ngroups = ngroups; c.total = total; c.jlo = jlo; c.jhi = jhi; c.out_w = p.out_w; c.any = ncols > 0;
  const int64_t ea0 = plane * plane_in + (int64_t)iy0 * p.in_w + cbase;   // element index of the first staged sample of row 0
#pragma unroll
  for (int r = 0; r < G; ++r) {
    const int ska = (int)((ea0 + (int64_t)r * p.in_w) & 3), skb = (int)((ea0 + plane_in + (int64_t)r * p.in_w) & 3);
    c.roff[r][0] = (uint32_t)((r * 2) * ROWB + (ix0 - cbase + ska) * 4);
    c.roff[r][1] = (uint32_t)((r * 2 + 1) * ROWB + (ix0 - cbase + skb) * 4);
  }
  if (!c.any) { c.jlo = 0; c.jhi = 0; }
  float* pa = out + plane * plane_out + ((int64_t)r0 - 3) * p.out_w + min(ox, p.out_w - 1);
  float* pb = pa + (has_b ? plane_out : 0);
  const bool wedge = __any_sync(0xffffffffu, c.tokmask != 0xfu);
  if (wedge) ring_consume<true>(c, pa, pb); else ring_consume<false>(c, pa, pb);
}

// up = 2, down = 1: a lane owns two adjacent output columns, 60 per warp.  Only every other tap meets a sample of the
// zero-stuffed signal: output (oy, ox) sums 2 x 2 input samples with the taps of its parity class.
constexpr int S21_W = 60;
struct S21Ctx {
  float ka[2][2][2], kb[2][2][2];  // [row class f][column class e][first / second sample]; a = first contributing row, b = second
  int in_w, in_h, out_w, iy0, total, rows, dx, dy;
  int colok, st0, st1;
};
template <bool FAST>
__device__ __forceinline__ void s21_load(float (&d)[STREAM_U], const S21Ctx& c, const float* __restrict__ col, int jb) {
#pragma unroll
  for (int u = 0; u < STREAM_U; ++u) {
    const int iy = c.iy0 + jb + u;
    const float* ptr = col + (jb + u) * c.in_w;
    if (FAST) d[u] = __ldg(ptr);   // out-of-image columns are zeroed where the row is consumed
    else d[u] = (c.colok && jb + u < c.total && iy >= 0 && iy < c.in_h) ? __ldg(ptr) : 0.f;
  }
}
// FULL: every input row of the batch completes two existing output rows (no row predicates).  VEC: float2 stores.
template <bool FULL, bool VEC>
__device__ __forceinline__ void s21_rows(const float (&v)[STREAM_U], const S21Ctx& c, float* (&prow)[2], int jb, float (&acc)[2][2]) {
  // prow[f]: output row (of class f) completed by the batch's first input row, at the lane's column pair; input row j
  // completes row f of output pair t = j - (f ? dy : 0) - 1
#pragma unroll
  for (int u = 0; u < STREAM_U; ++u) {
    const int j = jb + u;
    const float v0 = c.colok ? v[u] : 0.f;
    const float v1 = __shfl_down_sync(0xffffffffu, v0, 1);
    const float v2 = __shfl_down_sync(0xffffffffu, v0, 2);
    float xa[2], xb[2];
    xa[0] = v0; xb[0] = v1;
    xa[1] = c.dx ? v1 : v0; xb[1] = c.dx ? v2 : v1;
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      float done[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        done[e] = fmaf(xb[e], c.kb[f][e][1], fmaf(xa[e], c.kb[f][e][0], acc[f][e]));
        acc[f][e] = fmaf(xb[e], c.ka[f][e][1], xa[e] * c.ka[f][e][0]);
      }
      bool ok = true;
      if (!FULL) {
        const int t = j - (f ? c.dy : 0) - 1;
        ok = t >= 0 && 2 * t + f < c.rows && j < c.total;
      }
      if (VEC) {
        if (ok && c.st1) *reinterpret_cast<float2*>(prow[f]) = make_float2(done[0], done[1]);
      } else {
        if (ok && c.st0) prow[f][0] = done[0];
        if (ok && c.st1) prow[f][1] = done[1];
      }
      prow[f] += 2 * c.out_w;
    }
  }
}
template <bool VEC>
__global__ void __launch_bounds__(128, 5) upfirdn2d_stream21_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                                    float* __restrict__ out, UpfirdnParams p, int rh) {
  constexpr int U = STREAM_U;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 4 + warp) * S21_W;
  if (c0 >= p.out_w) return;
  S21Ctx c;
  const int r0 = blockIdx.y * rh;
  c.rows = min(rh, p.out_h - r0);
  const int tp = (c.rows + 1) >> 1;  // output row pairs
  // column classes e = 0, 1 (output column c0 + 2*lane + e) and row classes f = 0, 1 (output row r0 + 2*t + f):
  // first contributing tap q / pr, first contributing input sample cb / rb (relative to lane / pair index)
  int q[2], cb[2], pr[2], rb[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int bx = c0 + e - p.pad_x0;
    q[e] = bx & 1; cb[e] = (bx + q[e]) >> 1;
    const int by = r0 + e - p.pad_y0;
    pr[e] = by & 1; rb[e] = (by + pr[e]) >> 1;
  }
  c.dx = cb[1] - cb[0]; c.dy = rb[1] - rb[0];  // 0 or 1
  auto tap = [&](int ty, int tx) { return (ty < p.kh && tx < p.kw) ? __ldg(kernel + (p.kh - 1 - ty) * p.kw + (p.kw - 1 - tx)) : 0.f; };
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        c.ka[f][e][h] = tap(pr[f], q[e] + 2 * h);
        c.kb[f][e][h] = tap(pr[f] + 2, q[e] + 2 * h);
      }
  c.in_w = p.in_w; c.in_h = p.in_h; c.out_w = p.out_w;
  const int ix = cb[0] + lane;
  c.colok = ix >= 0 && ix < p.in_w;
  const int ixc = min(max(ix, 0), p.in_w - 1);
  c.iy0 = rb[0];
  c.total = tp + 1 + c.dy;  // input rows feeding this row block
  const int ox = c0 + 2 * lane;
  const bool lane_ok = lane < S21_W / 2;
  c.st0 = lane_ok && ox < p.out_w; c.st1 = lane_ok && ox + 1 < p.out_w;
  asm volatile("" : "+r"(c.colok), "+r"(c.st0), "+r"(c.st1));   // keep the lane flags in registers
  const int64_t plane_in = (int64_t)p.in_h * p.in_w, plane_out = (int64_t)p.out_h * p.out_w;
  auto load_fast = [&](int jb) { return c.iy0 + jb >= 0 && c.iy0 + jb + U <= c.in_h && jb + U <= c.total; };
  auto rows_full = [&](int jb) { return jb >= 2 && jb + U <= c.total && 2 * (jb + U - 2) + 1 < c.rows; };
  for (int64_t plane = blockIdx.z; plane < p.major; plane += gridDim.z) {
    const float* col = in + plane * plane_in + (int64_t)c.iy0 * p.in_w + ixc;
    float* dst = out + plane * plane_out + (int64_t)r0 * p.out_w + min(ox, p.out_w - 2 + (p.out_w & 1));
    float* prow[2] = {dst - 2 * p.out_w, dst + (1 - 2 * (c.dy + 1)) * p.out_w};   // rows 2*(0 - d_f - 1) + f of the band
    float v[U], nx[U];
    float acc[2][2] = {};
    if (load_fast(0)) s21_load<true>(v, c, col, 0); else s21_load<false>(v, c, col, 0);
    for (int jb = 0; jb < c.total; jb += U) {
      if (load_fast(jb + U)) s21_load<true>(nx, c, col, jb + U); else s21_load<false>(nx, c, col, jb + U);
      if (rows_full(jb)) s21_rows<true, VEC>(v, c, prow, jb, acc); else s21_rows<false, VEC>(v, c, prow, jb, acc);
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = nx[u];
    }
  }
}

// up = 1, down = 2: a lane owns one output column (30 per warp) and loads an aligned pair of input columns per row.
constexpr int S12_W = 30;
struct S12Ctx {
  float k[4][4];
  int in_w, in_h, out_w, iy0, total, rows, delta;
  bool ok0, ok1, vec, stok;
};
template <bool FAST>
__device__ __forceinline__ void s12_load(float2 (&d)[STREAM_U], const S12Ctx& c, const float* __restrict__ col, int jb) {
#pragma unroll
  for (int u = 0; u < STREAM_U; ++u) {
    const int iy = c.iy0 + jb + u;
    const float* ptr = col + (jb + u) * c.in_w;
    float2 t = make_float2(0.f, 0.f);
    if (FAST || (jb + u < c.total && iy >= 0 && iy < c.in_h)) {
      if (c.vec) t = __ldg(reinterpret_cast<const float2*>(ptr));
      else { if (c.ok0) t.x = __ldg(ptr); if (c.ok1) t.y = __ldg(ptr + 1); }
    }
    d[u] = t;
  }
}
template <int MODE>  // as s11_rows
__device__ __forceinline__ void s12_rows(const float2 (&v)[STREAM_U], const S12Ctx& c, float* __restrict__ dst, int jb, float& an,
                                         float& ao) {
#pragma unroll
  for (int u = 0; u < STREAM_U; ++u) {
    const int j = jb + u;
    const float x0 = v[u].x, y0 = v[u].y;
    const float x1 = __shfl_down_sync(0xffffffffu, x0, 1);
    const float y1 = __shfl_down_sync(0xffffffffu, y0, 1);
    const float x2 = __shfl_down_sync(0xffffffffu, x0, 2);
    const float w0 = c.delta ? y0 : x0, w1 = c.delta ? x1 : y0, w2 = c.delta ? y1 : x1, w3 = c.delta ? x2 : y1;
    if ((u & 1) == 0) {  // ty = 0 of the new output row, ty = 2 of the previous one
      an = w0 * c.k[0][0]; an = fmaf(w1, c.k[0][1], an); an = fmaf(w2, c.k[0][2], an); an = fmaf(w3, c.k[0][3], an);
      ao = fmaf(w0, c.k[2][0], ao); ao = fmaf(w1, c.k[2][1], ao); ao = fmaf(w2, c.k[2][2], ao); ao = fmaf(w3, c.k[2][3], ao);
    } else {             // ty = 1 / ty = 3; the previous output row is complete
      an = fmaf(w0, c.k[1][0], an); an = fmaf(w1, c.k[1][1], an); an = fmaf(w2, c.k[1][2], an); an = fmaf(w3, c.k[1][3], an);
      ao = fmaf(w0, c.k[3][0], ao); ao = fmaf(w1, c.k[3][1], ao); ao = fmaf(w2, c.k[3][2], ao); ao = fmaf(w3, c.k[3][3], ao);
      const int t = (j >> 1) - 1;
      float* ptr = dst + t * c.out_w;
      if (c.stok && (MODE == 1 || (MODE == 2 && u >= 3) || (MODE == 0 && t >= 0 && t < c.rows))) *ptr = ao;
      ao = an;
    }
  }
}
__global__ void __launch_bounds__(128) upfirdn2d_stream12_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                                 float* __restrict__ out, UpfirdnParams p, int rh) {
  constexpr int U = STREAM_U;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 4 + warp) * S12_W;
  if (c0 >= p.out_w) return;
  S12Ctx c;
  load_flipped_taps(kernel, p, c.k);
  const int r0 = blockIdx.y * rh;
  c.rows = min(rh, p.out_h - r0);
  c.total = 2 * c.rows + 2;              // input rows feeding this row block
  const int xs = 2 * c0 - p.pad_x0;      // first input column of the warp's first output
  const int ib = xs & ~1;                // floored to even: lanes load the pairs (ib + 2*lane, ib + 2*lane + 1)
  c.delta = xs - ib;                     // 0 or 1
  const int ix = ib + 2 * lane;
  c.ok0 = ix >= 0 && ix < p.in_w; c.ok1 = ix + 1 >= 0 && ix + 1 < p.in_w;
  // a lane whose pair straddles an image edge takes the scalar loads; its clamped pair start stays even
  const bool even_rows = (p.in_w & 1) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0;
  c.vec = even_rows && c.ok0 && c.ok1;
  const bool none = !c.ok0 && !c.ok1;
  const int ixc = none ? 0 : ix;         // lanes with no valid column never dereference; keep their pointer in range
  if (none) { c.vec = false; }
  c.in_w = p.in_w; c.in_h = p.in_h; c.out_w = p.out_w;
  c.iy0 = 2 * r0 - p.pad_y0;
  const int ox = c0 + lane;
  c.stok = lane < S12_W && ox < p.out_w;
  const int64_t plane_in = (int64_t)p.in_h * p.in_w, plane_out = (int64_t)p.out_h * p.out_w;
  auto load_fast = [&](int jb) { return c.iy0 + jb >= 0 && c.iy0 + jb + U <= c.in_h && jb + U <= c.total; };
  for (int64_t plane = blockIdx.z; plane < p.major; plane += gridDim.z) {
    const float* col = in + plane * plane_in + (int64_t)c.iy0 * p.in_w + ixc;
    float* dst = out + plane * plane_out + (int64_t)r0 * p.out_w + min(ox, p.out_w - 1);
    float2 v[U], nx[U];
    float an = 0.f, ao = 0.f;
    if (load_fast(0)) s12_load<true>(v, c, col, 0); else s12_load<false>(v, c, col, 0);
    for (int jb = 0; jb < c.total; jb += U) {
      if (load_fast(jb + U)) s12_load<true>(nx, c, col, jb + U); else s12_load<false>(nx, c, col, jb + U);
      if (jb + U <= c.total) {
        if (jb == 0) s12_rows<2>(v, c, dst, jb, an, ao); else s12_rows<1>(v, c, dst, jb, an, ao);
      } else {
        s12_rows<0>(v, c, dst, jb, an, ao);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = nx[u];
    }
  }
}

// ---- small planes: minor == 1, fp32, kernel <= 4x4, many planes of at most 34 x 34 samples -----------------------------
// The ring / streaming kernels give a plane (pair) its own CTAs; on 8 x 8 ... 33 x 33 maps most lanes of a strip are idle and
// the fixed per-CTA cost (barrier set-up, halo rows) dominates: 0.10-0.17 of the HBM rate at [64, 512, 16, 16]
// (profiles/r02_op_microbench.txt).  Here a CTA stages a run of PL consecutive planes - one contiguous piece of the input - in
// shared memory in a zero-haloed layout: row pitch in_w + 3 (the three zeros after a row are its right halo and the left halo
// of the next row), three zero rows above and `bot` zero rows below every plane.  With the halo in place no tap needs a range
// test: a thread produces a vertical strip of SP_R outputs from unpredicated shared-memory reads at immediate offsets, lanes
// follow consecutive output columns (conflict-free reads, coalesced stores), work items are decoded with multiply-high
// divisions.  A first version with range predicates and run-time divisions executed ~90 instructions per output and was
// slower than the kernels it replaced on 64 px planes; this one is used up to 34 x 34.  Taps are summed in ascending (ty, tx)
// order like every other kernel of this file; taps outside the kernel, on the zero halo or on stuffed zeros contribute + 0.
constexpr int SP_R = 4;                 // output rows per thread strip
constexpr int SP_HALO = 3;              // zero columns after every staged row, zero rows above every staged plane
constexpr int SP_STAGE_FLOATS = 10240;  // staged (haloed) floats per CTA: 40 KB, several CTAs per SM overlap staging and arithmetic
constexpr int SP_MAX_PLANE = 1200;      // samples of the larger of the input / output plane (34 x 34)

struct SpDiv { uint64_t M; int s; };    // n / d = (n * M) >> (32 + s) for 0 <= n < 2^31 (M = ceil(2^(32+s) / d), s = ceil(log2 d))
static inline SpDiv sp_make_div(int d) {
  SpDiv f; f.s = 0;
  while ((1ll << f.s) < d) ++f.s;
  f.M = (uint64_t)((((unsigned __int128)1 << (32 + f.s)) + (unsigned)d - 1) / (unsigned)d);
  return f;
}
__device__ __forceinline__ int sp_div(int n, const SpDiv& f) { return (int)(((uint64_t)(uint32_t)n * f.M) >> (32 + f.s)); }

struct SpGeom {
  int PL;       // planes per CTA
  int Wp, PP;   // staged row pitch (in_w + 3) and plane pitch ((3 + in_h + bot) * Wp) in floats
  int per_plane;  // work items per plane: ceil(out_h / SP_R) * out_w
  SpDiv d_in_hw, d_in_w, d_per_plane, d_out_w;
};

// up = 1 (down = 1 or 2): rows j = 0 .. DOWN (R - 1) + 3 below the strip's first tap row feed output r with tap row ty = j - DOWN r
template <int DOWN>
__device__ __forceinline__ void sp_strip_down(const float* __restrict__ tap00, int Wp, const float (&k)[4][4], float (&acc)[SP_R]) {
  constexpr int ROWS = DOWN * (SP_R - 1) + 4;
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    float v[4];
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) v[tx] = tap00[j * Wp + tx];
#pragma unroll
    for (int r = 0; r < SP_R; ++r) {
      const int ty = j - DOWN * r;   // compile-time after unrolling
      if (ty >= 0 && ty < 4) {
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) acc[r] = fmaf(v[tx], k[ty][tx], acc[r]);
      }
    }
  }
}

// up = 2: output (oy, ox) meets samples only at the taps with (oy + ty - pad_y0) and (ox + tx - pad_x0) even: two tap rows
// ty0, ty0 + 2 and two tap columns tx0, tx0 + 2.  PY = pad_y0 & 1 (the strip starts at a multiple of 4, so ty0 = (r + PY) & 1).
template <int PY>
__device__ __forceinline__ void sp_strip_up2(const float* __restrict__ plane, int Wp, const float (&k)[4][4], int oy0, int ox, int pad_x0,
                                             int pad_y0, float (&acc)[SP_R]) {
  const int tx0 = (ox + pad_x0) & 1;
  const int ix0 = (ox + tx0 - pad_x0) >> 1;   // even numerator: exact, also below zero
  float wa[4], wb[4];
#pragma unroll
  for (int ty = 0; ty < 4; ++ty) { wa[ty] = tx0 ? k[ty][1] : k[ty][0]; wb[ty] = tx0 ? k[ty][3] : k[ty][2]; }
#pragma unroll
  for (int r = 0; r < SP_R; ++r) {
    const int ty0 = (r + PY) & 1;             // compile-time
    const int iy0 = (oy0 + r + ty0 - pad_y0) >> 1;
    const float* q = plane + iy0 * Wp + ix0;
    float a = fmaf(q[0], wa[ty0], 0.f);
    a = fmaf(q[1], wb[ty0], a);
    a = fmaf(q[Wp], wa[ty0 + 2], a);
    acc[r] = fmaf(q[Wp + 1], wb[ty0 + 2], a);
  }
}

template <int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_planes_kernel(const float* __restrict__ in, const float* __restrict__ kernel,
                                                               float* __restrict__ out, UpfirdnParams p, SpGeom g) {
  extern __shared__ __align__(16) float sp_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p0 = (int64_t)blockIdx.x * g.PL;
  const int npl = (int)(p.major - p0 < g.PL ? p.major - p0 : g.PL);
  // 1. zero the staged run (halo included).  The run starts 4 floats into the buffer: the top-left taps of the first plane
  // reach up to 3 floats before its first halo row.
  float* const run = sp_smem + 4;
  {
    float4* z = reinterpret_cast<float4*>(sp_smem);
    const int nz = ((npl * g.PP + 3) >> 2) + 1;
    for (int i = threadIdx.x; i < nz; i += 256) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float k[4][4];
  load_flipped_taps(kernel, p, k);
  __syncthreads();
  // 2. the planes of a run are one contiguous piece of global memory: 128-bit loads, eight in flight per thread (a first
  // version copied row by row, one 68-byte request per warp and round trip, and spent its time waiting: 49 us for 67 MB), then
  // every element is scattered to its place in the haloed layout (element e of the run -> plane, row, column by multiply-high)
  {
    const int in_hw = p.in_h * p.in_w;
    const int n = npl * in_hw;
    const float* src = in + p0 * in_hw;
    auto place = [&](int e, float v) {
      const int pl = sp_div(e, g.d_in_hw);
      const int rem = e - pl * in_hw;
      const int r = sp_div(rem, g.d_in_w);
      run[pl * g.PP + (r + SP_HALO) * g.Wp + (rem - r * p.in_w)] = v;
    };
    const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    const int head = n < ((4 - mis) & 3) ? n : ((4 - mis) & 3);   // scalar elements before the first aligned vector
    const int nvec = (n - head) >> 2;
    if ((int)threadIdx.x < head) place(threadIdx.x, __ldg(src + threadIdx.x));
    for (int e = head + 4 * nvec + threadIdx.x; e < n; e += 256) place(e, __ldg(src + e));
    const float4* s4 = reinterpret_cast<const float4*>(src + head);
    constexpr int U = 8;
    for (int base = 0; base < nvec; base += 256 * U) {
      float4 buf[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * 256 + threadIdx.x;
        if (i < nvec) buf[u] = ldg_stream(s4 + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * 256 + threadIdx.x;
        if (i < nvec) {
          // first element by division, the other three by stepping (row / plane wraps)
          const int e = head + 4 * i;
          const int pl = sp_div(e, g.d_in_hw);
          const int rem = e - pl * in_hw;
          int r = sp_div(rem, g.d_in_w);
          int c = rem - r * p.in_w;
          float* d = run + pl * g.PP + (r + SP_HALO) * g.Wp + c;
          const float v4[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *d = v4[q];
            ++d; ++c;
            if (c == p.in_w) {
              c = 0; d += SP_HALO; ++r;
              if (r == p.in_h) { r = 0; d += g.PP - p.in_h * g.Wp; }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // 3. strips
  const int out_hw = p.out_h * p.out_w;
  const int items = npl * g.per_plane;
  for (int it = threadIdx.x; it < items; it += 256) {
    const int pli = sp_div(it, g.d_per_plane);
    const int rem = it - pli * g.per_plane;
    const int sy = sp_div(rem, g.d_out_w), ox = rem - sy * p.out_w;
    const int oy0 = sy * SP_R;
    const float* plane = run + pli * g.PP + SP_HALO * g.Wp;   // sample (0, 0) of the staged plane
    float acc[SP_R];
#pragma unroll
    for (int r = 0; r < SP_R; ++r) acc[r] = 0.f;
    if (UP == 2) {
      if (p.pad_y0 & 1) sp_strip_up2<1>(plane, g.Wp, k, oy0, ox, p.pad_x0, p.pad_y0, acc);
      else sp_strip_up2<0>(plane, g.Wp, k, oy0, ox, p.pad_x0, p.pad_y0, acc);
    } else {
      sp_strip_down<DOWN>(plane + (DOWN * oy0 - p.pad_y0) * g.Wp + (DOWN * ox - p.pad_x0), g.Wp, k, acc);
    }
    float* o = out + (p0 + pli) * out_hw + (int64_t)oy0 * p.out_w + ox;
#pragma unroll
    for (int r = 0; r < SP_R; ++r)
      if (oy0 + r < p.out_h) o[r * p.out_w] = acc[r];
  }
}

// The zero-haloed layout serves pads of 0 .. 3 whose taps stay within three samples of the plane; anything else (crops, wide
// pads) stays with the other kernels.  Returns false when the configuration does not qualify.
static bool upfirdn_planes_geom(int up, int down, const UpfirdnParams& p, SpGeom& g) {
  if (p.pad_x0 < 0 || p.pad_x0 > SP_HALO || p.pad_y0 < 0 || p.pad_y0 > SP_HALO) return false;
  const int nstrip = (p.out_h + SP_R - 1) / SP_R;
  int max_ix, max_iy;   // last staged column / row any tap of any strip (whole strips, also past out_h) reads
  if (up == 2) {
    max_ix = (p.out_w - 1 + 3 - p.pad_x0) >> 1;
    max_iy = ((nstrip * SP_R - 1 + 3 - p.pad_y0) >> 1) + 1;   // + 1: the second tap row is read unconditionally
    max_ix += 1;
  } else {
    max_ix = down * (p.out_w - 1) + 3 - p.pad_x0;
    max_iy = down * (nstrip * SP_R - 1) + 3 - p.pad_y0;
  }
  if (max_ix > p.in_w - 1 + SP_HALO) return false;
  int bot = max_iy - (p.in_h - 1);
  if (bot < SP_HALO) bot = SP_HALO;
  if (bot > 16) return false;
  g.Wp = p.in_w + SP_HALO;
  g.PP = (SP_HALO + p.in_h + bot) * g.Wp;
  if (g.PP > SP_STAGE_FLOATS) return false;
  static const int stage = getenv("LFP_SP_STAGE") ? atoi(getenv("LFP_SP_STAGE")) : SP_STAGE_FLOATS;   // tuning switch (floats, <= 10240)
  int64_t PL = (stage < SP_STAGE_FLOATS ? stage : SP_STAGE_FLOATS) / g.PP;
  const int64_t spread = ceil_div(p.major, (int64_t)num_sms() * 4);   // enough CTAs for every SM before the runs get long
  if (PL > spread) PL = spread;
  if (PL < 1) PL = 1;
  g.PL = (int)PL;
  g.per_plane = nstrip * p.out_w;
  g.d_in_hw = sp_make_div(p.in_h * p.in_w); g.d_in_w = sp_make_div(p.in_w); g.d_per_plane = sp_make_div(g.per_plane); g.d_out_w = sp_make_div(p.out_w);
  return true;
}

static int upfirdn_planes_launch(int up, int down, const float* in, const float* kernel, float* out, const UpfirdnParams& p,
                                 const SpGeom& g, cudaStream_t s) {
  const int64_t blocks = ceil_div(p.major, g.PL);
  const size_t smem = ((size_t)g.PL * g.PP + 12) * sizeof(float);
  if (up == 2) upfirdn2d_planes_kernel<2, 1><<<(unsigned)blocks, 256, smem, s>>>(in, kernel, out, p, g);
  else if (down == 2) upfirdn2d_planes_kernel<1, 2><<<(unsigned)blocks, 256, smem, s>>>(in, kernel, out, p, g);
  else upfirdn2d_planes_kernel<1, 1><<<(unsigned)blocks, 256, smem, s>>>(in, kernel, out, p, g);
  LFP_LAUNCH_CHECK();
  return 0;
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

// rows per warp task: long enough to amortise the halo rows re-read at the top of each block, short enough that
// there are several warps' worth of tasks per SM
static int upfirdn_stream_launch(int up, int down, const float* in, const float* kernel, float* out, const UpfirdnParams& p,
                                 cudaStream_t s) {
  const int wcols = up == 2 ? S21_W : down == 2 ? S12_W : S11_W;
  const int64_t strips = ceil_div(p.out_w, wcols);
  const int64_t units = up == 1 && down == 1 ? (p.major + 1) / 2 : p.major;  // the 1:1 kernel pairs planes
  const int64_t planes = units < 65535 ? units : 65535;
  // rows per warp task (rh): the input rows feeding a block (rh + 3, rh/2 + 2, 2*rh + 2) fill whole batches of STREAM_U
  const int halo = up == 2 ? 4 : down == 2 ? 1 : 3, unit = up == 2 ? 32 : down == 2 ? 4 : 8;
  int nb = 128 / unit;
  while (nb > 2 && strips * ceil_div(p.out_h, nb * unit - halo) * planes < (int64_t)num_sms() * 64) nb >>= 1;
  const int rh = nb * unit - halo;
  dim3 grid((unsigned)ceil_div(strips, 4), (unsigned)ceil_div(p.out_h, rh), (unsigned)planes);
  if (up == 2) {
    // output column pairs start at even columns: float2 stores need even rows and an 8-byte aligned base
    if ((p.out_w & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) upfirdn2d_stream21_kernel<true><<<grid, 128, 0, s>>>(in, kernel, out, p, rh);
    else upfirdn2d_stream21_kernel<false><<<grid, 128, 0, s>>>(in, kernel, out, p, rh);
  }
  else if (down == 2) upfirdn2d_stream12_kernel<<<grid, 128, 0, s>>>(in, kernel, out, p, rh);
  else upfirdn2d_stream11_kernel<<<grid, 128, 0, s>>>(in, kernel, out, p, rh);
  LFP_LAUNCH_CHECK();
  return 0;
}

template <int NCW>
static int upfirdn_ring_launch_t(const float* in, const float* kernel, float* out, const UpfirdnParams& p, cudaStream_t s) {
  using namespace ring;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    LFP_CUDA(cudaFuncSetAttribute(upfirdn2d_ring11_kernel<NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(NCW)));
    attr_done[dev] = true;
  }
  const int64_t pairs = (p.major + 1) / 2;
  // output rows per CTA: rb + 3 input rows fill whole groups of G; 61 (16 groups) unless the map is short or the grid small
  int rb = 61;
  if (p.out_h <= 125) rb = (p.out_h + 3 + G - 1) / G * G - 3;
  const dim3 grid0((unsigned)ceil_div(p.out_w, NCW * 32), (unsigned)ceil_div(p.out_h, rb), 1);
  for (int64_t z0 = 0; z0 < pairs; z0 += 65534) {  // even chunks keep the chunk base 16-byte aligned
    dim3 grid = grid0;
    grid.z = (unsigned)(pairs - z0 < 65534 ? pairs - z0 : 65534);
    UpfirdnParams pz = p;
    pz.major = p.major - 2 * z0;
    upfirdn2d_ring11_kernel<NCW><<<grid, (NCW + 1) * 32, smem_bytes(NCW), s>>>(in + 2 * z0 * (int64_t)p.in_h * p.in_w, kernel,
                                                                              out + 2 * z0 * (int64_t)p.out_h * p.out_w, pz, rb);
  }
  LFP_LAUNCH_CHECK();
  return 0;
}
static int upfirdn_ring_launch(const float* in, const float* kernel, float* out, const UpfirdnParams& p, cudaStream_t s) {
  // segment width: the widest that does not leave most of the last segment's warps idle
  if (p.out_w > 160) return upfirdn_ring_launch_t<8>(in, kernel, out, p, s);
  if (p.out_w > 64) return upfirdn_ring_launch_t<4>(in, kernel, out, p, s);
  return upfirdn_ring_launch_t<2>(in, kernel, out, p, s);
}

int upfirdn2d_dispatch(const void* input, const void* kernel, void* out, int dtype, int64_t major,
                       int in_h, int in_w, int64_t minor, int kh, int kw, int up_x, int up_y,
                       int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                       cudaStream_t s, bool allow_tiled, bool allow_planes) {
  UpfirdnParams p;
  p.major = major; p.minor = minor; p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  LFP_TRY(lfp_upfirdn2d_out_size(in_h, in_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1,
                                 pad_y0, pad_y1, &p.out_h, &p.out_w));
  if (major == 0 || minor == 0 || p.out_h <= 0 || p.out_w <= 0) return 0;
  const bool small_fir = kh <= 4 && kw <= 4 && up_x == up_y && down_x == down_y;
  // the bulk copies take 16-byte aligned supersets of the rows: they stay inside the tensor when its base and its size are
  // 16-byte multiples
  static const bool no_ring = getenv("LFP_FIR_NO_RING") && atoi(getenv("LFP_FIR_NO_RING")) != 0;
  const bool fast = allow_tiled && dtype == LFP_F32 && minor == 1 && small_fir && p.out_h < (1 << 30) / 2;
  static const bool no_planes = getenv("LFP_FIR_NO_PLANES") && atoi(getenv("LFP_FIR_NO_PLANES")) != 0;   // A/B switch
  if (fast && allow_planes && !no_planes && ((up_x == 1 && down_x == 1) || (up_x == 2 && down_x == 1) || (up_x == 1 && down_x == 2)) &&
      (int64_t)in_h * in_w <= SP_MAX_PLANE && (int64_t)p.out_h * p.out_w <= SP_MAX_PLANE && major >= 32 && major < (1ll << 31)) {
    SpGeom g;
    if (upfirdn_planes_geom(up_x, down_x, p, g))
      return upfirdn_planes_launch(up_x, down_x, (const float*)input, (const float*)kernel, (float*)out, p, g, s);
  }
  if (fast && !no_ring && up_x == 1 && down_x == 1 && p.out_w >= 16 && p.out_h >= 8 && (reinterpret_cast<uintptr_t>(input) & 15) == 0 &&
      ((major * in_h * (int64_t)in_w) & 3) == 0 && (int64_t)p.out_h * p.out_w < (1ll << 31) && (int64_t)in_h * in_w < (1ll << 31))
    return upfirdn_ring_launch((const float*)input, (const float*)kernel, (float*)out, p, s);
  if (fast && ((up_x == 1 && down_x == 1) || (up_x == 2 && down_x == 1) || (up_x == 1 && down_x == 2)))
    return upfirdn_stream_launch(up_x, down_x, (const float*)input, (const float*)kernel, (float*)out, p, s);
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
                            down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, (cudaStream_t)stream, true, true);
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
  HostStaging& hs = host_staging();
  void* din = hs.get(0, nin * es);
  void* dout = hs.get(1, nout * es);
  void* dk = hs.get(2, (size_t)kh * kw * es);
  if (!din || !dout || !dk) { set_error("upfirdn2d_host: device allocation failed"); return LFP_ENOMEM; }
  cudaError_t e = cudaMemcpy(din, input, nin * es, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dk, kernel, (size_t)kh * kw * es, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { set_error("upfirdn2d_host: %s", cudaGetErrorString(e)); return (int)e; }
  int rc = lfp_upfirdn2d(din, dk, dout, dtype, major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, nullptr);
  if (rc == 0) {
    e = cudaMemcpy(out, dout, nout * es, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("upfirdn2d_host: %s", cudaGetErrorString(e)); rc = (int)e; }
  }
  return rc;
}
