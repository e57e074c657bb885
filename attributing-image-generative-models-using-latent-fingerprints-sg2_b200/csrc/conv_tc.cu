// Implicit-GEMM modulated convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   D[128 pixels, BN channels] (fp32, TMEM)  +=  A[128 pixels, 32 ch] (tf32, smem) * B[BN, 32 ch]^T (tf32, smem)
//
// * One work item = one 8 (x) x 16 (y) tile of output-grid pixels of one sample x one BN-wide slice of the
//   output channels.  M = 128 is the UMMA M; a pixel is row m = 8*y + x of the accumulator.  A "tall" item
//   (Args::nhalf = 2) is two such tiles stacked in y with one accumulator each: every streamed weight slice
//   feeds both, which halves the L2 -> SM weight traffic of the launches that are bound by it.
// * A: TMA loads the haloed 10 x 18 (tall: 10 x 34) window of the NHWC activation (32 channels = one 128-byte
//   row per pixel, SWIZZLE_128B, out-of-range pixels zero-filled by TMA = the conv padding) ONCE per 32-channel
//   chunk; every filter tap (dy,dx) then reads the same image through a UMMA descriptor whose start
//   address is shifted by (dy*10+dx) rows and whose 8-row-group pitch (SBO) is 10 rows.  The swizzle
//   XOR is a function of the absolute smem address, so shifted starts stay consistent with what TMA
//   wrote (verified on hardware by tools/umma_probe.cu).  No im2col, no 9x re-read.
// * B: one TMA load of the [BN, 32] K-major weight slice per (chunk, tap); the weight is shared by
//   all samples (activation-modulated algebra, src/model.py:229-256).
// * Style modulation (src/model.py:259, folded onto the activations) is applied to the A image in
//   shared memory by the four transform warps between TMA arrival and MMA issue, together with the
//   round-to-nearest tf32 conversion; demodulation, noise, bias and leaky-ReLU are the epilogue
//   (src/model.py:261-263, 316, src/op/fused_act.py:110-127).
// * Persistent CTAs (one per SM) walk the (sample, tile, n-tile) list with a stride of gridDim.x.
//   Roles: warp 0 = TMA producer (runs ahead across work items), warp 1 = MMA issuer (warp-uniform control
//   flow, one elected lane, issue loop unrolled), warps 2-5 = A transform (a third epilogue set in the
//   unmodulated kernels), warps 6-9 / 10-13 = epilogue sets that alternate (virtual) tiles (TMEM ->
//   registers -> global); the modulated layout pads to 512 threads (register allocation granularity).
//   The accumulator has 1-4 TMEM stages, so the epilogue of item i overlaps the MMAs of items i+1...;
//   an epilogue set waits on mbarrier parities, which is only unambiguous while it steps at most nacc work
//   items (and at most XS ring stages) at a time - see tc_launch.
// * mbarrier rings: A (2-8 stages); B either streamed (3-12 stages) or, when the whole [taps, K, BN]
//   weight slice fits in shared memory (the C <= 64 layers), loaded once and kept resident for the
//   CTA's lifetime; XS = per-item ring of what the epilogue needs: saved forward input tiles of the
//   data-gradient epilogues (N <= 64), per-pixel noise / skip-gradient tiles and the per-channel
//   demodulation / bias / modulation vectors of the slice (PXS kernels).
// * Epilogues: tcgen05.ld.16x256b read-out for every slice width (thread = 4 pixels x 8 channels,
//   tools/tmem_layout_probe.cu; the 32x32b variant of round 1 is kept behind LFP_TC_E2=0).  EPI_ACT: demod,
//   noise, bias, lrelu (optionally the ToRGB dot product); EPI_STORE: raw, or the four sub-pixel phases
//   of the transposed conv from four accumulators; EPI_DGRAD(_ACT): x s, style-gradient pixel sums and,
//   fused, the backward through noise / bias / lrelu / ToRGB of the layer below; EPI_RELU / EPI_DGRAD_RELU
//   for the VGG16 backbone of the perceptual loss.
// * Measured floors (profiles/r02_umma_rate.md, r02_conv_l2_traffic.md): an M = 128, K = 8 tf32 MMA costs 52.4
//   cycles for every N <= 64 (N / 2 above); streamed weights run into the ~6300 B/clk L2 -> SM throughput cap.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include "synth_kernels.cuh"

namespace lfp {
namespace tc {

constexpr int TILE_W = 8, TILE_H = 16, HALO_W = TILE_W + 2;   // the haloed window is HALO_W x (TILE_H * nhalf + 2) pixels
// one haloed activation stage: 10 x 18 = 180 rows of 128 bytes (23040 B, stride 23552) for a 16-row tile, 10 x 34 = 340 rows
// (43520 B, stride 44032) for a tall one - Args::a_rows / a_stage
constexpr int MAX_SA = 8, MAX_SB = 12;
// warp 0 TMA, warp 1 MMA, warps 2-5 transform (or a third epilogue set when there is nothing to transform),
// warps 6-9 / 10-13 epilogue sets; warps 14-15 of the modulated (forward) kernels only pad the block to 512 threads
constexpr int NTHREADS_MOD = 512, NTHREADS_PLAIN = 448;   // both are charged as 512 threads: 128 registers
constexpr int PAD_WARP = 14;                               // first padding warp of the modulated layout
constexpr int MAX_ACC = 4;                // TMEM accumulator stages
constexpr int MAX_XS = 4;                 // stages of the saved-input (xsave) tile ring of the data-gradient epilogues
constexpr int XS_CHUNK = 128 * 128;       // one 128-pixel x 32-channel tile
constexpr int NBARS = 3 * MAX_SA + 2 * MAX_SB + 1 + 2 * MAX_ACC + 2 * MAX_XS;
constexpr int SMEM_OPTIN = 232448;        // 227 KB: static + dynamic shared memory available to one CTA
// static shared memory of the kernel (barriers, and for the data-gradient epilogue the reduction scratch), rounded up
constexpr int STATIC_SMEM_RESERVE = 1024;
constexpr int EPI_NRED(int epi) { return epi == EPI_DGRAD_ACT ? 3 : (epi == EPI_DGRAD ? 1 : 0); }
// per-warp 32x33 transpose scratch + per-set [kind][quarter][BN] column sums
__host__ __device__ constexpr int EPI_SCR(bool e2, int nsets) { return e2 ? 0 : nsets * 4 * 32 * 33; }   // floats; the 16x256b epilogue reduces with shuffles
constexpr int EPI_SMEM(int epi, int nsets, int bn, bool e2) { return EPI_NRED(epi) == 0 ? 0 : (EPI_SCR(e2, nsets) + nsets * EPI_NRED(epi) * 4 * bn) * 4; }

// Division by a launch constant as a multiply-high: q = (n * M) >> (32 + s), M = ceil(2^(32+s) / d), s = ceil(log2 d), exact
// for n < 2^31.  decode() runs per tile in every warp role; ncu attributed 13 % of all stall samples of the 32 -> 32 layer to
// the I2F / RCP / fix-up chains of its three runtime integer divisions.
struct FastDiv { uint64_t M; int s; int d; };
static inline FastDiv make_fastdiv(int d) {
  FastDiv f; f.d = d; f.s = 0;
  while ((1ll << f.s) < d) ++f.s;
  f.M = (uint64_t)((((unsigned __int128)1 << (32 + f.s)) + (unsigned)d - 1) / (unsigned)d);
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return f.M == 0 ? n / f.d : (int)(((uint64_t)(uint32_t)n * f.M) >> (32 + f.s)); }

struct Args {
  int batch, gh, gw, tiles_x, tiles_y, K, N, BN;
  FastDiv d_ntiles, d_tiles_per, d_tiles_x;
  int n_ntiles, total_work;   // work item = (sample, tile, n-tile)
  // Tall tiles (nhalf = 2): a work item covers 8 x 32 output pixels = two stacked 128-pixel halves that share every streamed
  // weight slice (two accumulators, two MMAs per tap and K-step reading the same B tile).  The layers whose weights do not fit
  // in shared memory re-read them from L2 for every tile and ran at 80-110 % of the L2 -> SM throughput cap with the tensor
  // pipe 50-69 % busy (tools/umma_rate.cu, profiles/r02_conv_l2_traffic.md); a tall tile halves that traffic.
  int nhalf;                  // 1 or 2 halves of 16 rows per work item
  int a_rows, a_stage;        // rows of one haloed activation stage (10 x (16 nhalf + 2)) and its stride in bytes (multiple of 1024)
  int vtiles_per;             // 16-row tiles per sample (tiles_x * ceil(gh / 16)): index space of the per-tile partial sums
  int SA, SB, b_resident;     // A stages; B stages (streaming) or 0 with the whole weight slice resident
  int nacc;                   // TMEM accumulator stages (1, 2 or 4)
  int epi_off;                // byte offset of the epilogue scratch in dynamic shared memory
  int XS, xs_off;             // xsave ring: stages (0 = epilogue reads xsave from global) and byte offset in dynamic smem
  int xs_has_x;               // the ring stages carry the saved-input chunks (0: per-pixel scalars only, saved input from global)
  int px_ok;                  // host: tensor maps for the per-pixel scalars exist (PXS kernel when a ring is planned)
  int xs_stride;              // bytes per ring stage: BN / 32 saved-input chunks (+ 2 KB of per-pixel scalars, PXS kernels)
  int xs_bcast;
  int bbatch;                 // streamed weight slices issued per elected region (3 when the B ring holds >= 6 slices, else 1)
  int nsets;                  // epilogue warp sets (2, or 3 when the transform warps are free and smem allows)
  int in_bcast;
  TcTaps taps;
  float* out;
  int out_planes, out_plane, out_h, out_w, out_stride, out_oy, out_ox;
  const float* mod;
  ConvEpiArgs e;
};

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
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// plain 1-D bulk copy global -> shared (size and both addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"((uint64_t)src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (tcgen05), see tools/umma_probe.cu
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// descriptor = {lo, hi}: lo = (addr >> 4) | LBO field, hi = SBO | version | SWIZZLE_128B (constant per operand)
__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ float lrelu(float v) { return (v > 0.f ? v : v * kLreluSlope) * kLreluGain; }

struct Work { int b, tile, nt, y0, x0, n0, ty, tx; };
__device__ __forceinline__ Work decode(const Args& a, int w) {
  Work r;
  const int bt = fdiv(w, a.d_ntiles);
  r.nt = w - bt * a.n_ntiles;
  const int tiles_per = a.tiles_x * a.tiles_y;
  r.b = fdiv(bt, a.d_tiles_per);
  r.tile = bt - r.b * tiles_per;
  const int ty = fdiv(r.tile, a.d_tiles_x);
  r.ty = ty;
  r.tx = r.tile - ty * a.tiles_x;
  r.y0 = ty * (TILE_H * a.nhalf);
  r.x0 = r.tx * TILE_W;
  r.n0 = r.nt * a.BN;
  return r;
}

// Persistent: CTA c processes work items c, c + gridDim.x, ...  (consecutive CTAs work on neighbouring
// tiles at the same time, so halo rows and weight slices are L2 hits).
// PXS (EPI_DGRAD_ACT, 16x256b epilogue, saved-input ring present): the per-pixel scalars of the fused act-backward - the
// noise map and the three planes of the skip-image gradient - arrive in the ring stage by TMA next to the saved input,
// instead of being prefetched into registers one tile ahead.  ncu on the stride-2 32 -> 64 layer at 1024 px: 22 % of all
// stall samples sat on those prefetch loads (the compiler merges the "next" and "current" registers, which collapses the
// prefetch distance to zero), and the prefetch registers are what pushed the RGB variant into 96 bytes of spills.
template <int EPI, bool MOD, bool RES, bool E2, bool RGB, bool PXS>
// Register budget: the register file is handed out to a CTA in units of four warps, so a 576-thread block is charged as
// 640 threads and gets 96 registers per thread.  At 96 the epilogues spilled to local memory - and with (almost) all of
// the SM's memory carved out as shared memory there is no L1 behind a spill: every STL / LDL was an L2 round trip (ncu:
// 30 % of the epilogue warps' stall samples on the 32 -> 32 layer at 1024 px were long-scoreboard waits on them, and
// the spill store of a prefetched noise value waited for the very load it was meant to overlap).  Both layouts therefore
// stay at <= 512 threads = 128 registers: the modulated kernels run two epilogue sets instead of three.
__global__ void __launch_bounds__(MOD ? NTHREADS_MOD : NTHREADS_PLAIN, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB,
                                                             const __grid_constant__ CUtensorMap tmX,
                                                             const __grid_constant__ CUtensorMap tmNz,
                                                             const __grid_constant__ CUtensorMap tmRg, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[NBARS];
  __shared__ uint32_t tmem_base_s;
  constexpr bool DG = EPI == EPI_DGRAD || EPI == EPI_DGRAD_ACT;     // style-gradient reductions
  constexpr bool DGX = DG || EPI == EPI_DGRAD_RELU;                  // epilogue reads the layer's saved forward input
  constexpr int NRED = EPI == EPI_DGRAD_ACT ? 3 : 1;   // column reductions per tile


  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const smem_al = smem_raw + (smem0 - smem_u32(smem_raw));
  // data-gradient epilogues: per-warp transpose scratch (8 x 32 x 33 floats) and per-set column sums, behind the A/B rings
  float* const s_scr = reinterpret_cast<float*>(smem_al + a.epi_off);
  float* const s_red = s_scr + EPI_SCR(E2, a.nsets);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int SA = a.SA, SB = a.SB;
  const uint32_t b_slice_bytes = (uint32_t)a.BN * 128u;
  const uint32_t a_base = smem0, b_base = smem0 + SA * a.a_stage;
  const uint32_t bar0 = smem_u32(bars);
  auto bar_a_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_a_ready = [&](int s) { return bar0 + 8u * (MAX_SA + s); };
  auto bar_a_empty = [&](int s) { return bar0 + 8u * (2 * MAX_SA + s); };
  auto bar_b_full = [&](int s) { return bar0 + 8u * (3 * MAX_SA + s); };
  auto bar_b_empty = [&](int s) { return bar0 + 8u * (3 * MAX_SA + MAX_SB + s); };
  const uint32_t bar_b_all = bar0 + 8u * (3 * MAX_SA + 2 * MAX_SB);
  auto bar_acc_full = [&](int s) { return bar0 + 8u * (3 * MAX_SA + 2 * MAX_SB + 1 + s); };
  auto bar_acc_empty = [&](int s) { return bar0 + 8u * (3 * MAX_SA + 2 * MAX_SB + 1 + MAX_ACC + s); };
  auto bar_xs_full = [&](int s) { return bar0 + 8u * (3 * MAX_SA + 2 * MAX_SB + 1 + 2 * MAX_ACC + s); };
  auto bar_xs_empty = [&](int s) { return bar0 + 8u * (3 * MAX_SA + 2 * MAX_SB + 1 + 2 * MAX_ACC + MAX_XS + s); };

  if (tid == 0) {
    for (int s = 0; s < MAX_SA; ++s) { mbar_init(bar_a_full(s), 1); mbar_init(bar_a_ready(s), 128); mbar_init(bar_a_empty(s), 1); }
    for (int s = 0; s < MAX_SB; ++s) { mbar_init(bar_b_full(s), 1); mbar_init(bar_b_empty(s), 1); }
    mbar_init(bar_b_all, 1);
    for (int s = 0; s < MAX_ACC; ++s) { mbar_init(bar_acc_full(s), 1); mbar_init(bar_acc_empty(s), 128 * a.nhalf); }
    for (int s = 0; s < MAX_XS; ++s) { mbar_init(bar_xs_full(s), 1); mbar_init(bar_xs_empty(s), 128 * a.nhalf); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem_cols = (uint32_t)(a.nacc * a.BN * a.taps.nphase * a.nhalf);
  const int acc_shift = a.nacc == 4 ? 2 : (a.nacc == 2 ? 1 : 0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int kchunks = a.K >> 5;
  const int ngroups = a.taps.ngroups;
  const int ntaps = a.taps.group_tap0[ngroups];
  const int tiles_per = a.tiles_x * a.tiles_y;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      if (RES) {
        // whole weight slice of this launch: slot (tap t, chunk kc) at b_base + (t*kchunks + kc) * slice
        mbar_expect_tx(bar_b_all, (uint32_t)(ntaps * kchunks) * b_slice_bytes);
        for (int t = 0; t < ntaps; ++t)
          for (int kc = 0; kc < kchunks; ++kc)
            tma_load_2d(b_base + (uint32_t)(t * kchunks + kc) * b_slice_bytes, &tmB, bar_b_all, kc * 32, (int)a.taps.widx[t] * a.N);
      }
      // (A second MMA-issuing warp with its own stage rings was tried - forward convolutions 7.6 -> 8.6 ms at 1024 px, B = 20:
      // each issuer gets half of the activation stages and the shallower prefetch costs more than the overlapped per-tile
      // issue overhead gains - and removed; one issuer keeps every value of the issue loop warp-uniform.)
      int sa = 0, sb = 0, sx = 0;
      uint32_t pa = 0, pb = 0, px = 0;
      int it = 0;
      for (int w = blockIdx.x; w < a.total_work; w += gridDim.x, ++it) {
        const Work wk = decode(a, w);
        const int bin = a.in_bcast ? 0 : wk.b;
        for (int kc = 0; kc < kchunks; ++kc)
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(bar_a_empty(sa), pa ^ 1u);
            mbar_expect_tx(bar_a_full(sa), (uint32_t)a.a_rows * 128u);
            tma_load_5d(a_base + sa * a.a_stage, &tmA, bar_a_full(sa), kc * 32, wk.x0 - 1, wk.y0 - 1, a.taps.group_plane[g], bin);
            if (!RES)
              for (int t = a.taps.group_tap0[g]; t < a.taps.group_tap0[g + 1]; ++t) {
                mbar_wait(bar_b_empty(sb), pb ^ 1u);
                mbar_expect_tx(bar_b_full(sb), b_slice_bytes);
                tma_load_2d(b_base + sb * b_slice_bytes, &tmB, bar_b_full(sb), kc * 32, (int)a.taps.widx[t] * a.N + wk.n0);
                if (++sb == SB) { sb = 0; pb ^= 1u; }
              }
            if (++sa == SA) { sa = 0; pa ^= 1u; }
          }
        if ((DGX || PXS) && a.XS > 0) {
          // saved forward input of this tile for the epilogue (data-gradient kernels) and / or its per-pixel scalars (PXS) (128 px x BN channels, one 16 KB box per 32 channels)
          const int nch = (DGX && a.xs_has_x) ? a.BN >> 5 : 0;
          const uint32_t st0 = smem0 + a.xs_off + (uint32_t)sx * (uint32_t)a.xs_stride;
          mbar_wait(bar_xs_empty(sx), px ^ 1u);
          const uint32_t vbytes = (uint32_t)a.BN * 4u;   // one per-channel vector of this CTA's slice
          const uint32_t nzb = 512u * (uint32_t)a.nhalf;   // one noise tile: [16 nhalf rows][8] floats
          mbar_expect_tx(bar_xs_full(sx), (uint32_t)nch * XS_CHUNK + (PXS ? nzb + (RGB ? 3u * nzb : 0u) + (DG ? 3u : 2u) * vbytes : 0u));
          for (int c = 0; c < nch; ++c)
            tma_load_5d(st0 + (uint32_t)c * XS_CHUNK, &tmX, bar_xs_full(sx), wk.n0 + c * 32, wk.x0, wk.y0, 0, a.xs_bcast ? 0 : wk.b);
          if (PXS) {
            // noise tile [16 rows][8] floats, then the skip-gradient planes [3][16][8]; pixels outside the map read as 0
            tma_load_3d(st0 + (uint32_t)nch * XS_CHUNK, &tmNz, bar_xs_full(sx), wk.x0, wk.y0, a.e.noise_bstride == 0 ? 0 : wk.b);
            if (RGB) tma_load_4d(st0 + (uint32_t)nch * XS_CHUNK + nzb, &tmRg, bar_xs_full(sx), wk.x0, wk.y0, 0, wk.b);
            // per-channel vectors of this (sample, slice): demodulation, bias, and (data gradient) the layer's modulation
            const uint32_t vb = st0 + (uint32_t)nch * XS_CHUNK + 4u * nzb;
            bulk_load(vb, a.e.demod + (int64_t)wk.b * a.N + wk.n0, vbytes, bar_xs_full(sx));
            bulk_load(vb + vbytes, a.e.bias + wk.n0, vbytes, bar_xs_full(sx));
            if (DG) bulk_load(vb + 2u * vbytes, a.e.mod_out + (int64_t)wk.b * a.N + wk.n0, vbytes, bar_xs_full(sx));
          }
          if (++sx == a.XS) { sx = 0; px ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // The whole warp runs the (warp-uniform) control flow; one elected lane issues tcgen05.mma / commit.
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = (uint32_t)((HALO_W * 128) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (a_base >> 4) | 0x10000u, b_lo0 = (b_base >> 4) | 0x10000u;
    const uint32_t b_slice16 = b_slice_bytes >> 4;
    // per-tap descriptor offsets (16-byte units), kept in registers: the issue loop below is fully unrolled
    uint32_t tap_a[9], tap_b[9], tap_p[9];
    const int nphase = a.taps.nphase;
    const uint32_t stage_cols = (uint32_t)(a.nhalf * nphase * a.BN);   // accumulators of one work item: [half][phase][BN]
    const uint32_t half_a = (uint32_t)(TILE_H * HALO_W * 128) >> 4;     // descriptor offset of the lower half's window (16 rows down)
    const uint32_t half_c = (uint32_t)(nphase * a.BN);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int tt = t < ntaps ? t : 0;
      tap_a[t] = (uint32_t)((((int)a.taps.dy[tt] + 1) * HALO_W + ((int)a.taps.dx[tt] + 1)) * 8);
      tap_b[t] = b_lo0 + (uint32_t)(tt * kchunks) * b_slice16;
      tap_p[t] = nphase > 1 ? (uint32_t)a.taps.acc[tt] : 0u;
    }
    if (RES) mbar_wait(bar_b_all, 0);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int w = blockIdx.x; w < a.total_work; w += gridDim.x, ++it) {
      const int as = it & (a.nacc - 1);
      mbar_wait(bar_acc_empty(as), (((uint32_t)it >> acc_shift) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc0 = tmem + (uint32_t)as * stage_cols;
      uint32_t started = 0;   // bit p: accumulator p of this tile has received its first MMA
      for (int kc = 0; kc < kchunks; ++kc)
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait(MOD ? bar_a_ready(sa) : bar_a_full(sa), pa);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_stage_lo = a_lo0 + (uint32_t)sa * ((uint32_t)a.a_stage >> 4);
          const int t0 = a.taps.group_tap0[g], t1 = a.taps.group_tap0[g + 1];
          if (RES) {
            // weights resident: every MMA of this activation stage is issued in one elected region
            const uint32_t kc_off = (uint32_t)kc * b_slice16;
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < 9; ++t)
                if (t >= t0 && t < t1) {
                  const uint32_t a_lo = a_stage_lo + tap_a[t], b_lo = tap_b[t] + kc_off;   // (resident weights: never a tall item)
                  const uint32_t tacc = tacc0 + tap_p[t] * (uint32_t)a.BN;
                  umma_tf32_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, (started >> tap_p[t]) & 1u);
                  umma_tf32_lohi(tacc, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                  umma_tf32_lohi(tacc, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                  umma_tf32_lohi(tacc, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                  started |= 1u << tap_p[t];
                }
              umma_commit(bar_a_empty(sa));
            }
            __syncwarp();
            started = (1u << nphase) - 1u;   // every accumulator has taps in every activation stage
          } else {
            // Weights stream through the B ring.  The taps of this activation stage are issued in batches of up to three
            // slices - one barrier-wait sequence, one elected region and one round of uniform-register set-up per batch.
            // ncu on the 64 -> 64 data gradient at 512 px (one slice per region): the issuing warp was busy ~80 % of the
            // time at ~60 SASS instructions per slice (R2UR / UMOV descriptor set-up, elect, reconvergence) for four MMAs
            // that occupy the tensor pipe for 4 x 52 cycles, i.e. the launch was bound by instruction issue, not by the pipe.
            // (Needs a ring of >= 6 slices so that the next batch loads while this one computes; otherwise one slice per region.)
            if (a.bbatch == 3) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int lo = t0 > 3 * j ? t0 : 3 * j, hi = t1 < 3 * j + 3 ? t1 : 3 * j + 3;   // active taps [lo, hi) of batch j
              if (lo < hi) {
                int sbi = sb;
                uint32_t pbi = pb;
                for (int i = lo; i < hi; ++i) {
                  mbar_wait(bar_b_full(sbi), pbi);
                  if (++sbi == SB) { sbi = 0; pbi ^= 1u; }
                }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                  int sbk = sb;
                  uint32_t st = started;
#pragma unroll
                  for (int i = 0; i < 3; ++i) {
                    const int t = 3 * j + i;
                    if (t >= lo && t < hi) {
                      const uint32_t b_lo = b_lo0 + (uint32_t)sbk * b_slice16;
                      const uint32_t a_lo = a_stage_lo + tap_a[t];
                      const uint32_t tacc = tacc0 + tap_p[t] * (uint32_t)a.BN;
                      umma_tf32_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, (st >> tap_p[t]) & 1u);
                      umma_tf32_lohi(tacc, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                      umma_tf32_lohi(tacc, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                      umma_tf32_lohi(tacc, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                      if (a.nhalf == 2) {   // lower half: the window 16 rows down, its own accumulator, the same weight slice
                        umma_tf32_lohi(tacc + half_c, a_lo + half_a, a_hi, b_lo, b_hi, idesc, (st >> tap_p[t]) & 1u);
                        umma_tf32_lohi(tacc + half_c, a_lo + half_a + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                        umma_tf32_lohi(tacc + half_c, a_lo + half_a + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                        umma_tf32_lohi(tacc + half_c, a_lo + half_a + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                      }
                      umma_commit(bar_b_empty(sbk));
                      st |= 1u << tap_p[t];
                      if (++sbk == SB) sbk = 0;
                    }
                  }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 3; ++i)
                  if (3 * j + i >= lo && 3 * j + i < hi) started |= 1u << tap_p[3 * j + i];
                sb = sbi; pb = pbi;
              }
            }
            } else {
#pragma unroll
            for (int t = 0; t < 9; ++t)
              if (t >= t0 && t < t1) {
                mbar_wait(bar_b_full(sb), pb);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_slice16;
                const uint32_t accumulate = (started >> tap_p[t]) & 1u;
                if (elect_one()) {
                  const uint32_t a_lo = a_stage_lo + tap_a[t];
                  const uint32_t tacc = tacc0 + tap_p[t] * (uint32_t)a.BN;
                  umma_tf32_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
                  umma_tf32_lohi(tacc, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                  umma_tf32_lohi(tacc, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                  umma_tf32_lohi(tacc, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                  if (a.nhalf == 2) {
                    umma_tf32_lohi(tacc + half_c, a_lo + half_a, a_hi, b_lo, b_hi, idesc, accumulate);
                    umma_tf32_lohi(tacc + half_c, a_lo + half_a + 2, a_hi, b_lo + 2, b_hi, idesc, 1u);
                    umma_tf32_lohi(tacc + half_c, a_lo + half_a + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
                    umma_tf32_lohi(tacc + half_c, a_lo + half_a + 6, a_hi, b_lo + 6, b_hi, idesc, 1u);
                  }
                  umma_commit(bar_b_empty(sb));
                }
                __syncwarp();
                started |= 1u << tap_p[t];
                if (++sb == SB) { sb = 0; pb ^= 1u; }
              }
            }
            if (elect_one()) umma_commit(bar_a_empty(sa));
            __syncwarp();
          }
          if (++sa == SA) { sa = 0; pa ^= 1u; }
        }
      if (elect_one()) umma_commit(bar_acc_full(as));
      __syncwarp();
    }
  } else if (warp < 6 && (MOD || a.nsets < 3)) {
    if (MOD) {
      // ---------------- A transform: x * s[b, k], rounded to tf32 ----------------
      const int et = tid - 64;  // 0..127
      int sa = 0;
      uint32_t pa = 0;
      int it = 0;
      for (int w = blockIdx.x; w < a.total_work; w += gridDim.x, ++it) {
        const Work wk = decode(a, w);
        const float* sm = a.mod + (int64_t)wk.b * a.K;
        for (int kc = 0; kc < kchunks; ++kc)
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(bar_a_full(sa), pa);
            // thread -> fixed 16-byte column `pos` of rows r0, r0+16, ...: (row & 7) is then constant, and so is
            // the channel quad this thread scales (SWIZZLE_128B stores quad c of row r at position c ^ (r & 7))
            const int pos = et & 7, r0 = et >> 3;
            const int ch = (pos ^ (r0 & 7)) << 2;
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(sm + kc * 32 + ch));
            // twelve rows (r0 + 16 i) per pass.  The plain 180-row stage is one pass with compile-time bounds (rows 176 + r0 exist
            // for r0 < 4 only): a version with run-time bounds for both stage heights cost the modulated streaming launches
            // 10-30 % (64 x 64 px conv 488 -> 531 us, the one-tap phase launches 80 -> 107 us).
            if (a.nhalf == 1) {
              float4* p = reinterpret_cast<float4*>(smem_al + sa * a.a_stage + r0 * 128 + pos * 16);
              float4 v[12];
#pragma unroll
              for (int i = 0; i < 12; ++i)
                if (i < 11 || r0 + 176 < 180) v[i] = p[i * 128];   // row r0 + 16 i, 128 float4 apart
#pragma unroll
              for (int i = 0; i < 12; ++i)
                if (i < 11 || r0 + 176 < 180) {
                  v[i].x = to_tf32(v[i].x * s4.x);
                  v[i].y = to_tf32(v[i].y * s4.y);
                  v[i].z = to_tf32(v[i].z * s4.z);
                  v[i].w = to_tf32(v[i].w * s4.w);
                  p[i * 128] = v[i];
                }
            } else {
              // tall stage, 340 rows: rows 0 .. 191 in full, then rows 192 + r0 + 16 i < 340 (i < 9, or i = 9 for r0 < 4)
#pragma unroll
              for (int ps = 0; ps < 2; ++ps) {
                float4* p = reinterpret_cast<float4*>(smem_al + sa * a.a_stage + (ps * 192 + r0) * 128 + pos * 16);
                float4 v[12];
#pragma unroll
                for (int i = 0; i < 12; ++i)
                  if (ps == 0 || i < 9 || (i == 9 && r0 < 4)) v[i] = p[i * 128];
#pragma unroll
                for (int i = 0; i < 12; ++i)
                  if (ps == 0 || i < 9 || (i == 9 && r0 < 4)) {
                    v[i].x = to_tf32(v[i].x * s4.x);
                    v[i].y = to_tf32(v[i].y * s4.y);
                    v[i].z = to_tf32(v[i].z * s4.z);
                    v[i].w = to_tf32(v[i].w * s4.w);
                    p[i * 128] = v[i];
                  }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_a_ready(sa));
            if (++sa == SA) { sa = 0; pa ^= 1u; }
          }
      }
    }
  } else if (warp >= PAD_WARP) {
    // padding warp of the 512-thread modulated layout: nothing to do
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    // two sets of four warps alternate tiles, so one tile's operand-load / store latency overlaps the next tile
    const int eset = warp >= 14 ? 2 : (warp >= 10 ? 1 : (warp >= 6 ? 0 : 2));
    const int et = (tid - 64) & 127;        // 0..127 within the set
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int m = q * 32 + lane;
    float* scr = s_scr + (eset * 4 + q) * (32 * 33);
    const int rs = 4 * a.BN;                // stride between reduction kinds
    float* red = s_red + eset * (NRED * rs);
    if constexpr (E2) {
      // ---- BN <= 64: accumulator read with tcgen05.ld.16x256b (layout verified by tools/tmem_layout_probe.cu) ----
      // Thread t of quarter q owns pixel column x = t >> 2 of tile rows 4q .. 4q+3 and, per 32-channel chunk, the eight
      // channels 8k + 2(t & 3) + e (k < 4, e < 2): per-channel vectors are eight values per thread, the pixel-sums of
      // the data-gradient epilogues are in-thread adds over four rows plus three xor-shuffles, and global traffic stays
      // in full 32-byte sectors (four lanes x 8 bytes per pixel).
      const int x = lane >> 2, cq = lane & 3;
      const bool need_nz2 = EPI == EPI_ACT || EPI == EPI_DGRAD_ACT;
      constexpr bool has_rgb2 = EPI == EPI_DGRAD_ACT && RGB;   // compile-time: the skip-gradient registers exist only where used
      const bool do_rgb2 = EPI == EPI_ACT && a.e.rgb_out != nullptr;
      const float nw2 = need_nz2 ? __ldg(a.e.noise_w) : 0.f;
      const int64_t hw = (int64_t)a.gh * a.gw;
      // per-pixel scalars of a work item (noise; skip-image gradient), fetched one work item ahead.  Plain register
      // arrays filled by a macro: a struct handed to a lambda ended up in local memory, and the spill store then waited
      // for the load it was meant to overlap.
#define LFP_FETCH2(VT, NZ, RG)                                                                        \
  {                                                                                                   \
    _Pragma("unroll") for (int r_ = 0; r_ < 4; ++r_) { NZ[r_] = 0.f; RG[r_][0] = RG[r_][1] = RG[r_][2] = 0.f; } \
    const int64_t w_ = (int64_t)blockIdx.x + (int64_t)((VT) >> (a.nhalf - 1)) * gridDim.x;           \
    if ((VT) < 0x40000000 && w_ < a.total_work) {                                                     \
      const Work k_ = decode(a, (int)w_);                                                             \
      const int gx_ = k_.x0 + x;                                                                      \
      _Pragma("unroll") for (int r_ = 0; r_ < 4; ++r_) {                                              \
        const int gy_ = k_.y0 + TILE_H * ((VT) & (a.nhalf - 1)) + 4 * q + r_;                         \
        if (gy_ < a.gh && gx_ < a.gw) {                                                               \
          const int pix_ = gy_ * a.gw + gx_;                                                          \
          if (need_nz2) NZ[r_] = __ldg(a.e.noise + (int64_t)k_.b * a.e.noise_bstride + pix_);         \
          if (has_rgb2) {                                                                             \
            RG[r_][0] = __ldg(a.e.drgb + ((int64_t)k_.b * 3 + 0) * hw + pix_);                        \
            RG[r_][1] = __ldg(a.e.drgb + ((int64_t)k_.b * 3 + 1) * hw + pix_);                        \
            RG[r_][2] = __ldg(a.e.drgb + ((int64_t)k_.b * 3 + 2) * hw + pix_);                        \
          }                                                                                           \
        }                                                                                             \
      }                                                                                               \
    }                                                                                                 \
  }
      constexpr float G = kLreluGain, GS = kLreluGain * kLreluSlope, IG = 1.f / kLreluGain, IGS = 1.f / (kLreluGain * kLreluSlope);
      const int nchunk = a.BN >> 5;
      const int nphase = a.taps.nphase;
      float nz_n[4], rg_n[4][3];
      if (!PXS) LFP_FETCH2(eset < a.nsets ? eset : 0x40000000, nz_n, rg_n)
      // Virtual tiles: half hf of the CTA's it-th work item is virtual tile vt = it * nhalf + hf; set e takes vt = e, e + nsets, ...
      // (nhalf = 1: vt = it, the plain alternation of work items).  Both halves of an item arrive on its accumulator / ring
      // "empty" barriers (initialised to 128 * nhalf arrivals).
      const int hshift = a.nhalf - 1;   // log2(nhalf), also the mask of hf
      for (int vt = eset < a.nsets ? eset : 0x40000000; (vt >> hshift) < 0x20000000; vt += a.nsets) {
        const int it = vt >> hshift, hf = vt & hshift;
        const int64_t w64 = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        if (w64 >= a.total_work) break;
        const int w = (int)w64;
        const Work wk = decode(a, w);
        const int as = it & (a.nacc - 1);
        const int b = wk.b, n0 = wk.n0;
        const int gx = wk.x0 + x;
        const int y0h = wk.y0 + TILE_H * hf;   // first row of this half
        float cnz[4], crg[4][3];
        if (!PXS) {
#pragma unroll
          for (int r = 0; r < 4; ++r) { cnz[r] = nw2 * nz_n[r]; crg[r][0] = rg_n[r][0]; crg[r][1] = rg_n[r][1]; crg[r][2] = rg_n[r][2]; }
          LFP_FETCH2(vt + a.nsets, nz_n, rg_n)
        }
        // Addresses are built once per tile (row pointers) and once per 32-channel chunk; inside the unrolled (k, row) body every
        // load / store is base register + immediate.  ncu on the 32 -> 32 act-backward launch at 1024 px: the epilogue sets were
        // busy > 90 % of the time at ~1140 warp instructions per tile, a third of them 64-bit address arithmetic, null-pointer
        // tests and the generic -> shared conversion repeated for every saved-input load; 17 % of their stall samples were
        // instruction-fetch misses of the ~25 KB unrolled body.
        bool valid[4], st_ok[4];
        float* outq[4];        // output row pointers at channel n0 + 2 cq (never dereferenced when !st_ok)
        const float* xg0;      // without the ring (BN = 128): the saved input straight from global memory (first row, same channel offset)
        int64_t xrow;
        {
          const int gy0 = y0h + 4 * q;
          float* const o0 = a.out + ((((int64_t)b * a.out_planes + a.out_plane) * a.out_h + (gy0 * a.out_stride + a.out_oy)) * a.out_w +
                                     (gx * a.out_stride + a.out_ox)) * a.N + n0 + 2 * cq;
          const int64_t orow = (int64_t)a.out_stride * a.out_w * a.N;
          xg0 = DGX ? a.e.xsave + (int64_t)b * a.e.xsave_bstride + ((int64_t)gy0 * a.gw + gx) * a.N + n0 + 2 * cq : nullptr;
          xrow = (int64_t)a.gw * a.N;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            valid[r] = gy0 + r < a.gh && gx < a.gw;
            st_ok[r] = valid[r] && a.out != nullptr;
            outq[r] = o0 + r * orow;
          }
        }
        const bool xs_smem2 = DGX && a.XS > 0 && a.xs_has_x;   // saved input comes from the ring
        const bool ring2 = (DGX || PXS) && a.XS > 0;       // the ring exists (saved input and / or per-pixel scalars)
        const int sx = ring2 ? it % a.XS : 0;
        // saved-input tile (TMA, SWIZZLE_128B): pixel m = 32q + 8r + x is row m, its 16-byte channel quad j sits at j ^ x;
        // quad of channel pair (k, cq) is j = 2k + (cq >> 1), so its byte position is xo0 ^ (k << 5)
        const uint32_t xs_stage = smem0 + (uint32_t)a.xs_off + (uint32_t)sx * (uint32_t)a.xs_stride;
        const uint32_t xs_row = xs_stage + (uint32_t)((32 * q + x) * 128 + (cq & 1) * 8);
        const uint32_t xo0 = (uint32_t)(((cq >> 1) ^ x) << 4);
        if (ring2) mbar_wait(bar_xs_full(sx), (uint32_t)(it / a.XS) & 1u);
        // per-pixel scalars of this tile from the ring stage: noise[row][x], skip gradient [plane][row][x]; behind them (+ 2 KB)
        // the per-channel vectors demod[BN], bias[BN], mod_out[BN]
        const float* const px = reinterpret_cast<const float*>(smem_al + a.xs_off + (size_t)sx * a.xs_stride + (size_t)((DGX && a.xs_has_x) ? nchunk : 0) * XS_CHUNK);
        const int npx = 128 * a.nhalf;   // floats per plane of per-pixel scalars
        const float* const pvec = px + 4 * npx + 2 * cq;
        if (PXS) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = (TILE_H * hf + 4 * q + r) * 8 + x;
            cnz[r] = nw2 * px[i];
            crg[r][0] = RGB ? px[npx + i] : 0.f; crg[r][1] = RGB ? px[2 * npx + i] : 0.f; crg[r][2] = RGB ? px[3 * npx + i] : 0.f;
          }
        }
        float rgbacc[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) rgbacc[r][0] = rgbacc[r][1] = rgbacc[r][2] = 0.f;
        constexpr bool VEC_DB = EPI == EPI_ACT || EPI == EPI_DGRAD_ACT;   // per-(sample, channel) demodulation and per-channel bias
        constexpr bool VEC_B = VEC_DB || EPI == EPI_RELU;
        for (int ph = 0; ph < nphase; ++ph)
          for (int c = 0; c < nchunk; ++c) {
            // per-channel vectors of this chunk, all four k at once and (first chunk) before the wait for the accumulator:
            // one L2 round trip per chunk instead of one per k behind the stores of the previous k
            const int64_t bn0 = (int64_t)b * a.N + n0 + c * 32 + 2 * cq;
            const int nc0 = n0 + c * 32 + 2 * cq;
            // (forward epilogues only: the data-gradient epilogues have no registers to spare for them and load per k)
            constexpr bool VPRE = !DG;
            float2 vD[4], vB[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              vD[k] = vB[k] = make_float2(0.f, 0.f);
              if (PXS) continue;   // staged in the ring
              if (VPRE && VEC_DB) vD[k] = __ldg(reinterpret_cast<const float2*>(a.e.demod + bn0) + 4 * k);
              if (VPRE && VEC_B) vB[k] = __ldg(reinterpret_cast<const float2*>(a.e.bias + nc0) + 4 * k);
            }
            if (ph == 0 && c == 0) {
              mbar_wait(bar_acc_full(as), ((uint32_t)it >> acc_shift) & 1u);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            uint32_t acc[2][16];   // [half][4k + 2 j2 + e]: row r = 2 half + j2
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t taddr = tmem + ((uint32_t)(q * 32 + hh * 16) << 16) + (uint32_t)(((as * a.nhalf + hf) * nphase + ph) * a.BN + c * 32);
              asm volatile(
                  "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                  : "=r"(acc[hh][0]), "=r"(acc[hh][1]), "=r"(acc[hh][2]), "=r"(acc[hh][3]), "=r"(acc[hh][4]), "=r"(acc[hh][5]),
                    "=r"(acc[hh][6]), "=r"(acc[hh][7]), "=r"(acc[hh][8]), "=r"(acc[hh][9]), "=r"(acc[hh][10]), "=r"(acc[hh][11]),
                    "=r"(acc[hh][12]), "=r"(acc[hh][13]), "=r"(acc[hh][14]), "=r"(acc[hh][15])
                  : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (c == nchunk - 1 && ph == nphase - 1) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              mbar_arrive(bar_acc_empty(as));
            }
            // store pointers of this (phase, chunk) at channel 2 cq of the chunk (+ 8 k floats per k): the row pointers themselves,
            // advanced by 32 channels at the end of every chunk
            if (EPI == EPI_STORE && nphase > 1 && c == 0) {
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const int gy = y0h + 4 * q + r;
                st_ok[r] = a.out != nullptr && gy < a.gh - (ph >> 1) && gx < a.gw - (ph & 1);
                if (a.out_stride == 2)   // phases interleaved into the [2H+1, 2W+1] image: (2 gy + a, 2 gx + b)
                  outq[r] = a.out + (((int64_t)b * (2 * a.out_h - 1) + (2 * gy + (ph >> 1))) * (2 * a.out_w - 1) + (2 * gx + (ph & 1))) * a.N + n0 + 2 * cq;
                else
                  outq[r] = a.out + ((((int64_t)b * a.out_planes + ph) * a.out_h + gy) * a.out_w + gx) * a.N + n0 + 2 * cq;
              }
            }
            const uint32_t xs_c = xs_row + (uint32_t)c * XS_CHUNK;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int ch = c * 32 + 8 * k + 2 * cq;       // channel offset inside this CTA's BN slice
              auto av = [&](int r, int e) { return __uint_as_float(acc[r >> 1][4 * k + 2 * (r & 1) + e]); };
              auto st2 = [&](int r, float v0, float v1) { if (st_ok[r]) *reinterpret_cast<float2*>(outq[r] + 8 * k) = make_float2(v0, v1); };
              // saved forward input of the four rows (channel pair of this k)
              float2 xk[4];
              if (DGX) {
                if (xs_smem2) {
                  const uint32_t ad = xs_c + (xo0 ^ (uint32_t)(k << 5));
#pragma unroll
                  for (int r = 0; r < 4; ++r)
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(xk[r].x), "=f"(xk[r].y) : "r"(ad + (uint32_t)(r * 8 * 128)));
                } else {
#pragma unroll
                  for (int r = 0; r < 4; ++r) {
                    xk[r] = make_float2(0.f, 0.f);
                    if (valid[r]) xk[r] = __ldg(reinterpret_cast<const float2*>(xg0 + r * xrow + c * 32) + 4 * k);
                  }
                }
              }
              if (EPI == EPI_STORE) {
#pragma unroll
                for (int r = 0; r < 4; ++r) st2(r, av(r, 0), av(r, 1));
              } else if (EPI == EPI_ACT) {
                const float2 d2 = PXS ? *reinterpret_cast<const float2*>(pvec + c * 32 + 8 * k) : vD[k];
                const float2 b2 = PXS ? *reinterpret_cast<const float2*>(pvec + a.BN + c * 32 + 8 * k) : vB[k];
                float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0;
                if (do_rgb2) {
                  const float2 s2 = __ldg(reinterpret_cast<const float2*>(a.e.s_rgb + bn0) + 4 * k);
                  const float2 w0 = __ldg(reinterpret_cast<const float2*>(a.e.wrgb + 0 * a.N + nc0) + 4 * k);
                  const float2 w1 = __ldg(reinterpret_cast<const float2*>(a.e.wrgb + 1 * a.N + nc0) + 4 * k);
                  const float2 w2 = __ldg(reinterpret_cast<const float2*>(a.e.wrgb + 2 * a.N + nc0) + 4 * k);
                  q0 = make_float2(s2.x * w0.x, s2.y * w0.y);
                  q1 = make_float2(s2.x * w1.x, s2.y * w1.y);
                  q2 = make_float2(s2.x * w2.x, s2.y * w2.y);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const float o0 = lrelu(fmaf(av(r, 0), d2.x, cnz[r]) + b2.x);
                  const float o1 = lrelu(fmaf(av(r, 1), d2.y, cnz[r]) + b2.y);
                  st2(r, o0, o1);
                  if (do_rgb2) {
                    rgbacc[r][0] = fmaf(o0, q0.x, fmaf(o1, q0.y, rgbacc[r][0]));
                    rgbacc[r][1] = fmaf(o0, q1.x, fmaf(o1, q1.y, rgbacc[r][1]));
                    rgbacc[r][2] = fmaf(o0, q2.x, fmaf(o1, q2.y, rgbacc[r][2]));
                  }
                }
              } else if (EPI == EPI_RELU) {
                const float2 b2 = vB[k];
#pragma unroll
                for (int r = 0; r < 4; ++r) st2(r, fmaxf(av(r, 0) + b2.x, 0.f), fmaxf(av(r, 1) + b2.y, 0.f));
              } else if (EPI == EPI_DGRAD_RELU) {
#pragma unroll
                for (int r = 0; r < 4; ++r) st2(r, xk[r].x > 0.f ? av(r, 0) : 0.f, xk[r].y > 0.f ? av(r, 1) : 0.f);
              } else {
                float2 m2, d2 = make_float2(0.f, 0.f), b2 = d2;
                if (PXS) {
                  d2 = *reinterpret_cast<const float2*>(pvec + c * 32 + 8 * k);
                  b2 = *reinterpret_cast<const float2*>(pvec + a.BN + c * 32 + 8 * k);
                  m2 = *reinterpret_cast<const float2*>(pvec + 2 * a.BN + c * 32 + 8 * k);
                } else {
                  m2 = __ldg(reinterpret_cast<const float2*>(a.e.mod_out + bn0) + 4 * k);
                  if (EPI == EPI_DGRAD_ACT) {
                    d2 = __ldg(reinterpret_cast<const float2*>(a.e.demod + bn0) + 4 * k);
                    b2 = __ldg(reinterpret_cast<const float2*>(a.e.bias + nc0) + 4 * k);
                  }
                }
                float2 s2 = make_float2(0.f, 0.f), w0 = s2, w1 = s2, w2 = s2;
                if (has_rgb2) {
                  s2 = __ldg(reinterpret_cast<const float2*>(a.e.s_rgb + bn0) + 4 * k);
                  w0 = __ldg(reinterpret_cast<const float2*>(a.e.wrgb + 0 * a.N + nc0) + 4 * k);
                  w1 = __ldg(reinterpret_cast<const float2*>(a.e.wrgb + 1 * a.N + nc0) + 4 * k);
                  w2 = __ldg(reinterpret_cast<const float2*>(a.e.wrgb + 2 * a.N + nc0) + 4 * k);
                }
                float X0 = 0.f, X1 = 0.f, T0 = 0.f, T1 = 0.f, R0 = 0.f, R1 = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  const float2 x2 = xk[r];
                  const float v0 = av(r, 0), v1 = av(r, 1);
                  X0 = fmaf(x2.x, v0, X0);
                  X1 = fmaf(x2.y, v1, X1);
                  if (EPI == EPI_DGRAD) {
                    st2(r, v0 * m2.x, v1 * m2.y);
                  } else {
                    float u0 = 0.f, u1 = 0.f;
                    if (has_rgb2) {
                      u0 = fmaf(crg[r][2], w2.x, fmaf(crg[r][1], w1.x, crg[r][0] * w0.x));
                      u1 = fmaf(crg[r][2], w2.y, fmaf(crg[r][1], w1.y, crg[r][0] * w0.y));
                    }
                    const float gt0 = fmaf(u0, s2.x, v0 * m2.x), gt1 = fmaf(u1, s2.y, v1 * m2.y);
                    const bool p0 = x2.x > 0.f, p1 = x2.y > 0.f;
                    const float gp0 = gt0 * (p0 ? G : GS), gp1 = gt1 * (p1 ? G : GS);
                    const float pre0 = x2.x * (p0 ? IG : IGS), pre1 = x2.y * (p1 ? IG : IGS);
                    if (valid[r]) {
                      T0 = fmaf(gp0, pre0 - cnz[r] - b2.x, T0);
                      T1 = fmaf(gp1, pre1 - cnz[r] - b2.y, T1);
                    }
                    R0 = fmaf(x2.x, u0, R0);
                    R1 = fmaf(x2.y, u1, R1);
                    st2(r, gp0 * d2.x, gp1 * d2.y);
                  }
                }
                // sum over the eight pixel columns of this quarter (lanes that share t & 3): fixed xor tree
#define LFP_XRED(v) v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                LFP_XRED(X0) LFP_XRED(X1)
                if (EPI == EPI_DGRAD_ACT) { LFP_XRED(T0) LFP_XRED(T1) if (has_rgb2) { LFP_XRED(R0) LFP_XRED(R1) } }
#undef LFP_XRED
                if (x == 0) {
                  red[q * a.BN + ch] = X0; red[q * a.BN + ch + 1] = X1;
                  if (EPI == EPI_DGRAD_ACT) {
                    red[rs + q * a.BN + ch] = T0; red[rs + q * a.BN + ch + 1] = T1;
                    if (has_rgb2) { red[2 * rs + q * a.BN + ch] = R0; red[2 * rs + q * a.BN + ch + 1] = R1; }
                  }
                }
              }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) outq[r] += 32;
          }
        if (ring2) mbar_arrive(bar_xs_empty(sx));
        if (do_rgb2) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              float v = rgbacc[r][o];
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              rgbacc[r][o] = v;
            }
          if (cq == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (valid[r]) {
                float* ro = a.e.rgb_out + (int64_t)b * 3 * hw + (int64_t)(y0h + 4 * q + r) * a.gw + gx;
                ro[0] = rgbacc[r][0] + __ldg(a.e.rgb_bias + 0);
                ro[hw] = rgbacc[r][1] + __ldg(a.e.rgb_bias + 1);
                ro[2 * hw] = rgbacc[r][2] + __ldg(a.e.rgb_bias + 2);
              }
          }
        }
        if (DG) {
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eset) : "memory");
          // per-tile partial sums are indexed by 16-row tiles (vtiles_per per sample), whatever the height of the work item;
          // a lower half that lies entirely below the image has no slot (and nothing to add)
          for (int n = et; n < a.BN && y0h < a.gh; n += 128) {
            const int64_t o = ((int64_t)b * a.vtiles_per + (wk.ty * a.nhalf + hf) * a.tiles_x + wk.tx) * a.N + n0 + n;
            const int bnn = a.BN;
            a.e.partial[o] = ((red[n] + red[bnn + n]) + red[2 * bnn + n]) + red[3 * bnn + n];
            if (EPI == EPI_DGRAD_ACT) {
              a.e.partial_T[o] = ((red[rs + n] + red[rs + bnn + n]) + red[rs + 2 * bnn + n]) + red[rs + 3 * bnn + n];
              if (has_rgb2) a.e.partial_R[o] = ((red[2 * rs + n] + red[2 * rs + bnn + n]) + red[2 * rs + 2 * bnn + n]) + red[2 * rs + 3 * bnn + n];
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eset) : "memory");
        }
      }
    } else {
    // per-pixel scalars (noise, skip-image gradient) of a work item; fetched one work item ahead so that their
    // DRAM latency is hidden behind the current tile
    const bool need_nz = EPI == EPI_ACT || EPI == EPI_DGRAD_ACT;
    const bool has_rgb = EPI == EPI_DGRAD_ACT && a.e.drgb != nullptr;
    const float nw = need_nz ? __ldg(a.e.noise_w) : 0.f;
    auto fetch = [&](int w, float& o_nz, float& o_r0, float& o_r1, float& o_r2) {
      o_nz = 0.f; o_r0 = 0.f; o_r1 = 0.f; o_r2 = 0.f;
      if (w >= a.total_work) return;
      const Work k = decode(a, w);
      const int gy = k.y0 + (m >> 3), gx = k.x0 + (m & 7);
      if (gy >= a.gh || gx >= a.gw) return;
      const int pix = gy * a.gw + gx;
      if (need_nz) o_nz = __ldg(a.e.noise + (int64_t)k.b * a.e.noise_bstride + pix);   // scaled at use
      if (has_rgb) {
        const int64_t hw = (int64_t)a.gh * a.gw;
        o_r0 = __ldg(a.e.drgb + ((int64_t)k.b * 3 + 0) * hw + pix);
        o_r1 = __ldg(a.e.drgb + ((int64_t)k.b * 3 + 1) * hw + pix);
        o_r2 = __ldg(a.e.drgb + ((int64_t)k.b * 3 + 2) * hw + pix);
      }
    };
    const int wstep = a.nsets * gridDim.x;
    float nz_n, rg0_n, rg1_n, rg2_n;
    fetch(blockIdx.x + eset * gridDim.x, nz_n, rg0_n, rg1_n, rg2_n);
    const bool xs_smem = DGX && a.XS > 0 && a.xs_has_x;
    const int nchunk = a.BN >> 5;
    int it = eset;
    for (int w = eset < a.nsets ? blockIdx.x + eset * gridDim.x : a.total_work; w < a.total_work; w += wstep, it += a.nsets) {
      const Work wk = decode(a, w);
      const int as = it & (a.nacc - 1);
      const int b = wk.b, n0 = wk.n0;
      const int gy = wk.y0 + (m >> 3), gx = wk.x0 + (m & 7);
      const bool valid = gy < a.gh && gx < a.gw;
      const int pix = gy * a.gw + gx;
      const float nz = nw * nz_n, rg0 = rg0_n, rg1 = rg1_n, rg2 = rg2_n;
      fetch(w + wstep, nz_n, rg0_n, rg1_n, rg2_n);
      float* outp0 = nullptr;
      if (a.out != nullptr && valid)
        outp0 = a.out + ((((int64_t)b * a.out_planes + a.out_plane) * a.out_h + (gy * a.out_stride + a.out_oy)) * a.out_w +
                        (gx * a.out_stride + a.out_ox)) * a.N + n0;
      float* const outp = outp0;
      const float* xs = nullptr;
      if (DGX && valid && !xs_smem) xs = a.e.xsave + (int64_t)b * a.e.xsave_bstride + (int64_t)pix * a.N + n0;
      // saved-input tile in shared memory (TMA, SWIZZLE_128B): pixel m is row m, channel quad j at position j ^ (m & 7)
      const int sx = xs_smem ? it % a.XS : 0;
      const uint8_t* xrow = smem_al + a.xs_off + (size_t)sx * a.xs_stride + m * 128;
      if (xs_smem) mbar_wait(bar_xs_full(sx), (uint32_t)(it / a.XS) & 1u);
      auto load_x = [&](int c, int j) -> float4 {
        if (xs_smem) return *reinterpret_cast<const float4*>(xrow + c * XS_CHUNK + ((j ^ (m & 7)) << 4));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xs) v = __ldg(reinterpret_cast<const float4*>(xs + c * 32 + j * 4));
        return v;
      };
      const bool do_rgb = EPI == EPI_ACT && a.e.rgb_out != nullptr;
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      mbar_wait(bar_acc_full(as), ((uint32_t)it >> acc_shift) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nphase = a.taps.nphase;
      for (int ph = 0; ph < nphase; ++ph)
      for (int c = 0; c < nchunk; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * nphase + ph) * a.BN + c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
            "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == nchunk - 1 && ph == nphase - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(bar_acc_empty(as));
        }
        const int nc = n0 + c * 32;
        if (EPI == EPI_ACT) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.e.demod + (int64_t)b * a.N + nc + j * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.e.bias + nc + j * 4));
            float4 v;
            v.x = lrelu(fmaf(__uint_as_float(r[j * 4 + 0]), d4.x, nz) + b4.x);
            v.y = lrelu(fmaf(__uint_as_float(r[j * 4 + 1]), d4.y, nz) + b4.y);
            v.z = lrelu(fmaf(__uint_as_float(r[j * 4 + 2]), d4.z, nz) + b4.z);
            v.w = lrelu(fmaf(__uint_as_float(r[j * 4 + 3]), d4.w, nz) + b4.w);
            if (outp) *reinterpret_cast<float4*>(outp + c * 32 + j * 4) = v;
            if (do_rgb) {   // ToRGB: 1x1 modulated conv, no demodulation (src/model.py:379-383)
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.e.s_rgb + (int64_t)b * a.N + nc + j * 4));
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.e.wrgb + 0 * a.N + nc + j * 4));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.e.wrgb + 1 * a.N + nc + j * 4));
              const float4 w2 = __ldg(reinterpret_cast<const float4*>(a.e.wrgb + 2 * a.N + nc + j * 4));
              const float4 vs = make_float4(v.x * s4.x, v.y * s4.y, v.z * s4.z, v.w * s4.w);
              rgb0 = fmaf(vs.x, w0.x, fmaf(vs.y, w0.y, fmaf(vs.z, w0.z, fmaf(vs.w, w0.w, rgb0))));
              rgb1 = fmaf(vs.x, w1.x, fmaf(vs.y, w1.y, fmaf(vs.z, w1.z, fmaf(vs.w, w1.w, rgb1))));
              rgb2 = fmaf(vs.x, w2.x, fmaf(vs.y, w2.y, fmaf(vs.z, w2.z, fmaf(vs.w, w2.w, rgb2))));
            }
          }
        } else if (EPI == EPI_STORE) {
          float* outp = outp0;
          if (nphase > 1) {   // fused sub-pixel phases: accumulator ph -> output plane ph over its own (smaller) grid
            outp = nullptr;
            if (a.out != nullptr && gy < a.gh - (ph >> 1) && gx < a.gw - (ph & 1)) {
              if (a.out_stride == 2)
                outp = a.out + (((int64_t)b * (2 * a.out_h - 1) + (2 * gy + (ph >> 1))) * (2 * a.out_w - 1) + (2 * gx + (ph & 1))) * a.N + n0;
              else
                outp = a.out + ((((int64_t)b * a.out_planes + ph) * a.out_h + gy) * a.out_w + gx) * a.N + n0;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (outp)
              *reinterpret_cast<float4*>(outp + c * 32 + j * 4) =
                  make_float4(__uint_as_float(r[j * 4 + 0]), __uint_as_float(r[j * 4 + 1]), __uint_as_float(r[j * 4 + 2]),
                              __uint_as_float(r[j * 4 + 3]));
        } else if (EPI == EPI_RELU || EPI == EPI_DGRAD_RELU) {
          // the VGG epilogues exist in the 16x256b form only (tc_launch2 refuses them when LFP_TC_E2=0)
        } else if (EPI == EPI_DGRAD) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x4 = load_x(c, j);
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.e.mod_out + (int64_t)b * a.N + nc + j * 4));
            const float v0 = __uint_as_float(r[j * 4 + 0]), v1 = __uint_as_float(r[j * 4 + 1]);
            const float v2 = __uint_as_float(r[j * 4 + 2]), v3 = __uint_as_float(r[j * 4 + 3]);
            scr[lane * 33 + j * 4 + 0] = x4.x * v0;
            scr[lane * 33 + j * 4 + 1] = x4.y * v1;
            scr[lane * 33 + j * 4 + 2] = x4.z * v2;
            scr[lane * 33 + j * 4 + 3] = x4.w * v3;
            if (outp) *reinterpret_cast<float4*>(outp + c * 32 + j * 4) = make_float4(v0 * m4.x, v1 * m4.y, v2 * m4.z, v3 * m4.w);
          }
          __syncwarp();
          float sum = 0.f;
#pragma unroll 8
          for (int l = 0; l < 32; ++l) sum += scr[l * 33 + lane];  // fixed order: deterministic
          __syncwarp();
          red[q * a.BN + c * 32 + lane] = sum;
        } else {
          // data gradient + backward through noise / bias / lrelu (+ ToRGB branch) of the layer that produced xsave
          constexpr float G = kLreluGain, GS = kLreluGain * kLreluSlope, IG = 1.f / kLreluGain, IGS = 1.f / (kLreluGain * kLreluSlope);
          float tv[32], rv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 x4 = load_x(c, j);
            const int64_t bn = (int64_t)b * a.N + nc + j * 4;
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(a.e.mod_out + bn));
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.e.demod + bn));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.e.bias + nc + j * 4));
            float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = u4;
            if (has_rgb) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.e.wrgb + 0 * a.N + nc + j * 4));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.e.wrgb + 1 * a.N + nc + j * 4));
              const float4 w2 = __ldg(reinterpret_cast<const float4*>(a.e.wrgb + 2 * a.N + nc + j * 4));
              s4 = __ldg(reinterpret_cast<const float4*>(a.e.s_rgb + bn));
              u4.x = fmaf(rg2, w2.x, fmaf(rg1, w1.x, rg0 * w0.x));
              u4.y = fmaf(rg2, w2.y, fmaf(rg1, w1.y, rg0 * w0.y));
              u4.z = fmaf(rg2, w2.z, fmaf(rg1, w1.z, rg0 * w0.z));
              u4.w = fmaf(rg2, w2.w, fmaf(rg1, w1.w, rg0 * w0.w));
            }
            float4 o4;
#define LFP_DGA(comp, k)                                                        \
  {                                                                             \
    const float v = __uint_as_float(r[j * 4 + k]);                              \
    const float x = x4.comp;                                                    \
    scr[lane * 33 + j * 4 + k] = x * v;                                         \
    const float gt = fmaf(u4.comp, s4.comp, v * m4.comp);                       \
    const bool pos = x > 0.f;                                                   \
    const float gpre = gt * (pos ? G : GS);                                     \
    const float pre = x * (pos ? IG : IGS);                                     \
    tv[j * 4 + k] = valid ? gpre * (pre - nz - b4.comp) : 0.f;                  \
    rv[j * 4 + k] = x * u4.comp;                                                \
    o4.comp = gpre * d4.comp;                                                   \
  }
            LFP_DGA(x, 0) LFP_DGA(y, 1) LFP_DGA(z, 2) LFP_DGA(w, 3)
#undef LFP_DGA
            if (outp) *reinterpret_cast<float4*>(outp + c * 32 + j * 4) = o4;
          }
          __syncwarp();
          float sum = 0.f;
#pragma unroll 8
          for (int l = 0; l < 32; ++l) sum += scr[l * 33 + lane];
          __syncwarp();
          red[q * a.BN + c * 32 + lane] = sum;
#pragma unroll
          for (int k = 0; k < 32; ++k) scr[lane * 33 + k] = tv[k];
          __syncwarp();
          sum = 0.f;
#pragma unroll 8
          for (int l = 0; l < 32; ++l) sum += scr[l * 33 + lane];
          __syncwarp();
          red[rs + q * a.BN + c * 32 + lane] = sum;
          if (has_rgb) {
#pragma unroll
            for (int k = 0; k < 32; ++k) scr[lane * 33 + k] = rv[k];
            __syncwarp();
            sum = 0.f;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) sum += scr[l * 33 + lane];
            __syncwarp();
            red[2 * rs + q * a.BN + c * 32 + lane] = sum;
          }
        }
      }
      if (xs_smem) mbar_arrive(bar_xs_empty(sx));   // this thread has read its row of the saved-input tile
      if (do_rgb && valid) {
        const int64_t hw = (int64_t)a.gh * a.gw;
        float* ro = a.e.rgb_out + (int64_t)b * 3 * hw + pix;
        ro[0] = rgb0 + __ldg(a.e.rgb_bias + 0);
        ro[hw] = rgb1 + __ldg(a.e.rgb_bias + 1);
        ro[2 * hw] = rgb2 + __ldg(a.e.rgb_bias + 2);
      }
      if (DG) {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eset) : "memory");
        for (int n = et; n < a.BN; n += 128) {
          const int64_t o = ((int64_t)b * tiles_per + wk.tile) * a.N + n0 + n;
          const int bn = a.BN;
          a.e.partial[o] = ((red[n] + red[bn + n]) + red[2 * bn + n]) + red[3 * bn + n];
          if (EPI == EPI_DGRAD_ACT) {
            a.e.partial_T[o] = ((red[rs + n] + red[rs + bn + n]) + red[rs + 2 * bn + n]) + red[rs + 3 * bn + n];
            if (has_rgb) a.e.partial_R[o] = ((red[2 * rs + n] + red[2 * rs + bn + n]) + red[2 * rs + 2 * bn + n]) + red[2 * rs + 3 * bn + n];
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eset) : "memory");
      }
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static int encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
  if (fn == nullptr) { set_error("conv_tc: cuTensorMapEncodeTiled is not available from this driver"); return LFP_EUNSUPPORTED; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return LFP_EINVAL; }
  return 0;
}

// plain (un-swizzled) tile map: the per-pixel scalar planes of the PXS kernels
static int encode_plain(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                        const cuuint32_t* box) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
  if (fn == nullptr) { set_error("conv_tc: cuTensorMapEncodeTiled is not available from this driver"); return LFP_EUNSUPPORTED; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled (plain) failed with CUresult %d", (int)r); return LFP_EINVAL; }
  return 0;
}

}  // namespace tc

int tc_block_n(int N) { return N >= 256 ? 256 : N; }
// Output-channel slice of one launch: the widest slice (best MMA shape) unless that leaves SMs idle - then halve it, down
// to 32, until there is one work item per SM (every CTA also streams 1/n of the weights, which is what the low-resolution
// 512-channel layers and small batches are bound by).  If that lands just above one item per SM, step back up: the second
// round would run on a handful of SMs (4 x 4 and 8 x 8 at B = 20: 160 items of 64 channels in two rounds -> 80 items of 128
// in one; forward 72 -> 46 us, data gradient 107 -> 67 us per launch).  A full cost model (rounds x MMA cycles) was tried and
// lost on the 16 - 32 px layers, whose time is set by the weight stream per CTA rather than by the round count.
static int tc_pick_bn(int N, int64_t tiles_times_batch) {
  static const bool legacy = getenv("LFP_TC_PICK_BN") != nullptr && atoi(getenv("LFP_TC_PICK_BN")) == 0;   // A/B knob
  int bn = tc_block_n(N);
  const int sms = num_sms();
  while (bn > 32 && tiles_times_batch * (N / bn) < sms) bn >>= 1;
  const int64_t items = tiles_times_batch * (N / bn);
  if (!legacy && bn < tc_block_n(N) && items > sms && items * 4 <= (int64_t)sms * 5) bn <<= 1;
  return bn;
}
static int tc_bn_index(int bn) { return bn == 256 ? 0 : (bn == 128 ? 1 : (bn == 64 ? 2 : 3)); }
int tc_tiles_per_sample(int gh, int gw) { return (int)(ceil_div(gw, tc::TILE_W) * ceil_div(gh, tc::TILE_H)); }
bool tc_supported(int K, int N, int gh, int gw) {
  const int bn = tc_block_n(N);
  const bool n_ok = (bn == 32 || bn == 64 || bn == 128 || bn == 256) && N % bn == 0;
  return K % 32 == 0 && K <= 512 && n_ok && gh >= 4 && gw >= 4;
}

int tc_make_weight_maps(void* maps_out, const float* table, int rows, int K, int N) {
  // four maps, box rows 256 / 128 / 64 / 32 (the ones wider than N are left unused)
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  for (int i = 0; i < 4; ++i) {
    const int bn = 256 >> i;
    if (bn > N || N % bn != 0) continue;
    const cuuint32_t box[2] = {32, (cuuint32_t)bn};
    LFP_TRY(tc::encode(reinterpret_cast<CUtensorMap*>(reinterpret_cast<unsigned char*>(maps_out) + 128 * i), table, 2, dims, strides, box));
  }
  return 0;
}

template <int EPI, bool MOD, bool RES, bool E2, bool RGB, bool PXS>
static int tc_launch4(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const CUtensorMap& tmNz, const CUtensorMap& tmRg,
                      const tc::Args& a, size_t dyn_smem, cudaStream_t s) {
  // the opt-in shared-memory limit is a per-device function attribute
  static unsigned long long attr_done_mask = 0;
  int dev = 0;
  LFP_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_done_mask >> dev) & 1ull)) {
    LFP_CUDA(cudaFuncSetAttribute(tc::conv_tc_kernel<EPI, MOD, RES, E2, RGB, PXS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(tc::SMEM_OPTIN - tc::STATIC_SMEM_RESERVE)));
    if (dev < 64) attr_done_mask |= 1ull << dev;
  }
  const int max_ctas = num_sms();
  const dim3 grid((unsigned)(a.total_work < max_ctas ? a.total_work : max_ctas));
  tc::conv_tc_kernel<EPI, MOD, RES, E2, RGB, PXS><<<grid, MOD ? tc::NTHREADS_MOD : tc::NTHREADS_PLAIN, dyn_smem, s>>>(tmA, tmB, tmX, tmNz, tmRg, a);
  LFP_LAUNCH_CHECK();
  return 0;
}

template <int EPI, bool MOD, bool RES, bool E2>
static int tc_launch3(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const CUtensorMap& tmNz, const CUtensorMap& tmRg,
                      const tc::Args& a, size_t dyn_smem, cudaStream_t s) {
  constexpr bool da = (EPI == EPI_DGRAD_ACT || EPI == EPI_ACT) && E2;   // epilogues with per-pixel scalars
  const bool pxs = da && a.px_ok && a.XS > 0;
  constexpr bool rgbv = EPI == EPI_DGRAD_ACT && E2;   // only the fused act-backward has a skip-gradient variant
  if (rgbv && a.e.drgb != nullptr)
    return pxs ? tc_launch4<EPI, MOD, RES, E2, rgbv, da>(tmA, tmB, tmX, tmNz, tmRg, a, dyn_smem, s)
               : tc_launch4<EPI, MOD, RES, E2, rgbv, false>(tmA, tmB, tmX, tmNz, tmRg, a, dyn_smem, s);
  return pxs ? tc_launch4<EPI, MOD, RES, E2, false, da>(tmA, tmB, tmX, tmNz, tmRg, a, dyn_smem, s)
             : tc_launch4<EPI, MOD, RES, E2, false, false>(tmA, tmB, tmX, tmNz, tmRg, a, dyn_smem, s);
}

static bool tc_use_e2(int bn) {
  // 16x256b epilogue for every slice width: the slice width is chosen per launch from the amount of work, and a
  // trajectory's result must not depend on it - both epilogues add the same numbers, but in a different order
  static const bool e2_off = getenv("LFP_TC_E2") != nullptr && atoi(getenv("LFP_TC_E2")) == 0;
  (void)bn;
  return !e2_off;
}

template <int EPI, bool MOD, bool RES>
static int tc_launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const CUtensorMap& tmNz, const CUtensorMap& tmRg,
                      const tc::Args& a, size_t dyn_smem, cudaStream_t s) {
  const bool e2 = tc_use_e2(a.BN);
  LFP_CHECK_ARG(e2 || (EPI != EPI_RELU && EPI != EPI_DGRAD_RELU), "conv_tc: the ReLU epilogues need the 16x256b epilogue (LFP_TC_E2=0 is set)");
  return e2 ? tc_launch3<EPI, MOD, RES, true>(tmA, tmB, tmX, tmNz, tmRg, a, dyn_smem, s) : tc_launch3<EPI, MOD, RES, false>(tmA, tmB, tmX, tmNz, tmRg, a, dyn_smem, s);
}

template <int EPI, bool MOD>
static int tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, tc::Args& a, int ntaps, cudaStream_t s,
                     const CUtensorMap* tmNz = nullptr, const CUtensorMap* tmRg = nullptr) {
  // shared-memory plan: the weight slice stays resident when it fits beside >= 3 activation stages
  const size_t b_all = (size_t)ntaps * (a.K / 32) * a.BN * 128;
  // an epilogue set that finished tile i waits next for tile i + nsets; the parity wait on that TMEM stage is only
  // unambiguous when the stage's previous use (tile i + nsets - nacc) is already known to be complete, i.e.
  // nsets <= nacc (BN = 256 leaves room for two accumulator stages only)
  a.nsets = (a.BN <= 128 && !MOD) ? 3 : 2;   // the modulated kernels have two epilogue sets (512-thread layout)
  constexpr bool dg = EPI == EPI_DGRAD || EPI == EPI_DGRAD_ACT || EPI == EPI_DGRAD_RELU;
  // data-gradient epilogues of the HBM-bound layers (N <= 64) get their saved-input tiles through a TMA ring
  const bool px_kernel = (EPI == EPI_DGRAD_ACT || EPI == EPI_ACT) && a.px_ok && tc_use_e2(a.BN);
  // the ring carries the saved-input tiles of the HBM-bound data-gradient layers (N <= 64) and / or, for the PXS kernels
  // of any width, the per-pixel scalars
  a.xs_has_x = (dg && a.BN <= 64 && a.n_ntiles == 1) ? 1 : 0;
  // LFP_TC_XS_WIDE: ring stages when a stage carries two saved-input chunks (BN = 64: 34 KB per stage) - A/B knob
  static const int xs_wide = getenv("LFP_TC_XS_WIDE") ? atoi(getenv("LFP_TC_XS_WIDE")) : 2;
  const int xs_max = (a.xs_has_x || px_kernel) ? ((a.xs_has_x && a.BN >= 64) ? xs_wide : 3) : 0;
  // ring stage = BN / 32 saved-input chunks (when carried), plus 2 KB of per-pixel scalars (PXS kernels)
  // (PXS kernels: 2 KB of per-pixel scalars + the three per-channel vectors of the slice, rounded up to 1 KB)
  const size_t xs_stride = (a.xs_has_x ? (size_t)(a.BN / 32) * tc::XS_CHUNK : 0) + (px_kernel ? 2048 * (size_t)a.nhalf + ((size_t)a.BN * 12 + 1023) / 1024 * 1024 : 0);
  const size_t a_stage = (size_t)a.a_stage;
  const size_t min_sa = a.nhalf == 2 ? 2 : 3;   // a tall stage feeds twice the MMAs: double buffering is enough
  a.xs_stride = (int)xs_stride;
  size_t xs_smem = 0, epi_smem = 0;
  // LFP_TC_SMEM_CAP (bytes): cap the shared memory a launch asks for, which leaves the rest of the 228 KB to the L1
  static const size_t smem_cap = getenv("LFP_TC_SMEM_CAP") ? (size_t)atol(getenv("LFP_TC_SMEM_CAP")) : (size_t)tc::SMEM_OPTIN;
  const size_t optin = smem_cap < (size_t)tc::SMEM_OPTIN && smem_cap >= 98304 ? smem_cap : (size_t)tc::SMEM_OPTIN;
  const int nsets0 = a.nsets;
  // Resident weights first: when the whole weight slice fits once the saved-input ring is given up (the 64 -> 64 data
  // gradients: 147 KB of weights against a 34 KB-per-stage ring), keep the weights and let the epilogue read the saved input
  // from global memory.  With the round-1 epilogue this was a wash (1.43 ms either way at 512 px: the streamed weights cost
  // 147 KB of L2 traffic per tile, the global loads 43 % of the epilogue's stall samples); with the rebuilt epilogue
  // addressing it wins: 64 -> 64 at 512 px 1356 -> 1127 us, the LPIPS step at 1024 px 80.6 -> 79.3 ms (same box).
  // LFP_TC_PREFER_RES=0 restores ring-first.
  static const bool prefer_res = getenv("LFP_TC_PREFER_RES") == nullptr || atoi(getenv("LFP_TC_PREFER_RES")) != 0;
  bool planned = false;
  if (prefer_res && a.n_ntiles == 1 && a.nhalf == 1)
    for (int xs = xs_max; xs >= 0 && !planned; --xs) {
      if (xs == 1) continue;                       // a one-stage ring serialises the epilogue sets
      int nsets = nsets0;
      if (xs > 0 && nsets > xs) nsets = xs;
      const size_t xsm = (size_t)xs * xs_stride;
      const size_t epi = tc::EPI_SMEM(EPI, nsets, a.BN, tc_use_e2(a.BN)) + xsm;
      if (optin < tc::STATIC_SMEM_RESERVE + 1024 + epi) continue;
      const size_t budget = optin - tc::STATIC_SMEM_RESERVE - 1024 - epi;
      if (b_all + 3 * a_stage <= budget) {
        a.XS = xs; a.nsets = nsets; xs_smem = xsm; epi_smem = epi;
        a.b_resident = 1; a.SB = 0; a.SA = (int)((budget - b_all) / a_stage);
        planned = true;
      }
    }
  for (int xs = xs_max; !planned; --xs) {
    a.XS = xs;
    // same parity argument for the saved-input ring: a set may only wait on stage (i + nsets) % XS when the stage's
    // previous use (tile i + nsets - XS) is one it has already seen complete, i.e. nsets <= XS
    if (xs > 0 && a.nsets > xs * a.nhalf) a.nsets = xs * a.nhalf;   // (a set steps ceil(nsets / nhalf) work items at a time)
    xs_smem = (size_t)xs * xs_stride;
    epi_smem = tc::EPI_SMEM(EPI, a.nsets, a.BN, tc_use_e2(a.BN)) + xs_smem;
    const size_t budget = optin - tc::STATIC_SMEM_RESERVE - 1024 - epi_smem;
    if (a.n_ntiles == 1 && a.nhalf == 1 && b_all + 3 * a_stage <= budget) {
      a.b_resident = 1; a.SB = 0;
      a.SA = (int)((budget - b_all) / a_stage);
    } else {
      // enough weight slices in flight to cover the L2 latency: one slice feeds 4 MMAs of BN/2 cycles each
      a.b_resident = 0; a.SB = a.BN >= 256 ? 4 : (a.BN >= 128 ? 8 : 12);
      if (const char* e = getenv("LFP_TC_SB")) { const int v = atoi(e); if (v >= 2 && v <= tc::MAX_SB) a.SB = v; }
      while (a.SB > 2 && (size_t)a.SB * a.BN * 128 + min_sa * a_stage > budget) --a.SB;
      a.SA = (int)((budget - (size_t)a.SB * a.BN * 128) / a_stage);
    }
    if ((a.SA >= (int)min_sa && (a.b_resident || a.SB >= 3)) || xs <= (xs_max > 0 ? 2 : 0)) break;
  }
  if (a.SA > tc::MAX_SA) a.SA = tc::MAX_SA;
  // resident-weight launches: more than five activation stages buy nothing, while keeping the request under the 196 KB
  // carve-out leaves 32 KB of L1 for the per-pixel noise / skip-gradient rows every tile re-reads (32 -> 32 at 1024 px:
  // 1649 -> 1396 us with six stages instead of eight)
  if (a.b_resident && smem_cap >= (size_t)tc::SMEM_OPTIN) {
    const size_t fixed = (size_t)b_all + epi_smem + 1024 + tc::STATIC_SMEM_RESERVE;
    static const int l1_min_sa = getenv("LFP_TC_L1_MIN_SA") ? atoi(getenv("LFP_TC_L1_MIN_SA")) : 5;
    while (a.SA > l1_min_sa && fixed + (size_t)a.SA * a_stage > 196u * 1024u) --a.SA;
  }
  if (const char* e = getenv("LFP_TC_SA_MAX")) { const int v = atoi(e); if (v >= 2 && a.SA > v) a.SA = v; }
  const int acc_cols = a.BN * a.taps.nphase * a.nhalf;   // TMEM columns of one work item
  a.nacc = acc_cols <= 128 ? 4 : (acc_cols <= 256 ? 2 : 1);
  if (a.nsets > a.nacc * a.nhalf) a.nsets = a.nacc * a.nhalf;
  static const int bbatch_env = getenv("LFP_TC_BBATCH") ? atoi(getenv("LFP_TC_BBATCH")) : 3;   // A/B knob
  a.bbatch = (!a.b_resident && a.SB >= 6 && bbatch_env == 3) ? 3 : 1;
  LFP_CHECK_ARG(a.SA >= 2, "conv_tc: shared-memory plan failed (BN=%d)", a.BN);
  // layout: [A ring][B ring or resident slice][xsave ring (1024-aligned)][epilogue scratch]
  a.xs_off = (int)((size_t)a.SA * a_stage + (a.b_resident ? b_all : (size_t)a.SB * a.BN * 128));
  a.epi_off = (int)((size_t)a.xs_off + xs_smem);
  const size_t smem = (size_t)a.xs_off + epi_smem + 1024;
  static const bool verbose = getenv("LFP_TC_VERBOSE") != nullptr;
  if (verbose)
    fprintf(stderr, "conv_tc plan: epi %d mod %d K %d N %d BN %d grid %dx%d taps %d phases %d halves %d | SA %d SB %d resident %d XS %d (x %d, stride %d) nsets %d nacc %d smem %zu\n",
            EPI, (int)MOD, a.K, a.N, a.BN, a.gh, a.gw, ntaps, a.taps.nphase, a.nhalf, a.SA, a.SB, a.b_resident, a.XS, a.xs_has_x, a.xs_stride, a.nsets, a.nacc, smem);
  const CUtensorMap& nzm = tmNz ? *tmNz : tmA;
  const CUtensorMap& rgm = tmRg ? *tmRg : tmA;
  return a.b_resident ? tc_launch2<EPI, MOD, true>(tmA, tmB, tmX, nzm, rgm, a, smem, s) : tc_launch2<EPI, MOD, false>(tmA, tmB, tmX, nzm, rgm, a, smem, s);
}

int launch_conv_tc(const TcConv& c, cudaStream_t s) {
  LFP_CHECK_ARG(tc_supported(c.K, c.N, c.gh, c.gw), "conv_tc: unsupported shape K=%d N=%d grid %dx%d", c.K, c.N, c.gh, c.gw);
  LFP_CHECK_ARG(c.taps.ngroups >= 1 && c.taps.ngroups <= 4 && c.taps.group_tap0[c.taps.ngroups] <= 9, "conv_tc: bad tap table");
  LFP_CHECK_ARG(c.taps.nphase <= 1 || (c.taps.nphase == 4 && c.epi == EPI_STORE && c.taps.ngroups == 1),
                "conv_tc: fused phases need EPI_STORE and one tap group");
  LFP_CHECK_ARG(((uintptr_t)c.in & 15) == 0 && c.wmap != nullptr, "conv_tc: input must be 16-byte aligned");
  // tall work items (two 16-row halves sharing the streamed weight slices): wide slices whose weights can never be resident,
  // enough work left for every SM, and little padding below the image
  // (fused ToRGB keeps all channels in one CTA; the four accumulators of fused phases cap the slice at 128 channels)
  int bn0 = c.e.rgb_out != nullptr ? tc_block_n(c.N) : tc_pick_bn(c.N, (int64_t)ceil_div(c.gw, tc::TILE_W) * ceil_div(c.gh, tc::TILE_H) * c.batch);
  if (c.taps.nphase > 1 && bn0 > 128) bn0 = 128;
  // LFP_TC_TALL: 0 = off, 1 (default) = launches whose slice is 128 wide, 2 = 256-wide slices are also split into two 128-wide
  // tall ones.  Measured at 1024 px, B = 20, same box: 128 -> 128 at 256 px 742 -> 615 us (forward) / 714 -> 598 us (data
  // gradient) with tall items.  A 256-wide tall item needs all 512 TMEM columns for its two accumulators, which serialises its
  // epilogue with the next item's MMAs (+8-17 %); splitting it into 128-wide tall items cuts the L2 -> SM traffic per
  // (pixel x channel) by 40 % but measured 3-10 % slower on five of the six 256 / 512-channel layers - at 730 TFLOP/s those are
  // already above the sustained tensor peak of MEASURED_PEAKS.json, i.e. bound by power, not by L2 - so mode 2 stays off.
  static const int tall_mode = getenv("LFP_TC_TALL") != nullptr ? atoi(getenv("LFP_TC_TALL")) : 1;
  int nhalf = 1;
  {
    const int nph = c.taps.nphase > 0 ? c.taps.nphase : 1;
    const int64_t rows16 = ceil_div(c.gh, 16) * 16, rows32 = ceil_div(c.gh, 32) * 32;
    const int64_t tall_items = ceil_div(c.gw, tc::TILE_W) * ceil_div(c.gh, 32) * c.batch * (c.N / 128);
    const size_t w_bytes = (size_t)c.taps.group_tap0[c.taps.ngroups] * (c.K / 32) * 128 * 128;
    const bool ok128 = tall_mode > 0 && tc_use_e2(128) && nph == 1 && c.e.rgb_out == nullptr && c.N % 128 == 0 &&
                       rows32 * 100 <= rows16 * 107 && tall_items >= 4 * (int64_t)num_sms() && w_bytes > 160u * 1024u;
    if (ok128 && (bn0 == 128 || (bn0 == 256 && tall_mode >= 2))) { bn0 = 128; nhalf = 2; }
  }
  alignas(64) CUtensorMap tmA;
  const int nb = c.in_bcast ? 1 : c.batch;
  const cuuint64_t dims[5] = {(cuuint64_t)c.K, (cuuint64_t)c.in_w, (cuuint64_t)c.in_h, (cuuint64_t)c.in_planes, (cuuint64_t)nb};
  const cuuint64_t strides[4] = {(cuuint64_t)c.K * 4, (cuuint64_t)c.in_w * c.K * 4, (cuuint64_t)c.in_h * c.in_w * c.K * 4,
                                 (cuuint64_t)c.in_planes * c.in_h * c.in_w * c.K * 4};
  const cuuint32_t box[5] = {32, tc::HALO_W, (cuuint32_t)(tc::TILE_H * nhalf + 2), 1, 1};
  LFP_TRY(tc::encode(&tmA, c.in, 5, dims, strides, box));
  tc::Args a{};
  a.batch = c.batch; a.gh = c.gh; a.gw = c.gw;
  a.nhalf = nhalf;
  a.a_rows = tc::HALO_W * (tc::TILE_H * nhalf + 2);
  a.a_stage = (a.a_rows * 128 + 1023) / 1024 * 1024;
  a.tiles_x = (int)ceil_div(c.gw, tc::TILE_W); a.tiles_y = (int)ceil_div(c.gh, tc::TILE_H * nhalf);
  a.vtiles_per = a.tiles_x * (int)ceil_div(c.gh, tc::TILE_H);
  a.K = c.K; a.N = c.N;
  a.BN = bn0;   // (fused phases / fused ToRGB keep all channels in one CTA)
  LFP_CHECK_ARG(c.e.rgb_out == nullptr || a.BN == c.N, "conv_tc: the fused ToRGB epilogue needs all %d channels in one CTA", c.N);
  a.in_bcast = c.in_bcast ? 1 : 0;
  a.taps = c.taps;
  if (a.taps.nphase <= 0) a.taps.nphase = 1;
  a.out = c.out; a.out_planes = c.out_planes; a.out_plane = c.out_plane; a.out_h = c.out_h; a.out_w = c.out_w;
  a.out_stride = c.out_stride > 0 ? c.out_stride : 1; a.out_oy = c.out_oy; a.out_ox = c.out_ox;
  a.mod = c.mod; a.e = c.e;
  a.n_ntiles = c.N / a.BN;
  a.d_ntiles = tc::make_fastdiv(a.n_ntiles); a.d_tiles_per = tc::make_fastdiv(a.tiles_x * a.tiles_y); a.d_tiles_x = tc::make_fastdiv(a.tiles_x);
  static const bool plain_div = getenv("LFP_TC_FASTDIV") != nullptr && atoi(getenv("LFP_TC_FASTDIV")) == 0;   // A/B switch
  if (plain_div) a.d_ntiles.M = a.d_tiles_per.M = a.d_tiles_x.M = 0;
  const int64_t total = (int64_t)a.tiles_x * a.tiles_y * c.batch * a.n_ntiles;
  LFP_CHECK_ARG(total < (1ll << 31), "conv_tc: too many tiles");
  a.total_work = (int)total;
  const int ntaps = c.taps.group_tap0[c.taps.ngroups];
  const CUtensorMap& tmB = *reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const unsigned char*>(c.wmap) + 128 * tc_bn_index(a.BN));
  const bool mod = c.mod != nullptr;
  if (c.epi == EPI_ACT) {
    // the noise tile of every work item by TMA (PXS): needs a 16-byte aligned map with rows that are multiples of 16 bytes
    alignas(64) CUtensorMap tmNz;
    static const bool pxs_off = getenv("LFP_TC_PXS") != nullptr && atoi(getenv("LFP_TC_PXS")) == 0;
    static const bool pxs_fwd_off = getenv("LFP_TC_PXS_FWD") != nullptr && atoi(getenv("LFP_TC_PXS_FWD")) == 0;
    const bool nb1 = c.e.noise_bstride == 0;
    bool ok = !pxs_off && !pxs_fwd_off && c.e.rgb_out == nullptr && c.e.noise != nullptr && (c.gw % 4) == 0 && ((uintptr_t)c.e.noise & 15) == 0 &&
              (nb1 || c.e.noise_bstride == (int64_t)c.gh * c.gw) && c.e.demod != nullptr && c.e.bias != nullptr &&
              (((uintptr_t)c.e.demod | (uintptr_t)c.e.bias) & 15) == 0;
    if (ok) {
      const cuuint64_t nd[3] = {(cuuint64_t)c.gw, (cuuint64_t)c.gh, (cuuint64_t)(nb1 ? 1 : c.batch)};
      const cuuint64_t ns[2] = {(cuuint64_t)c.gw * 4, (cuuint64_t)c.gh * c.gw * 4};
      const cuuint32_t nbx[3] = {tc::TILE_W, (cuuint32_t)(tc::TILE_H * nhalf), 1};
      ok = tc::encode_plain(&tmNz, c.e.noise, 3, nd, ns, nbx) == 0;
    }
    a.px_ok = ok ? 1 : 0;
    return mod ? tc_launch<EPI_ACT, true>(tmA, tmB, tmA, a, ntaps, s, ok ? &tmNz : nullptr) : tc_launch<EPI_ACT, false>(tmA, tmB, tmA, a, ntaps, s, ok ? &tmNz : nullptr);
  }
  if (c.epi == EPI_STORE) return mod ? tc_launch<EPI_STORE, true>(tmA, tmB, tmA, a, ntaps, s) : tc_launch<EPI_STORE, false>(tmA, tmB, tmA, a, ntaps, s);
  LFP_CHECK_ARG(!mod, "conv_tc: this epilogue takes an unmodulated input");
  if (c.epi == EPI_RELU) return tc_launch<EPI_RELU, false>(tmA, tmB, tmA, a, ntaps, s);
  // saved forward input [B or 1, gh, gw, N]: 128-pixel x 32-channel boxes for the epilogue
  alignas(64) CUtensorMap tmX;
  a.xs_bcast = c.e.xsave_bstride == 0 ? 1 : 0;
  {
    const int xb = a.xs_bcast ? 1 : c.batch;
    const cuuint64_t xd[5] = {(cuuint64_t)c.N, (cuuint64_t)c.gw, (cuuint64_t)c.gh, 1, (cuuint64_t)xb};
    const cuuint64_t xst[4] = {(cuuint64_t)c.N * 4, (cuuint64_t)c.gw * c.N * 4, (cuuint64_t)c.gh * c.gw * c.N * 4, (cuuint64_t)c.gh * c.gw * c.N * 4};
    const cuuint32_t xbox[5] = {32, tc::TILE_W, tc::TILE_H, 1, 1};
    LFP_TRY(tc::encode(&tmX, c.e.xsave, 5, xd, xst, xbox));
  }
  if (c.epi == EPI_DGRAD_ACT) {
    // per-pixel scalars (noise, skip-image gradient) as TMA tiles: needs 16-byte aligned planes and rows
    alignas(64) CUtensorMap tmNz, tmRg;
    static const bool pxs_off = getenv("LFP_TC_PXS") != nullptr && atoi(getenv("LFP_TC_PXS")) == 0;
    const bool nb1 = c.e.noise_bstride == 0;
    bool ok = !pxs_off && c.e.noise != nullptr && (c.gw % 4) == 0 && ((uintptr_t)c.e.noise & 15) == 0 &&
              (c.e.drgb == nullptr || ((uintptr_t)c.e.drgb & 15) == 0) && (nb1 || c.e.noise_bstride == (int64_t)c.gh * c.gw) &&
              c.e.demod != nullptr && c.e.bias != nullptr && c.e.mod_out != nullptr &&
              (((uintptr_t)c.e.demod | (uintptr_t)c.e.bias | (uintptr_t)c.e.mod_out) & 15) == 0;
    if (ok) {
      const cuuint64_t nd[3] = {(cuuint64_t)c.gw, (cuuint64_t)c.gh, (cuuint64_t)(nb1 ? 1 : c.batch)};
      const cuuint64_t ns[2] = {(cuuint64_t)c.gw * 4, (cuuint64_t)c.gh * c.gw * 4};
      const cuuint32_t nbx[3] = {tc::TILE_W, (cuuint32_t)(tc::TILE_H * nhalf), 1};
      ok = tc::encode_plain(&tmNz, c.e.noise, 3, nd, ns, nbx) == 0;
    }
    if (ok && c.e.drgb != nullptr) {
      const cuuint64_t rd[4] = {(cuuint64_t)c.gw, (cuuint64_t)c.gh, 3, (cuuint64_t)c.batch};
      const cuuint64_t rs[3] = {(cuuint64_t)c.gw * 4, (cuuint64_t)c.gh * c.gw * 4, (cuuint64_t)3 * c.gh * c.gw * 4};
      const cuuint32_t rbx[4] = {tc::TILE_W, (cuuint32_t)(tc::TILE_H * nhalf), 3, 1};
      ok = tc::encode_plain(&tmRg, c.e.drgb, 4, rd, rs, rbx) == 0;
    }
    a.px_ok = ok ? 1 : 0;
    return tc_launch<EPI_DGRAD_ACT, false>(tmA, tmB, tmX, a, ntaps, s, ok ? &tmNz : nullptr, (ok && c.e.drgb != nullptr) ? &tmRg : nullptr);
  }
  if (c.epi == EPI_DGRAD_RELU) return tc_launch<EPI_DGRAD_RELU, false>(tmA, tmB, tmX, a, ntaps, s);
  return tc_launch<EPI_DGRAD, false>(tmA, tmB, tmX, a, ntaps, s);
}

}  // namespace lfp
