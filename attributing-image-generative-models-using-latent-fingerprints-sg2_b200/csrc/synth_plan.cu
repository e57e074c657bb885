// Whole-synthesis plan: owns the prepared generator weights and sequences the kernels of one
// forward / backward pass (the additive C-ABI group 3 of include/lfp_sg2.h).
//
// Replaces the module chain of Generator.forward(input_is_latent=True), src/model.py:551-566.
// Layer order, channel table and latent-slot indexing follow src/model.py:418-474, 551-564.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "synth_kernels.cuh"

#define LFP_PROF(h, kind, flops, bytes, s, call)            \
  do {                                                      \
    const bool _p = (h)->prof_begin(kind, flops, bytes, s); \
    const int _r = (call);                                  \
    (h)->prof_end(_p, s);                                   \
    if (_r != 0) return _r;                                 \
    if ((h)->debug_sync) {                                  \
      fprintf(stderr, "[lfp] kind %d flops %.3g bytes %.3g ...", (int)(kind), (double)(flops), (double)(bytes)); \
      const cudaError_t _e = cudaStreamSynchronize(s);      \
      fprintf(stderr, " %s\n", cudaGetErrorString(_e));     \
    }                                                       \
  } while (0)

namespace lfp {

static double conv_flops(const ConvGeom& g) { return 2.0 * g.batch * g.gh * g.gw * g.ntaps * (double)g.K * g.N; }
static double conv_bytes(const ConvGeom& g) {
  // algorithmic: the input window the grid touches once + the output once + the taps used
  const double in_px = (double)g.batch * (g.in_bstride == 0 ? 1.0 / g.batch : 1.0) * g.in_h * g.in_w;
  return 4.0 * (in_px * g.K + (double)g.batch * g.gh * g.gw * g.N + (double)g.ntaps * g.K * g.N);
}

struct alignas(64) TmapBuf { unsigned char bytes[4 * 128]; };  // four CUtensorMaps (weight-slice widths 256/128/64/32)

struct ConvLayer {
  std::string name;
  int cin = 0, cout = 0, res_in = 0, res_out = 0;
  bool up = false;
  int slot = 0, noise_idx = 0;
  float *W = nullptr, *modw = nullptr, *modb = nullptr, *noise_w = nullptr, *act_bias = nullptr;
  float *wf = nullptr, *wg = nullptr, *wsq = nullptr;
  float *wf_t = nullptr, *wg_t = nullptr;   // tf32-rounded copies for the tensor-core path
  TmapBuf map_fwd, map_bwd;                 // weight tensor maps: forward reads wg_t [t][co][ci], dgrad wf_t [t][ci][co]
  bool tc_fwd = false, tc_bwd = false;
  int row0 = 0;       // first row in the stacked modulation matrix
  int demod_off = 0;  // offset into the demod table (units of cout, times batch at run time)
};
struct RgbLayer {
  std::string name;
  int cin = 0, res = 0, slot = 0;
  float *W = nullptr, *modw = nullptr, *modb = nullptr, *bias = nullptr, *wrgb = nullptr;
  int row0 = 0;
};

// Workspace layout for one batch size (float offsets).  Full layout: every StyledConv output is kept for the backward,
// and every layer owns its partial-sum regions so that all reductions of a backward pass can run in one launch at its end.
// Forward-only layout (generation, no backward): activations and skip images ping-pong between two buffers.
struct Layout {
  size_t s_all = 0, d_all = 0, scratchT = 0, bufA = 0, bufB = 0, dskipA = 0, dskipB = 0, T_all = 0, R1_all = 0, ds_all = 0, total = 0;
  std::vector<size_t> act_off, skip_off;      // per conv layer / per ToRGB layer
  std::vector<size_t> pX_off, pT_off, pR_off; // per conv layer: dgrad style partials, act-backward partials (T and ToRGB R)
};

// what one backward pass does per conv layer for a given precision (decided once, used by the launch loop and by the
// builder of the batched-reduction tables)
struct BwdStep { bool use_tc = false, fuse = false; int Qx = 0, QT = 0; };

// device tables of the batched kernels for one (batch, precision)
struct Tables {
  ReduceItem* ritems = nullptr; int2* rblocks = nullptr; int nrblocks = 0;
  GradItem* gitems = nullptr; int2* gblocks = nullptr; int ngblocks = 0;
  DemodItem* ditems = nullptr; int2* dblocks = nullptr; int ndblocks = 0;
};

// forward bookkeeping per workspace: the backward validates against the workspace it is given
struct FwdRec { int batch = -1, precision = -1; bool fwd_only = false; std::vector<const float*> noise; std::vector<int> noise_batch; };

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace lfp

using namespace lfp;

struct lfp_synth {
  int size = 0, log_size = 0, style_dim = 0, cm = 2, n_latent = 0, num_noise = 0;
  std::vector<ConvLayer> convs;  // [0]=conv1, [1+2j]=convs.2j (up), [2+2j]=convs.2j+1
  std::vector<RgbLayer> rgbs;    // [0]=to_rgb1, [1+j]=to_rgbs.j
  float* const_nchw = nullptr;
  float* const_nhwc = nullptr;
  float *A_all = nullptr, *b_all = nullptr;
  int *row_slot = nullptr, *row_base = nullptr, *row_cin = nullptr, *slot_begin = nullptr, *slot_end = nullptr;
  int rows = 0, demod_total = 0;
  float* fir = nullptr;  // 4 x 16 floats: blur fwd coef, blur bwd coef, upsample kernel, flipped upsample kernel
  float blur1d[4] = {1, 3, 3, 1};
  bool finalized = false;
  int tc_min_res = 4;
  bool debug_sync = false;   // env LFP_DEBUG_SYNC=1: synchronise and log after every profiled launch
  int fuse_phases_max_c = 128;   // env LFP_FUSE_PHASES_MAXC
  // wider layers are fused too while the whole layer is a single wave of work items (input <= 8 px): one launch instead of four
  // latency-bound ones (4 -> 8 px: 136 -> 72 us, 8 -> 16 px: 138 -> 84 us at B = 20).  From 16 px on the four accumulators of a
  // 128-wide slice fill the TMEM, the epilogue no longer overlaps the next item's MMAs, and the separate launches win.
  int fuse_phases_max_res = 8;   // env LFP_FUSE_PHASES_MAXRES
  bool fuse_phases = true;   // one launch for the four sub-pixel phases of the C <= 64 transposed convs (env LFP_FUSE_PHASES=0 disables)
  bool fuse_rgb = false;     // ToRGB inside the forward conv epilogue on the tensor-core path (env LFP_FUSE_RGB=1 enables;
                             // measured slower than the separate kernel: the epilogue is the longer pole at N <= 64)
  bool fuse_actbwd = true;   // run act_bwd inside the upstream dgrad epilogue on the tensor-core path (env LFP_FUSE_ACTBWD=0 disables)   // smallest output grid the tensor-core kernel is used for (env LFP_TC_MIN_RES)
  std::mutex mu;                              // guards the caches below (a plan may be shared by host threads)
  std::map<const void*, FwdRec> fwd_recs;     // last forward per workspace
  std::map<long long, Tables> tables;         // key = batch * 8 + precision * 2 + forward-only
  std::vector<void*> owned;
  // grow-only device buffers of the host-pointer entry point (no allocation per call)
  struct HostBufs { void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; size_t cap[5] = {0, 0, 0, 0, 0}; std::vector<void*> noise; std::vector<size_t> noise_cap; } hb;

  // optional per-kernel-class timing with CUDA events on the launching stream
  bool prof_on = false;
  int prof_mask = 0;
  std::vector<cudaEvent_t> prof_ev;       // pairs
  std::vector<int> prof_kind;
  std::vector<double> prof_lflops, prof_lbytes;   // per recorded launch
  std::vector<float> prof_lms;
  std::vector<int> prof_lkind;
  size_t prof_used = 0;
  double prof_flops[LFP_KIND_COUNT] = {0}, prof_bytes[LFP_KIND_COUNT] = {0};
  int64_t prof_launches[LFP_KIND_COUNT] = {0};

  ~lfp_synth() {
    for (void* p : owned) cudaFree(p);
    for (int i = 0; i < 5; ++i) if (hb.p[i]) cudaFree(hb.p[i]);
    for (void* p : hb.noise) if (p) cudaFree(p);
    for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
  }
  bool prof_begin(int kind, double flops, double bytes, cudaStream_t s) {
    if (!prof_on || !((prof_mask >> kind) & 1)) return false;
    prof_flops[kind] += flops; prof_bytes[kind] += bytes; prof_launches[kind] += 1;
    if (prof_used + 2 > prof_ev.size()) {
      if (prof_ev.size() >= 400000) return false;
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return false; prof_ev.push_back(e); }
    }
    if (prof_kind.size() < prof_ev.size() / 2) { prof_kind.resize(prof_ev.size() / 2); prof_lflops.resize(prof_ev.size() / 2); prof_lbytes.resize(prof_ev.size() / 2); }
    prof_kind[prof_used / 2] = kind;
    prof_lflops[prof_used / 2] = flops; prof_lbytes[prof_used / 2] = bytes;
    cudaEventRecord(prof_ev[prof_used], s);
    return true;
  }
  void prof_end(bool started, cudaStream_t s) {
    if (!started) return;
    cudaEventRecord(prof_ev[prof_used + 1], s);
    prof_used += 2;
  }
  int alloc(float** p, size_t n) {
    LFP_CUDA(cudaMalloc((void**)p, n * sizeof(float)));
    owned.push_back(*p);
    return 0;
  }
  int alloc_i(int** p, size_t n) {
    LFP_CUDA(cudaMalloc((void**)p, n * sizeof(int)));
    owned.push_back(*p);
    return 0;
  }
  int channels(int res) const {
    // src/model.py:418-428
    switch (res) {
      case 4: case 8: case 16: case 32: return 512;
      case 64: return 256 * cm;
      case 128: return 128 * cm;
      case 256: return 64 * cm;
      case 512: return 32 * cm;
      case 1024: return 16 * cm;
    }
    return 0;
  }
  Layout layout(int B, bool fwd_only = false) const;
  std::vector<BwdStep> bwd_steps(int precision) const;
  int get_tables(int B, int precision, const Layout& L, const std::vector<BwdStep>& st, const Tables** out);
};

Layout lfp_synth::layout(int B, bool fwd_only) const {
  Layout L;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 64); return o; };
  L.s_all = take((size_t)B * rows);
  L.d_all = take((size_t)B * demod_total);
  size_t max_act = 0, max_T = 0;
  for (const ConvLayer& c : convs) {
    const size_t n = (size_t)B * c.res_out * c.res_out * c.cout;
    if (n > max_act) max_act = n;
    const size_t nin = (size_t)B * c.res_in * c.res_in * c.cin;
    if (nin > max_act) max_act = nin;
    if (c.up) {
      // interleaved [2H+1, 2H+1] (fp32 path, fused phases) or phase-major [4, H+1, H+1] (tensor-core path)
      const size_t t = (size_t)B * (c.res_out + 2) * (c.res_out + 2) * c.cout;
      if (t > max_T) max_T = t;
    }
  }
  L.scratchT = take(max_T);
  L.bufA = take(max_act);
  L.bufB = take(max_act);
  L.act_off.resize(convs.size());
  L.skip_off.resize(rgbs.size());
  if (fwd_only) {
    // layer li reads the output of li - 1 and writes its own: two buffers; the skip image likewise
    for (size_t li = 0; li < convs.size(); ++li) L.act_off[li] = (li & 1) ? L.bufB : L.bufA;
    const size_t full = (size_t)B * 3 * size * size;
    const size_t s0 = take(full), s1 = take(full);
    for (size_t ri = 0; ri < rgbs.size(); ++ri) L.skip_off[ri] = (ri & 1) ? s1 : s0;
    L.total = off;
    return L;
  }
  for (size_t li = 0; li < convs.size(); ++li)
    L.act_off[li] = take((size_t)B * convs[li].res_out * convs[li].res_out * convs[li].cout);
  for (size_t ri = 0; ri < rgbs.size(); ++ri) L.skip_off[ri] = take((size_t)B * 3 * rgbs[ri].res * rgbs[ri].res);
  const size_t half = (size_t)B * 3 * (size / 2) * (size / 2);
  L.dskipA = take(half);
  L.dskipB = take(half);
  L.pX_off.resize(convs.size()); L.pT_off.resize(convs.size()); L.pR_off.resize(convs.size());
  for (size_t li = 0; li < convs.size(); ++li) {
    const ConvLayer& c = convs[li];
    const int hw = c.res_out * c.res_out, hwi = c.res_in * c.res_in;
    // rows per sample: the larger of what the CUDA-core and the tensor-core kernels write
    int qx = hwi / (hwi < 128 ? hwi : 128);
    const int qx_tc = tc_tiles_per_sample(c.res_in, c.res_in);
    if (qx_tc > qx) qx = qx_tc;
    int qt = hw / actbwd_seglen(hw, c.cout);
    const int qt_tc = tc_tiles_per_sample(c.res_out, c.res_out);
    if (qt_tc > qt) qt = qt_tc;
    L.pX_off[li] = take((size_t)B * qx * c.cin);
    L.pT_off[li] = take((size_t)B * qt * c.cout);
    L.pR_off[li] = (li == 0 || (li % 2) == 0) ? take((size_t)B * qt * c.cout) : 0;
  }
  L.T_all = take((size_t)B * demod_total);
  L.R1_all = take((size_t)B * rows);
  L.ds_all = take((size_t)B * rows);
  L.total = off;
  return L;
}

std::vector<BwdStep> lfp_synth::bwd_steps(int precision) const {
  std::vector<BwdStep> st(convs.size());
  for (size_t li = 0; li < convs.size(); ++li) {
    const ConvLayer& c = convs[li];
    const int hwi = c.res_in * c.res_in;
    st[li].use_tc = precision == LFP_PREC_TF32 && c.tc_bwd && c.res_in >= tc_min_res;
    st[li].fuse = st[li].use_tc && li >= 1 && fuse_actbwd;
    st[li].Qx = st[li].use_tc ? tc_tiles_per_sample(c.res_in, c.res_in) : hwi / (hwi < 128 ? hwi : 128);
  }
  for (size_t li = 0; li < convs.size(); ++li) {
    const ConvLayer& c = convs[li];
    const int hw = c.res_out * c.res_out;
    const bool from_above = li + 1 < convs.size() && st[li + 1].fuse;   // the dgrad of the layer above ran this layer's act backward
    st[li].QT = from_above ? st[li + 1].Qx : hw / actbwd_seglen(hw, c.cout);
  }
  return st;
}

template <class T>
static int upload(std::vector<void*>& owned, const std::vector<T>& v, T** out) {
  *out = nullptr;
  if (v.empty()) return 0;
  LFP_CUDA(cudaMalloc((void**)out, v.size() * sizeof(T)));
  owned.push_back(*out);
  LFP_CUDA(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int lfp_synth::get_tables(int B, int precision, const Layout& L, const std::vector<BwdStep>& st, const Tables** out) {
  std::lock_guard<std::mutex> lock(mu);
  const long long key = (long long)B * 8 + precision * 2 + (L.pX_off.empty() ? 1 : 0);
  auto it = tables.find(key);
  if (it != tables.end()) { *out = &it->second; return 0; }
  std::vector<ReduceItem> ri; std::vector<int2> rb;
  std::vector<GradItem> gi; std::vector<int2> gb;
  std::vector<DemodItem> di; std::vector<int2> db;
  auto add_reduce = [&](size_t src, size_t dst, int Q, int C) {
    ri.push_back(ReduceItem{(int64_t)src, (int64_t)dst, Q, C});
    for (int cb = 0; cb < (C + 31) / 32; ++cb) rb.push_back(make_int2((int)ri.size() - 1, cb));
  };
  for (size_t li = 0; li < convs.size(); ++li) {
    const ConvLayer& c = convs[li];
    di.push_back(DemodItem{(int64_t)(L.s_all + (size_t)B * c.row0), (int64_t)(L.d_all + (size_t)B * c.demod_off), c.wsq, c.cin, c.cout});
    for (int cb = 0; cb < (c.cout + 7) / 8; ++cb) db.push_back(make_int2((int)di.size() - 1, cb));
    if (L.pX_off.empty()) continue;   // forward-only layout has no backward tables
    add_reduce(L.pX_off[li], L.R1_all + (size_t)B * c.row0, st[li].Qx, c.cin);
    add_reduce(L.pT_off[li], L.T_all + (size_t)B * c.demod_off, st[li].QT, c.cout);
    if (li == 0 || (li % 2) == 0) add_reduce(L.pR_off[li], L.ds_all + (size_t)B * rgbs[li / 2].row0, st[li].QT, c.cout);
    gi.push_back(GradItem{(int64_t)(L.R1_all + (size_t)B * c.row0), (int64_t)(L.s_all + (size_t)B * c.row0),
                          (int64_t)(L.T_all + (size_t)B * c.demod_off), (int64_t)(L.d_all + (size_t)B * c.demod_off),
                          (int64_t)(L.ds_all + (size_t)B * c.row0), c.wsq, c.cin, c.cout});
    for (int cb = 0; cb < (c.cin + 31) / 32; ++cb) gb.push_back(make_int2((int)gi.size() - 1, cb));
  }
  Tables t;
  LFP_TRY(upload(owned, ri, &t.ritems)); LFP_TRY(upload(owned, rb, &t.rblocks)); t.nrblocks = (int)rb.size();
  LFP_TRY(upload(owned, gi, &t.gitems)); LFP_TRY(upload(owned, gb, &t.gblocks)); t.ngblocks = (int)gb.size();
  LFP_TRY(upload(owned, di, &t.ditems)); LFP_TRY(upload(owned, db, &t.dblocks)); t.ndblocks = (int)db.size();
  *out = &(tables[key] = t);
  return 0;
}

extern "C" int lfp_synth_create(lfp_synth** out, int size, int style_dim, int channel_multiplier,
                                const float* blur_kernel_1d, int blur_taps) {
  LFP_CHECK_ARG(out != nullptr, "synth_create: null out");
  const int ls = (int)lround(log2((double)size));
  LFP_CHECK_ARG(size >= 8 && size <= 1024 && (1 << ls) == size, "synth_create: size %d must be a power of two in [8,1024]", size);
  LFP_CHECK_ARG(style_dim >= 4 && style_dim % 4 == 0, "synth_create: style_dim %d must be a multiple of 4", style_dim);
  LFP_CHECK_ARG(channel_multiplier >= 1, "synth_create: bad channel multiplier");
  if (blur_kernel_1d != nullptr && blur_taps != 4) {
    set_error("synth_create: only 4-tap blur kernels are supported by the fused path (got %d)", blur_taps);
    return LFP_EUNSUPPORTED;
  }
  lfp_synth* h = new lfp_synth();
  h->size = size; h->log_size = ls; h->style_dim = style_dim; h->cm = channel_multiplier;
  h->n_latent = ls * 2 - 2;
  h->num_noise = (ls - 2) * 2 + 1;
  if (blur_kernel_1d) memcpy(h->blur1d, blur_kernel_1d, 4 * sizeof(float));
  if (const char* e = getenv("LFP_DEBUG_SYNC")) h->debug_sync = atoi(e) != 0;
  if (const char* e = getenv("LFP_FUSE_PHASES")) h->fuse_phases = atoi(e) != 0;
  if (const char* e = getenv("LFP_FUSE_PHASES_MAXC")) h->fuse_phases_max_c = atoi(e);
  if (const char* e = getenv("LFP_FUSE_PHASES_MAXRES")) h->fuse_phases_max_res = atoi(e);
  if (const char* e = getenv("LFP_FUSE_RGB")) h->fuse_rgb = atoi(e) != 0;
  if (const char* e = getenv("LFP_FUSE_ACTBWD")) h->fuse_actbwd = atoi(e) != 0;
  if (const char* e = getenv("LFP_TC_MIN_RES")) { const int v = atoi(e); if (v >= 4) h->tc_min_res = v; }
  for (int r = 4; r <= size; r *= 2) {
    const int c = h->channels(r);
    if (c % 16 != 0 || 1024 % c != 0) {
      set_error("synth_create: %d channels at %d px not supported (need a power of two >= 16)", c, r);
      delete h;
      return LFP_EUNSUPPORTED;
    }
  }
  ConvLayer c1;
  c1.name = "conv1"; c1.cin = c1.cout = h->channels(4); c1.res_in = c1.res_out = 4; c1.slot = 0; c1.noise_idx = 0;
  h->convs.push_back(c1);
  RgbLayer r1;
  r1.name = "to_rgb1"; r1.cin = h->channels(4); r1.res = 4; r1.slot = 1;
  h->rgbs.push_back(r1);
  int cin = h->channels(4);
  for (int j = 0; j + 3 <= ls; ++j) {
    const int res = 8 << j, cout = h->channels(res);
    ConvLayer u;
    u.name = "convs." + std::to_string(2 * j); u.cin = cin; u.cout = cout; u.res_in = res / 2; u.res_out = res;
    u.up = true; u.slot = 1 + 2 * j; u.noise_idx = 1 + 2 * j;
    ConvLayer p;
    p.name = "convs." + std::to_string(2 * j + 1); p.cin = cout; p.cout = cout; p.res_in = p.res_out = res;
    p.slot = 2 + 2 * j; p.noise_idx = 2 + 2 * j;
    RgbLayer r;
    r.name = "to_rgbs." + std::to_string(j); r.cin = cout; r.res = res; r.slot = 3 + 2 * j;
    h->convs.push_back(u); h->convs.push_back(p); h->rgbs.push_back(r);
    cin = cout;
  }
  // raw parameter storage
  int rc = 0;
  const int c4 = h->channels(4);
  rc |= h->alloc(&h->const_nchw, (size_t)c4 * 16);
  rc |= h->alloc(&h->const_nhwc, (size_t)c4 * 16);
  for (ConvLayer& c : h->convs) {
    rc |= h->alloc(&c.W, (size_t)c.cout * c.cin * 9);
    rc |= h->alloc(&c.modw, (size_t)c.cin * style_dim);
    rc |= h->alloc(&c.modb, c.cin);
    rc |= h->alloc(&c.noise_w, 1);
    rc |= h->alloc(&c.act_bias, c.cout);
    rc |= h->alloc(&c.wf, (size_t)9 * c.cin * c.cout);
    rc |= h->alloc(&c.wg, (size_t)9 * c.cin * c.cout);
    rc |= h->alloc(&c.wsq, (size_t)c.cin * c.cout);
    rc |= h->alloc(&c.wf_t, (size_t)9 * c.cin * c.cout);
    rc |= h->alloc(&c.wg_t, (size_t)9 * c.cin * c.cout);
  }
  for (RgbLayer& r : h->rgbs) {
    rc |= h->alloc(&r.W, (size_t)3 * r.cin);
    rc |= h->alloc(&r.modw, (size_t)r.cin * style_dim);
    rc |= h->alloc(&r.modb, r.cin);
    rc |= h->alloc(&r.bias, 3);
    rc |= h->alloc(&r.wrgb, (size_t)3 * r.cin);
  }
  // modulation rows, ordered by latent slot so every slot owns one contiguous row range
  std::vector<int> row_slot, row_base, row_cin, sb(h->n_latent, 0), se(h->n_latent, 0);
  int rows = 0, demod = 0;
  for (int slot = 0; slot < h->n_latent; ++slot) {
    sb[slot] = rows;
    for (RgbLayer& r : h->rgbs)
      if (r.slot == slot) { r.row0 = rows; for (int i = 0; i < r.cin; ++i) { row_slot.push_back(slot); row_base.push_back(r.row0); row_cin.push_back(r.cin); } rows += r.cin; }
    for (ConvLayer& c : h->convs)
      if (c.slot == slot) { c.row0 = rows; for (int i = 0; i < c.cin; ++i) { row_slot.push_back(slot); row_base.push_back(c.row0); row_cin.push_back(c.cin); } rows += c.cin; }
    se[slot] = rows;
  }
  for (ConvLayer& c : h->convs) { c.demod_off = demod; demod += c.cout; }
  h->rows = rows; h->demod_total = demod;
  rc |= h->alloc(&h->A_all, (size_t)rows * style_dim);
  rc |= h->alloc(&h->b_all, rows);
  rc |= h->alloc_i(&h->row_slot, rows);
  rc |= h->alloc_i(&h->row_base, rows);
  rc |= h->alloc_i(&h->row_cin, rows);
  rc |= h->alloc_i(&h->slot_begin, h->n_latent);
  rc |= h->alloc_i(&h->slot_end, h->n_latent);
  rc |= h->alloc(&h->fir, 80);
  if (rc != 0) { delete h; return rc; }
  auto up = [&](void* d, const void* s, size_t n) { return cudaMemcpy(d, s, n, cudaMemcpyHostToDevice); };
  cudaError_t e = up(h->row_slot, row_slot.data(), rows * sizeof(int));
  if (e == cudaSuccess) e = up(h->row_base, row_base.data(), rows * sizeof(int));
  if (e == cudaSuccess) e = up(h->row_cin, row_cin.data(), rows * sizeof(int));
  if (e == cudaSuccess) e = up(h->slot_begin, sb.data(), h->n_latent * sizeof(int));
  if (e == cudaSuccess) e = up(h->slot_end, se.data(), h->n_latent * sizeof(int));
  // FIR tables.  k2 = outer(k,k)/sum * 4 is both Blur(upsample_factor=2) (src/model.py:75-86) and
  // Upsample (src/model.py:37-38).
  float k2[16], tab[80];
  float sum = 0.f;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { k2[i * 4 + j] = h->blur1d[i] * h->blur1d[j]; sum += k2[i * 4 + j]; }
  for (int i = 0; i < 16; ++i) k2[i] = k2[i] / sum * 4.f;
  for (int ty = 0; ty < 4; ++ty)
    for (int tx = 0; tx < 4; ++tx) {
      tab[0 + ty * 4 + tx] = k2[(3 - ty) * 4 + (3 - tx)];   // blur forward: flipped taps, pad 1
      tab[16 + ty * 4 + tx] = k2[ty * 4 + tx];              // blur adjoint: un-flipped taps, pad 2
      tab[32 + ty * 4 + tx] = k2[ty * 4 + tx];              // upsample kernel as stored by the module
      tab[48 + ty * 4 + tx] = k2[(3 - ty) * 4 + (3 - tx)];  // flipped, for the skip-upsample adjoint
    }
  {
    // separable factors of the same tables: k2 = outer(k, k) / sum(k)^2 * 4 = outer(g, g), g = 2 k / sum(k)
    float s1 = 0.f;
    for (int i = 0; i < 4; ++i) s1 += h->blur1d[i];
    for (int i = 0; i < 4; ++i) {
      tab[64 + i] = 2.f * h->blur1d[3 - i] / s1;   // blur forward (flipped taps)
      tab[68 + i] = 2.f * h->blur1d[i] / s1;       // blur adjoint (un-flipped taps)
    }
  }
  if (e == cudaSuccess) e = up(h->fir, tab, sizeof(tab));
  if (e != cudaSuccess) { set_error("synth_create: %s", cudaGetErrorString(e)); delete h; return (int)e; }
  *out = h;
  return 0;
}

extern "C" void lfp_synth_destroy(lfp_synth* h) { delete h; }
extern "C" int lfp_synth_n_latent(const lfp_synth* h) { return h ? h->n_latent : 0; }
extern "C" int lfp_synth_num_noise(const lfp_synth* h) { return h ? h->num_noise : 0; }

extern "C" int lfp_synth_set_param(lfp_synth* h, const char* name, const float* data, int64_t numel, void* stream) {
  LFP_CHECK_ARG(h && name && data, "synth_set_param: null argument");
  const std::string n(name);
  float* dst = nullptr;
  int64_t want = -1;
  auto match = [&](const std::string& prefix, const char* suffix, float* p, int64_t cnt) {
    if (n == prefix + suffix) { dst = p; want = cnt; }
  };
  if (n == "input.input") { dst = h->const_nchw; want = (int64_t)h->channels(4) * 16; }
  for (ConvLayer& c : h->convs) {
    match(c.name, ".conv.weight", c.W, (int64_t)c.cout * c.cin * 9);
    match(c.name, ".conv.modulation.weight", c.modw, (int64_t)c.cin * h->style_dim);
    match(c.name, ".conv.modulation.bias", c.modb, c.cin);
    match(c.name, ".noise.weight", c.noise_w, 1);
    match(c.name, ".activate.bias", c.act_bias, c.cout);
  }
  for (RgbLayer& r : h->rgbs) {
    match(r.name, ".conv.weight", r.W, (int64_t)3 * r.cin);
    match(r.name, ".conv.modulation.weight", r.modw, (int64_t)r.cin * h->style_dim);
    match(r.name, ".conv.modulation.bias", r.modb, r.cin);
    match(r.name, ".bias", r.bias, 3);
  }
  LFP_CHECK_ARG(dst != nullptr, "synth_set_param: unknown parameter '%s'", name);
  LFP_CHECK_ARG(want == numel, "synth_set_param: '%s' expects %lld elements, got %lld", name, (long long)want, (long long)numel);
  LFP_CUDA(cudaMemcpyAsync(dst, data, numel * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  h->finalized = false;
  return 0;
}

extern "C" int lfp_synth_finalize(lfp_synth* h, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "synth_finalize: null plan");
  cudaStream_t s = (cudaStream_t)stream;
  const float mscale = 1.f / sqrtf((float)h->style_dim);  // EqualLinear scale, lr_mul = 1 (src/model.py:148)
  for (ConvLayer& c : h->convs) {
    const float wscale = 1.f / sqrtf((float)(c.cin * 9));  // src/model.py:208-209
    LFP_TRY(launch_prep_conv3x3(c.W, wscale, c.wf, c.wg, c.wsq, c.cin, c.cout, s));
    LFP_TRY(launch_round_tf32(c.wf, c.wf_t, (int64_t)9 * c.cin * c.cout, s));
    LFP_TRY(launch_round_tf32(c.wg, c.wg_t, (int64_t)9 * c.cin * c.cout, s));
    c.tc_fwd = tc_supported(c.cin, c.cout, 4, 4);
    c.tc_bwd = tc_supported(c.cout, c.cin, 4, 4);
    if (c.tc_fwd) LFP_TRY(tc_make_weight_maps(c.map_fwd.bytes, c.wg_t, 9 * c.cout, c.cin, c.cout));
    if (c.tc_bwd) LFP_TRY(tc_make_weight_maps(c.map_bwd.bytes, c.wf_t, 9 * c.cin, c.cout, c.cin));
    LFP_TRY(launch_scale_copy(c.modw, h->A_all + (size_t)c.row0 * h->style_dim, mscale, (int64_t)c.cin * h->style_dim, s));
    LFP_TRY(launch_scale_copy(c.modb, h->b_all + c.row0, 1.f, c.cin, s));
  }
  for (RgbLayer& r : h->rgbs) {
    LFP_TRY(launch_scale_copy(r.W, r.wrgb, 1.f / sqrtf((float)r.cin), (int64_t)3 * r.cin, s));
    LFP_TRY(launch_scale_copy(r.modw, h->A_all + (size_t)r.row0 * h->style_dim, mscale, (int64_t)r.cin * h->style_dim, s));
    LFP_TRY(launch_scale_copy(r.modb, h->b_all + r.row0, 1.f, r.cin, s));
  }
  LFP_TRY(launch_nchw_to_nhwc(h->const_nchw, h->const_nhwc, 1, h->channels(4), 16, s));
  h->finalized = true;
  return 0;
}

extern "C" size_t lfp_synth_workspace_bytes(const lfp_synth* h, int batch) {
  if (!h || batch <= 0) return 0;
  return h->layout(batch).total * sizeof(float);
}
extern "C" size_t lfp_synth_generate_workspace_bytes(const lfp_synth* h, int batch) {
  if (!h || batch <= 0) return 0;
  return h->layout(batch, true).total * sizeof(float);
}

static void plain_taps(ConvGeom& g) {  // out[y,x] += in[y+ky-1, x+kx-1] * W[ky,kx]
  g.ntaps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) { const int t = ky * 3 + kx; g.dy[t] = (signed char)(ky - 1); g.dx[t] = (signed char)(kx - 1); g.widx[t] = (signed char)t; }
}

static int check_run(const lfp_synth* h, int batch, const void* ws, size_t ws_bytes, int precision, const Layout& L) {
  LFP_CHECK_ARG(h != nullptr && ws != nullptr, "synth: null plan or workspace");
  LFP_CHECK_ARG(batch >= 1 && batch <= 65535, "synth: batch %d out of range", batch);
  if (!h->finalized) { set_error("synth: lfp_synth_finalize has not been called since the last set_param"); return LFP_ESTATE; }
  if (ws_bytes < L.total * sizeof(float)) { set_error("synth: workspace too small (%zu < %zu bytes)", ws_bytes, L.total * sizeof(float)); return LFP_ENOMEM; }
  LFP_CHECK_ARG(((uintptr_t)ws & 255) == 0, "synth: workspace must be 256-byte aligned");
  if (precision != LFP_PREC_FP32 && precision != LFP_PREC_TF32) { set_error("synth: unknown precision mode %d", precision); return LFP_EINVAL; }
  return 0;
}

static int synth_forward_impl(lfp_synth* h, int batch, const float* latent, const float* const* noise,
                              const int* noise_batch, float* image, void* workspace, size_t workspace_bytes,
                              int precision, void* stream, bool fwd_only) {
  LFP_CHECK_ARG(h != nullptr && batch >= 1 && batch <= 65535, "synth_forward: null plan or batch out of range");
  const Layout L = h->layout(batch, fwd_only);
  LFP_TRY(check_run(h, batch, workspace, workspace_bytes, precision, L));
  LFP_CHECK_ARG(latent && noise && noise_batch && image, "synth_forward: null argument");
  for (int i = 0; i < h->num_noise; ++i)
    LFP_CHECK_ARG(noise[i] != nullptr && (noise_batch[i] == 1 || noise_batch[i] == batch), "synth_forward: noise %d must have batch 1 or %d", i, batch);
  cudaStream_t s = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  float* s_all = ws + L.s_all;
  float* d_all = ws + L.d_all;
  const int B = batch;
  const Tables* tb = nullptr;
  LFP_TRY(h->get_tables(B, precision, L, h->bwd_steps(precision), &tb));

  LFP_TRY(launch_style_affine(latent, h->A_all, h->b_all, h->row_slot, h->row_base, h->row_cin, s_all, B, h->rows, h->n_latent, h->style_dim, s));
  LFP_TRY(launch_batched_demod(tb->ditems, tb->dblocks, tb->ndblocks, ws, B, s));   // all 2 log2(size) - 3 layers in one launch

  const float* x = h->const_nhwc;
  int64_t x_bstride = 0;
  for (size_t li = 0; li < h->convs.size(); ++li) {
    const ConvLayer& c = h->convs[li];
    float* act = ws + L.act_off[li];
    const float* smod = s_all + (size_t)B * c.row0;
    const float* dmod = d_all + (size_t)B * c.demod_off;
    const int nb = noise_batch[c.noise_idx];
    const int64_t nstride = nb == 1 ? 0 : (int64_t)c.res_out * c.res_out;
    const bool use_tc = precision == LFP_PREC_TF32 && c.tc_fwd && c.res_in >= h->tc_min_res;
    bool rgb_fused = false;
    if (!c.up) {
      ConvGeom g{};
      g.batch = B; g.gh = g.gw = c.res_out; g.in_h = g.in_w = c.res_in; g.in_bstride = x_bstride; g.in_stride = 1;
      g.out_h = g.out_w = c.res_out; g.out_stride = 1; g.out_oy = g.out_ox = 0; g.K = c.cin; g.N = c.cout;
      plain_taps(g);
      ConvEpiArgs e;
      e.demod = dmod; e.noise = noise[c.noise_idx]; e.noise_bstride = nstride; e.noise_w = c.noise_w; e.bias = c.act_bias;
      if (use_tc) {
        TcConv t{};
        t.in = x; t.in_planes = 1; t.in_h = t.in_w = c.res_in; t.in_bcast = x_bstride == 0; t.mod = smod; t.wmap = c.map_fwd.bytes;
        t.out = act; t.out_planes = 1; t.out_plane = 0; t.out_h = t.out_w = c.res_out;
        t.batch = B; t.gh = t.gw = c.res_out; t.K = c.cin; t.N = c.cout;
        t.taps.ngroups = 1; t.taps.group_plane[0] = 0; t.taps.group_tap0[0] = 0; t.taps.group_tap0[1] = 9;
        for (int i = 0; i < 9; ++i) { t.taps.dy[i] = g.dy[i]; t.taps.dx[i] = g.dx[i]; t.taps.widx[i] = g.widx[i]; }
        t.epi = EPI_ACT; t.e = e;
        // ToRGB of this layer inside the conv epilogue (all channels of a pixel are in one CTA when Cout <= 256)
        if (h->fuse_rgb && (li == 0 || (li % 2) == 0) && c.cout <= 256) {
          const size_t ri = li / 2;
          const RgbLayer& r = h->rgbs[ri];
          t.e.s_rgb = s_all + (size_t)B * r.row0; t.e.wrgb = r.wrgb; t.e.rgb_bias = r.bias;
          t.e.rgb_out = (ri + 1 == h->rgbs.size()) ? image : ws + L.skip_off[ri];
          rgb_fused = true;
        }
        LFP_PROF(h, LFP_KIND_CONV_FWD, conv_flops(g), conv_bytes(g), s, launch_conv_tc(t, s));
      } else {
        LFP_PROF(h, LFP_KIND_CONV_FWD, conv_flops(g), conv_bytes(g), s, launch_conv_simt(x, smod, c.wf, act, g, EPI_ACT, e, s));
      }
    } else {
      // stride-2 transposed conv as four sub-pixel phases into [B, 2H+1, 2W+1, Cout], then blur+epilogue
      float* T = ws + L.scratchT;
      const int H = c.res_in;
      const bool fuse_phases = use_tc && (c.cout <= h->fuse_phases_max_c || c.res_in <= h->fuse_phases_max_res) && h->fuse_phases;
      if (fuse_phases) {
        // all four sub-pixel phases in one launch: the activation tile is loaded and modulated once, every tap
        // accumulates into its phase's TMEM accumulator, and the epilogue writes the four planes of [B, 4, H+1, H+1, Cout]
        TcConv q{};
        q.in = x; q.in_planes = 1; q.in_h = q.in_w = H; q.in_bcast = x_bstride == 0; q.mod = smod; q.wmap = c.map_fwd.bytes;
        q.out = T; q.out_planes = 4; q.out_plane = 0; q.out_h = q.out_w = H + 1;
        q.out_stride = 2;   // the four phases are written interleaved: the blur then reads an ordinary [2H+1, 2H+1] image
        q.batch = B; q.gh = q.gw = H + 1; q.K = c.cin; q.N = c.cout;
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
        const double fl = 2.0 * B * 9.0 * c.cin * c.cout * H * H;   // the transposed conv's own MACs
        const double by = 4.0 * ((double)B * H * H * c.cin + (double)B * 4 * (H + 1) * (H + 1) * c.cout + 9.0 * c.cin * c.cout);
        LFP_PROF(h, LFP_KIND_CONV_FWD, fl, by, s, launch_conv_tc(q, s));
      } else
      for (int a = 0; a < 2; ++a)
        for (int bb = 0; bb < 2; ++bb) {
          ConvGeom g{};
          g.batch = B; g.gh = a == 0 ? H + 1 : H; g.gw = bb == 0 ? H + 1 : H;
          g.in_h = g.in_w = H; g.in_bstride = x_bstride; g.in_stride = 1;
          g.out_h = g.out_w = 2 * H + 1; g.out_stride = 2; g.out_oy = a; g.out_ox = bb; g.K = c.cin; g.N = c.cout;
          int t = 0;
          for (int ky = a == 0 ? 0 : 1; ky < 3; ky += 2)
            for (int kx = bb == 0 ? 0 : 1; kx < 3; kx += 2) {
              g.dy[t] = (signed char)(ky == 2 ? -1 : 0); g.dx[t] = (signed char)(kx == 2 ? -1 : 0); g.widx[t] = (signed char)(ky * 3 + kx); ++t;
            }
          g.ntaps = t;
          if (use_tc) {
            // same phase, written dense into plane a*2+bb of the phase-major intermediate [B, 4, H+1, H+1, Cout]
            TcConv q{};
            q.in = x; q.in_planes = 1; q.in_h = q.in_w = H; q.in_bcast = x_bstride == 0; q.mod = smod; q.wmap = c.map_fwd.bytes;
            q.out = T; q.out_planes = 4; q.out_plane = a * 2 + bb; q.out_h = q.out_w = H + 1;
            q.batch = B; q.gh = g.gh; q.gw = g.gw; q.K = c.cin; q.N = c.cout;
            q.taps.ngroups = 1; q.taps.group_plane[0] = 0; q.taps.group_tap0[0] = 0; q.taps.group_tap0[1] = t;
            for (int i = 0; i < t; ++i) { q.taps.dy[i] = g.dy[i]; q.taps.dx[i] = g.dx[i]; q.taps.widx[i] = g.widx[i]; }
            q.epi = EPI_STORE;
            LFP_PROF(h, LFP_KIND_CONV_FWD, conv_flops(g), conv_bytes(g), s, launch_conv_tc(q, s));
          } else {
            LFP_PROF(h, LFP_KIND_CONV_FWD, conv_flops(g), conv_bytes(g), s, launch_conv_simt(x, smod, c.wf, T, g, EPI_STORE, ConvEpiArgs{}, s));
          }
        }
      FirArgs f{};
      f.in_planar = use_tc && !fuse_phases;
      f.kx = f.ky = h->fir + 64;
      f.batch = B; f.in_h = f.in_w = 2 * H + 1; f.out_h = f.out_w = 2 * H; f.C = c.cout; f.pad = 1; f.coef = h->fir + 0;
      f.act = true; f.demod = dmod; f.noise = noise[c.noise_idx]; f.noise_bstride = nstride; f.noise_w = c.noise_w; f.bias = c.act_bias;
      LFP_PROF(h, LFP_KIND_FIR, 0.0, 4.0 * B * c.cout * ((double)f.in_h * f.in_w + (double)f.out_h * f.out_w), s, launch_fir4x4_nhwc(T, act, f, s));
    }
    x = act;
    x_bstride = (int64_t)c.res_out * c.res_out * c.cout;
    if (li == 0 || (li % 2) == 0) {  // conv1 and every second conv of a block feed a ToRGB
      const size_t ri = li / 2;
      const RgbLayer& r = h->rgbs[ri];
      const bool last = ri + 1 == h->rgbs.size();
      float* dst = last ? image : ws + L.skip_off[ri];
      const float* skip = ri == 0 ? nullptr : ws + L.skip_off[ri - 1];
      if (rgb_fused) {
        if (skip) LFP_PROF(h, LFP_KIND_TORGB, 0.0, 4.0 * B * r.res * r.res * 6.75, s, launch_skip_add(dst, skip, h->fir + 32, B, r.res, r.res, s));
      } else
      LFP_PROF(h, LFP_KIND_TORGB, 0.0, 4.0 * B * r.res * r.res * (r.cin + 3.75), s,
               launch_torgb_fwd(act, s_all + (size_t)B * r.row0, r.wrgb, r.bias, skip, h->fir + 32, dst, B, r.res, r.res, r.cin, s));
    }
  }
  {
    std::lock_guard<std::mutex> lock(h->mu);
    if (h->fwd_recs.size() > 256) h->fwd_recs.clear();   // stale workspaces of long-lived plans
    FwdRec& rec = h->fwd_recs[workspace];
    rec.batch = batch; rec.precision = precision; rec.fwd_only = fwd_only;
    rec.noise.assign(noise, noise + h->num_noise);
    rec.noise_batch.assign(noise_batch, noise_batch + h->num_noise);
  }
  return 0;
}

extern "C" int lfp_synth_forward(lfp_synth* h, int batch, const float* latent, const float* const* noise,
                                 const int* noise_batch, float* image, void* workspace, size_t workspace_bytes,
                                 int precision, void* stream) {
  return synth_forward_impl(h, batch, latent, noise, noise_batch, image, workspace, workspace_bytes, precision, stream, false);
}

extern "C" int lfp_synth_generate(lfp_synth* h, int batch, const float* latent, const float* const* noise,
                                  const int* noise_batch, float* image, void* workspace, size_t workspace_bytes,
                                  int precision, void* stream) {
  return synth_forward_impl(h, batch, latent, noise, noise_batch, image, workspace, workspace_bytes, precision, stream, true);
}

extern "C" int lfp_synth_backward(lfp_synth* h, int batch, const float* d_image, float* d_latent,
                                  void* workspace, size_t workspace_bytes, int precision, void* stream) {
  LFP_CHECK_ARG(h != nullptr && batch >= 1 && batch <= 65535, "synth_backward: null plan or batch out of range");
  const Layout L = h->layout(batch);
  LFP_TRY(check_run(h, batch, workspace, workspace_bytes, precision, L));
  LFP_CHECK_ARG(d_image && d_latent, "synth_backward: null argument");
  FwdRec rec;
  {
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->fwd_recs.find(workspace);
    if (it != h->fwd_recs.end()) rec = it->second;
  }
  // validated against the forward that ran on THIS workspace (its batch, arithmetic and noise pointers)
  if (rec.batch != batch || rec.fwd_only || rec.precision != precision) {
    set_error("synth_backward: no matching forward on this workspace (forward batch %d%s precision %d; backward batch %d precision %d)",
              rec.batch, rec.fwd_only ? " forward-only," : ",", rec.precision, batch, precision);
    return LFP_ESTATE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const std::vector<BwdStep> steps = h->bwd_steps(precision);
  const Tables* tb = nullptr;
  LFP_TRY(h->get_tables(batch, precision, L, steps, &tb));
  float* ws = (float*)workspace;
  float* s_all = ws + L.s_all;
  float* d_all = ws + L.d_all;
  float* ds_all = ws + L.ds_all;
  float* bufs[2] = {ws + L.bufA, ws + L.bufB};
  float* dskips[2] = {ws + L.dskipA, ws + L.dskipB};
  const int B = batch;

  const float* dskip = d_image;       // gradient wrt the running skip image at the current level
  float* g = nullptr;                 // gradient wrt the current layer's output activation (null: none yet)
  int cur = 0, dcur = 0;
  bool act_done = false;   // backward through this layer's noise/bias/lrelu already applied by the upstream dgrad epilogue
  for (int li = (int)h->convs.size() - 1; li >= 0; --li) {
    const ConvLayer& c = h->convs[li];
    const int hw = c.res_out * c.res_out;
    const bool feeds_rgb = li == 0 || (li % 2) == 0;
    const RgbLayer* r = feeds_rgb ? &h->rgbs[li / 2] : nullptr;
    float* gbuf = g != nullptr ? g : bufs[cur];
    if (!act_done) {
    ActBwdArgs ab{};
    ab.batch = B; ab.hw = hw; ab.C = c.cout; ab.act = ws + L.act_off[li]; ab.g = gbuf; ab.g_has_input = g != nullptr;
    ab.demod = d_all + (size_t)B * c.demod_off;
    ab.noise = rec.noise[c.noise_idx];
    ab.noise_bstride = rec.noise_batch[c.noise_idx] == 1 ? 0 : hw;
    ab.noise_w = c.noise_w; ab.bias = c.act_bias;
    if (r) { ab.drgb = dskip; ab.s_rgb = s_all + (size_t)B * r->row0; ab.wrgb = r->wrgb; }
    ab.pT = ws + L.pT_off[li]; ab.pR = ws + L.pR_off[li];
    LFP_PROF(h, LFP_KIND_ACTBWD, 0.0, 4.0 * B * hw * c.cout * (g != nullptr ? 3.0 : 2.0), s, launch_act_bwd(ab, s));
    }
    if (g == nullptr) { g = gbuf; }
    // g now holds dRaw = d(loss)/d(conv output before demod) * demod, at the layer's output resolution
    float* other = (g == bufs[0]) ? bufs[1] : bufs[0];
    const float* xin = li == 0 ? h->const_nhwc : ws + L.act_off[li - 1];
    const int64_t xin_bstride = li == 0 ? 0 : (int64_t)c.res_in * c.res_in * c.cin;
    ConvGeom gg{};
    gg.batch = B; gg.gh = gg.gw = c.res_in; gg.in_bstride = 0; gg.K = c.cout; gg.N = c.cin;
    gg.out_h = gg.out_w = c.res_in; gg.out_stride = 1; gg.out_oy = gg.out_ox = 0;
    const float* gin = g;
    const bool use_tc = steps[li].use_tc;
    TcConv tq{};
    tq.in_bcast = false; tq.mod = nullptr; tq.wmap = c.map_bwd.bytes;
    tq.out_planes = 1; tq.out_plane = 0; tq.out_h = tq.out_w = c.res_in;
    tq.batch = B; tq.gh = tq.gw = c.res_in; tq.K = c.cout; tq.N = c.cin; tq.epi = EPI_DGRAD;
    if (!c.up) {
      tq.in_planes = 1; tq.in_h = tq.in_w = c.res_out;
      tq.taps.ngroups = 1; tq.taps.group_plane[0] = 0; tq.taps.group_tap0[0] = 0; tq.taps.group_tap0[1] = 9;
      for (int i = 0; i < 9; ++i) { tq.taps.dy[i] = (signed char)(1 - i / 3); tq.taps.dx[i] = (signed char)(1 - i % 3); tq.taps.widx[i] = (signed char)i; }
      gg.in_h = gg.in_w = c.res_out; gg.in_stride = 1; gg.in_bstride = (int64_t)hw * c.cout;
      gg.ntaps = 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) { const int t = ky * 3 + kx; gg.dy[t] = (signed char)(1 - ky); gg.dx[t] = (signed char)(1 - kx); gg.widx[t] = (signed char)t; }
    } else {
      // adjoint of the blur: [2H,2H] -> [2H+1,2H+1], un-flipped taps, pad 2 (src/op/upfirdn2d.py:112-115)
      float* T = ws + L.scratchT;
      FirArgs f{};
      f.batch = B; f.in_h = f.in_w = c.res_out; f.out_h = f.out_w = c.res_out + 1; f.C = c.cout; f.pad = 2; f.coef = h->fir + 16;
      f.out_planar = use_tc;
      f.kx = f.ky = h->fir + 68;
      // tensor-core path: T' is phase-major [B, 4, H+1, H+1, Cout]; dx(y,x) = sum T'(2y+ky, 2x+kx) W[ky,kx] reads plane
      // (ky&1, kx&1) at (y + (ky>>1), x + (kx>>1))
      tq.in_planes = 4; tq.in_h = tq.in_w = c.res_in + 1;
      tq.taps.ngroups = 4;
      {
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
      LFP_PROF(h, LFP_KIND_FIR, 0.0, 4.0 * B * c.cout * ((double)f.in_h * f.in_w + (double)f.out_h * f.out_w), s, launch_fir4x4_nhwc(g, T, f, s));
      gin = T;
      gg.in_h = gg.in_w = c.res_out + 1; gg.in_stride = 2; gg.in_bstride = (int64_t)(c.res_out + 1) * (c.res_out + 1) * c.cout;
      gg.ntaps = 9;
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) { const int t = ky * 3 + kx; gg.dy[t] = (signed char)ky; gg.dx[t] = (signed char)kx; gg.widx[t] = (signed char)t; }
    }
    ConvEpiArgs e;
    e.mod_out = s_all + (size_t)B * c.row0; e.xsave = xin; e.xsave_bstride = xin_bstride; e.partial = ws + L.pX_off[li];
    float* dx = li == 0 ? nullptr : other;
    // the skip gradient crosses a resolution boundary through the Upsample adjoint; done before the dgrad because the
    // fused epilogue below needs it at this layer's input resolution
    if (c.up) {
      float* nd = dskips[dcur];
      LFP_TRY(upfirdn2d_dispatch(dskip, h->fir + 48, nd, LFP_F32, (int64_t)B * 3, c.res_out, c.res_out, 1, 4, 4, 1, 1, 2, 2, 1, 1, 1, 1, s, true, false));
      dskip = nd;
      dcur ^= 1;
    }
    // tensor-core path: the dgrad epilogue also runs the backward through noise/bias/lrelu (+ ToRGB branch) of the
    // layer below (what act_bwd_kernel does as a separate pass) - it already holds that layer's saved output
    const bool fuse = steps[li].fuse;
    const ConvLayer* below = fuse ? &h->convs[li - 1] : nullptr;
    const RgbLayer* rbelow = (fuse && ((li - 1) % 2 == 0)) ? &h->rgbs[(li - 1) / 2] : nullptr;
    if (fuse) {
      tq.epi = EPI_DGRAD_ACT;
      e.demod = d_all + (size_t)B * below->demod_off;
      e.noise = rec.noise[below->noise_idx];
      e.noise_bstride = rec.noise_batch[below->noise_idx] == 1 ? 0 : (int64_t)below->res_out * below->res_out;
      e.noise_w = below->noise_w; e.bias = below->act_bias;
      e.partial_T = ws + L.pT_off[li - 1]; e.partial_R = ws + L.pR_off[li - 1];
      if (rbelow) { e.drgb = dskip; e.s_rgb = s_all + (size_t)B * rbelow->row0; e.wrgb = rbelow->wrgb; }
    }
    if (use_tc) {
      tq.in = gin; tq.out = dx; tq.e = e;
      LFP_PROF(h, LFP_KIND_CONV_DGRAD, conv_flops(gg), conv_bytes(gg) + 4.0 * B * c.res_in * c.res_in * c.cin, s, launch_conv_tc(tq, s));
    } else {
      LFP_PROF(h, LFP_KIND_CONV_DGRAD, conv_flops(gg), conv_bytes(gg) + 4.0 * B * c.res_in * c.res_in * c.cin, s,
               launch_conv_simt(gin, nullptr, c.wg, dx, gg, EPI_DGRAD, e, s));
    }
    g = dx;
    act_done = fuse;
  }
  // every partial-sum reduction of the pass (style-gradient sums, act-backward T and ToRGB R sums: 43 launches before),
  // then every layer's style gradient (17 launches before), then the backward of the modulation linears
  LFP_TRY(launch_batched_partial_reduce(tb->ritems, tb->rblocks, tb->nrblocks, ws, B, s));
  LFP_TRY(launch_batched_style_grad(tb->gitems, tb->gblocks, tb->ngblocks, ws, B, s));
  LFP_TRY(launch_style_affine_bwd(ds_all, h->A_all, h->slot_begin, h->slot_end, h->row_base, h->row_cin, d_latent, B, h->rows, h->n_latent, h->style_dim, s));
  return 0;
}

extern "C" int lfp_synth_num_convs(const lfp_synth* h) { return h ? (int)h->convs.size() : 0; }

extern "C" int lfp_synth_read_activation(lfp_synth* h, int batch, int conv_index, const void* workspace,
                                         float* out, int* channels, int* res, void* stream) {
  LFP_CHECK_ARG(h != nullptr, "read_activation: null plan");
  LFP_CHECK_ARG(conv_index >= 0 && conv_index < (int)h->convs.size(), "read_activation: conv index %d out of range", conv_index);
  const ConvLayer& c = h->convs[conv_index];
  if (channels) *channels = c.cout;
  if (res) *res = c.res_out;
  if (out == nullptr) return 0;
  LFP_CHECK_ARG(workspace != nullptr && batch >= 1, "read_activation: null workspace or bad batch");
  {
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->fwd_recs.find(workspace);
    if (it == h->fwd_recs.end() || it->second.batch != batch || it->second.fwd_only) {
      set_error("read_activation: no full forward with batch %d on this workspace", batch);
      return LFP_ESTATE;
    }
  }
  const Layout L = h->layout(batch);
  return launch_nhwc_to_nchw((const float*)workspace + L.act_off[conv_index], out, batch, c.cout, c.res_out * c.res_out, (cudaStream_t)stream);
}

// grow-only device buffer `slot` of the host-pointer entry point
static void* host_buf(lfp_synth* h, int slot, size_t bytes) {
  if (h->hb.cap[slot] < bytes) {
    if (h->hb.p[slot]) cudaFree(h->hb.p[slot]);
    h->hb.p[slot] = nullptr; h->hb.cap[slot] = 0;
    if (cudaMalloc(&h->hb.p[slot], bytes) != cudaSuccess) return nullptr;
    h->hb.cap[slot] = bytes;
  }
  return h->hb.p[slot];
}

extern "C" int lfp_synth_forward_backward_host(lfp_synth* h, int batch, const float* latent,
                                               const float* const* noise, const int* noise_batch,
                                               float* image, const float* d_image, float* d_latent,
                                               int precision) {
  LFP_CHECK_ARG(h && latent && noise && noise_batch && image, "synth_host: null argument");
  LFP_CHECK_ARG(batch >= 1, "synth_host: bad batch");
  const bool with_bwd = d_image != nullptr && d_latent != nullptr;
  const size_t nlat = (size_t)batch * h->n_latent * h->style_dim;
  const size_t nimg = (size_t)batch * 3 * h->size * h->size;
  const size_t wsb = with_bwd ? lfp_synth_workspace_bytes(h, batch) : lfp_synth_generate_workspace_bytes(h, batch);
  // device staging buffers are owned by the plan and only ever grow: a loop over this call allocates nothing after its
  // first iteration (round 1 allocated and freed ~30 GB per call at B = 20)
  float* dlat = (float*)host_buf(h, 0, nlat * 4);
  float* dimg = (float*)host_buf(h, 1, nimg * 4);
  void* ws = host_buf(h, 2, wsb);
  float* dg = with_bwd ? (float*)host_buf(h, 3, nimg * 4) : nullptr;
  float* dl = with_bwd ? (float*)host_buf(h, 4, nlat * 4) : nullptr;
  bool ok = dlat && dimg && ws && (!with_bwd || (dg && dl));
  if ((int)h->hb.noise.size() < h->num_noise) { h->hb.noise.resize(h->num_noise, nullptr); h->hb.noise_cap.resize(h->num_noise, 0); }
  std::vector<const float*> dn(h->num_noise);
  for (int i = 0; ok && i < h->num_noise; ++i) {
    const int res = i == 0 ? 4 : (8 << ((i - 1) / 2));
    const size_t n = (size_t)noise_batch[i] * res * res * 4;
    if (h->hb.noise_cap[i] < n) {
      if (h->hb.noise[i]) cudaFree(h->hb.noise[i]);
      h->hb.noise[i] = nullptr; h->hb.noise_cap[i] = 0;
      ok = cudaMalloc(&h->hb.noise[i], n) == cudaSuccess;
      if (ok) h->hb.noise_cap[i] = n;
    }
    ok = ok && cudaMemcpy(h->hb.noise[i], noise[i], n, cudaMemcpyHostToDevice) == cudaSuccess;
    dn[i] = (const float*)h->hb.noise[i];
  }
  if (!ok) { set_error("synth_host: device allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError())); return LFP_ENOMEM; }
  int rc = 0;
  if (cudaMemcpy(dlat, latent, nlat * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = LFP_EINVAL;
  if (rc == 0) rc = with_bwd ? lfp_synth_forward(h, batch, dlat, dn.data(), noise_batch, dimg, ws, wsb, precision, nullptr)
                             : lfp_synth_generate(h, batch, dlat, dn.data(), noise_batch, dimg, ws, wsb, precision, nullptr);
  if (rc == 0 && cudaMemcpy(image, dimg, nimg * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = LFP_EINVAL;
  if (rc == 0 && with_bwd) {
    if (cudaMemcpy(dg, d_image, nimg * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = LFP_EINVAL;
    if (rc == 0) rc = lfp_synth_backward(h, batch, dg, dl, ws, wsb, precision, nullptr);
    if (rc == 0 && cudaMemcpy(d_latent, dl, nlat * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = LFP_EINVAL;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (rc == 0 && e != cudaSuccess) { set_error("synth_host: %s", cudaGetErrorString(e)); rc = (int)e; }
  return rc;
}

extern "C" int lfp_synth_profile_begin(lfp_synth* h, int kind_mask) {
  LFP_CHECK_ARG(h != nullptr, "profile_begin: null plan");
  h->prof_on = true; h->prof_mask = kind_mask; h->prof_used = 0;
  for (int k = 0; k < LFP_KIND_COUNT; ++k) { h->prof_flops[k] = 0; h->prof_bytes[k] = 0; h->prof_launches[k] = 0; }
  return 0;
}

extern "C" int lfp_synth_profile_end(lfp_synth* h, double* ms, int64_t* launches, double* flops, double* bytes) {
  LFP_CHECK_ARG(h != nullptr && ms && launches && flops && bytes, "profile_end: null argument");
  h->prof_on = false;
  for (int k = 0; k < LFP_KIND_COUNT; ++k) { ms[k] = 0; launches[k] = h->prof_launches[k]; flops[k] = h->prof_flops[k]; bytes[k] = h->prof_bytes[k]; }
  h->prof_lms.clear(); h->prof_lkind.clear();
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    LFP_CUDA(cudaEventSynchronize(h->prof_ev[i + 1]));
    float t = 0.f;
    LFP_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
    ms[h->prof_kind[i / 2]] += t;
    h->prof_lms.push_back(t); h->prof_lkind.push_back(h->prof_kind[i / 2]);
  }
  h->prof_used = 0;
  return 0;
}

extern "C" int lfp_synth_profile_launches(lfp_synth* h, int max_launches, int* kinds, float* ms, double* flops, double* bytes) {
  LFP_CHECK_ARG(h != nullptr, "profile_launches: null plan");
  const int n = (int)h->prof_lms.size();
  for (int i = 0; i < n && i < max_launches; ++i) {
    if (kinds) kinds[i] = h->prof_lkind[i];
    if (ms) ms[i] = h->prof_lms[i];
    if (flops) flops[i] = h->prof_lflops[i];
    if (bytes) bytes[i] = h->prof_lbytes[i];
  }
  return n;
}
