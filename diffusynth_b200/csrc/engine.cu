// Module-level C ABI: the network plans, the weight packing and the sampling graph behind opaque handles, so
// that a host in any language can load a reference checkpoint by its state_dict names and run
//   ds_unet_forward            = ConditionedUnet.forward                      model/diffusion.py:187-258
//   ds_sample_graph_build/run  = DiffSynthSampler.p_sample_loop + the tail    model/DiffSynthSampler.py:425-517, text2sound.py:112-134
//   ds_vqgan_quantize/decode/encode = VectorQuantizerEMA / Decoder / Encoder  model/VQGAN.py:98-146, 390-400, 323-326
// without any Python.  Everything here is host code: it packs the fp32 parameters once (GroupNorm(1,C) affine folded into the
// following convolution, border-class tables, K-major 16-bit weight matrices, sub-pixel phases of the transposed convolutions,
// fused time / condition projections), lays out the activation buffers of one (N, H, W) evaluation, and replays a fixed list of
// calls into the operator-level entry points of this library (ds_conv_gemm, ds_dwconv7, ds_attn_qkv_ctx, ...).  The sampling
// graph captures n_iter x (U-Net evaluation + fused CFG / DDIM update [+ inpaint blend]) and the VQ -> decoder -> iSTFT tail in
// one CUDA graph.  A handle belongs to one device and one host thread at a time; distinct handles are independent.
#include "common.cuh"
#include "../../include/diffusynth_b200.h"
#include <cmath>
#include <cstdarg>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace ds {
namespace eng {

struct Err { int code; std::string msg; };
[[noreturn]] static void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Err{code, buf};
}
#define ENG_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) ds::eng::fail(DS_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); } while (0)
#define ENG_CALL(expr) do { int rc_ = (expr); if (rc_ != 0) throw ds::eng::Err{rc_, std::string(ds::get_error())}; } while (0)
#define ENG_REQUIRE(cond, ...) do { if (!(cond)) ds::eng::fail(DS_ERR_INVALID, __VA_ARGS__); } while (0)

template <typename F>
static int guarded(F&& f) {
  try { f(); return DS_OK; }
  catch (const Err& e) { set_error("%s", e.msg.c_str()); return e.code; }
  catch (const std::exception& e) { set_error("%s", e.what()); return DS_ERR_INVALID; }
}

// ---- host-side 16-bit conversion (round to nearest even, as torch's .to(float16 / bfloat16)) ----
#ifdef DS_OPERANDS_BF16
static inline act_t h_f2act(float v) { return __float2bfloat16_rn(v); }
static inline float h_act2f(act_t v) { return __bfloat162float(v); }
#else
static inline act_t h_f2act(float v) { return __float2half_rn(v); }
static inline float h_act2f(act_t v) { return __half2float(v); }
#endif

struct HostTensor {
  std::vector<float> v;
  std::vector<long long> shape;
  long long dim(int i) const { return shape[i]; }
};
typedef std::map<std::string, HostTensor> StateDict;

static const HostTensor& need(const StateDict& sd, const std::string& name, std::initializer_list<long long> shape) {
  auto it = sd.find(name);
  ENG_REQUIRE(it != sd.end(), "missing parameter '%s'", name.c_str());
  const std::vector<long long> want(shape);
  ENG_REQUIRE(it->second.shape == want, "size mismatch for '%s' (rank %d, first extent %lld)", name.c_str(), (int)it->second.shape.size(),
              it->second.shape.empty() ? 0ll : it->second.shape[0]);
  return it->second;
}

// Device allocations of one handle.
struct Arena {
  std::vector<void*> ptrs;
  ~Arena() { for (void* p : ptrs) cudaFree(p); }
  void* raw(size_t bytes, bool zero) {
    void* p = nullptr;
    ENG_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    ptrs.push_back(p);
    if (zero) ENG_CUDA(cudaMemset(p, 0, bytes ? bytes : 16));
    return p;
  }
  template <typename T> T* alloc(size_t n, bool zero = false) { return static_cast<T*>(raw(n * sizeof(T), zero)); }
  template <typename T> T* upload(const std::vector<T>& h) {
    T* p = alloc<T>(h.size());
    ENG_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
};

static inline int pad16(int c) { return (c + 15) / 16 * 16; }
static inline int pad32(int c) { return (c + 31) / 32 * 32; }
static inline int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

// ---- packed convolution constants (one ds_conv_gemm call site) ----
enum { KIND_S1 = 0, KIND_DOWN = 1, KIND_UP = 2 };
struct PackedConv {
  act_t* weight = nullptr;   // [groups][cout_pad][K]
  float* e1 = nullptr;       // [ncls][cout_pad] or null
  float* e2 = nullptr;
  std::vector<std::vector<ds_conv_tap>> taps;
  int cin = 0, cout = 0, cout_pad = 0, ncls = 1, kind = KIND_S1;
};

static ds_conv_tap mk_tap(int dy, int dx, int view) { ds_conv_tap t; t.dy = (int8_t)dy; t.dx = (int8_t)dx; t.view = (int8_t)view; t.pad_ = 0; return t; }

// Stride-1 'same' conv (1x1 or 3x3), optionally preceded by GroupNorm(1, Cin): gamma folds into the weights, beta into e2, the
// per-sample (mean, rstd) are applied by the epilogue through e1 (one row per border class of a zero-padded 3x3).
static PackedConv pack_conv_s1(Arena& A, const HostTensor& w, const float* bias, const float* gamma, const float* beta, int cin_pad = 0) {
  const int O = (int)w.dim(0), Cin = (int)w.dim(1), k = (int)w.dim(2);
  ENG_REQUIRE(w.shape.size() == 4 && w.dim(3) == k && (k == 1 || k == 3), "pack_conv_s1: kernel %dx%lld", k, w.dim(3));
  const int cp = cin_pad ? cin_pad : Cin, cout_pad = pad16(O), K = k * k * cp;
  std::vector<act_t> wp((size_t)cout_pad * K, h_f2act(0.f));
  std::vector<double> s1((size_t)O * k * k, 0.0), s2((size_t)O * k * k, 0.0);
  for (int o = 0; o < O; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) {
          const float wv = w.v[(((size_t)o * Cin + c) * k + ky) * k + kx];
          const act_t h = h_f2act(gamma ? wv * gamma[c] : wv);
          wp[(size_t)o * K + (size_t)(ky * k + kx) * cp + c] = h;
          s1[((size_t)o * k + ky) * k + kx] += (double)h_act2f(h);
          if (beta) s2[((size_t)o * k + ky) * k + kx] += (double)(wv * beta[c]);
        }
  const int ncls = k == 3 ? 9 : 1;
  std::vector<float> e1((size_t)ncls * cout_pad, 0.f), e2((size_t)ncls * cout_pad, 0.f);
  for (int cls = 0; cls < ncls; ++cls) {
    const int rc = k == 3 ? cls / 3 : 1, cc = k == 3 ? cls % 3 : 1;
    for (int o = 0; o < O; ++o) {
      double a = 0.0, b = 0.0;
      for (int ky = 0; ky < k; ++ky) {
        if (k == 3 && ((rc == 0 && ky == 0) || (rc == 2 && ky == 2))) continue;
        for (int kx = 0; kx < k; ++kx) {
          if (k == 3 && ((cc == 0 && kx == 0) || (cc == 2 && kx == 2))) continue;
          a += s1[((size_t)o * k + ky) * k + kx];
          b += s2[((size_t)o * k + ky) * k + kx];
        }
      }
      e1[(size_t)cls * cout_pad + o] = (float)a;
      e2[(size_t)cls * cout_pad + o] = (float)b + (bias ? bias[o] : 0.f);
    }
  }
  PackedConv pc;
  pc.weight = A.upload(wp);
  pc.e2 = A.upload(e2);
  pc.e1 = gamma ? A.upload(e1) : nullptr;
  pc.taps.resize(1);
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) pc.taps[0].push_back(mk_tap(ky - k / 2, kx - k / 2, 0));
  pc.cin = cp; pc.cout = O; pc.cout_pad = cout_pad; pc.ncls = ncls; pc.kind = KIND_S1;
  return pc;
}

// Conv2d(k=4, stride=2, padding=1): tap (ky, kx) reads input pixel (2h + ky - 1, 2w + kx - 1) = parity view ((ky+1)%2, (kx+1)%2)
// at half-resolution offset floor((k-1)/2).
static PackedConv pack_conv_down(Arena& A, const HostTensor& w, const float* bias, int cin_pad = 0) {
  const int O = (int)w.dim(0), Cin = (int)w.dim(1);
  ENG_REQUIRE(w.shape.size() == 4 && w.dim(2) == 4 && w.dim(3) == 4, "pack_conv_down: kernel must be 4x4");
  const int cp = cin_pad ? cin_pad : Cin, cout_pad = pad16(O), K = 16 * cp;
  std::vector<act_t> wp((size_t)cout_pad * K, h_f2act(0.f));
  for (int o = 0; o < O; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < 16; ++t) wp[(size_t)o * K + (size_t)t * cp + c] = h_f2act(w.v[((size_t)o * Cin + c) * 16 + t]);
  std::vector<float> e2(cout_pad, 0.f);
  for (int o = 0; o < O; ++o) e2[o] = bias[o];
  PackedConv pc;
  pc.weight = A.upload(wp); pc.e2 = A.upload(e2);
  pc.taps.resize(1);
  for (int ky = 0; ky < 4; ++ky)
    for (int kx = 0; kx < 4; ++kx) pc.taps[0].push_back(mk_tap(floordiv2(ky - 1), floordiv2(kx - 1), ((ky + 1) % 2) * 2 + (kx + 1) % 2));
  pc.cin = cp; pc.cout = O; pc.cout_pad = cout_pad; pc.ncls = 1; pc.kind = KIND_DOWN;
  return pc;
}

// ConvTranspose2d(k=4, stride=2, padding=1), weight [Cin, Cout, 4, 4]: output phase (py, px) is a 2x2 conv of the input:
// py = 0 uses ky = 1 (dy 0), ky = 3 (dy -1); py = 1 uses ky = 0 (dy +1), ky = 2 (dy 0).
static PackedConv pack_conv_up(Arena& A, const HostTensor& w, const float* bias, int cin_pad = 0) {
  const int Cin = (int)w.dim(0), O = (int)w.dim(1);
  ENG_REQUIRE(w.shape.size() == 4 && w.dim(2) == 4 && w.dim(3) == 4, "pack_conv_up: kernel must be 4x4");
  const int cp = cin_pad ? cin_pad : Cin, cout_pad = pad16(O), K = 4 * cp;
  static const int sel[2][2][2] = {{{1, 0}, {3, -1}}, {{0, 1}, {2, 0}}};      // [phase][i] = {k, d}
  std::vector<act_t> wp((size_t)4 * cout_pad * K, h_f2act(0.f));
  PackedConv pc;
  pc.taps.resize(4);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int g = py * 2 + px;
      int t = 0;
      for (int iy = 0; iy < 2; ++iy)
        for (int ix = 0; ix < 2; ++ix) {
          const int ky = sel[py][iy][0], dy = sel[py][iy][1], kx = sel[px][ix][0], dx = sel[px][ix][1];
          for (int o = 0; o < O; ++o)
            for (int c = 0; c < Cin; ++c)
              wp[((size_t)g * cout_pad + o) * K + (size_t)t * cp + c] = h_f2act(w.v[(((size_t)c * O + o) * 4 + ky) * 4 + kx]);
          pc.taps[g].push_back(mk_tap(dy, dx, 0));
          ++t;
        }
    }
  std::vector<float> e2(cout_pad, 0.f);
  for (int o = 0; o < O; ++o) e2[o] = bias[o];
  pc.weight = A.upload(wp); pc.e2 = A.upload(e2);
  pc.cin = cp; pc.cout = O; pc.cout_pad = cout_pad; pc.ncls = 1; pc.kind = KIND_UP;
  return pc;
}

// A per-sample-weight 1x1 call site (to_out of the linear attention): only e2 (= bias) and the sizes; the weights are produced
// on the device by ds_attn_finalize.
static PackedConv pack_bias_only(Arena& A, const float* bias, int cin, int cout) {
  PackedConv pc;
  const int cout_pad = pad16(cout);
  std::vector<float> e2(cout_pad, 0.f);
  for (int o = 0; o < cout; ++o) e2[o] = bias[o];
  pc.e2 = A.upload(e2);
  pc.taps.resize(1);
  pc.taps[0].push_back(mk_tap(0, 0, 0));
  pc.cin = cin; pc.cout = cout; pc.cout_pad = cout_pad; pc.ncls = 1; pc.kind = KIND_S1;
  return pc;
}

// ---- statistics buffers and conv call construction ----
struct Stats { float* buf = nullptr; int slots = 0; long long count = 0; };
static Stats new_stats(Arena& A, int N, int slots, long long count) {
  Stats s;
  s.buf = A.alloc<float>((size_t)N * (slots + 2) * 2, /*zero=*/true);
  s.slots = slots; s.count = count;
  return s;
}

static void choose_tile(int H, int W, int& Hb, int& Wb) {
  long long best = -1;
  for (int wb = 128; wb >= 1; wb /= 2) {
    const int hb = 128 / wb;
    const long long area = (long long)((W + wb - 1) / wb) * wb * ((H + hb - 1) / hb) * hb;
    if (best < 0 || area < best) { best = area; Hb = hb; Wb = wb; }
  }
}
// N-tile width.  Large jobs: the widest divisor of Cout_pad (<= 256): fewest tcgen05.mma instructions per FLOP.  Jobs whose widest tiling
// leaves more than half of the SMs idle (m_units = samples x groups x 128-pixel tiles: one-prompt sampling, the deep levels of small
// batches) take the widest divisor <= 64 instead: one thread cannot issue tcgen05.mma faster than one per ~55 cycles
// (profiles/r02_umma_rate.md), so a narrower tile costs the same per K-step while the K loop, the weight fetch and the epilogue
// spread over 3-4x as many SMs.  Every output element keeps its K order: values do not depend on the choice (the statistics
// partials are grouped per tile, their fixed-order double-precision merge differs in the last bits only).
static int choose_bn(int cout_pad, long long m_units = 1 << 30) {
  int wide = 0;
  for (int bn = 256; bn >= 16 && !wide; bn -= 16)
    if (cout_pad % bn == 0) wide = bn;
  if (!wide) fail(DS_ERR_INVALID, "choose_bn(%d)", cout_pad);
  if (m_units * (cout_pad / wide) > 74) return wide;
  for (int bn = 64; bn >= 32; bn -= 32)
    if (bn < wide && cout_pad % bn == 0) return bn;
  return wide;
}

struct ConvOpts {
  void* out = nullptr; int out_c = 0; int out_hp = 0, out_wp = 0;      // 16-bit NHWC output (out_hp x out_wp: padded map of an "up" conv)
  float* out_f32 = nullptr;
  const Stats* stats_in = nullptr;
  const float* sbias = nullptr; int sbias_stride = 0;
  int act = 0;
  const void* residual = nullptr; int res_c = 0;
  bool want_stats = false;
  int src_batch_mod = 0;
  const void* weight_override = nullptr; bool per_sample_weights = false;
  float eps = 1e-5f;
  bool wide_tiles = false;      // batch-invariant tiling: always the widest N tile (ds_unet_config / ds_vqgan_config .batch_invariant)
};

// the argument block of one ds_conv_gemm call; src* 16-bit NHWC [N, Hin, Win, C]
static ds_conv_gemm_args conv_args(Arena& A, const PackedConv& pc, const void* src0, int C0, const void* src1, int C1, int N, int Hin, int Win,
                                   const ConvOpts& o, Stats* stats_out) {
  ds_conv_gemm_args a;
  memset(&a, 0, sizeof(a));
  ENG_REQUIRE(C0 + C1 == pc.cin, "conv_args: C0=%d C1=%d but the packed conv expects %d input channels", C0, C1, pc.cin);
  a.d_src0 = src0; a.d_src1 = src1; a.C0 = C0; a.C1 = C1; a.N = N; a.src_batch_mod = o.src_batch_mod;
  int Hg, Wg, Ho, Wo;
  if (pc.kind == KIND_DOWN) {
    Hg = Hin / 2; Wg = Win / 2;
    a.Hv = Hg; a.Wv = Wg;
    a.view_sn = (long long)Hin * Win; a.view_sh = 2ll * Win; a.view_sw = 2;
    for (int v = 0; v < 4; ++v) {
      a.view_off[v] = (long long)(v / 2) * Win + (v % 2);
      a.view_hv[v] = (Hin - v / 2 + 1) / 2; a.view_wv[v] = (Win - v % 2 + 1) / 2;
    }
    a.num_views = 4;
    Ho = Hg; Wo = Wg;
  } else {
    Hg = Hin; Wg = Win;
    a.Hv = Hin; a.Wv = Win;
    a.view_sn = (long long)Hin * Win; a.view_sh = Win; a.view_sw = 1;
    a.num_views = 1;
    Ho = pc.kind == KIND_UP ? 2 * Hin : Hin; Wo = pc.kind == KIND_UP ? 2 * Win : Win;
  }
  a.H = Hg; a.W = Wg;
  choose_tile(Hg, Wg, a.Hb, a.Wb);
  a.d_weight = o.weight_override ? o.weight_override : pc.weight;
  a.Cout_pad = pc.cout_pad; a.Cout = pc.cout;
  a.BN = o.wide_tiles ? choose_bn(pc.cout_pad)
                      : choose_bn(pc.cout_pad, (long long)N * (long long)pc.taps.size() * ((Hg + a.Hb - 1) / a.Hb) * ((Wg + a.Wb - 1) / a.Wb));
  a.BK = (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32;
  a.ntaps = (int)pc.taps[0].size(); a.groups = (int)pc.taps.size();
  a.per_sample_weights = o.per_sample_weights ? 1 : 0;
  for (size_t g = 0; g < pc.taps.size(); ++g)
    for (size_t t = 0; t < pc.taps[g].size(); ++t) a.taps[g][t] = pc.taps[g][t];
  if (o.stats_in) {
    ENG_REQUIRE(pc.e1 != nullptr, "conv_args: stats_in on a conv packed without a GroupNorm fold");
    a.d_stats_in = o.stats_in->buf; a.stats_in_slots = o.stats_in->slots; a.d_e1 = pc.e1;
  }
  a.eps = o.eps;
  a.d_e2 = pc.e2; a.ncls = pc.ncls;
  if (o.sbias) { a.d_sbias = o.sbias; a.sbias_stride = o.sbias_stride; }
  a.act = o.act;
  if (o.residual) { a.d_residual = o.residual; a.res_sn = (long long)Ho * Wo * o.res_c; a.res_sh = (long long)Wo * o.res_c; a.res_sw = o.res_c; }
  if (o.out) {
    const long long Co = o.out_c;
    a.d_out = o.out;
    if (pc.kind == KIND_UP) {
      const int Hp = o.out_hp ? o.out_hp : Ho, Wp = o.out_wp ? o.out_wp : Wo;
      ENG_REQUIRE(Hp >= Ho && Wp >= Wo, "conv_args: padded output %dx%d smaller than %dx%d", Hp, Wp, Ho, Wo);
      const long long org = (long long)((Hp - Ho) / 2) * Wp + (Wp - Wo) / 2;
      a.out_sn = (long long)Hp * Wp * Co; a.out_sh = 2ll * Wp * Co; a.out_sw = 2 * Co;
      for (int g = 0; g < 4; ++g) a.out_goff[g] = (org + (long long)(g / 2) * Wp + (g % 2)) * Co;
    } else {
      a.out_sn = (long long)Ho * Wo * Co; a.out_sh = (long long)Wo * Co; a.out_sw = Co;
    }
  }
  a.d_out_f32_nchw = o.out_f32;
  if (o.want_stats) {
    ENG_REQUIRE(stats_out != nullptr, "conv_args: want_stats without a destination");
    const int slots = ds_conv_gemm_stats_slots(&a);
    *stats_out = new_stats(A, N, slots, (long long)Ho * Wo * pc.cout);
    a.d_stats_out = stats_out->buf;
    a.stats_out_inv_count = 1.0f / (float)stats_out->count;
  }
  return a;
}

struct Run { const float* x; const long long* t; cudaStream_t s; bool no_cond = false; };
typedef std::function<void(const Run&)> Op;

static const int HEADS = 4, DHEAD = 32, HID = HEADS * DHEAD;

// =====================================================================================================================
// U-Net
// =====================================================================================================================
struct Block {
  std::string p;
  int dim = 0, dim_out = 0, t_off = 0;
  bool has_time = true, has_res = false;
  int t_dim = 0;                  // width of this block's slice of the fused time projection (dim for ConvNeXt, dim_out for ResNet blocks)
  float* dw = nullptr;            // [49][C]
  PackedConv conv1, conv2, res;
  float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;      // ResNet-block GroupNorm(groups) affines
};
struct Attn {
  std::string p;
  int dim = 0, c_off = 0;
  PackedConv qkv, out;
  float *wout = nullptr, *gamma = nullptr, *beta = nullptr;
};

struct UnetPlan;

struct Unet {
  ds_unet_config cfg;
  int device = 0;
  StateDict sd;
  bool packed = false;
  std::unique_ptr<Arena> weights;
  std::vector<int> dd, ud;
  std::map<std::string, Block> blocks;
  std::map<std::string, Attn> attns;
  std::map<std::string, PackedConv> samplers;
  std::vector<std::string> block_order, attn_order;
  PackedConv final_conv, stem;
  int t_total = 0, c_total = 0;
  float *t_w = nullptr, *t_b = nullptr, *c_w = nullptr, *c_b = nullptr, *lab_w = nullptr, *lab_b = nullptr;
  float *tm1_w = nullptr, *tm1_b = nullptr, *tm3_w = nullptr, *tm3_b = nullptr;
  std::vector<std::unique_ptr<UnetPlan>> plans;
  void pack();
};

struct UnetPlan {
  Unet* net;
  int N, H, W, x_batch_mod, uniform_time, nb, shared = 0;
  Arena A;
  std::vector<Op> ops, cond_ops;
  int launches = 0;
  float* cond = nullptr;      // [N][L] input
  float* eps = nullptr;       // [N][out_dim][H][W] output
  size_t cond_bytes = 0;      // N * L floats, or N int64 labels (condition_type "instrument_family")
  float* sbias = nullptr;     // [N][c_total]
  float* tbias = nullptr;
  std::map<std::string, void*> scratch;
  UnetPlan(Unet* n, int N_, int H_, int W_, int mod, int ut);
  void run_cond(cudaStream_t s) { Run r{nullptr, nullptr, s}; for (auto& f : cond_ops) f(r); }
  void run(const float* x, const long long* t, cudaStream_t s, bool no_cond = false) { Run r{x, t, s, no_cond}; for (auto& f : ops) f(r); }
};

void Unet::pack() {
  ENG_CUDA(cudaGetDevice(&device));
  const int L = cfg.label_emb_dim, td = cfg.time_dim, mult = cfg.convnext_mult;
  dd.assign(cfg.down_dims, cfg.down_dims + cfg.n_levels);
  ud.assign(cfg.up_dims, cfg.up_dims + cfg.n_levels);
  weights.reset(new Arena());
  plans.clear();
  blocks.clear(); attns.clear(); samplers.clear(); block_order.clear(); attn_order.clear();
  Arena& A = *weights;

  auto blk = [&](const std::string& p, int dim, int dim_out, bool has_time) {
    Block b;
    b.p = p; b.dim = dim; b.dim_out = dim_out; b.has_time = has_time && cfg.with_time_emb;
    has_time = b.has_time;
    if (!cfg.use_convnext) {
      // ResnetBlock (diffusion_components.py:59-104): two (conv3x3 -> GroupNorm(groups) -> SiLU) blocks, the time embedding added in
      // between; the residual projection is a 1x1 conv or -- where the reference uses nn.Identity -- an identity matrix, so that the
      // residual add and the statistics of the block's output stay in the conv epilogue
      b.t_dim = dim_out;
      b.conv1 = pack_conv_s1(A, need(sd, p + "block1.proj.weight", {dim_out, dim, 3, 3}), need(sd, p + "block1.proj.bias", {dim_out}).v.data(), nullptr, nullptr);
      b.conv2 = pack_conv_s1(A, need(sd, p + "block2.proj.weight", {dim_out, dim_out, 3, 3}), need(sd, p + "block2.proj.bias", {dim_out}).v.data(), nullptr, nullptr);
      b.g1 = A.upload(need(sd, p + "block1.norm.weight", {dim_out}).v); b.b1 = A.upload(need(sd, p + "block1.norm.bias", {dim_out}).v);
      b.g2 = A.upload(need(sd, p + "block2.norm.weight", {dim_out}).v); b.b2 = A.upload(need(sd, p + "block2.norm.bias", {dim_out}).v);
      b.has_res = true;
      if (dim != dim_out) {
        b.res = pack_conv_s1(A, need(sd, p + "res_conv.weight", {dim_out, dim, 1, 1}), need(sd, p + "res_conv.bias", {dim_out}).v.data(), nullptr, nullptr);
      } else {
        HostTensor eye;
        eye.shape = {dim, dim, 1, 1};
        eye.v.assign((size_t)dim * dim, 0.f);
        for (int c = 0; c < dim; ++c) eye.v[(size_t)c * dim + c] = 1.f;
        std::vector<float> zero(dim, 0.f);
        b.res = pack_conv_s1(A, eye, zero.data(), nullptr, nullptr);
      }
      if (has_time) { need(sd, p + "mlp.1.weight", {dim_out, td}); need(sd, p + "mlp.1.bias", {dim_out}); }
      blocks[p] = b;
      block_order.push_back(p);
      return;
    }
    b.t_dim = dim;
    const HostTensor& dw = need(sd, p + "ds_conv.weight", {dim, 1, 7, 7});
    need(sd, p + "ds_conv.bias", {dim});
    std::vector<float> dwt((size_t)49 * dim);
    for (int c = 0; c < dim; ++c)
      for (int t = 0; t < 49; ++t) dwt[(size_t)t * dim + c] = dw.v[(size_t)c * 49 + t];
    b.dw = A.upload(dwt);
    const int hid = dim_out * mult;
    b.conv1 = pack_conv_s1(A, need(sd, p + "net.1.weight", {hid, dim, 3, 3}), need(sd, p + "net.1.bias", {hid}).v.data(),
                           need(sd, p + "net.0.weight", {dim}).v.data(), need(sd, p + "net.0.bias", {dim}).v.data());
    b.conv2 = pack_conv_s1(A, need(sd, p + "net.4.weight", {dim_out, hid, 3, 3}), need(sd, p + "net.4.bias", {dim_out}).v.data(),
                           need(sd, p + "net.3.weight", {hid}).v.data(), need(sd, p + "net.3.bias", {hid}).v.data());
    b.has_res = dim != dim_out;
    if (b.has_res) b.res = pack_conv_s1(A, need(sd, p + "res_conv.weight", {dim_out, dim, 1, 1}), need(sd, p + "res_conv.bias", {dim_out}).v.data(), nullptr, nullptr);
    if (has_time) { need(sd, p + "mlp.1.weight", {dim, td}); need(sd, p + "mlp.1.bias", {dim}); }
    blocks[p] = b;
    block_order.push_back(p);
  };
  auto att = [&](const std::string& p, int dim) {
    Attn a;
    a.p = p; a.dim = dim;
    a.qkv = pack_conv_s1(A, need(sd, p + "fn.fn.to_qkv.weight", {3 * HID, dim, 1, 1}), nullptr, need(sd, p + "fn.norm.weight", {dim}).v.data(),
                         need(sd, p + "fn.norm.bias", {dim}).v.data());
    a.wout = A.upload(need(sd, p + "fn.fn.to_out.0.weight", {dim, HID, 1, 1}).v);
    a.out = pack_bias_only(A, need(sd, p + "fn.fn.to_out.0.bias", {dim}).v.data(), HID, dim);
    a.gamma = A.upload(need(sd, p + "fn.fn.to_out.1.weight", {dim}).v);
    a.beta = A.upload(need(sd, p + "fn.fn.to_out.1.bias", {dim}).v);
    const char* second = cfg.attn_type == 0 ? "label_query" : "label_value";      // LinearCrossAttentionAdd / LinearCrossAttention
    need(sd, p + "fn.fn." + second + ".weight", {HID, L}); need(sd, p + "fn.fn." + second + ".bias", {HID});
    need(sd, p + "fn.fn.label_key.weight", {HID, L}); need(sd, p + "fn.fn.label_key.bias", {HID});
    attns[p] = a;
    attn_order.push_back(p);
  };

  std::vector<int> skips;
  const int n_stage = cfg.n_levels - 1;
  for (int i = 0; i < n_stage; ++i) {
    const int cin = dd[i], cout = dd[i + 1];
    const std::string p = "downs." + std::to_string(i) + ".";
    blk(p + "0.", cin, cout, true); att(p + "1.", cout); blk(p + "2.", cout, cout, true); att(p + "3.", cout);
    samplers[p + "4."] = pack_conv_down(A, need(sd, p + "4.weight", {cout, cout, 4, 4}), need(sd, p + "4.bias", {cout}).v.data());
    skips.push_back(cout);
  }
  const int mid = dd.back();
  for (int j = 0; j < cfg.mid_depth - 1; ++j) blk("mid_left." + std::to_string(j) + ".", mid, mid, true);
  blk("mid_mid.0.", mid, mid, true); att("mid_mid.1.", mid); blk("mid_mid.2.", mid, mid, true);
  for (int j = 0; j < cfg.mid_depth - 1; ++j) blk("mid_right." + std::to_string(j) + ".", 2 * mid, mid, true);
  for (int i = 0; i < n_stage; ++i) {
    const int cin = ud[i], cout = ud[i + 1];
    const int s = skips.back(); skips.pop_back();
    const std::string p = "ups." + std::to_string(i) + ".";
    blk(p + "0.", cin + s, cin, true); att(p + "1.", cin);
    samplers[p + "2."] = pack_conv_up(A, need(sd, p + "2.weight", {cin, cin, 4, 4}), need(sd, p + "2.bias", {cin}).v.data());
    blk(p + "3.", cin + s, cout, true); att(p + "4.", cout);
    blk(p + "5.", cout + s, cout, true); att(p + "6.", cout);
  }
  blk("final_conv.0.", dd[0] + ud.back(), ud.back(), false);
  final_conv = pack_conv_s1(A, need(sd, "final_conv.1.weight", {cfg.out_dim, ud.back(), 3, 3}), need(sd, "final_conv.1.bias", {cfg.out_dim}).v.data(), nullptr, nullptr);

  // fused time projection: rows = concat over blocks of mlp.1 (GELU on the input); bias += ds_conv.bias
  {
    std::vector<float> w, b;
    int off = 0;
    for (const std::string& p : block_order) {
      Block& bk = blocks[p];
      bk.t_off = off;
      const std::vector<float> dwb = cfg.use_convnext ? sd[p + "ds_conv.bias"].v : std::vector<float>((size_t)bk.t_dim, 0.f);
      if (bk.has_time) {
        const std::vector<float>& mw = sd[p + "mlp.1.weight"].v;
        const std::vector<float>& mb = sd[p + "mlp.1.bias"].v;
        w.insert(w.end(), mw.begin(), mw.end());
        for (int c = 0; c < bk.t_dim; ++c) b.push_back(mb[c] + dwb[c]);
      } else {
        w.insert(w.end(), (size_t)bk.t_dim * td, 0.f);
        b.insert(b.end(), dwb.begin(), dwb.end());
      }
      off += bk.t_dim;
    }
    t_total = off;
    t_w = A.upload(w); t_b = A.upload(b);
  }
  // fused condition projection: per attention site [label_query | label_key | zeros(v)], added to q, k in the to_qkv epilogue
  {
    std::vector<float> w, b;
    int off = 0;
    for (const std::string& p : attn_order) {
      Attn& a = attns[p];
      a.c_off = off;
      // linear_add: [label_query | label_key | 0] added to q, k in the to_qkv epilogue; linear_cat: [label_key | label_value | 0] = the
      // extra key / value token consumed by ds_attn_finalize_cat
      const char* first_nm = cfg.attn_type == 0 ? "label_query" : "label_key";
      const char* second_nm = cfg.attn_type == 0 ? "label_key" : "label_value";
      for (const char* nm : {first_nm, second_nm}) {
        const std::vector<float>& lw = sd[p + "fn.fn." + nm + ".weight"].v;
        const std::vector<float>& lb = sd[p + "fn.fn." + nm + ".bias"].v;
        w.insert(w.end(), lw.begin(), lw.end());
        b.insert(b.end(), lb.begin(), lb.end());
      }
      w.insert(w.end(), (size_t)HID * L, 0.f);
      b.insert(b.end(), (size_t)HID, 0.f);
      off += 3 * HID;
    }
    c_total = off;
    c_w = A.upload(w); c_b = A.upload(b);
  }
  if (cfg.condition_type == 0) {            // ConditionalEmbedding: nn.Linear for text embeddings, nn.Embedding for class labels (diffusion_components.py:155-168)
    lab_w = A.upload(need(sd, "label_embedding.embedding.weight", {L, L}).v);
    lab_b = A.upload(need(sd, "label_embedding.embedding.bias", {L}).v);
  } else {
    lab_w = A.upload(need(sd, "label_embedding.embedding.weight", {cfg.n_label_class + 1, L}).v);
    lab_b = nullptr;
  }
  if (cfg.with_time_emb) {
    tm1_w = A.upload(need(sd, "time_mlp.1.weight", {td, dd[0]}).v); tm1_b = A.upload(need(sd, "time_mlp.1.bias", {td}).v);
    tm3_w = A.upload(need(sd, "time_mlp.3.weight", {td, td}).v); tm3_b = A.upload(need(sd, "time_mlp.3.bias", {td}).v);
  }
  // stem as a GEMM over im2col patches: [Cout, Cin, 7, 7] -> [Cout_pad][ky*32 + kx*4 + ci] (8th pixel slot and ci >= Cin are zero)
  {
    const HostTensor& w0 = need(sd, "init_conv.weight", {dd[0], cfg.in_dim, 7, 7});
    const std::vector<float>& b0 = need(sd, "init_conv.bias", {dd[0]}).v;
    ENG_REQUIRE(cfg.in_dim <= 4, "in_dim=%d (the im2col stem packs at most 4 input channels)", cfg.in_dim);
    const int cout_pad = pad16(dd[0]);
    std::vector<act_t> wp((size_t)cout_pad * 224, h_f2act(0.f));
    for (int o = 0; o < dd[0]; ++o)
      for (int ci = 0; ci < cfg.in_dim; ++ci)
        for (int ky = 0; ky < 7; ++ky)
          for (int kx = 0; kx < 7; ++kx)
            wp[(size_t)o * 224 + ky * 32 + kx * 4 + ci] = h_f2act(w0.v[(((size_t)o * cfg.in_dim + ci) * 7 + ky) * 7 + kx]);
    stem = pack_bias_only(A, b0.data(), 224, dd[0]);
    stem.weight = A.upload(wp);
  }
  packed = true;
}

UnetPlan::UnetPlan(Unet* n, int N_, int H_, int W_, int mod, int ut) : net(n), N(N_), H(H_), W(W_), x_batch_mod(mod), uniform_time(ut) {
  const ds_unet_config& cfg = net->cfg;
  const std::vector<int>& dd = net->dd;
  const int td = cfg.time_dim, L = cfg.label_emb_dim;
  nb = x_batch_mod > 0 ? x_batch_mod : N;
  const int NT = uniform_time ? 1 : N;
  const long long t_stride = (uniform_time || !cfg.with_time_emb) ? 0 : net->t_total;
  const bool labels = cfg.condition_type == 1;      // integer class labels [N] (int64) instead of embeddings [N][L]
  cond = static_cast<float*>(A.raw(labels ? (size_t)N * sizeof(long long) : (size_t)N * L * sizeof(float), true));
  cond_bytes = labels ? (size_t)N * sizeof(long long) : (size_t)N * L * sizeof(float);
  eps = A.alloc<float>((size_t)N * cfg.out_dim * H * W, true);
  float* cemb = A.alloc<float>((size_t)N * L);
  sbias = A.alloc<float>((size_t)N * net->c_total);
  float* sin = A.alloc<float>((size_t)NT * dd[0]);
  float* t1 = A.alloc<float>((size_t)NT * td);
  float* temb = A.alloc<float>((size_t)NT * td);
  tbias = A.alloc<float>((size_t)NT * net->t_total);
  Unet* nt = net;
  const int Nn = N;
  const int c_total = net->c_total, t_total = net->t_total;

  auto act = [&](int n_, int h, int w, int c) { return A.alloc<act_t>((size_t)n_ * h * w * c); };
  auto scr = [&](const char* role, int n_, int h, int w, int c) -> act_t* {
    char key[96];
    snprintf(key, sizeof(key), "%s/%d/%d/%d/%d", role, n_, h, w, c);
    auto it = scratch.find(key);
    if (it != scratch.end()) return static_cast<act_t*>(it->second);
    act_t* p = act(n_, h, w, c);
    scratch[key] = p;
    return p;
  };
  auto add = [&](Op f, int nl = 1) { ops.push_back(std::move(f)); launches += nl; };

  // ---- condition path (step-invariant: once per sampling call) ----
  {
    float* cnd = cond; float* sb = sbias;
    if (labels) {
      const int rows = cfg.n_label_class + 1;
      cond_ops.push_back([=](const Run& r) { ENG_CALL(ds_embedding_gather(nt->lab_w, reinterpret_cast<const long long*>(cnd), cemb, Nn, L, rows, r.s)); });
    } else {
      cond_ops.push_back([=](const Run& r) { ENG_CALL(ds_linear(cnd, L, nt->lab_w, nt->lab_b, cemb, L, Nn, L, L, 0, 0, r.s)); });
    }
    cond_ops.push_back([=](const Run& r) { ENG_CALL(ds_linear(cemb, L, nt->c_w, nt->c_b, sb, c_total, Nn, L, c_total, 0, 0, r.s)); });
  }
  // ---- time path ----
  if (!cfg.with_time_emb) {
    // diffusion.py:107-109,211: no time terms; one row of per-channel biases (ds_conv.bias of every block), read with stride 0
    ENG_CUDA(cudaMemcpy(tbias, net->t_b, (size_t)net->t_total * sizeof(float), cudaMemcpyDeviceToDevice));
  } else {
    const int d0 = dd[0];
    float* tb = tbias;
    add([=](const Run& r) { ENG_CALL(ds_sinusoidal_embedding(r.t, sin, NT, d0, r.s)); });
    add([=](const Run& r) { ENG_CALL(ds_linear(sin, d0, nt->tm1_w, nt->tm1_b, t1, td, NT, d0, td, 0, 1, r.s)); });
    add([=](const Run& r) { ENG_CALL(ds_linear(t1, td, nt->tm3_w, nt->tm3_b, temb, td, NT, td, td, 0, 0, r.s)); });
    const int act_in = cfg.use_convnext ? 1 : 2;      // per-block mlp: GELU (ConvNextBlock) or SiLU (ResnetBlock) on the time embedding
    add([=](const Run& r) { ENG_CALL(ds_linear(temb, td, nt->t_w, nt->t_b, tb, t_total, NT, td, t_total, act_in, 0, r.s)); });
  }

  const bool wide_tiles = cfg.batch_invariant != 0;
  auto conv = [&](const PackedConv& pc, const void* s0, int C0, const void* s1, int C1, int n_, int h, int w, const ConvOpts& o_, Stats* st_out) {
    ConvOpts o = o_; o.wide_tiles = wide_tiles;
    const ds_conv_gemm_args a = conv_args(A, pc, s0, C0, s1, C1, n_, h, w, o, st_out);
    add([a](const Run& r) { ENG_CALL(ds_conv_gemm(&a, r.s)); });
  };

  struct Act { act_t* p; int c; };
  // One ConvNextBlock over n_ samples; s0_mod > 0: source 0 holds s0_mod samples shared by the guidance halves.
  auto block = [&](const std::string& p, Act s0, Act s1, int h, int w, int n_, int s0_mod, Stats* st_o_out) -> Act {
    const Block& b = net->blocks.at(p);
    ENG_REQUIRE(s0.c + s1.c == b.dim, "block %s: %d + %d input channels, expected %d", p.c_str(), s0.c, s1.c, b.dim);
    act_t* hbuf = scr("dw", n_, h, w, b.dim);
    const float* tb = tbias + b.t_off;
    Stats st_h = new_stats(A, n_, ds_dwconv7_stats_slots(b.dim, h, w), (long long)h * w * b.dim);
    {
      const float* dw = b.dw;
      const int C0 = s0.c, C1 = s1.c;
      const act_t *p0 = s0.p, *p1 = s1.p;
      add([=](const Run& r) { ENG_CALL(ds_dwconv7(p0, p1, C0, C1, s0_mod, dw, tb, t_stride, hbuf, st_h.buf, 1e-5f, n_, h, w, r.s)); });
    }
    act_t* y = scr("hid", n_, h, w, b.conv1.cout);
    Stats st_y;
    { ConvOpts o; o.out = y; o.out_c = b.conv1.cout; o.stats_in = &st_h; o.act = 1; o.want_stats = true; conv(b.conv1, hbuf, b.dim, nullptr, 0, n_, h, w, o, &st_y); }
    const act_t* rsd;
    if (b.has_res) {
      act_t* rr = scr("res", n_, h, w, b.dim_out);
      ConvOpts o; o.out = rr; o.out_c = b.dim_out; o.src_batch_mod = s0_mod;
      conv(b.res, s0.p, s0.c, s1.p, s1.c, n_, h, w, o, nullptr);
      rsd = rr;
    } else {
      ENG_REQUIRE(s1.p == nullptr && s0_mod == 0, "block %s: identity residual with two sources", p.c_str());
      rsd = s0.p;
    }
    act_t* ob = act(n_, h, w, b.dim_out);
    { ConvOpts o; o.out = ob; o.out_c = b.dim_out; o.stats_in = &st_y; o.residual = rsd; o.res_c = b.dim_out; o.want_stats = true;
      conv(b.conv2, y, b.conv1.cout, nullptr, 0, n_, h, w, o, st_o_out); }
    return Act{ob, b.dim_out};
  };
  const int RG = cfg.resnet_block_groups > 0 ? cfg.resnet_block_groups : 8;
  auto gn_silu = [&](const act_t* xin, int c, int h, int w, int n_, const float* gamma, const float* beta, const char* role) -> act_t* {
    // GroupNorm(groups, eps 1e-5) + SiLU as statistics + apply passes (the VQGAN GroupNorm kernels)
    float* part = A.alloc<float>((size_t)n_ * RG * 32 * 2);
    act_t* out = scr(role, n_, h, w, c);
    const long long hw = (long long)h * w;
    add([=](const Run& r) { ENG_CALL(ds_group_stats(xin, part, n_, c, c, RG, hw, 32, r.s)); });
    add([=](const Run& r) { ENG_CALL(ds_gn_act(xin, out, part, 32, gamma, beta, n_, c, c, RG, hw, 1e-5f, 2, r.s)); });
    return out;
  };
  auto res_block = [&](const std::string& p, Act s0, Act s1, int h, int w, int n_, int s0_mod, Stats* st_o_out) -> Act {
    const Block& b = net->blocks.at(p);
    ENG_REQUIRE(s0.c + s1.c == b.dim, "block %s: %d + %d input channels, expected %d", p.c_str(), s0.c, s1.c, b.dim);
    act_t* y1 = scr("rb_y1", n_, h, w, b.dim_out);
    { ConvOpts o; o.out = y1; o.out_c = b.dim_out; o.src_batch_mod = s0_mod; conv(b.conv1, s0.p, s0.c, s1.p, s1.c, n_, h, w, o, nullptr); }
    act_t* h1 = gn_silu(y1, b.dim_out, h, w, n_, b.g1, b.b1, "rb_h1");
    if (b.has_time) {
      const float* tb = tbias + b.t_off;
      const int co = b.dim_out; const long long hw = (long long)h * w;
      add([=](const Run& r) { ENG_CALL(ds_add_channel_bias(h1, tb, t_stride, n_, co, hw, r.s)); });
    }
    act_t* y2 = scr("rb_y2", n_, h, w, b.dim_out);
    { ConvOpts o; o.out = y2; o.out_c = b.dim_out; conv(b.conv2, h1, b.dim_out, nullptr, 0, n_, h, w, o, nullptr); }
    act_t* h2 = gn_silu(y2, b.dim_out, h, w, n_, b.g2, b.b2, "rb_h2");
    act_t* ob = act(n_, h, w, b.dim_out);
    { ConvOpts o; o.out = ob; o.out_c = b.dim_out; o.residual = h2; o.res_c = b.dim_out; o.want_stats = true; o.src_batch_mod = s0_mod;
      conv(b.res, s0.p, s0.c, s1.p, s1.c, n_, h, w, o, st_o_out); }
    return Act{ob, b.dim_out};
  };
  auto any_block = [&](const std::string& p, Act s0, Act s1, int h, int w, int n_, int s0_mod, Stats* st_o_out) -> Act {
    return cfg.use_convnext ? block(p, s0, s1, h, w, n_, s0_mod, st_o_out) : res_block(p, s0, s1, h, w, n_, s0_mod, st_o_out);
  };
  // Residual(PreNorm(LinearCrossAttentionAdd)); x_mod > 0: x (and its statistics) hold x_mod samples shared by the guidance halves.
  auto attn = [&](const std::string& p, Act x, const Stats& st_x, int h, int w, int x_mod) -> Act {
    const Attn& a = net->attns.at(p);
    const long long npix = (long long)h * w;
    const float* sb = sbias + a.c_off;
    act_t* qp = scr("qp", Nn, h, w, HID);
    float* part = A.alloc<float>((size_t)ds_attn_part_floats(Nn, HEADS, npix));
    act_t* M = A.alloc<act_t>((size_t)Nn * a.out.cout_pad * HID);
    {
      const act_t* xp = x.p; const int dim = a.dim;
      const Stats sx = st_x;
      const PackedConv qkv = a.qkv;
      const bool cat = cfg.attn_type == 1;      // LinearCrossAttention: the condition is one extra key / value token, merged by the finalize
      const float* sb_qk = cat ? nullptr : sb;
      add([=](const Run& r) { ENG_CALL(ds_attn_qkv_ctx(xp, dim, x_mod, sx.buf, sx.slots, qkv.weight, qkv.e1, qkv.e2, sb_qk, c_total, qp, part, Nn, HEADS, npix,
                                                       1.0f / sqrtf((float)DHEAD), r.s)); });
      const float* wout = a.wout; const int cop = a.out.cout_pad;
      if (cat) add([=](const Run& r) {      // condition=None: no extra token (diffusion_components.py:195-200)
        if (r.no_cond) ENG_CALL(ds_attn_finalize(part, wout, M, Nn, HEADS, npix, dim, cop, r.s));
        else ENG_CALL(ds_attn_finalize_cat(part, sb, sb + HID, c_total, wout, M, Nn, HEADS, npix, dim, cop, r.s)); }, 2);
      else add([=](const Run& r) { ENG_CALL(ds_attn_finalize(part, wout, M, Nn, HEADS, npix, dim, cop, r.s)); }, 2);
    }
    act_t* y = scr("atty", Nn, h, w, a.dim);
    Stats st_y;
    { ConvOpts o; o.out = y; o.out_c = a.dim; o.want_stats = true; o.weight_override = M; o.per_sample_weights = true;
      conv(a.out, qp, HID, nullptr, 0, Nn, h, w, o, &st_y); }
    act_t* ob = act(Nn, h, w, a.dim);
    {
      const act_t* xp = x.p; const float *g = a.gamma, *bt = a.beta; const int dim = a.dim;
      add([=](const Run& r) { ENG_CALL(ds_gn_apply_residual(y, xp, ob, st_y.buf, st_y.slots, g, bt, Nn, dim, npix, x_mod, r.s)); });
    }
    return Act{ob, a.dim};
  };

  // ---- network (diffusion.py:187-258) ----
  const int n_stage = cfg.n_levels - 1;
  ENG_REQUIRE((H >> n_stage) >= 1 && (W >> n_stage) >= 1, "H=%d, W=%d: the map vanishes after %d stride-2 stages", H, W, n_stage);
  // the GroupNorm fold's border classes assume every level is >= 2 pixels in both directions
  if ((H >> n_stage) < 2 || (W >> n_stage) < 2) fail(DS_ERR_UNSUPPORTED, "H=%d, W=%d: a level of the U-Net is 1 pixel wide or high (minimum input %d x %d)", H, W, 2 << n_stage, 2 << n_stage);
  int h = H, w = W;
  // Classifier-free guidance inside the sampling loop: both halves of the doubled batch share the latent and the timestep and
  // differ only through the condition, which first enters in downs.0.1: init_conv and downs.0.0 are evaluated once for the nb
  // distinct latents and read with a batch modulus afterwards.
  shared = (x_batch_mod > 0 && uniform_time && nb < N) ? nb : 0;
  const int n0 = shared ? shared : N;
  act_t* x0 = act(n0, h, w, dd[0]);
  act_t* col = act(nb, h, w, 224);
  {
    const int nbb = nb, cin = cfg.in_dim, HH = H, WW = W;
    add([=](const Run& r) { ENG_CALL(ds_stem_im2col(r.x, col, nbb, cin, HH, WW, r.s)); });
    ConvOpts o; o.out = x0; o.out_c = dd[0]; o.src_batch_mod = shared ? 0 : x_batch_mod;
    conv(net->stem, col, 224, nullptr, 0, n0, h, w, o, nullptr);
  }
  const Act none{nullptr, 0};
  std::vector<Act> hs;
  std::vector<std::pair<int, int>> sizes;
  hs.push_back(Act{x0, dd[0]});
  Act x{x0, dd[0]};
  Stats st;
  for (int i = 0; i < n_stage; ++i) {
    const std::string p = "downs." + std::to_string(i) + ".";
    if (i == 0 && shared) {
      x = any_block(p + "0.", x, none, h, w, shared, 0, &st);
      x = attn(p + "1.", x, st, h, w, shared); hs.push_back(x);
    } else {
      x = any_block(p + "0.", x, none, h, w, N, 0, &st);
      x = attn(p + "1.", x, st, h, w, 0); hs.push_back(x);
    }
    x = any_block(p + "2.", x, none, h, w, N, 0, &st);
    x = attn(p + "3.", x, st, h, w, 0); hs.push_back(x);
    const PackedConv& pc = net->samplers.at(p + "4.");
    act_t* d = act(N, h / 2, w / 2, pc.cout);
    { ConvOpts o; o.out = d; o.out_c = pc.cout; conv(pc, x.p, x.c, nullptr, 0, N, h, w, o, nullptr); }
    sizes.push_back({h, w});
    x = Act{d, pc.cout}; h /= 2; w /= 2; hs.push_back(x);
  }
  for (int j = 0; j < cfg.mid_depth - 1; ++j) { x = any_block("mid_left." + std::to_string(j) + ".", x, none, h, w, N, 0, &st); hs.push_back(x); }
  x = any_block("mid_mid.0.", x, none, h, w, N, 0, &st);
  x = attn("mid_mid.1.", x, st, h, w, 0);
  x = any_block("mid_mid.2.", x, none, h, w, N, 0, &st);
  for (int j = 0; j < cfg.mid_depth - 1; ++j) { Act s = hs.back(); hs.pop_back(); x = any_block("mid_right." + std::to_string(j) + ".", s, x, h, w, N, 0, &st); }
  for (int i = 0; i < n_stage; ++i) {
    const std::string p = "ups." + std::to_string(i) + ".";
    { Act s = hs.back(); hs.pop_back(); x = any_block(p + "0.", s, x, h, w, N, 0, &st); }
    x = attn(p + "1.", x, st, h, w, 0);
    const PackedConv& pc = net->samplers.at(p + "2.");
    const int hp = sizes.back().first, wp = sizes.back().second;     // the skip's size: 2h / 2w, or one more where the level was odd (pad_to_match)
    sizes.pop_back();
    const bool exact = hp == 2 * h && wp == 2 * w;
    act_t* u = exact ? act(N, hp, wp, pc.cout) : A.alloc<act_t>((size_t)N * hp * wp * pc.cout, /*zero=*/true);     // the conv never writes the padding
    { ConvOpts o; o.out = u; o.out_c = pc.cout; o.out_hp = hp; o.out_wp = wp; conv(pc, x.p, x.c, nullptr, 0, N, h, w, o, nullptr); }
    x = Act{u, pc.cout}; h = hp; w = wp;
    { Act s = hs.back(); hs.pop_back(); x = any_block(p + "3.", s, x, h, w, N, 0, &st); }
    x = attn(p + "4.", x, st, h, w, 0);
    { Act s = hs.back(); hs.pop_back(); x = any_block(p + "5.", s, x, h, w, N, 0, &st); }
    x = attn(p + "6.", x, st, h, w, 0);
  }
  { Act s = hs.back(); hs.pop_back(); x = any_block("final_conv.0.", s, x, h, w, N, shared, &st); }      // the last skip is init_conv's output
  { ConvOpts o; o.out_f32 = eps; conv(net->final_conv, x.p, x.c, nullptr, 0, N, h, w, o, nullptr); }
}

static UnetPlan* get_plan(Unet* u, int N, int H, int W, int mod, int ut) {
  ENG_REQUIRE(u->packed, "ds_unet: ds_unet_finalize() has not been called");
  for (auto& p : u->plans)
    if (p->N == N && p->H == H && p->W == W && p->x_batch_mod == mod && p->uniform_time == ut) return p.get();
  u->plans.emplace_back(new UnetPlan(u, N, H, W, mod, ut));
  return u->plans.back().get();
}

// =====================================================================================================================
// VQGAN: quantiser, decoder, encoder (layer stacks of model/VQGAN.py:278-321, 332-387)
// =====================================================================================================================
enum { L_DOWN, L_UP, L_RES, L_ATTN, L_NORM, L_RELU, L_CONV1, L_CONV1_NOBIAS };
struct LayerSpec { int idx, kind, cin, cout; };
struct Layer {
  LayerSpec s;
  PackedConv conv, shortc, qkv, out;
  bool has_short = false;
  float *gamma = nullptr, *beta = nullptr, *wout = nullptr;
};
static const int VQ_DH = 32, GN_CHUNKS = 64;

struct StackPlan;
struct Stack {
  std::string prefix;
  bool is_decoder = false, wide_tiles = false;
  int G = 16, res_act = 2;
  std::vector<Layer> layers;
  std::vector<std::unique_ptr<StackPlan>> plans;
};
struct StackPlan {
  int B, H, W;
  Arena A;
  std::vector<Op> ops;
  int launches = 0;
  float* out_f32 = nullptr;
  int out_c = 0, out_h = 0, out_w = 0;
  StackPlan(const Stack& st, int B_, int H_, int W_);
  void run(const float* in, cudaStream_t s) { Run r{in, nullptr, s}; for (auto& f : ops) f(r); }
};

struct Vqgan {
  ds_vqgan_config cfg;
  StateDict sd;
  bool packed = false;
  std::unique_ptr<Arena> weights;
  float* codebook = nullptr;
  Stack enc, dec;
  void pack();
};

static void layer_plans(const ds_vqgan_config& cfg, std::vector<LayerSpec>& enc, std::vector<LayerSpec>& dec) {
  std::vector<int> hc(cfg.hidden_channels, cfg.hidden_channels + cfg.n_hidden);
  auto in_attn = [&](int c) { for (int i = 0; i < cfg.n_attn_pos; ++i) if (cfg.attn_pos[i] == c) return true; return false; };
  const int depth = cfg.block_depth;
  int idx = 1, cur = hc[0];
  enc.push_back({0, L_DOWN, cfg.in_channels, hc[0]});
  auto enc_blocks = [&]() {
    for (int d = 0; d < depth - 1; ++d) {
      enc.push_back({idx++, L_RES, cur, cur});
      if (in_attn(cur)) enc.push_back({idx++, L_ATTN, cur, cur});
    }
  };
  for (size_t i = 1; i < hc.size(); ++i) {
    enc_blocks();
    enc.push_back({idx++, L_NORM, cur, cur});
    enc.push_back({idx++, L_RELU, cur, cur});
    enc.push_back({idx++, L_DOWN, cur, hc[i]});
    cur = hc[i];
  }
  enc_blocks();
  enc.push_back({idx++, L_NORM, cur, cur});
  enc.push_back({idx++, L_RELU, cur, cur});
  enc.push_back({idx, L_CONV1, cur, cfg.embedding_dim});

  std::vector<int> rc(hc.rbegin(), hc.rend());
  dec.push_back({0, L_CONV1_NOBIAS, cfg.embedding_dim, rc[0]});
  idx = 1; cur = rc[0];
  auto dec_blocks = [&]() {
    for (int d = 0; d < depth - 1; ++d) {
      if (in_attn(cur)) dec.push_back({idx++, L_ATTN, cur, cur});
      dec.push_back({idx++, L_RES, cur, cur});
    }
  };
  dec_blocks();
  for (size_t i = 1; i < rc.size(); ++i) {
    dec.push_back({idx++, L_NORM, cur, cur});
    dec.push_back({idx++, L_RELU, cur, cur});
    dec.push_back({idx++, L_UP, cur, rc[i]});
    cur = rc[i];
    dec_blocks();
  }
  dec.push_back({idx++, L_NORM, cur, cur});
  dec.push_back({idx++, L_RELU, cur, cur});
  dec.push_back({idx++, L_UP, cur, cur});
  dec.push_back({idx, L_RES, cur, cfg.out_channels});
}

void Vqgan::pack() {
  weights.reset(new Arena());
  Arena& A = *weights;
  enc.plans.clear(); dec.plans.clear(); enc.layers.clear(); dec.layers.clear();
  std::vector<LayerSpec> es, dsp;
  layer_plans(cfg, es, dsp);
  codebook = A.upload(need(sd, "_vq_vae._embedding.weight", {cfg.num_embeddings, cfg.embedding_dim}).v);
  auto build = [&](Stack& st, const std::string& prefix, const std::vector<LayerSpec>& specs, bool is_dec) {
    st.prefix = prefix; st.is_decoder = is_dec; st.G = cfg.num_groups; st.wide_tiles = cfg.batch_invariant != 0;
    // the encoder's ResnetBlocks get the literal string "act_type" (VQGAN.py:441) -> swish; decoder: the configured act_type
    st.res_act = (!is_dec || cfg.act_relu == 0) ? 2 : 1;
    for (const LayerSpec& s : specs) {
      Layer Ly;
      Ly.s = s;
      const std::string p = prefix + std::to_string(s.idx) + ".";
      const int cin = s.cin, cout = s.cout, cp = pad32(cin);
      switch (s.kind) {
        case L_DOWN: Ly.conv = pack_conv_down(A, need(sd, p + "_conv2d.weight", {cout, cin, 4, 4}), need(sd, p + "_conv2d.bias", {cout}).v.data(), cp); break;
        case L_UP: Ly.conv = pack_conv_up(A, need(sd, p + "_conv2d.weight", {cin, cout, 4, 4}), need(sd, p + "_conv2d.bias", {cout}).v.data(), cp); break;
        case L_RES:
          Ly.gamma = A.upload(need(sd, p + "norm1.weight", {cin}).v); Ly.beta = A.upload(need(sd, p + "norm1.bias", {cin}).v);
          Ly.conv = pack_conv_s1(A, need(sd, p + "conv1.weight", {cout, cin, 3, 3}), need(sd, p + "conv1.bias", {cout}).v.data(), nullptr, nullptr, cp);
          if (cin != cout) {
            Ly.has_short = true;
            Ly.shortc = pack_conv_s1(A, need(sd, p + "nin_shortcut.weight", {cout, cin, 1, 1}), need(sd, p + "nin_shortcut.bias", {cout}).v.data(), nullptr, nullptr, cp);
          }
          break;
        case L_ATTN:
          Ly.qkv = pack_conv_s1(A, need(sd, p + "to_qkv.weight", {3 * VQ_DH, cin, 1, 1}), nullptr, nullptr, nullptr, cp);
          Ly.wout = A.upload(need(sd, p + "to_out.weight", {cin, VQ_DH, 1, 1}).v);
          Ly.out = pack_bias_only(A, need(sd, p + "to_out.bias", {cin}).v.data(), VQ_DH, cin);
          if (cfg.attn_with_skip) {
            Ly.has_short = true;
            Ly.shortc = pack_conv_s1(A, need(sd, p + "nin_shortcut.weight", {cin, cin, 1, 1}), need(sd, p + "nin_shortcut.bias", {cin}).v.data(), nullptr, nullptr, cp);
          }
          break;
        case L_NORM: Ly.gamma = A.upload(need(sd, p + "weight", {cin}).v); Ly.beta = A.upload(need(sd, p + "bias", {cin}).v); break;
        case L_RELU: break;
        case L_CONV1: Ly.conv = pack_conv_s1(A, need(sd, p + "weight", {cout, cin, 1, 1}), need(sd, p + "bias", {cout}).v.data(), nullptr, nullptr, cp); break;
        case L_CONV1_NOBIAS: Ly.conv = pack_conv_s1(A, need(sd, p + "weight", {cout, cin, 1, 1}), nullptr, nullptr, nullptr, cp); break;
      }
      st.layers.push_back(Ly);
    }
  };
  build(enc, "_encoder._layers.", es, false);
  build(dec, "_decoder._layers.", dsp, true);
  packed = true;
}

StackPlan::StackPlan(const Stack& st, int B_, int H_, int W_) : B(B_), H(H_), W(W_) {
  const int Bn = B;
  auto act = [&](int h, int w, int c) { return A.alloc<act_t>((size_t)Bn * h * w * pad32(c), /*zero=*/true); };     // padded channels stay zero
  auto add = [&](Op f, int nl = 1) { ops.push_back(std::move(f)); launches += nl; };
  const bool wide_tiles = st.wide_tiles;
  auto conv = [&](const PackedConv& pc, const act_t* src, int h, int w, const ConvOpts& o_) {
    ConvOpts o = o_; o.wide_tiles = wide_tiles;
    const ds_conv_gemm_args a = conv_args(A, pc, src, pc.cin, nullptr, 0, Bn, h, w, o, nullptr);
    add([a](const Run& r) { ENG_CALL(ds_conv_gemm(&a, r.s)); });
  };
  const int G = st.G;
  auto gn_act = [&](const act_t* x, int c, int h, int w, const float* gamma, const float* beta, int actv) -> act_t* {
    float* part = A.alloc<float>((size_t)Bn * G * GN_CHUNKS * 2);
    act_t* out = act(h, w, c);
    const int cp = pad32(c);
    const long long hw = (long long)h * w;
    add([=](const Run& r) { ENG_CALL(ds_group_stats(x, part, Bn, c, cp, G, hw, GN_CHUNKS, r.s)); });
    add([=](const Run& r) { ENG_CALL(ds_gn_act(x, out, part, GN_CHUNKS, gamma, beta, Bn, c, cp, G, hw, 1e-6f, actv, r.s)); });
    return out;
  };
  int h = H, w = W;
  const int first_c = st.layers.front().s.cin;
  act_t* x = act(h, w, first_c);
  {
    act_t* x_ = x; const int HH = H, WW = W;
    add([=](const Run& r) { ENG_CALL(ds_nchw_f32_to_nhwc_bf16(r.x, x_, Bn, first_c, pad32(first_c), (long long)HH * WW, r.s)); });
  }
  const Layer* pending = nullptr;
  const int last_idx = st.layers.back().s.idx;
  for (const Layer& Ly : st.layers) {
    const int cin = Ly.s.cin, cout = Ly.s.cout;
    switch (Ly.s.kind) {
      case L_DOWN: {
        act_t* o_ = act(h / 2, w / 2, cout);
        ConvOpts o; o.out = o_; o.out_c = pad32(cout);
        conv(Ly.conv, x, h, w, o);
        x = o_; h /= 2; w /= 2;
      } break;
      case L_UP: {
        act_t* o_ = act(2 * h, 2 * w, cout);
        ConvOpts o; o.out = o_; o.out_c = pad32(cout);
        conv(Ly.conv, x, h, w, o);
        x = o_; h *= 2; w *= 2;
      } break;
      case L_RES: {
        act_t* t = gn_act(x, cin, h, w, Ly.gamma, Ly.beta, st.res_act);
        if (cout % 16 == 0) {
          const act_t* r_ = x;
          if (Ly.has_short) {
            act_t* rr = act(h, w, cout);
            ConvOpts o; o.out = rr; o.out_c = pad32(cout);
            conv(Ly.shortc, x, h, w, o);
            r_ = rr;
          }
          act_t* o_ = act(h, w, cout);
          ConvOpts o; o.out = o_; o.out_c = pad32(cout); o.residual = r_; o.res_c = pad32(cout);
          conv(Ly.conv, t, h, w, o);
          x = o_;
        } else {
          // final decoder block (80 -> 3): both branches as fp32 NCHW, summed inside the head kernel
          ENG_REQUIRE(Ly.s.idx == last_idx && st.is_decoder && Ly.has_short, "a ResnetBlock with %d output channels is only supported as the decoder's last layer", cout);
          float* a32 = A.alloc<float>((size_t)Bn * cout * h * w);
          float* b32 = A.alloc<float>((size_t)Bn * cout * h * w);
          { ConvOpts o; o.out_f32 = a32; conv(Ly.conv, t, h, w, o); }
          { ConvOpts o; o.out_f32 = b32; conv(Ly.shortc, x, h, w, o); }
          out_f32 = A.alloc<float>((size_t)Bn * cout * h * w);
          float* of = out_f32; const long long hw = (long long)h * w;
          ENG_REQUIRE(cout == 3, "decoder heads expect 3 output channels (got %d)", cout);
          add([=](const Run& r) { ENG_CALL(ds_decoder_head(a32, b32, of, Bn, hw, r.s)); });
          out_c = cout; out_h = h; out_w = w;
          x = nullptr;
        }
      } break;
      case L_ATTN: {
        const long long npix = (long long)h * w;
        act_t* qkv = act(h, w, 3 * VQ_DH);
        { ConvOpts o; o.out = qkv; o.out_c = 3 * VQ_DH; conv(Ly.qkv, x, h, w, o); }
        act_t* qp = act(h, w, VQ_DH);
        float* part = A.alloc<float>((size_t)ds_attn_part_floats(Bn, 1, npix));
        act_t* M = A.alloc<act_t>((size_t)Bn * Ly.out.cout_pad * VQ_DH);
        add([=](const Run& r) { ENG_CALL(ds_attn_ctx_partial(qkv, qp, part, Bn, 1, npix, 1, 1.0f, r.s)); });
        const float* wout = Ly.wout; const int cop = Ly.out.cout_pad;
        add([=](const Run& r) { ENG_CALL(ds_attn_finalize(part, wout, M, Bn, 1, npix, cin, cop, r.s)); }, 2);
        const act_t* r_ = nullptr;
        if (Ly.has_short) {
          act_t* rr = act(h, w, cin);
          ConvOpts o; o.out = rr; o.out_c = pad32(cin);
          conv(Ly.shortc, x, h, w, o);
          r_ = rr;
        }
        act_t* o_ = act(h, w, cin);
        ConvOpts o; o.out = o_; o.out_c = pad32(cin); o.residual = r_; o.res_c = pad32(cin); o.weight_override = M; o.per_sample_weights = true;
        conv(Ly.out, qp, h, w, o);
        x = o_;
      } break;
      case L_NORM: pending = &Ly; break;
      case L_RELU:
        ENG_REQUIRE(pending != nullptr, "ReLU without a preceding Normalize");
        x = gn_act(x, pending->s.cin, h, w, pending->gamma, pending->beta, 1);       // Normalize + nn.ReLU fused into one apply pass
        pending = nullptr;
        break;
      case L_CONV1: case L_CONV1_NOBIAS:
        if (cout % 16 == 0) {
          act_t* o_ = act(h, w, cout);
          ConvOpts o; o.out = o_; o.out_c = pad32(cout);
          conv(Ly.conv, x, h, w, o);
          x = o_;
        } else {
          out_f32 = A.alloc<float>((size_t)Bn * cout * h * w);
          ConvOpts o; o.out_f32 = out_f32;
          conv(Ly.conv, x, h, w, o);
          out_c = cout; out_h = h; out_w = w;
          x = nullptr;
        }
        break;
    }
  }
  ENG_REQUIRE(out_f32 != nullptr, "layer stack must end in an fp32 output (decoder heads or the encoder's 1x1 conv)");
}

static StackPlan* stack_plan(Stack& st, int B, int H, int W) {
  for (auto& p : st.plans)
    if (p->B == B && p->H == H && p->W == W) return p.get();
  st.plans.emplace_back(new StackPlan(st, B, H, W));
  return st.plans.back().get();
}

// =====================================================================================================================
// Sampling graph
// =====================================================================================================================
struct SampleGraph {
  Unet* unet; Vqgan* vq;
  ds_sample_buffers b;
  int B, C, H, W, n_iter, cfg_on, N;
  UnetPlan* plan = nullptr;
  StackPlan* dec = nullptr;
  Arena A;
  float* frames = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
  ~SampleGraph() { if (exec) cudaGraphExecDestroy(exec); if (graph) cudaGraphDestroy(graph); }
  void body(cudaStream_t s, int steps) {
    const size_t per = (size_t)B * C * H * W;
    for (int k = 0; k < steps; ++k) {
      plan->run(b.d_imgs + k * per, b.d_ttab + k, s);
      const float* eps_u = cfg_on ? plan->eps : nullptr;
      const float* eps_c = cfg_on ? plan->eps + per : plan->eps;
      ENG_CALL(ds_ddim_step(eps_u, eps_c, b.d_imgs + k * per, b.d_noise ? b.d_noise + k * per : nullptr, b.d_coef + 8 * k, b.d_imgs + (k + 1) * per,
                            (long long)per, s));
      if (b.d_masks)
        ENG_CALL(ds_mask_blend(b.d_guide, b.d_init_noise, b.d_masks + k * per, C, b.d_blend_coef + 2 * k, b.d_imgs + (k + 1) * per, B, C, (long long)H * W, s));
    }
  }
  void tail(cudaStream_t s) {
    if (!vq) return;
    const size_t per = (size_t)B * C * H * W;
    ENG_CALL(ds_vq_quantize(b.d_imgs + (size_t)n_iter * per, vq->codebook, vq->cfg.num_embeddings, b.d_quantized, b.d_indices, B, (long long)H * W, s));
    dec->run(b.d_quantized, s);
    const int T = dec->out_w;
    if (b.d_spec) ENG_CUDA(cudaMemcpyAsync(b.d_spec, dec->out_f32, (size_t)B * dec->out_c * dec->out_h * dec->out_w * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (b.d_wave) ENG_CALL(ds_stft_decode_istft(dec->out_f32, frames, b.d_wave, B, T, s));
  }
};

}  // namespace eng
}  // namespace ds

using namespace ds;
using namespace ds::eng;

struct ds_unet { Unet u; };
struct ds_vqgan { Vqgan v; };
struct ds_sample_graph { SampleGraph g; };

static void store_param(StateDict& sd, const char* name, const float* data, const long long* shape, int ndim) {
  ENG_REQUIRE(name && data && (shape || ndim == 0) && ndim >= 0 && ndim <= 8, "load: bad arguments");
  HostTensor t;
  long long n = 1;
  for (int i = 0; i < ndim; ++i) { ENG_REQUIRE(shape[i] > 0, "load(%s): extent %lld", name, shape[i]); t.shape.push_back(shape[i]); n *= shape[i]; }
  t.v.resize((size_t)n);
  ENG_CUDA(cudaMemcpy(t.v.data(), data, (size_t)n * sizeof(float), cudaMemcpyDefault));      // host or device source
  sd[name] = std::move(t);
}

extern "C" {

int ds_unet_create(const ds_unet_config* cfg, ds_unet** out) {
  return guarded([&] {
    ENG_REQUIRE(cfg && out, "ds_unet_create: null argument");
    ENG_REQUIRE(cfg->n_levels >= 2 && cfg->n_levels <= DS_MAX_LEVELS, "ds_unet_create: n_levels=%d", cfg->n_levels);
    if (cfg->attn_type < 0 || cfg->attn_type > 1) fail(DS_ERR_UNSUPPORTED, "ds_unet_create: attn_type %d (0 = 'linear_add', 1 = 'linear_cat')", cfg->attn_type);      // diffusion.py:96
    if (cfg->condition_type < 0 || cfg->condition_type > 1) fail(DS_ERR_UNSUPPORTED, "ds_unet_create: condition_type %d", cfg->condition_type);                // diffusion_components.py:165
    ENG_REQUIRE(cfg->condition_type == 0 || cfg->n_label_class >= 1, "ds_unet_create: n_label_class=%d", cfg->n_label_class);
    ENG_REQUIRE(cfg->down_dims[0] == cfg->up_dims[cfg->n_levels - 1] && cfg->up_dims[0] == cfg->down_dims[cfg->n_levels - 1], "ds_unet_create: down_dims / up_dims do not mirror");
    for (int i = 0; i < cfg->n_levels; ++i)
      ENG_REQUIRE(cfg->down_dims[i] > 0 && cfg->down_dims[i] % 32 == 0 && cfg->up_dims[i] > 0 && cfg->up_dims[i] % 32 == 0, "ds_unet_create: channel widths must be multiples of 32");
    std::unique_ptr<ds_unet> h(new ds_unet());
    h->u.cfg = *cfg;
    if (h->u.cfg.out_dim <= 0) h->u.cfg.out_dim = cfg->in_dim;
    if (h->u.cfg.time_dim <= 0) h->u.cfg.time_dim = cfg->down_dims[0] * 4;
    if (h->u.cfg.mid_depth <= 0) h->u.cfg.mid_depth = 3;
    if (h->u.cfg.convnext_mult <= 0) h->u.cfg.convnext_mult = 2;
    *out = h.release();
  });
}
void ds_unet_destroy(ds_unet* h) { delete h; }
int ds_unet_load(ds_unet* h, const char* name, const float* data, const long long* shape, int ndim) {
  return guarded([&] { ENG_REQUIRE(h, "ds_unet_load: null handle"); store_param(h->u.sd, name, data, shape, ndim); h->u.packed = false; });
}
int ds_unet_finalize(ds_unet* h) {
  return guarded([&] { ENG_REQUIRE(h, "ds_unet_finalize: null handle"); h->u.pack(); });
}
int ds_unet_forward(ds_unet* h, const float* d_x, const long long* d_t, const void* d_cond, float* d_out, int N, int H, int W, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h && d_x && d_t && d_out && N > 0 && H > 0 && W > 0, "ds_unet_forward: bad arguments");
    UnetPlan* p = get_plan(&h->u, N, H, W, 0, 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (d_cond == nullptr) {      // condition=None (diffusion.py:199-202, diffusion_components.py:279-283): no label_query / label_key terms
      ENG_CUDA(cudaMemsetAsync(p->sbias, 0, (size_t)N * h->u.c_total * sizeof(float), s));
    } else {
      ENG_CUDA(cudaMemcpyAsync(p->cond, d_cond, p->cond_bytes, cudaMemcpyDeviceToDevice, s));
      p->run_cond(s);
    }
    p->run(d_x, d_t, s, d_cond == nullptr);
    ENG_CUDA(cudaMemcpyAsync(d_out, p->eps, (size_t)N * h->u.cfg.out_dim * H * W * sizeof(float), cudaMemcpyDeviceToDevice, s));
  });
}
int ds_unet_plan_get(ds_unet* h, int N, int H, int W, int x_batch_mod, int uniform_time, ds_unet_plan_io* io) {
  return guarded([&] {
    ENG_REQUIRE(h && io && N > 0 && H > 0 && W > 0 && x_batch_mod >= 0, "ds_unet_plan_get: bad arguments");
    UnetPlan* p = get_plan(&h->u, N, H, W, x_batch_mod, uniform_time ? 1 : 0);
    for (size_t i = 0; i < h->u.plans.size(); ++i)
      if (h->u.plans[i].get() == p) io->plan = (int)i;
    io->d_cond = p->cond; io->d_eps = p->eps; io->launches = p->launches; io->cond_launches = (int)p->cond_ops.size();
  });
}
int ds_unet_plan_run_cond(ds_unet* h, int plan, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h && plan >= 0 && plan < (int)h->u.plans.size(), "ds_unet_plan_run_cond: bad plan");
    h->u.plans[plan]->run_cond((cudaStream_t)stream);
  });
}
int ds_unet_plan_run(ds_unet* h, int plan, const float* d_x, const long long* d_t, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h && plan >= 0 && plan < (int)h->u.plans.size() && d_x && d_t, "ds_unet_plan_run: bad arguments");
    h->u.plans[plan]->run(d_x, d_t, (cudaStream_t)stream);
  });
}

int ds_vqgan_create(const ds_vqgan_config* cfg, ds_vqgan** out) {
  return guarded([&] {
    ENG_REQUIRE(cfg && out, "ds_vqgan_create: null argument");
    ENG_REQUIRE(cfg->n_hidden >= 1 && cfg->n_hidden <= DS_MAX_LEVELS && cfg->n_attn_pos >= 0 && cfg->n_attn_pos <= DS_MAX_LEVELS, "ds_vqgan_create: bad channel lists");
    if (cfg->embedding_dim != 4) fail(DS_ERR_UNSUPPORTED, "ds_vqgan_create: embedding_dim must be 4 (deployed VQGAN, app.py:32)");
    ENG_REQUIRE(cfg->block_depth >= 1 && cfg->num_groups >= 1 && cfg->num_embeddings >= 1, "ds_vqgan_create: bad sizes");
    std::unique_ptr<ds_vqgan> h(new ds_vqgan());
    h->v.cfg = *cfg;
    *out = h.release();
  });
}
void ds_vqgan_destroy(ds_vqgan* h) { delete h; }
int ds_vqgan_load(ds_vqgan* h, const char* name, const float* data, const long long* shape, int ndim) {
  return guarded([&] { ENG_REQUIRE(h, "ds_vqgan_load: null handle"); store_param(h->v.sd, name, data, shape, ndim); h->v.packed = false; });
}
int ds_vqgan_finalize(ds_vqgan* h) {
  return guarded([&] { ENG_REQUIRE(h, "ds_vqgan_finalize: null handle"); h->v.pack(); });
}
int ds_vqgan_quantize(ds_vqgan* h, const float* d_x, float* d_out, long long* d_idx, int B, long long hw, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h && h->v.packed, "ds_vqgan_quantize: ds_vqgan_finalize() has not been called");
    ENG_CALL(ds_vq_quantize(d_x, h->v.codebook, h->v.cfg.num_embeddings, d_out, d_idx, B, hw, stream));
  });
}
static void run_stack(Stack& st, const float* d_in, float* d_out, int B, int H, int W, cudaStream_t s) {
  StackPlan* p = stack_plan(st, B, H, W);
  p->run(d_in, s);
  ENG_CUDA(cudaMemcpyAsync(d_out, p->out_f32, (size_t)B * p->out_c * p->out_h * p->out_w * sizeof(float), cudaMemcpyDeviceToDevice, s));
}
int ds_vqgan_decode(ds_vqgan* h, const float* d_latent, float* d_spec, int B, int H, int W, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h && h->v.packed && d_latent && d_spec && B > 0 && H > 0 && W > 0, "ds_vqgan_decode: bad arguments or ds_vqgan_finalize() not called");
    run_stack(h->v.dec, d_latent, d_spec, B, H, W, (cudaStream_t)stream);
  });
}
int ds_vqgan_encode(ds_vqgan* h, const float* d_spec, float* d_latent, int B, int H, int W, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h && h->v.packed && d_spec && d_latent && B > 0 && H > 0 && W > 0, "ds_vqgan_encode: bad arguments or ds_vqgan_finalize() not called");
    run_stack(h->v.enc, d_spec, d_latent, B, H, W, (cudaStream_t)stream);
  });
}

int ds_sample_graph_build(ds_unet* unet, ds_vqgan* vqgan, const ds_sample_buffers* bufs, int B, int H, int W, int n_iter, int cfg_on, int use_graph,
                          void* stream, ds_sample_graph** out) {
  return guarded([&] {
    ENG_REQUIRE(unet && bufs && out && B > 0 && H > 0 && W > 0 && n_iter > 0, "ds_sample_graph_build: bad arguments");
    ENG_REQUIRE(bufs->d_imgs && bufs->d_coef && bufs->d_ttab && bufs->d_cond, "ds_sample_graph_build: imgs, coef, ttab and cond are required");
    ENG_REQUIRE(!bufs->d_masks || (bufs->d_guide && bufs->d_init_noise && bufs->d_blend_coef), "ds_sample_graph_build: inpainting needs guide, init_noise, masks and blend_coef");
    ENG_REQUIRE(!vqgan || (vqgan->v.packed && bufs->d_quantized && bufs->d_indices), "ds_sample_graph_build: the tail needs a finalized VQGAN, d_quantized and d_indices");
    std::unique_ptr<ds_sample_graph> h(new ds_sample_graph());
    SampleGraph& g = h->g;
    g.unet = &unet->u; g.vq = vqgan ? &vqgan->v : nullptr; g.b = *bufs;
    g.B = B; g.C = unet->u.cfg.in_dim; g.H = H; g.W = W; g.n_iter = n_iter; g.cfg_on = cfg_on ? 1 : 0;
    g.N = cfg_on ? 2 * B : B;
    g.plan = get_plan(&unet->u, g.N, H, W, cfg_on ? B : 0, 1);
    g.launches = n_iter * (g.plan->launches + 1 + (bufs->d_masks ? 1 : 0));
    if (g.vq) {
      g.dec = stack_plan(g.vq->dec, B, H, W);
      g.frames = g.A.alloc<float>((size_t)B * g.dec->out_w * 1024);
      g.launches += 1 + g.dec->launches + (bufs->d_wave ? 2 : 0);
    }
    cudaStream_t s = (cudaStream_t)stream;
    // warm-up outside capture (lazy one-time initialisations: function attributes, twiddle tables), then capture
    ENG_CUDA(cudaMemcpyAsync(g.plan->cond, bufs->d_cond, g.plan->cond_bytes, cudaMemcpyDeviceToDevice, s));
    g.plan->run_cond(s);
    g.body(s, 1);
    g.tail(s);
    ENG_CUDA(cudaStreamSynchronize(s));
    if (use_graph) {
      cudaStream_t cap;
      ENG_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
      try {
        ENG_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        g.body(cap, n_iter);
        g.tail(cap);
        ENG_CUDA(cudaStreamEndCapture(cap, &g.graph));
      } catch (...) {
        cudaGraph_t junk = nullptr;
        cudaStreamEndCapture(cap, &junk);
        if (junk) cudaGraphDestroy(junk);
        cudaStreamDestroy(cap);
        throw;
      }
      ENG_CUDA(cudaStreamDestroy(cap));
      ENG_CUDA(cudaGraphInstantiate(&g.exec, g.graph, 0));
    }
    *out = h.release();
  });
}
int ds_sample_graph_run(ds_sample_graph* h, void* stream) {
  return guarded([&] {
    ENG_REQUIRE(h, "ds_sample_graph_run: null handle");
    SampleGraph& g = h->g;
    cudaStream_t s = (cudaStream_t)stream;
    // the condition projections are step-invariant: once per call, outside the graph
    ENG_CUDA(cudaMemcpyAsync(g.plan->cond, g.b.d_cond, g.plan->cond_bytes, cudaMemcpyDeviceToDevice, s));
    g.plan->run_cond(s);
    if (g.exec) ENG_CUDA(cudaGraphLaunch(g.exec, s));
    else { g.body(s, g.n_iter); g.tail(s); }
  });
}
int ds_sample_graph_launches(const ds_sample_graph* h) { return h ? h->g.launches + (int)h->g.plan->cond_ops.size() : -1; }
void ds_sample_graph_destroy(ds_sample_graph* h) { delete h; }

}  // extern "C"
