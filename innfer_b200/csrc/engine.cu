// C-ABI of the RRDB path (include/innfer_b200.h): network handle, weight loading under the
// reference's state-dict key names, forward schedule over planar-chunk buffers, chop_forward.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/innfer_b200.h"
#include "color_fix.cuh"
#include "conv_direct.cuh"
#include "i2i.cuh"
#include "layers.cuh"
#include "pan_ops.cuh"
#include "pixel_ops.cuh"
#include "sync_ops.cuh"

using namespace innfer;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

thread_local int g_code = 0;

int fail(int code, const std::string& msg) {
  g_err = msg;
  g_code = code;
  return code;
}
int g_err_code() { return g_code; }   // code of the last fail() on this thread
int cuda_fail(cudaError_t e, const char* what) {
  return fail(INNFER_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU_TRY(expr)                                          \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr);       \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  unsigned gen = 0;   // bumped whenever the allocation changes (what is known about the old contents is void)
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    ++gen;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) return (int)e;
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    ++gen;
  }
};

struct Param {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

PixelDType to_pix(int dtype) { return dtype == INNFER_F16 ? kF16 : (dtype == INNFER_F32 ? kF32 : kU8); }

}  // namespace

struct innfer_rrdb {
  innfer_rrdb_cfg cfg;
  int arch = 0;            // 0: RRDBNet (ESRGAN), 1: SRResNet (SRGAN generator), 2: PPON, 3: PAN, 4: pix2pix UNet / CycleGAN ResNet
  // image-to-image generators (arch 4): the whole network lives in csrc/i2i.cu
  I2ICfg i2i_cfg;
  std::unique_ptr<I2INet> i2i;
  bool i2i_per_sample = false;   // the batch is a list of tiles the reference would run one by one (chop_forward)
  float res_scale = 1.f;   // SRResNet residual scaling
  bool ps_mode = true;     // SRResNet upsampler: pixelshuffle (default) or upconv
  int device = 0;
  int num_sms = 148;
  int n_up = 0, up_factor = 2;
  bool finalized = false;
  int max_batch = 95;
  std::map<std::string, Param> params;
  // layers in execution order
  ConvLayer fea, lr_conv, hr0, hr1;
  std::vector<ConvLayer> rdb;  // [nb][3][5]
  std::vector<ConvLayer> c1x1; // [nb][3] ESRGAN+ conv1x1 (cfg.plus)
  std::vector<ConvLayer> ups;
  // PPON (arch 2): residual blocks of 10 convs (c1, d1..d8, c2) -- content module nb*3, then SFEM 6, PFEM 6 --
  // and the three reconstruction tails (ups..., HR_conv0, HR_conv1) CRM / SRM / PRM
  float ppon_alpha = 1.f;
  std::vector<ConvLayer> prb;
  std::vector<ConvLayer> ptail[3];
  // PAN (arch 3): conv_first = fea, conv_last = hr1; prb = SCPA convs [trunk][nb][conv1ab, k1, k3|k2, k4, conv3];
  // ptail[0] = trunk_conv(s); ups = [stage][upconv, PA conv, HRconv]; the FSA projections as one fp32 matrix
  int pan_unf = 24;
  bool pan_sa = true, pan_double = false, pan_hr_act = false;
  float pan_gamma = 0.f;
  float* d_pan_w = nullptr;
  float* d_pan_b = nullptr;
  TmapCache cache;
  // workspace
  DevBuf in_tiles, feat, xbuf[3], hrbuf[2], out_tiles, img_in, img_out;
  // uint8 frames with <= 8 channels fill only chunk 0 of every input tile: the other chunks of in_tiles are known to
  // hold zeros for allocation generation `in_pad_gen` and tile size `in_pad_p` (zero-filled once, invalidated by
  // other writers)
  unsigned in_pad_gen = ~0u;
  int in_pad_p = 0;
  DevBuf pbuf[9], paux[2];   // PPON: F, X0..X2, T1, S, E1, E2 (64 channels each), CAT (256); out_c / out_s at HR
  // optional device-side timing of the conv sequence (bench.py roofline)
  int profiling = 0;   // 1: events around every per-batch conv sequence; 2: around every conv launch, per kernel family
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  uint64_t prof_conv_launches = 0;
  struct FamStat {
    uint64_t launches = 0;
    double flop = 0, bytes = 0;   // algorithmic: real channels only, reference conv arithmetic
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  };
  std::map<std::string, FamStat> fam;
  // fp32-mode workspace lives in the same buffers (sized in bytes)
  ~innfer_rrdb() {
    for (auto& e : prof_events) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    for (auto& f : fam)
      for (auto& e : f.second.ev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
      }
    conv_layer_free(fea);
    conv_layer_free(lr_conv);
    conv_layer_free(hr0);
    conv_layer_free(hr1);
    for (auto& l : rdb) conv_layer_free(l);
    for (auto& l : c1x1) conv_layer_free(l);
    for (auto& l : ups) conv_layer_free(l);
    for (auto& l : prb) conv_layer_free(l);
    for (auto& t : ptail)
      for (auto& l : t) conv_layer_free(l);
    for (auto& b : pbuf) b.release();
    for (auto& b : paux) b.release();
    if (d_pan_w) cudaFree(d_pan_w);
    if (d_pan_b) cudaFree(d_pan_b);
    in_tiles.release();
    feat.release();
    for (auto& b : xbuf) b.release();
    for (auto& b : hrbuf) b.release();
    out_tiles.release();
    img_in.release();
    img_out.release();
  }
  int in_ct() const { return (cfg.in_nc + 15) / 16 * 2; }
  int nf_ct() const { return cfg.nf / 8; }
  // concat buffer: x, x1..x4 (+ 4 chunks for conv1x1(x) in ESRGAN+ mode)
  int cat_ct() const { return arch == 1 ? cfg.nf / 8 : (cfg.nf + 4 * 32) / 8 + (cfg.plus ? 4 : 0); }
  size_t esz() const { return cfg.fp16 ? 2 : 4; }
};

namespace {

// Every exported entry point that needs the handle's device makes it current for the duration of the call and
// restores the caller's device afterwards (a process may drive several GPUs; torch keeps its own notion of the
// current device and must not find it changed behind its back).
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int d) : dev(d) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define GUARD_DEVICE(d)  \
  DeviceGuard _guard(d); \
  if (_guard.err != cudaSuccess) return cuda_fail(_guard.err, "cudaSetDevice")

int build_layer(innfer_rrdb* h, ConvLayer& L, const std::string& prefix, int Cout, int Cin, int up, int ksize = 3,
                bool has_bias = true, int dil = 1) {
  auto wi = h->params.find(prefix + ".weight");
  auto bi = h->params.find(prefix + ".bias");
  if (wi == h->params.end()) return fail(INNFER_E_STATE, "missing key " + prefix + ".weight");
  if (has_bias && bi == h->params.end()) return fail(INNFER_E_STATE, "missing key " + prefix + ".bias");
  const auto& ws = wi->second.shape;
  if (ws.size() != 4 || ws[0] != Cout || ws[1] != Cin || ws[2] != ksize || ws[3] != ksize)
    return fail(INNFER_E_INVALID, "size mismatch for " + prefix + ".weight");
  if (has_bias && (bi->second.shape.size() != 1 || bi->second.shape[0] != Cout))
    return fail(INNFER_E_INVALID, "size mismatch for " + prefix + ".bias");
  std::string err;
  int rc = conv_layer_build(L, wi->second.data.data(), has_bias ? bi->second.data.data() : nullptr, Cout, Cin, up, err,
                            ksize, dil);
  if (rc) return fail(rc == -2 ? INNFER_E_UNSUPPORTED : INNFER_E_CUDA, prefix + ": " + err);
  if (!h->cfg.fp16) {
    rc = conv_direct_upload(L);
    if (rc) return fail(INNFER_E_CUDA, prefix + ": fp32 weight upload failed");
  }
  return 0;
}

int build_ps_layer(innfer_rrdb* h, ConvLayer& L, const std::string& prefix, int Cout, int Cin, int r) {
  auto wi = h->params.find(prefix + ".weight");
  auto bi = h->params.find(prefix + ".bias");
  if (wi == h->params.end()) return fail(INNFER_E_STATE, "missing key " + prefix + ".weight");
  if (bi == h->params.end()) return fail(INNFER_E_STATE, "missing key " + prefix + ".bias");
  const auto& ws = wi->second.shape;
  if (ws.size() != 4 || ws[0] != Cout * r * r || ws[1] != Cin || ws[2] != 3 || ws[3] != 3)
    return fail(INNFER_E_INVALID, "size mismatch for " + prefix + ".weight");
  if (bi->second.shape.size() != 1 || bi->second.shape[0] != Cout * r * r)
    return fail(INNFER_E_INVALID, "size mismatch for " + prefix + ".bias");
  std::string err;
  int rc = conv_layer_build_ps(L, wi->second.data.data(), bi->second.data.data(), Cout, Cin, r, err);
  if (rc) return fail(rc == -2 ? INNFER_E_UNSUPPORTED : INNFER_E_CUDA, prefix + ": " + err);
  if (!h->cfg.fp16) {
    rc = conv_direct_upload(L);
    if (rc) return fail(INNFER_E_CUDA, prefix + ": fp32 weight upload failed");
  }
  return 0;
}

// layer from weights assembled on the host (PAN: merged / re-indexed 1x1 convs); `bias` may be null
int build_custom(innfer_rrdb* h, ConvLayer& L, const std::string& what, const std::vector<float>& w, const float* bias,
                 int Cout, int Cin, int ksize) {
  std::string err;
  int rc = conv_layer_build(L, w.data(), bias, Cout, Cin, 1, err, ksize, 1);
  if (rc) return fail(rc == -2 ? INNFER_E_UNSUPPORTED : INNFER_E_CUDA, what + ": " + err);
  if (!h->cfg.fp16) {
    rc = conv_direct_upload(L);
    if (rc) return fail(INNFER_E_CUDA, what + ": fp32 weight upload failed");
  }
  return 0;
}

// a loaded parameter with exactly this shape, or null (error recorded)
const Param* need_param(innfer_rrdb* h, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = h->params.find(key);
  if (it == h->params.end()) {
    fail(INNFER_E_STATE, "missing key " + key);
    return nullptr;
  }
  if (it->second.shape != std::vector<int64_t>(shape)) {
    fail(INNFER_E_INVALID, "size mismatch for " + key);
    return nullptr;
  }
  return &it->second;
}

// PAN keys (PAN_arch.py:107-169; upsampler indices through block.sequential, block.py:197-210)
int finalize_pan(innfer_rrdb* h) {
  const auto& c = h->cfg;
  const int nf = c.nf, gw = nf / 2, G = (gw + 15) / 16 * 16, unf = h->pan_unf;
  int rc;
  size_t expected = 0;
  if ((rc = build_layer(h, h->fea, "conv_first", nf, c.in_nc, 1))) return rc;
  expected += 2;
  const int ntrunk = h->pan_double ? 2 : 1;
  h->prb.resize((size_t)ntrunk * c.nb * 5);
  h->ptail[0].resize(ntrunk);
  for (int t = 0; t < ntrunk; ++t) {
    const std::string tn = t ? "SCPA_trunk2." : "SCPA_trunk.";
    for (int b = 0; b < c.nb; ++b) {
      const std::string pre = tn + std::to_string(b) + ".";
      ConvLayer* L = &h->prb[((size_t)t * c.nb + b) * 5];
      const Param* wa = need_param(h, pre + "conv1_a.weight", {gw, nf, 1, 1});
      const Param* wb = need_param(h, pre + "conv1_b.weight", {gw, nf, 1, 1});
      const Param* w3 = need_param(h, pre + "conv3.weight", {nf, 2 * gw, 1, 1});
      if (!wa || !wb || !w3) return g_err_code();
      // both 1x1 input convs as one: rows [0, gw) = conv1_a, [G, G + gw) = conv1_b, zero rows in between
      std::vector<float> wab((size_t)2 * G * nf, 0.f);
      for (int r = 0; r < gw; ++r)
        for (int k = 0; k < nf; ++k) {
          wab[(size_t)r * nf + k] = wa->data[(size_t)r * nf + k];
          wab[(size_t)(G + r) * nf + k] = wb->data[(size_t)r * nf + k];
        }
      if ((rc = build_custom(h, L[0], pre + "conv1_a|b", wab, nullptr, 2 * G, nf, 1))) return rc;
      if ((rc = build_layer(h, L[1], pre + "k1.0", gw, gw, 1, 3, false))) return rc;
      // PACnv's k3 (3x3) and k2 (1x1, bias) read the same tensor: one conv with 2 * G outputs, k2 as the centre tap of
      // rows [G, G + gw); the epilogue multiplies column c by sigmoid(column G + c) (Epilogue.self_gate)
      const Param* w_k3 = need_param(h, pre + "PACnv.k3.weight", {gw, gw, 3, 3});
      const Param* w_k2 = need_param(h, pre + "PACnv.k2.weight", {gw, gw, 1, 1});
      const Param* b_k2 = need_param(h, pre + "PACnv.k2.bias", {gw});
      if (!w_k3 || !w_k2 || !b_k2) return g_err_code();
      std::vector<float> wpa((size_t)2 * G * gw * 9, 0.f), bpa((size_t)2 * G, 0.f);
      for (int r = 0; r < gw; ++r) {
        for (int k = 0; k < gw; ++k) {
          for (int t = 0; t < 9; ++t) wpa[((size_t)r * gw + k) * 9 + t] = w_k3->data[((size_t)r * gw + k) * 9 + t];
          wpa[((size_t)(G + r) * gw + k) * 9 + 4] = w_k2->data[(size_t)r * gw + k];
        }
        bpa[G + r] = b_k2->data[r];
      }
      if ((rc = build_custom(h, L[2], pre + "PACnv.k3|k2", wpa, bpa.data(), 2 * G, gw, 3))) return rc;
      if ((rc = build_layer(h, L[3], pre + "PACnv.k4", gw, gw, 1, 3, false))) return rc;
      // conv3 reads the padded concat: input channel k of branch a sits at k, of branch b at G + k
      std::vector<float> wc((size_t)nf * 2 * G, 0.f);
      for (int r = 0; r < nf; ++r)
        for (int k = 0; k < gw; ++k) {
          wc[(size_t)r * 2 * G + k] = w3->data[(size_t)r * 2 * gw + k];
          wc[(size_t)r * 2 * G + G + k] = w3->data[(size_t)r * 2 * gw + gw + k];
        }
      if ((rc = build_custom(h, L[4], pre + "conv3", wc, nullptr, nf, 2 * G, 1))) return rc;
      expected += 8;
    }
    if ((rc = build_layer(h, h->ptail[0][t], t ? "trunk_conv2" : "trunk_conv", nf, nf, 1))) return rc;
    expected += 2;
  }
  if (h->pan_sa) {
    const int cq = nf / 8;
    const Param* ga = need_param(h, "FSA.gamma", {1});
    const Param* wf = need_param(h, "FSA.conv_f.weight", {cq, nf, 1});
    const Param* bf = need_param(h, "FSA.conv_f.bias", {cq});
    const Param* wg = need_param(h, "FSA.conv_g.weight", {cq, nf, 1});
    const Param* bg = need_param(h, "FSA.conv_g.bias", {cq});
    const Param* wh = need_param(h, "FSA.conv_h.weight", {nf, nf, 1});
    const Param* bh = need_param(h, "FSA.conv_h.bias", {nf});
    if (!ga || !wf || !bf || !wg || !bg || !wh || !bh) return g_err_code();
    expected += 7;
    h->pan_gamma = ga->data[0];
    constexpr int R = 2 * kPanQK + kPanRow;
    std::vector<float> wcat((size_t)R * kPanRow, 0.f), bcat(R, 0.f);
    for (int r = 0; r < cq; ++r) {
      for (int k = 0; k < nf; ++k) {
        wcat[(size_t)k * R + r] = wf->data[(size_t)r * nf + k];   // transposed: [input channel][row]
        wcat[(size_t)k * R + kPanQK + r] = wg->data[(size_t)r * nf + k];
      }
      bcat[r] = bf->data[r];
      bcat[kPanQK + r] = bg->data[r];
    }
    for (int r = 0; r < nf; ++r) {
      for (int k = 0; k < nf; ++k) wcat[(size_t)k * R + 2 * kPanQK + r] = wh->data[(size_t)r * nf + k];
      bcat[2 * kPanQK + r] = bh->data[r];
    }
    CU_TRY(cudaMalloc(&h->d_pan_w, wcat.size() * sizeof(float)));
    CU_TRY(cudaMalloc(&h->d_pan_b, bcat.size() * sizeof(float)));
    CU_TRY(cudaMemcpy(h->d_pan_w, wcat.data(), wcat.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_pan_b, bcat.data(), bcat.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  // one stage: its Sequential is used as is and keeps both activations; several: flattened, five entries per stage
  h->pan_hr_act = h->n_up == 1;
  h->ups.resize((size_t)h->n_up * 3);
  for (int i = 0; i < h->n_up; ++i) {
    const int base = h->n_up == 1 ? 0 : 5 * i;
    ConvLayer* L = &h->ups[(size_t)i * 3];
    if ((rc = build_layer(h, L[0], "upsample." + std::to_string(base + 1), unf, i ? unf : nf, h->up_factor))) return rc;
    if ((rc = build_layer(h, L[1], "upsample." + std::to_string(base + 2) + ".conv", unf, unf, 1, 1))) return rc;
    if ((rc = build_layer(h, L[2], "upsample." + std::to_string(base + 4), unf, unf, 1))) return rc;
    expected += 6;
  }
  if ((rc = build_layer(h, h->hr1, "conv_last", c.out_nc, h->n_up ? unf : nf, 1))) return rc;
  expected += 2;
  if (h->params.size() != expected)
    return fail(INNFER_E_INVALID, "unexpected keys in state dict (" + std::to_string(h->params.size()) + " loaded, " +
                                      std::to_string(expected) + " expected)");
  h->params.clear();
  h->finalized = true;
  return 0;
}

// run one conv in the handle's precision mode
int run_conv(innfer_rrdb* h, const ConvLayer& L, ChunkView in, int B, int H, int W, ChunkView out,
             int out_nchunks, const Epilogue& ep, cudaStream_t st) {
  int rc;
  cudaEvent_t ea = nullptr, eb = nullptr;
  if (h->profiling == 2) {
    // per-launch events: they also defeat the programmatic-dependent-launch overlap of consecutive convs, so the
    // times of this mode are those of isolated launches (the per-batch mode 1 times the real schedule)
    if (cudaEventCreate(&ea) != cudaSuccess || cudaEventCreate(&eb) != cudaSuccess) return fail(INNFER_E_CUDA, "cudaEventCreate");
    cudaEventRecord(ea, st);
  }
  if (h->cfg.fp16) {
    rc = conv_layer_run(L, h->cache, in, B, H, W, out, out_nchunks, ep, h->num_sms, st);
  } else {
    g_last_conv_kernel = "conv_direct";
    rc = conv_direct_run(L, in, B, H, W, out, out_nchunks, ep, st);
  }
  if (h->profiling == 2) {
    cudaEventRecord(eb, st);
    char name[96];
    snprintf(name, sizeof name, "%s %d->%d%s%s%s", g_last_conv_kernel, L.Cin, L.Cout, L.up > 1 ? (L.up == 2 ? " up2" : " up3") : "",
             (ep.res1.base || ep.res2.base) ? " +res" : "", ep.compact4 ? " compact" : "");
    auto& f = h->fam[name];
    f.launches += 1;
    // the reference conv: k*k*Cin*Cout MACs per output pixel (after the nearest upsample / pixel shuffle), real pixels only
    const double opx = (double)B * H * W * L.up * L.up;
    const int taps = (L.max_taps == 1 && L.up == 1) ? 1 : 9;
    f.flop += 2.0 * taps * L.Cin * (double)L.Cout * opx;
    const double e = (double)h->esz();
    f.bytes += ((double)L.Cin * B * H * W + (double)L.Cout * opx * (1 + (ep.res1.base ? 1 : 0) + (ep.res2.base ? 1 : 0))) * e;
    f.ev.emplace_back(ea, eb);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc != 0) {
    char msg[160];
    snprintf(msg, sizeof msg, "conv launch failed (rc=%d, Cin=%d Cout=%d up=%d H=%d W=%d B=%d): %s", rc,
             L.Cin, L.Cout, L.up, H, W, B, rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "launcher error");
    return fail(INNFER_E_CUDA, msg);
  }
  return 0;
}

// Wide layout geometry (layers.cuh: ChunkView): images `wid + sep` columns apart, rows padded to 16 columns.  The
// separator must be as wide as the horizontal reach of the widest conv of the net: 1 column for the plain 3x3 convs,
// 8 for PPON's dilated convs (a single zero column let the taps at distance 2..8 read the neighbouring tile).
// Set per call from the handle (ensure_workspace / forward_tiles); the separator columns are kept zero by every
// kernel that writes a wide tensor.
thread_local int g_wide_sep = 1;
int wide_sep_of(const innfer_rrdb* h) { return h->arch == 2 ? 8 : 1; }
int wide_pitch(int wid) { return wid + g_wide_sep; }
int wide_cols(int B, int wid) { return (B * wide_pitch(wid) + 15) / 16 * 16; }

ChunkView wview(DevBuf& b, int CT, int chunk0, int B, int wid, int up) {
  ChunkView v;
  v.base = reinterpret_cast<__half*>(b.p);
  v.CT = CT;
  v.chunk0 = chunk0;
  v.pitch = wide_pitch(wid) * up;
  v.Wtot = wide_cols(B, wid) * up;
  return v;
}

// INNFER_WIDE=0 keeps the fp16 path on the tiled layout (9-tap kernel only), for A/B measurements
bool wide_enabled() {
  static const int v = getenv("INNFER_WIDE") ? atoi(getenv("INNFER_WIDE")) : 1;
  return v != 0;
}

ChunkView view(DevBuf& b, int CT, int chunk0) {
  ChunkView v;
  v.base = reinterpret_cast<__half*>(b.p);
  v.CT = CT;
  v.chunk0 = chunk0;
  return v;
}

int ensure_workspace(innfer_rrdb* h, int B, int hgt, int wid) {
  g_wide_sep = wide_sep_of(h);
  // sized for the wide layout of the fp16 path (one separator column per image, row padded to 16)
  const size_t px = (size_t)hgt * wide_cols(B, wid);
  const size_t e8 = 8 * h->esz();
  const int s = h->cfg.scale;
  int rc = 0;
  rc |= h->in_tiles.ensure(px * h->in_ct() * e8);
  if (h->arch == 4) return rc ? fail(INNFER_E_NOMEM, "workspace allocation failed") : 0;   // I2INet sizes its own buffers
  if (h->arch == 3) {
    // PAN: trunk buffers (channels padded to 16), the two SCPA branch pairs, PACnv scratch, HR ping-pong, ILR, attention
    const int nfC = (h->cfg.nf + 15) / 16 * 2, gc = (h->cfg.nf / 2 + 15) / 16 * 2, ufC = (h->pan_unf + 15) / 16 * 2;
    for (int i = 0; i < 3; ++i) rc |= h->pbuf[i].ensure(px * nfC * e8);
    for (int i = 3; i < 5; ++i) rc |= h->pbuf[i].ensure(px * 2 * gc * e8);
    rc |= h->pbuf[5].ensure(px * gc * e8);
    const size_t hpx3 = px * s * s;
    if (s > 1)
      for (auto& b : h->hrbuf) rc |= b.ensure(hpx3 * ufC * e8);
    rc |= h->paux[0].ensure(hpx3 * e8);
    const size_t npool = (size_t)B * (hgt / 4) * (wid / 4);
    rc |= h->paux[1].ensure((npool + 1) * (3 * kPanRow + 2 * kPanQK) * sizeof(float));
    if (rc) {
      h->in_tiles.release();
      for (auto& b : h->hrbuf) b.release();
      for (auto& b : h->pbuf) b.release();
      for (auto& b : h->paux) b.release();
      return fail(INNFER_E_NOMEM, "workspace allocation failed");
    }
    return 0;
  }
  rc |= h->feat.ensure(px * h->nf_ct() * e8);
  if (h->arch != 2)
    for (auto& b : h->xbuf) rc |= b.ensure(px * h->cat_ct() * e8);
  const size_t hpx = px * s * s;
  // HR ping-pong buffers: the last upconv output and HR_conv0 output are both full resolution
  for (auto& b : h->hrbuf) rc |= b.ensure(hpx * h->nf_ct() * e8);
  if (h->arch == 2) {
    for (int i = 0; i < 8; ++i) rc |= h->pbuf[i].ensure(px * h->nf_ct() * e8);
    rc |= h->pbuf[8].ensure(px * 32 * e8);
    for (auto& b : h->paux) rc |= b.ensure(hpx * e8);
  }
  if (rc) {
    // drop partial allocations so that a retry with a smaller batch starts clean
    h->in_tiles.release();
    h->feat.release();
    for (auto& b : h->xbuf) b.release();
    for (auto& b : h->hrbuf) b.release();
    for (auto& b : h->pbuf) b.release();
    for (auto& b : h->paux) b.release();
    return fail(INNFER_E_NOMEM, "workspace allocation failed");
  }
  return 0;
}

// Forward B tiles that already sit in h->in_tiles ([B][in_ct][hgt][wid][8]); the result
// ([B][1 or 2][s*hgt][s*wid][8], out_nc channels in chunk 0..) is written to `dst`.
int forward_tiles_impl(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st);

int forward_tiles(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st) {
  g_wide_sep = wide_sep_of(h);
  if (h->profiling != 1) return forward_tiles_impl(h, B, hgt, wid, dst, compact, st);
  cudaEvent_t a, b;
  CU_TRY(cudaEventCreate(&a));
  CU_TRY(cudaEventCreate(&b));
  CU_TRY(cudaEventRecord(a, st));
  const uint64_t l0 = g_launches.load();
  int rc = forward_tiles_impl(h, B, hgt, wid, dst, compact, st);
  CU_TRY(cudaEventRecord(b, st));
  h->prof_events.emplace_back(a, b);
  h->prof_conv_launches += g_launches.load() - l0;
  return rc;
}

// SRResNet.forward (SRResNet_arch.py:15-62): fea_conv, ShortcutBlock(ResNetBlock x nb, LR_conv),
// pixel-shuffle (or upconv) blocks + ReLU, HR_conv0 + ReLU, HR_conv1.  ResNetBlock (CNA, no norm):
// x + res_scale * conv1(relu(conv0(x))) -- the scaled add is conv1's residual epilogue.
int forward_tiles_srresnet(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st) {
  const int nfc = h->nf_ct();
  int rc;
  Epilogue plain, relu;
  relu.lrelu = true;
  relu.slope = 0.f;
  const bool wide = h->cfg.fp16 && wide_enabled();
  int lvl = 1;
  auto view = [&](DevBuf& b, int CT, int chunk0) {
    return wide ? wview(b, CT, chunk0, B, wid, lvl) : ::view(b, CT, chunk0);
  };
  if (wide) CU_TRY(cudaMemsetAsync(h->feat.p, 0, (size_t)nfc * hgt * wide_cols(B, wid) * 8 * h->esz(), st));
  if ((rc = run_conv(h, h->fea, ::view(h->in_tiles, h->in_ct(), 0), B, hgt, wid, view(h->feat, nfc, 0), nfc, plain, st))) return rc;
  int X = 0, T = 1, Y = 2;
  ChunkView cur = view(h->feat, nfc, 0);
  for (int b = 0; b < h->cfg.nb; ++b) {
    const ConvLayer* L = &h->rdb[(size_t)b * 2];
    if ((rc = run_conv(h, L[0], cur, B, hgt, wid, view(h->xbuf[T], nfc, 0), nfc, relu, st))) return rc;
    // INNFER_SRRES_SPLIT=1 (experiment): the second conv without its residual epilogue + the scaled add as one elementwise
    // pass (one more fp16 rounding per block).  Measured per 1080p frame: 9-tap kernel with residual epilogue 53.3 ms,
    // this split 51.0 ms, row kernel WITH the residual epilogue (INNFER_ROWS64_RES, the default) 49.4 ms.
    static const int split = getenv("INNFER_SRRES_SPLIT") ? atoi(getenv("INNFER_SRRES_SPLIT")) : 0;
    if (split && wide) {
      if ((rc = run_conv(h, L[1], view(h->xbuf[T], nfc, 0), B, hgt, wid, view(h->xbuf[Y], nfc, 0), nfc, plain, st))) return rc;
      const size_t n16 = (size_t)nfc * hgt * wide_cols(B, wid);
      if (launch_axpy_f16(reinterpret_cast<__half*>(h->xbuf[Y].p), cur.base, reinterpret_cast<const __half*>(h->xbuf[Y].p),
                          h->res_scale, n16, st))
        return fail(INNFER_E_CUDA, "axpy launch failed");
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
      Epilogue e;
      e.res1 = cur;
      e.alpha1 = h->res_scale;
      if ((rc = run_conv(h, L[1], view(h->xbuf[T], nfc, 0), B, hgt, wid, view(h->xbuf[Y], nfc, 0), nfc, e, st))) return rc;
    }
    cur = view(h->xbuf[Y], nfc, 0);
    const int t = X;
    X = Y;
    Y = t;
  }
  {
    Epilogue e;
    e.res1 = view(h->feat, nfc, 0);
    e.alpha1 = 1.0f;
    if ((rc = run_conv(h, h->lr_conv, cur, B, hgt, wid, view(h->xbuf[T], nfc, 0), nfc, e, st))) return rc;
    cur = view(h->xbuf[T], nfc, 0);
  }
  int ch = hgt, cw = wid, pp = 0;
  for (size_t i = 0; i < h->ups.size(); ++i) {
    lvl *= h->ups[i].up;
    ChunkView o = view(h->hrbuf[pp], nfc, 0);
    if ((rc = run_conv(h, h->ups[i], cur, B, ch, cw, o, nfc, relu, st))) return rc;
    ch *= h->ups[i].up;
    cw *= h->ups[i].up;
    cur = o;
    pp ^= 1;
  }
  ChunkView o = view(h->hrbuf[pp], nfc, 0);
  if ((rc = run_conv(h, h->hr0, cur, B, ch, cw, o, nfc, relu, st))) return rc;
  Epilogue last;
  last.compact4 = compact;
  return run_conv(h, h->hr1, o, B, ch, cw, dst, (h->cfg.out_nc + 7) / 8, last, st);
}

// PPON.forward (PPON_arch.py:64-76), third output only (out_p; run.py:191-192).  _ResBlock_32 (78-114):
//   o1 = lrelu(c1(x)); d_k = conv(o1, dilation k); add_k = d_1 + .. + d_{k+1};
//   out = x + 0.2 * c2(lrelu(cat[d_1, add_1..add_7]))              (c2 is a 1x1 conv over 256 channels)
// The running sums are the residual epilogue of the dilated convs: conv d_k stores add_{k-1} + d_k twice, raw into a
// 32-channel scratch (the next conv's residual input) and activated into its slice of the concat buffer.
// RRBlock_32 (116-127) = three blocks and "*0.2 + x", folded into the third c2's second residual.
int forward_tiles_ppon(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st) {
  const int nfc = h->nf_ct();
  int rc;
  const bool wide = h->cfg.fp16 && wide_enabled();
  int lvl = 1;
  auto view = [&](DevBuf& b, int CT, int chunk0) {
    return wide ? wview(b, CT, chunk0, B, wid, lvl) : ::view(b, CT, chunk0);
  };
  DevBuf &F = h->pbuf[0], &T1 = h->pbuf[4], &S = h->pbuf[5], &E1 = h->pbuf[6], &E2 = h->pbuf[7], &CAT = h->pbuf[8];
  Epilogue plain, act;
  act.lrelu = true;
  if (wide) CU_TRY(cudaMemsetAsync(F.p, 0, (size_t)nfc * hgt * wide_cols(B, wid) * 8 * h->esz(), st));
  if ((rc = run_conv(h, h->fea, ::view(h->in_tiles, h->in_ct(), 0), B, hgt, wid, view(F, nfc, 0), nfc, plain, st))) return rc;

  // one _ResBlock_32: in -> out (+ the RRBlock-level residual for the third block of a RRBlock)
  auto res_block = [&](const ConvLayer* L, DevBuf& in, DevBuf& out, DevBuf* rrb_in) -> int {
    int r;
    if ((r = run_conv(h, L[0], view(in, nfc, 0), B, hgt, wid, view(T1, nfc, 0), nfc, act, st))) return r;
    for (int k = 0; k < 8; ++k) {
      Epilogue e;
      e.lrelu = true;
      e.act_after_res = true;
      // fp16: the running sum add_{k-1} is read back from its ACTIVATED copy in the concat buffer (LeakyReLU is
      // invertible: v > 0 ? v : v / 0.2), which saves the raw second store -- these convs are bandwidth-bound.
      // INNFER_PPON_RAW=1 (and the fp32 mode) keep the raw scratch copy.
      static const bool keep_raw = getenv("INNFER_PPON_RAW") && atoi(getenv("INNFER_PPON_RAW"));
      const bool unact = h->cfg.fp16 && !keep_raw;
      if (k > 0) {
        if (unact) {
          e.res1 = view(CAT, 32, 4 * (k - 1));
          e.res1_unact = true;
        } else {
          e.res1 = view(S, nfc, 4 * ((k - 1) & 1));
        }
        e.alpha1 = 1.0f;
      }
      if (k < 7 && !unact) e.raw_out = view(S, nfc, 4 * (k & 1));
      if ((r = run_conv(h, L[1 + k], view(T1, nfc, 0), B, hgt, wid, view(CAT, 32, 4 * k), 4, e, st))) return r;
    }
    Epilogue e2;
    e2.res1 = view(in, nfc, 0);
    e2.alpha1 = 0.2f;
    if (rrb_in) {
      e2.res2 = view(*rrb_in, nfc, 0);
      e2.alpha2 = 0.2f;
    }
    return run_conv(h, L[9], view(CAT, 32, 0), B, hgt, wid, view(out, nfc, 0), nfc, e2, st);
  };
  // a chain of RRBlock_32 starting from `src` (not modified); returns the buffer holding the result
  auto rr_chain = [&](const ConvLayer* L, int nblocks, DevBuf* src, DevBuf** result) -> int {
    DevBuf* X[3] = {&h->pbuf[1], &h->pbuf[2], &h->pbuf[3]};
    DevBuf* cur = src;
    for (int b = 0; b < nblocks; ++b) {
      // free rotating buffers: the two that are not `cur`
      DevBuf* f[2];
      int n = 0;
      for (auto* x : X)
        if (x != cur && n < 2) f[n++] = x;
      int r;
      const ConvLayer* Lb = L + (size_t)b * 30;
      if ((r = res_block(Lb, *cur, *f[0], nullptr))) return r;
      if ((r = res_block(Lb + 10, *f[0], *f[1], nullptr))) return r;
      if ((r = res_block(Lb + 20, *f[1], *f[0], cur))) return r;
      cur = f[0];
    }
    *result = cur;
    return 0;
  };
  // reconstruction tail: ups..., HR_conv0, HR_conv1 [* alpha + res]
  auto tail = [&](const std::vector<ConvLayer>& T, DevBuf& src, const ChunkView* res, float alpha, ChunkView out,
                  bool out_compact) -> int {
    int r;
    lvl = 1;
    ChunkView cur = view(src, nfc, 0);
    int ch = hgt, cw = wid, pp = 0;
    const size_t nups = T.size() - 2;
    for (size_t i = 0; i < nups; ++i) {
      lvl *= T[i].up;
      ChunkView o = view(h->hrbuf[pp], nfc, 0);
      if ((r = run_conv(h, T[i], cur, B, ch, cw, o, nfc, act, st))) return r;
      ch *= T[i].up;
      cw *= T[i].up;
      cur = o;
      pp ^= 1;
    }
    ChunkView o = view(h->hrbuf[pp], nfc, 0);
    if ((r = run_conv(h, T[nups], cur, B, ch, cw, o, nfc, act, st))) return r;
    Epilogue last;
    last.compact4 = out_compact;
    if (res) {
      last.res1 = *res;
      last.alpha1 = alpha;
    }
    r = run_conv(h, T[nups + 1], o, B, ch, cw, out, (h->cfg.out_nc + 7) / 8, last, st);
    return r;
  };

  const ConvLayer* L = h->prb.data();
  DevBuf* res = nullptr;
  if ((rc = rr_chain(L, h->cfg.nb, &F, &res))) return rc;
  {  // LR_conv + shortcut -> out_CFEM
    Epilogue e;
    e.res1 = view(F, nfc, 0);
    e.alpha1 = 1.0f;
    if ((rc = run_conv(h, h->lr_conv, view(*res, nfc, 0), B, hgt, wid, view(E1, nfc, 0), nfc, e, st))) return rc;
  }
  const int s = h->cfg.scale;
  auto hr_view = [&](DevBuf& b) {  // one-chunk tensor at output resolution
    lvl = s;
    ChunkView v = view(b, 1, 0);
    lvl = 1;
    return v;
  };
  const ChunkView out_c = hr_view(h->paux[0]), out_s = hr_view(h->paux[1]);
  if ((rc = tail(h->ptail[0], E1, nullptr, 1.f, out_c, false))) return rc;               // out_c = CRM(out_CFEM)
  lvl = 1;
  if ((rc = rr_chain(L + (size_t)h->cfg.nb * 30, 2, &E1, &res))) return rc;             // SFEM
  if (res != &E2) {
    CU_TRY(cudaMemcpyAsync(E2.p, res->p, (size_t)nfc * hgt * (wide ? (size_t)wide_cols(B, wid) : (size_t)B * wid) * 8 * h->esz(),
                           cudaMemcpyDeviceToDevice, st));
  }
  if ((rc = tail(h->ptail[1], E2, &out_c, 1.f, out_s, false))) return rc;               // out_s = SRM(out_SFEM) + out_c
  lvl = 1;
  if ((rc = rr_chain(L + (size_t)(h->cfg.nb + 2) * 30, 2, &E2, &res))) return rc;       // PFEM
  return tail(h->ptail[2], *res, &out_s, h->ppon_alpha, dst, compact);                  // out_p = alpha * PRM(..) + out_s
}

// PAN.forward (PAN_arch.py:171-222) on the tiled layout.  SCPA (86-103) as five convs:
//   AB = lrelu(conv1_a | conv1_b (x))            one 1x1 conv, each branch padded to a multiple of 16 channels
//   CD[a] = lrelu(k1(AB[a]));  G = k3(AB[b]) * sigmoid(k2(AB[b]))   (k3 | k2 as one conv, self-gate epilogue);
//   CD[b] = lrelu(k4(G))
//   x' = conv3(CD) + x                           1x1 conv over the padded concat (weights re-indexed at load time)
// then trunk_conv + fea, the max-pooled self-attention block (pan_ops.cu), the pixel-attention upsampling stages
// (upconv with the nearest upsample folded in, u * sigmoid(conv1x1(u)) -> lrelu, HRconv) and conv_last + bilinear ILR.
int forward_tiles_pan(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st) {
  const auto& c = h->cfg;
  const int nfc = c.nf / 8, nfC = (c.nf + 15) / 16 * 2, gc = (c.nf / 2 + 15) / 16 * 2;
  const int ufc = (h->pan_unf + 7) / 8, ufC = (h->pan_unf + 15) / 16 * 2;
  const size_t e8 = 8 * h->esz();
  const bool f16 = c.fp16 != 0;
  int rc;
  DevBuf *X0 = &h->pbuf[0], *X1 = &h->pbuf[1], *X2 = &h->pbuf[2];
  DevBuf &AB = h->pbuf[3], &CD = h->pbuf[4], &G = h->pbuf[5];
  // chunks [first, CT) of every image are read (against zero weights) but never written: keep them finite
  auto zero_pad = [&](DevBuf& b, int CT, int first, size_t plane_px) -> int {
    if (first >= CT) return 0;
    CU_TRY(cudaMemset2DAsync(reinterpret_cast<uint8_t*>(b.p) + (size_t)first * plane_px * e8, (size_t)CT * plane_px * e8, 0,
                             (size_t)(CT - first) * plane_px * e8, (size_t)B, st));
    return 0;
  };
  const size_t lr_px = (size_t)hgt * wid;
  for (DevBuf* x : {X0, X1, X2})
    if ((rc = zero_pad(*x, nfC, nfc, lr_px))) return rc;
  Epilogue plain, act;
  act.lrelu = true;
  if ((rc = run_conv(h, h->fea, view(h->in_tiles, h->in_ct(), 0), B, hgt, wid, view(*X0, nfC, 0), nfc, plain, st))) return rc;
  DevBuf* cur = X0;
  const int ntrunk = h->pan_double ? 2 : 1;
  for (int t = 0; t < ntrunk; ++t) {
    for (int b = 0; b < c.nb; ++b) {
      const ConvLayer* L = &h->prb[((size_t)t * c.nb + b) * 5];
      DevBuf* nxt = (cur == X1) ? X2 : X1;
      if ((rc = run_conv(h, L[0], view(*cur, nfC, 0), B, hgt, wid, view(AB, 2 * gc, 0), 2 * gc, act, st))) return rc;
      if ((rc = run_conv(h, L[1], view(AB, 2 * gc, 0), B, hgt, wid, view(CD, 2 * gc, 0), gc, act, st))) return rc;
      Epilogue gate;
      gate.self_gate = true;
      if ((rc = run_conv(h, L[2], view(AB, 2 * gc, gc), B, hgt, wid, view(G, gc, 0), gc, gate, st))) return rc;
      if ((rc = run_conv(h, L[3], view(G, gc, 0), B, hgt, wid, view(CD, 2 * gc, gc), gc, act, st))) return rc;
      Epilogue res;
      res.res1 = view(*cur, nfC, 0);
      if ((rc = run_conv(h, L[4], view(CD, 2 * gc, 0), B, hgt, wid, view(*nxt, nfC, 0), nfc, res, st))) return rc;
      cur = nxt;
    }
    DevBuf* nxt = (cur == X1) ? X2 : X1;
    Epilogue e;
    if (t == ntrunk - 1) e.res1 = view(*X0, nfC, 0);   // fea + trunk (PAN_arch.py:180-185)
    if ((rc = run_conv(h, h->ptail[0][t], view(*cur, nfC, 0), B, hgt, wid, view(*nxt, nfC, 0), nfc, e, st))) return rc;
    cur = nxt;
  }
  if (h->pan_sa) {
    const int hp = hgt / 4, wp = wid / 4;
    if (hp < 1 || wp < 1) return fail(INNFER_E_UNSUPPORTED, "PAN self-attention needs images of at least 4x4 pixels");
    const size_t n = (size_t)B * hp * wp;
    float* pooled = reinterpret_cast<float*>(h->paux[1].p);
    float* hv = pooled + n * kPanRow;
    float* att = hv + n * kPanRow;
    float* fq = att + n * kPanRow;
    float* gk = fq + n * kPanQK;
    DevBuf* nxt = (cur == X1) ? X2 : X1;
    int r;
    if (f16) r = launch_pan_maxpool(reinterpret_cast<const __half*>(cur->p), nfC, nfc, B, hgt, wid, 4, pooled, st);
    else r = launch_pan_maxpool(reinterpret_cast<const float*>(cur->p), nfC, nfc, B, hgt, wid, 4, pooled, st);
    if (!r) r = launch_pan_proj(pooled, (long long)n, nfc * 8, h->d_pan_w, h->d_pan_b, fq, gk, hv, st);
    // fp16 mode: the contraction on the tensor cores (INNFER_PAN_ATT_TC=0: the fp32 CUDA-core kernel, as in fp32 mode)
    static const int att_tc = getenv("INNFER_PAN_ATT_TC") ? atoi(getenv("INNFER_PAN_ATT_TC")) : 1;
    if (!r) r = (f16 && att_tc) ? launch_pan_attention_tc(fq, gk, hv, B, hp * wp, nfc * 8, att, st)
                                : launch_pan_attention(fq, gk, hv, B, hp * wp, nfc * 8, att, st);
    if (!r) {
      if (f16)
        r = launch_pan_bicubic_add(att, hp, wp, reinterpret_cast<const __half*>(cur->p), nfC,
                                   reinterpret_cast<__half*>(nxt->p), nfC, nfc, B, hgt, wid, h->pan_gamma, st);
      else
        r = launch_pan_bicubic_add(att, hp, wp, reinterpret_cast<const float*>(cur->p), nfC,
                                   reinterpret_cast<float*>(nxt->p), nfC, nfc, B, hgt, wid, h->pan_gamma, st);
    }
    g_launches.fetch_add(4, std::memory_order_relaxed);
    if (r) return fail(INNFER_E_CUDA, "PAN self-attention launch failed");
    cur = nxt;
  }
  ChunkView curv = view(*cur, nfC, 0);
  int ch = hgt, cw = wid;
  DevBuf *P = &h->hrbuf[0], *Q = &h->hrbuf[1];
  for (int i = 0; i < h->n_up; ++i) {
    const ConvLayer* L = &h->ups[(size_t)i * 3];
    const int f = L[0].up;
    const size_t px = (size_t)(ch * f) * (cw * f);
    if ((rc = zero_pad(*P, ufC, ufc, px))) return rc;
    if ((rc = run_conv(h, L[0], curv, B, ch, cw, view(*P, ufC, 0), ufc, plain, st))) return rc;
    ch *= f;
    cw *= f;
    if ((rc = zero_pad(*Q, ufC, ufc, px))) return rc;   // Q held this stage's input until the upconv above
    Epilogue pa;                                        // lrelu(u * sigmoid(conv1x1(u)))
    pa.gate = true;
    pa.res1 = view(*P, ufC, 0);
    pa.lrelu = true;
    pa.act_after_res = true;
    if ((rc = run_conv(h, L[1], view(*P, ufC, 0), B, ch, cw, view(*Q, ufC, 0), ufc, pa, st))) return rc;
    if ((rc = run_conv(h, L[2], view(*Q, ufC, 0), B, ch, cw, view(*P, ufC, 0), ufc, h->pan_hr_act ? act : plain, st))) return rc;
    curv = view(*P, ufC, 0);
    DevBuf* tmp = P;
    P = Q;
    Q = tmp;
  }
  {
    int r;
    if (f16) r = launch_pan_bilinear(reinterpret_cast<const __half*>(h->in_tiles.p), h->in_ct(), B, hgt, wid, c.scale,
                                     reinterpret_cast<__half*>(h->paux[0].p), st);
    else r = launch_pan_bilinear(reinterpret_cast<const float*>(h->in_tiles.p), h->in_ct(), B, hgt, wid, c.scale,
                                 reinterpret_cast<float*>(h->paux[0].p), st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (r) return fail(INNFER_E_CUDA, "PAN bilinear launch failed");
  }
  Epilogue last;
  last.compact4 = compact;
  last.res1 = view(h->paux[0], 1, 0);
  return run_conv(h, h->hr1, curv, B, ch, cw, dst, (c.out_nc + 7) / 8, last, st);
}

int i2i_code(int rc) {   // I2INet's codes (i2i.cuh) -> INNFER_E_*
  switch (rc) {
    case -1: return INNFER_E_INVALID;
    case -2: return INNFER_E_UNSUPPORTED;
    case -3: return INNFER_E_CUDA;
    case -4: return INNFER_E_STATE;
    default: return INNFER_E_NOMEM;
  }
}

int forward_tiles_i2i(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st) {
  std::string err;
  const uint64_t l0 = h->i2i->launches();
  const int rc = h->i2i->forward(h->in_tiles.p, h->in_ct(), B, hgt, wid, GenView{dst.base, dst.CT, dst.chunk0}, compact,
                                 h->i2i_per_sample, st, err);
  g_launches.fetch_add(h->i2i->launches() - l0, std::memory_order_relaxed);
  return rc ? fail(i2i_code(rc), err) : 0;
}

int forward_tiles_impl(innfer_rrdb* h, int B, int hgt, int wid, ChunkView dst, bool compact, cudaStream_t st) {
  if (h->arch == 4) return forward_tiles_i2i(h, B, hgt, wid, dst, compact, st);
  if (h->arch == 3) return forward_tiles_pan(h, B, hgt, wid, dst, compact, st);
  if (h->arch == 1) return forward_tiles_srresnet(h, B, hgt, wid, dst, compact, st);
  if (h->arch == 2) return forward_tiles_ppon(h, B, hgt, wid, dst, compact, st);
  const int nfc = h->nf_ct(), catc = h->cat_ct();
  const int gcc = 32 / 8;
  int rc;
  Epilogue plain;
  Epilogue act;
  act.lrelu = true;
  // fp16 mode runs on the wide layout: the batch is one image with the tiles side by side, which lets
  // the row-streaming kernel use full 128-pixel MMA tiles whatever the tile width is
  const bool wide = h->cfg.fp16 && wide_enabled();
  int lvl = 1;  // resolution of the buffer a view is made for (1 = LR, scale = HR)
  auto view = [&](DevBuf& b, int CT, int chunk0) {
    return wide ? wview(b, CT, chunk0, B, wid, lvl) : ::view(b, CT, chunk0);
  };
  // fea_conv -> feat (kept for the ShortcutBlock) and copy into xbuf[0][0:nfc]
  if (wide) {
    // fea_conv reads tiled input and skips the separator columns of its wide destination
    const size_t plane = (size_t)nfc * hgt * wide_cols(B, wid) * 8 * h->esz();
    CU_TRY(cudaMemsetAsync(h->feat.p, 0, plane, st));
    if ((rc = run_conv(h, h->fea, ::view(h->in_tiles, h->in_ct(), 0), B, hgt, wid, view(h->feat, nfc, 0), nfc, plain, st))) return rc;
    CU_TRY(cudaMemcpyAsync(h->xbuf[0].p, h->feat.p, plane, cudaMemcpyDeviceToDevice, st));
  } else {
    if ((rc = run_conv(h, h->fea, ::view(h->in_tiles, h->in_ct(), 0), B, hgt, wid, ::view(h->feat, nfc, 0), nfc, plain, st))) return rc;
    const size_t row = (size_t)nfc * hgt * wid * 8 * h->esz();
    CU_TRY(cudaMemcpy2DAsync(h->xbuf[0].p, (size_t)catc * hgt * wid * 8 * h->esz(), h->feat.p, row, row, B,
                             cudaMemcpyDeviceToDevice, st));
  }
  int P = 0, Q = 1, R = 2;
  for (int b = 0; b < h->cfg.nb; ++b) {
    const int order[3] = {P, Q, R};
    for (int r = 0; r < 3; ++r) {
      const int cur = order[r];
      const ConvLayer* L = &h->rdb[((size_t)b * 3 + r) * 5];
      if (h->cfg.plus) {
        // ESRGAN+ (RRDBNet_arch.py:155-160): t = conv1x1(x) parked behind x4
        if ((rc = run_conv(h, h->c1x1[(size_t)b * 3 + r], view(h->xbuf[cur], catc, 0), B, hgt, wid,
                           view(h->xbuf[cur], catc, nfc + gcc * 4), gcc, plain, st)))
          return rc;
      }
      for (int k = 0; k < 4; ++k) {
        Epilogue e = act;
        if (h->cfg.plus && k == 1) {         // x2 = lrelu(conv2(..)) + conv1x1(x)
          e.res1 = view(h->xbuf[cur], catc, nfc + gcc * 4);
          e.alpha1 = 1.0f;
        } else if (h->cfg.plus && k == 3) {  // x4 = lrelu(conv4(..)) + x2
          e.res1 = view(h->xbuf[cur], catc, nfc + gcc * 1);
          e.alpha1 = 1.0f;
        }
        if ((rc = run_conv(h, L[k], view(h->xbuf[cur], catc, 0), B, hgt, wid,
                           view(h->xbuf[cur], catc, nfc + gcc * k), gcc, e, st)))
          return rc;
      }
      Epilogue e5;
      e5.res1 = view(h->xbuf[cur], catc, 0);
      e5.alpha1 = 0.2f;
      int dstbuf;
      if (r < 2) {
        dstbuf = order[r + 1];
      } else {
        dstbuf = Q;  // RRDB output: (rdb3*0.2 + x_rdb3)*0.2 + x_rrdb  (RRDBNet_arch.py:98,165)
        e5.res2 = view(h->xbuf[P], catc, 0);
        e5.alpha2 = 0.2f;
      }
      if ((rc = run_conv(h, L[4], view(h->xbuf[cur], catc, 0), B, hgt, wid, view(h->xbuf[dstbuf], catc, 0), nfc, e5, st)))
        return rc;
    }
    const int nP = Q, nQ = R, nR = P;
    P = nP;
    Q = nQ;
    R = nR;
  }
  // LR_conv + ShortcutBlock: fea + LR_conv(trunk)  (block.py:189-191)
  {
    Epilogue e;
    e.res1 = view(h->feat, nfc, 0);
    e.alpha1 = 1.0f;
    if ((rc = run_conv(h, h->lr_conv, view(h->xbuf[P], catc, 0), B, hgt, wid, view(h->xbuf[Q], catc, 0), nfc, e, st)))
      return rc;
  }
  ChunkView cur = view(h->xbuf[Q], catc, 0);
  int ch = hgt, cw = wid, pp = 0;
  for (size_t i = 0; i < h->ups.size(); ++i) {
    lvl *= h->ups[i].up;
    ChunkView o = view(h->hrbuf[pp], nfc, 0);
    if ((rc = run_conv(h, h->ups[i], cur, B, ch, cw, o, nfc, act, st))) return rc;
    ch *= h->ups[i].up;
    cw *= h->ups[i].up;
    cur = o;
    pp ^= 1;
  }
  ChunkView o = view(h->hrbuf[pp], nfc, 0);
  if ((rc = run_conv(h, h->hr0, cur, B, ch, cw, o, nfc, act, st))) return rc;
  const int out_chunks = (h->cfg.out_nc + 7) / 8;
  Epilogue last;
  last.compact4 = compact;
  if ((rc = run_conv(h, h->hr1, o, B, ch, cw, dst, out_chunks, last, st))) return rc;
  return 0;
}

// Split ntiles into the fewest batches <= max_batch, evenly sized (the persistent kernels split a batch
// into equal per-SM ranges, so only the number of launches matters).
int pick_batch(const innfer_rrdb* h, int ntiles, int p) {
  (void)p;
  if (ntiles <= 1) return 1;
  const int nbatch = (ntiles + h->max_batch - 1) / h->max_batch;
  return (ntiles + nbatch - 1) / nbatch;
}

// Tile outputs of the chop path use the compact [tile][P][P][4] fp16 layout when it applies.
bool compact_tiles(const innfer_rrdb* h) { return h->cfg.fp16 && h->cfg.out_nc <= 4; }

// Compute tiles [t_begin, t_end) of the plan from `src` and store them at
// tiles_base + t * tile_bytes.  tiles_base may be peer memory of another GPU (P2P stores over
// NVLink): the HR_conv1 epilogue writes there directly, no staging copy.
int compute_tiles(innfer_rrdb* h, const void* src, PixelDType st, const TilePlan& plan, int t_begin, int t_end,
                  void* tiles_base, cudaStream_t stream) {
  const int p = plan.p, s = h->cfg.scale;
  const int n = t_end - t_begin;
  if (n <= 0) return 0;
  h->i2i_per_sample = true;   // the reference runs the tiles one by one (run.py:187-197)
  int B = pick_batch(h, n, p);
  int rc;
  // Workspace grows with the batch (about 215 MB per 200x200 tile at 4x): when the device cannot
  // hold it (small GPU, several handles alive) shrink the batch instead of failing.
  while ((rc = ensure_workspace(h, B, p, p)) == INNFER_E_NOMEM && B > 1) {
    cudaGetLastError();  // clear the allocation error
    h->max_batch = B / 2;
    B = pick_batch(h, n, p);
  }
  if (rc) return rc;
  const int oct = (h->cfg.out_nc + 7) / 8;
  const bool compact = compact_tiles(h);
  const size_t tile_out_elems = compact ? (size_t)(s * p) * (s * p) * 4 : (size_t)oct * (s * p) * (s * p) * 8;
  const bool skip_pad = st == kU8 && h->cfg.in_nc <= 8 && h->in_ct() > 1;
  if (skip_pad && (h->in_pad_gen != h->in_tiles.gen || h->in_pad_p != p)) {
    if (cudaMemsetAsync(h->in_tiles.p, 0, h->in_tiles.bytes, stream) != cudaSuccess)
      return fail(INNFER_E_CUDA, "input tile zero fill failed");
    h->in_pad_gen = h->in_tiles.gen;
    h->in_pad_p = p;
  }
  for (int t0 = t_begin; t0 < t_end; t0 += B) {
    const int nb = (t_end - t0) < B ? (t_end - t0) : B;
    if (h->cfg.fp16) {
      rc = launch_image_to_tiles(src, st, h->cfg.in_nc, plan, t0, nb, reinterpret_cast<__half*>(h->in_tiles.p),
                                 h->in_ct(), stream, skip_pad);
    } else {
      rc = launch_image_to_tiles_f32(src, st, h->cfg.in_nc, plan, t0, nb, reinterpret_cast<float*>(h->in_tiles.p),
                                     h->in_ct(), stream, skip_pad);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (rc) return fail(INNFER_E_CUDA, "image_to_tiles launch failed");
    ChunkView dv;
    dv.base = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(tiles_base) + (size_t)t0 * tile_out_elems * h->esz());
    dv.CT = oct;
    dv.chunk0 = 0;
    if ((rc = forward_tiles(h, nb, p, p, dv, compact, stream))) return rc;
  }
  return 0;
}

int blend_tiles(innfer_rrdb* h, const void* tiles_base, const TilePlan& plan, void* dst, PixelDType dt,
                cudaStream_t stream) {
  const int oct = (h->cfg.out_nc + 7) / 8;
  int rc;
  if (h->cfg.fp16)
    rc = launch_blend(reinterpret_cast<const __half*>(tiles_base), compact_tiles(h) ? 0 : oct, plan, h->cfg.scale,
                      h->cfg.out_nc, dst, dt, stream);
  else
    rc = launch_blend_f32(reinterpret_cast<const float*>(tiles_base), oct, plan, h->cfg.scale, h->cfg.out_nc, dst, dt, stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc == -2) return fail(INNFER_E_UNSUPPORTED, "blend geometry not reproducible: negative blend core or an HR tile grid that differs from the LR one (utils.py:396-411)");
  if (rc) return fail(INNFER_E_CUDA, "blend launch failed");
  return 0;
}

int make_plan_checked(innfer_rrdb* h, int H, int W, int patch, double step, TilePlan& plan) {
  if (!h->finalized) return fail(INNFER_E_STATE, "innfer_rrdb_finalize has not been called");
  // recompose_tensor's own assertion (utils.py:391); Model.chop_forward defaults to 1.0, __call__ passes 0.5
  if (!(step >= 0.5 && step <= 1.0)) return fail(INNFER_E_INVALID, "step must be in [0.5, 1.0]");
  if (make_tile_plan(H, W, patch, step, plan)) return fail(INNFER_E_INVALID, "cannot tile this image size");
  return 0;
}

size_t tile_bytes(const innfer_rrdb* h, const TilePlan& plan) {
  const int oct = (h->cfg.out_nc + 7) / 8;
  const size_t P = (size_t)h->cfg.scale * plan.p;
  if (compact_tiles(h)) return P * P * 4 * sizeof(__half);
  return (size_t)oct * P * P * 8 * h->esz();
}

int chop_impl(innfer_rrdb* h, const void* src, PixelDType st, int H, int W, int patch, double step, void* dst,
              PixelDType dt, cudaStream_t stream) {
  TilePlan plan;
  int rc;
  if ((rc = make_plan_checked(h, H, W, patch, step, plan))) return rc;
  const int ntiles = plan.nty * plan.ntx;
  if (h->out_tiles.ensure((size_t)ntiles * tile_bytes(h, plan)))
    return fail(INNFER_E_NOMEM, "tile output allocation failed");
  if ((rc = compute_tiles(h, src, st, plan, 0, ntiles, h->out_tiles.p, stream))) return rc;
  return blend_tiles(h, h->out_tiles.p, plan, dst, dt, stream);
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* innfer_last_error(void) { return g_err.c_str(); }
const char* innfer_version(void) { return "innfer_b200 0.1 (sm_100a)"; }
uint64_t innfer_kernel_launches(void) { return g_launches.load(); }

int innfer_rrdb_create(const innfer_rrdb_cfg* cfg, int device, innfer_rrdb** out) {
  if (!cfg || !out) return fail(INNFER_E_INVALID, "null argument");
  if (cfg->nf != 64 && cfg->nf != 32) return fail(INNFER_E_UNSUPPORTED, "nf must be 32 or 64");
  if (cfg->in_nc < 1 || cfg->in_nc > 16 || cfg->out_nc < 1 || cfg->out_nc > 8)
    return fail(INNFER_E_UNSUPPORTED, "in_nc must be <= 16 and out_nc <= 8");
  if (cfg->nb < 1) return fail(INNFER_E_INVALID, "nb must be >= 1");
  int n_up = 0, f = 2;
  switch (cfg->scale) {
    case 1: n_up = 0; break;
    case 2: n_up = 1; break;
    case 3: n_up = 1; f = 3; break;
    case 4: n_up = 2; break;
    case 8: n_up = 3; break;
    default: return fail(INNFER_E_UNSUPPORTED, "scale must be 1, 2, 3, 4 or 8");
  }
  GUARD_DEVICE(device);
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(INNFER_E_UNSUPPORTED, "this library contains sm_100a code only; device is not compute capability 10.x");
  innfer_rrdb* h = new innfer_rrdb();
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->n_up = n_up;
  h->up_factor = f;
  *out = h;
  return 0;
}

int innfer_srresnet_create(const innfer_srresnet_cfg* cfg, int device, innfer_rrdb** out) {
  if (!cfg || !out) return fail(INNFER_E_INVALID, "null argument");
  if (cfg->upsample_mode != 0 && cfg->upsample_mode != 1) return fail(INNFER_E_INVALID, "upsample_mode must be 0 or 1");
  innfer_rrdb_cfg base = {cfg->in_nc, cfg->out_nc, cfg->nf, cfg->nb, 32, cfg->scale, 0, cfg->fp16};
  int rc = innfer_rrdb_create(&base, device, out);
  if (rc) return rc;
  (*out)->arch = 1;
  (*out)->res_scale = cfg->res_scale;
  (*out)->up_factor = cfg->scale == 3 ? 3 : 2;
  (*out)->ps_mode = cfg->upsample_mode == 0;
  return 0;
}

int innfer_ppon_create(const innfer_ppon_cfg* cfg, int device, innfer_rrdb** out) {
  if (!cfg || !out) return fail(INNFER_E_INVALID, "null argument");
  if (cfg->nf != 64) return fail(INNFER_E_UNSUPPORTED, "PPON: nf must be 64 (RRBlock_32 is hard-wired to 64 channels)");
  innfer_rrdb_cfg base = {cfg->in_nc, cfg->out_nc, cfg->nf, cfg->nb, 32, cfg->scale, 0, cfg->fp16};
  int rc = innfer_rrdb_create(&base, device, out);
  if (rc) return rc;
  (*out)->arch = 2;
  (*out)->ppon_alpha = cfg->alpha;
  return 0;
}

int innfer_pan_create(const innfer_pan_cfg* cfg, int device, innfer_rrdb** out) {
  if (!cfg || !out) return fail(INNFER_E_INVALID, "null argument");
  if (cfg->nf < 8 || cfg->nf > 64 || cfg->nf % 8) return fail(INNFER_E_UNSUPPORTED, "PAN: nf must be a multiple of 8, at most 64");
  const int unf = cfg->scale == 1 ? cfg->nf : cfg->unf;   // PAN_arch.py:111-112
  if (unf < 1 || unf > 64) return fail(INNFER_E_UNSUPPORTED, "PAN: unf must be in 1..64");
  if (cfg->in_nc != cfg->out_nc || cfg->in_nc > 8)
    return fail(INNFER_E_UNSUPPORTED, "PAN: in_nc must equal out_nc (the bilinear skip is added to the output) and be <= 8");
  innfer_rrdb_cfg base = {cfg->in_nc, cfg->out_nc, 64, cfg->nb, 32, cfg->scale, 0, cfg->fp16};
  int rc = innfer_rrdb_create(&base, device, out);
  if (rc) return rc;
  (*out)->cfg.nf = cfg->nf;
  (*out)->arch = 3;
  (*out)->pan_unf = unf;
  (*out)->pan_sa = cfg->self_attention != 0;
  (*out)->pan_double = cfg->double_scpa != 0;
  return 0;
}

int innfer_i2i_create(const innfer_i2i_cfg* cfg, int device, innfer_rrdb** out) {
  if (!cfg || !out) return fail(INNFER_E_INVALID, "null argument");
  if (cfg->kind != 0 && cfg->kind != 1) return fail(INNFER_E_INVALID, "kind must be 0 (UnetGenerator) or 1 (ResnetGenerator)");
  if (cfg->norm != 0 && cfg->norm != 1) return fail(INNFER_E_INVALID, "norm must be 0 (batch) or 1 (instance)");
  if (cfg->ngf < 8 || cfg->ngf % 8) return fail(INNFER_E_UNSUPPORTED, "ngf must be a multiple of 8");
  if (cfg->kind == 0 && (cfg->depth < 5 || cfg->depth > 12)) return fail(INNFER_E_UNSUPPORTED, "num_downs must be in [5, 12]");
  if (cfg->kind == 1 && (cfg->depth < 0 || cfg->depth > 64)) return fail(INNFER_E_UNSUPPORTED, "n_blocks must be in [0, 64]");
  if (cfg->unit_io && cfg->kind != 1)
    return fail(INNFER_E_UNSUPPORTED, "unit_io needs reflection padding at the first conv: ResnetGenerator only");
  innfer_rrdb_cfg base = {cfg->in_nc, cfg->out_nc, 64, 1, 32, 1, 0, cfg->fp16};
  int rc = innfer_rrdb_create(&base, device, out);
  if (rc) return rc;
  innfer_rrdb* h = *out;
  h->arch = 4;
  h->cfg.nf = cfg->ngf;
  h->cfg.nb = cfg->depth;
  h->i2i_cfg.kind = cfg->kind;
  h->i2i_cfg.in_nc = cfg->in_nc;
  h->i2i_cfg.out_nc = cfg->out_nc;
  h->i2i_cfg.ngf = cfg->ngf;
  h->i2i_cfg.depth = cfg->depth;
  h->i2i_cfg.norm = cfg->norm;
  h->i2i_cfg.train = cfg->train != 0;
  h->i2i_cfg.fp16 = cfg->fp16 != 0;
  h->i2i_cfg.unit_io = cfg->unit_io != 0;
  return 0;
}

int innfer_rrdb_load(innfer_rrdb* h, const char* key, const float* host_data, const int64_t* shape, int ndim) {
  if (!h || !key || !host_data || !shape || ndim < 1 || ndim > 4) return fail(INNFER_E_INVALID, "bad argument");
  if (h->finalized) return fail(INNFER_E_STATE, "handle already finalized");
  Param p;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    p.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  p.data.assign(host_data, host_data + n);
  h->params[key] = std::move(p);
  return 0;
}

int innfer_rrdb_finalize(innfer_rrdb* h) {
  if (!h) return fail(INNFER_E_INVALID, "null handle");
  if (h->finalized) return 0;
  int rc;
  GUARD_DEVICE(h->device);
  const auto& c = h->cfg;
  size_t expected = 0;
  if (h->arch == 3) return finalize_pan(h);
  if (h->arch == 4) {
    // keys of UnetGenerator / ResnetGenerator (UNet_arch.py:120-158, ResNet_arch.py:55-91); entries the network does not
    // use (InstanceNorm running statistics of old checkpoints, num_batches_tracked) are the host's strict / non-strict
    // business (run.py:299-309) and are ignored here
    h->i2i.reset(new I2INet(h->i2i_cfg, h->num_sms));
    std::string err;
    size_t consumed = 0;
    auto get = [h](const std::string& key, std::vector<int64_t>& shape) -> const float* {
      auto it = h->params.find(key);
      if (it == h->params.end()) return nullptr;
      shape = it->second.shape;
      return it->second.data.data();
    };
    if ((rc = h->i2i->build(get, err, &consumed))) {
      h->i2i.reset();
      return fail(i2i_code(rc), err);
    }
    h->params.clear();
    h->finalized = true;
    return 0;
  }
  if (h->arch == 2) {
    // PPON keys (PPON_arch.py:24-63 through block.sequential's flattening)
    if ((rc = build_layer(h, h->fea, "CFEM.0", c.nf, c.in_nc, 1))) return rc;
    expected += 2;
    const int nblk = (c.nb + 4) * 3;
    h->prb.resize((size_t)nblk * 10);
    for (int i = 0; i < c.nb + 4; ++i)
      for (int r = 0; r < 3; ++r) {
        char pre[96];
        if (i < c.nb) snprintf(pre, sizeof pre, "CFEM.1.sub.%d.RB%d", i, r + 1);
        else if (i < c.nb + 2) snprintf(pre, sizeof pre, "SFEM.%d.RB%d", i - c.nb, r + 1);
        else snprintf(pre, sizeof pre, "PFEM.%d.RB%d", i - c.nb - 2, r + 1);
        ConvLayer* L = &h->prb[((size_t)i * 3 + r) * 10];
        if ((rc = build_layer(h, L[0], std::string(pre) + ".c1", c.nf, c.nf, 1))) return rc;
        for (int k = 1; k <= 8; ++k)
          if ((rc = build_layer(h, L[k], std::string(pre) + ".d" + std::to_string(k), c.nf / 2, c.nf, 1, 3, true, k))) return rc;
        if ((rc = build_layer(h, L[9], std::string(pre) + ".c2", c.nf, c.nf * 4, 1, 1))) return rc;
        expected += 20;
      }
    char key[64];
    snprintf(key, sizeof key, "CFEM.1.sub.%d", c.nb);
    if ((rc = build_layer(h, h->lr_conv, key, c.nf, c.nf, 1))) return rc;
    expected += 2;
    const char* names[3] = {"CRM", "SRM", "PRM"};
    for (int t = 0; t < 3; ++t) {
      h->ptail[t].resize((size_t)h->n_up + 2);
      for (int i = 0; i < h->n_up; ++i) {
        snprintf(key, sizeof key, "%s.%d", names[t], 1 + 3 * i);
        if ((rc = build_layer(h, h->ptail[t][i], key, c.nf, c.nf, h->up_factor))) return rc;
      }
      snprintf(key, sizeof key, "%s.%d", names[t], 3 * h->n_up);
      if ((rc = build_layer(h, h->ptail[t][h->n_up], key, c.nf, c.nf, 1))) return rc;
      snprintf(key, sizeof key, "%s.%d", names[t], 3 * h->n_up + 2);
      if ((rc = build_layer(h, h->ptail[t][h->n_up + 1], key, c.out_nc, c.nf, 1))) return rc;
      expected += 2 * ((size_t)h->n_up + 2);
    }
    if (h->params.size() != expected)
      return fail(INNFER_E_INVALID, "unexpected keys in state dict (" + std::to_string(h->params.size()) +
                                        " loaded, " + std::to_string(expected) + " expected)");
    h->params.clear();
    h->finalized = true;
    return 0;
  }
  if ((rc = build_layer(h, h->fea, "model.0", c.nf, c.in_nc, 1))) return rc;
  expected += 2;
  if (h->arch == 1) {
    // SRResNet keys (block.py:197-210 flattening): model.1.sub.<i>.res.{0,2}, model.1.sub.<nb>,
    // then per upsampler block conv at 2+3i (pixelshuffle: conv, PixelShuffle, ReLU) or 3+3i (upconv:
    // Upsample, conv, ReLU), HR_conv0 at 2+3*n_up, HR_conv1 at 4+3*n_up
    h->rdb.resize((size_t)c.nb * 2);
    for (int b = 0; b < c.nb; ++b)
      for (int k = 0; k < 2; ++k) {
        char key[96];
        snprintf(key, sizeof key, "model.1.sub.%d.res.%d", b, 2 * k);
        if ((rc = build_layer(h, h->rdb[(size_t)b * 2 + k], key, c.nf, c.nf, 1))) return rc;
        expected += 2;
      }
    char key[64];
    snprintf(key, sizeof key, "model.1.sub.%d", c.nb);
    if ((rc = build_layer(h, h->lr_conv, key, c.nf, c.nf, 1))) return rc;
    expected += 2;
    h->ups.resize(h->n_up);
    for (int i = 0; i < h->n_up; ++i) {
      snprintf(key, sizeof key, "model.%d", (h->ps_mode ? 2 : 3) + 3 * i);
      rc = h->ps_mode ? build_ps_layer(h, h->ups[i], key, c.nf, c.nf, h->up_factor)
                      : build_layer(h, h->ups[i], key, c.nf, c.nf, h->up_factor);
      if (rc) return rc;
      expected += 2;
    }
    snprintf(key, sizeof key, "model.%d", 2 + 3 * h->n_up);
    if ((rc = build_layer(h, h->hr0, key, c.nf, c.nf, 1))) return rc;
    snprintf(key, sizeof key, "model.%d", 4 + 3 * h->n_up);
    if ((rc = build_layer(h, h->hr1, key, c.out_nc, c.nf, 1))) return rc;
    expected += 4;
    if (h->params.size() != expected)
      return fail(INNFER_E_INVALID, "unexpected keys in state dict (" + std::to_string(h->params.size()) +
                                        " loaded, " + std::to_string(expected) + " expected)");
    h->params.clear();
    h->finalized = true;
    return 0;
  }
  h->rdb.resize((size_t)c.nb * 15);
  if (c.plus) h->c1x1.resize((size_t)c.nb * 3);
  for (int b = 0; b < c.nb; ++b)
    for (int r = 0; r < 3; ++r) {
      if (c.plus) {
        char key[96];
        snprintf(key, sizeof key, "model.1.sub.%d.RDB%d.conv1x1", b, r + 1);
        if ((rc = build_layer(h, h->c1x1[(size_t)b * 3 + r], key, 32, c.nf, 1, 1, false))) return rc;
        expected += 1;
      }
      for (int k = 0; k < 5; ++k) {
        char key[96];
        snprintf(key, sizeof key, "model.1.sub.%d.RDB%d.conv%d.0", b, r + 1, k + 1);
        const int cin = c.nf + 32 * k, cout = k < 4 ? 32 : c.nf;
        if ((rc = build_layer(h, h->rdb[((size_t)b * 3 + r) * 5 + k], key, cout, cin, 1))) return rc;
        expected += 2;
      }
    }
  {
    char key[64];
    snprintf(key, sizeof key, "model.1.sub.%d", c.nb);
    if ((rc = build_layer(h, h->lr_conv, key, c.nf, c.nf, 1))) return rc;
    expected += 2;
  }
  h->ups.resize(h->n_up);
  for (int i = 0; i < h->n_up; ++i) {
    char key[64];
    snprintf(key, sizeof key, "model.%d", 3 + 3 * i);
    if ((rc = build_layer(h, h->ups[i], key, c.nf, c.nf, h->up_factor))) return rc;
    expected += 2;
  }
  {
    char key[64];
    snprintf(key, sizeof key, "model.%d", 2 + 3 * h->n_up);
    if ((rc = build_layer(h, h->hr0, key, c.nf, c.nf, 1))) return rc;
    snprintf(key, sizeof key, "model.%d", 4 + 3 * h->n_up);
    if ((rc = build_layer(h, h->hr1, key, c.out_nc, c.nf, 1))) return rc;
    expected += 4;
  }
  if (h->params.size() != expected) {
    // strict load (run.py:93): unexpected keys are an error
    return fail(INNFER_E_INVALID, "unexpected keys in state dict (" + std::to_string(h->params.size()) +
                                      " loaded, " + std::to_string(expected) + " expected)");
  }
  h->params.clear();
  h->finalized = true;
  return 0;
}

void innfer_rrdb_destroy(innfer_rrdb* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  delete h;
}

int innfer_rrdb_set_max_batch(innfer_rrdb* h, int max_tiles) {
  if (!h || max_tiles < 1) return fail(INNFER_E_INVALID, "bad argument");
  h->max_batch = max_tiles;
  return 0;
}

int innfer_rrdb_profile(innfer_rrdb* h, int enable) {
  if (!h) return fail(INNFER_E_INVALID, "null handle");
  for (auto& e : h->prof_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  h->prof_events.clear();
  h->prof_conv_launches = 0;
  for (auto& f : h->fam)
    for (auto& e : f.second.ev) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
  h->fam.clear();
  h->profiling = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  return 0;
}

int innfer_rrdb_profile_families(innfer_rrdb* h, char* buf, uint64_t cap, uint64_t* needed) {
  if (!h || !needed) return fail(INNFER_E_INVALID, "null argument");
  std::string out;
  for (auto& kv : h->fam) {
    double ms = 0.0;
    for (auto& e : kv.second.ev) {
      CU_TRY(cudaEventSynchronize(e.second));
      float t = 0.f;
      CU_TRY(cudaEventElapsedTime(&t, e.first, e.second));
      ms += t;
    }
    char line[256];
    snprintf(line, sizeof line, "%s\t%llu\t%.6f\t%.17g\t%.17g\n", kv.first.c_str(), (unsigned long long)kv.second.launches, ms,
             kv.second.flop, kv.second.bytes);
    out += line;
  }
  *needed = out.size() + 1;
  if (buf && cap >= out.size() + 1) std::memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

int innfer_rrdb_profile_read(innfer_rrdb* h, double* conv_ms, uint64_t* conv_launches) {
  if (!h || !conv_ms || !conv_launches) return fail(INNFER_E_INVALID, "null argument");
  double total = 0.0;
  for (auto& e : h->prof_events) {
    CU_TRY(cudaEventSynchronize(e.second));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, e.first, e.second));
    total += ms;
  }
  *conv_ms = total;
  *conv_launches = h->prof_conv_launches;
  return 0;
}

int innfer_rrdb_forward(innfer_rrdb* h, const void* x, int n, int hgt, int wid, void* y, int dtype, void* stream) {
  if (!h || !x || !y) return fail(INNFER_E_INVALID, "null argument");
  if (!h->finalized) return fail(INNFER_E_STATE, "innfer_rrdb_finalize has not been called");
  if (dtype != INNFER_F16 && dtype != INNFER_F32) return fail(INNFER_E_INVALID, "dtype must be F16 or F32");
  if (n < 1 || hgt < 1 || wid < 1) return fail(INNFER_E_INVALID, "bad shape");
  int rc;
  GUARD_DEVICE(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int s = h->cfg.scale;
  const int oct = (h->cfg.out_nc + 7) / 8;
  h->i2i_per_sample = false;   // one reference call on the whole batch: train-mode BatchNorm sees all of it
  if (h->arch == 4 && n > h->max_batch && h->i2i_cfg.norm == 0 && h->i2i_cfg.train)
    return fail(INNFER_E_UNSUPPORTED, "batch statistics need the whole batch in one pass: raise max_batch");
  // batch images one at a time through max_batch-sized groups
  for (int b0 = 0; b0 < n; b0 += h->max_batch) {
    const int nb = (n - b0) < h->max_batch ? (n - b0) : h->max_batch;
    if ((rc = ensure_workspace(h, nb, hgt, wid))) return rc;
    if (h->out_tiles.ensure((size_t)nb * oct * s * hgt * s * wid * 8 * h->esz()))
      return fail(INNFER_E_NOMEM, "output tile allocation failed");
    const size_t in_off = (size_t)b0 * h->cfg.in_nc * hgt * wid * (dtype == INNFER_F16 ? 2 : 4);
    const size_t out_off = (size_t)b0 * h->cfg.out_nc * s * hgt * s * wid * (dtype == INNFER_F16 ? 2 : 4);
    const void* xs = reinterpret_cast<const uint8_t*>(x) + in_off;
    void* ys = reinterpret_cast<uint8_t*>(y) + out_off;
    h->in_pad_gen = ~0u;   // another tile geometry is written over the buffer
    if (h->cfg.fp16)
      rc = launch_nchw_to_chunks(xs, to_pix(dtype), nb, h->cfg.in_nc, hgt, wid, reinterpret_cast<__half*>(h->in_tiles.p), h->in_ct(), st);
    else
      rc = launch_nchw_to_chunks_f32(xs, to_pix(dtype), nb, h->cfg.in_nc, hgt, wid, reinterpret_cast<float*>(h->in_tiles.p), h->in_ct(), st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (rc) return fail(INNFER_E_CUDA, "nchw_to_chunks launch failed");
    ChunkView dv = view(h->out_tiles, oct, 0);
    if ((rc = forward_tiles(h, nb, hgt, wid, dv, false, st))) return rc;
    if (h->cfg.fp16)
      rc = launch_chunks_to_nchw(reinterpret_cast<const __half*>(h->out_tiles.p), oct, nb, h->cfg.out_nc, s * hgt, s * wid, ys, to_pix(dtype), st);
    else
      rc = launch_chunks_to_nchw_f32(reinterpret_cast<const float*>(h->out_tiles.p), oct, nb, h->cfg.out_nc, s * hgt, s * wid, ys, to_pix(dtype), st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (rc) return fail(INNFER_E_CUDA, "chunks_to_nchw launch failed");
  }
  return 0;
}

int innfer_rrdb_chop_forward(innfer_rrdb* h, const void* x, int H, int W, int patch_size, double step, void* y,
                             int dtype, void* stream) {
  if (!h || !x || !y) return fail(INNFER_E_INVALID, "null argument");
  if (dtype != INNFER_F16 && dtype != INNFER_F32) return fail(INNFER_E_INVALID, "dtype must be F16 or F32");
  int rc;
  GUARD_DEVICE(h->device);
  return chop_impl(h, x, to_pix(dtype), H, W, patch_size, step, y, to_pix(dtype), reinterpret_cast<cudaStream_t>(stream));
}

int innfer_rrdb_chop_forward_ex(innfer_rrdb* h, const void* x, int x_dtype, int H, int W, int patch_size, double step,
                                void* y, int y_dtype, void* stream) {
  if (!h || !x || !y) return fail(INNFER_E_INVALID, "null argument");
  for (int d : {x_dtype, y_dtype}) {
    if (d != INNFER_F16 && d != INNFER_F32 && d != INNFER_U8) return fail(INNFER_E_INVALID, "dtype must be F16, F32 or U8");
  }
  if (x_dtype == INNFER_U8 && h->cfg.in_nc != 3) return fail(INNFER_E_INVALID, "uint8 input needs a 3-channel model");
  if (y_dtype == INNFER_U8 && h->cfg.out_nc != 3) return fail(INNFER_E_INVALID, "uint8 output needs a 3-channel model");
  GUARD_DEVICE(h->device);
  return chop_impl(h, x, to_pix(x_dtype), H, W, patch_size, step, y, to_pix(y_dtype), reinterpret_cast<cudaStream_t>(stream));
}

int innfer_rrdb_upscale_u8_device(innfer_rrdb* h, const uint8_t* img, int H, int W, int patch_size, double step,
                                  uint8_t* out, void* stream) {
  if (!h || !img || !out) return fail(INNFER_E_INVALID, "null argument");
  if (h->cfg.in_nc != 3 || h->cfg.out_nc != 3) return fail(INNFER_E_INVALID, "uint8 path needs 3-channel models");
  int rc;
  GUARD_DEVICE(h->device);
  return chop_impl(h, img, kU8, H, W, patch_size, step, out, kU8, reinterpret_cast<cudaStream_t>(stream));
}

int innfer_rrdb_upscale_u8(innfer_rrdb* h, const uint8_t* img, int H, int W, int patch_size, double step,
                           uint8_t* out, void* stream) {
  if (!h || !img || !out) return fail(INNFER_E_INVALID, "null argument");
  int rc;
  GUARD_DEVICE(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int s = h->cfg.scale;
  const size_t ib = (size_t)H * W * 3, ob = ib * s * s;
  if (h->img_in.ensure(ib) || h->img_out.ensure(ob)) return fail(INNFER_E_NOMEM, "image buffer allocation failed");
  CU_TRY(cudaMemcpyAsync(h->img_in.p, img, ib, cudaMemcpyHostToDevice, st));
  rc = innfer_rrdb_upscale_u8_device(h, reinterpret_cast<const uint8_t*>(h->img_in.p), H, W, patch_size, step,
                                     reinterpret_cast<uint8_t*>(h->img_out.p), stream);
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(out, h->img_out.p, ob, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  return 0;
}

int innfer_rrdb_tile_buffer(innfer_rrdb* h, int H, int W, int patch_size, double step, void** ptr, uint64_t* bytes,
                            uint64_t* bytes_per_tile) {
  if (!h || !ptr || !bytes) return fail(INNFER_E_INVALID, "null argument");
  int rc;
  GUARD_DEVICE(h->device);
  TilePlan plan;
  if ((rc = make_plan_checked(h, H, W, patch_size, step, plan))) return rc;
  const size_t tb = tile_bytes(h, plan), total = (size_t)plan.nty * plan.ntx * tb;
  if (h->out_tiles.ensure(total)) return fail(INNFER_E_NOMEM, "tile output allocation failed");
  *ptr = h->out_tiles.p;
  *bytes = h->out_tiles.bytes;
  if (bytes_per_tile) *bytes_per_tile = tb;
  return 0;
}

int innfer_rrdb_tile_bytes(innfer_rrdb* h, int H, int W, int patch_size, double step, uint64_t* bytes_total,
                           uint64_t* bytes_per_tile, int* ntiles) {
  if (!h || !bytes_total) return fail(INNFER_E_INVALID, "null argument");
  TilePlan plan;
  int rc;
  if ((rc = make_plan_checked(h, H, W, patch_size, step, plan))) return rc;
  const size_t tb = tile_bytes(h, plan);
  *bytes_total = (uint64_t)plan.nty * plan.ntx * tb;
  if (bytes_per_tile) *bytes_per_tile = tb;
  if (ntiles) *ntiles = plan.nty * plan.ntx;
  return 0;
}

int innfer_rrdb_forward_tile_range(innfer_rrdb* h, const void* img, int img_dtype, int H, int W, int patch_size,
                                   double step, int t_begin, int t_end, void* tiles_base, void* stream) {
  if (!h || !img || !tiles_base) return fail(INNFER_E_INVALID, "null argument");
  int rc;
  GUARD_DEVICE(h->device);
  TilePlan plan;
  if ((rc = make_plan_checked(h, H, W, patch_size, step, plan))) return rc;
  if (t_begin < 0 || t_end > plan.nty * plan.ntx || t_begin > t_end) return fail(INNFER_E_INVALID, "bad tile range");
  if (img_dtype == INNFER_U8 && h->cfg.in_nc != 3) return fail(INNFER_E_INVALID, "uint8 input needs a 3-channel model");
  return compute_tiles(h, img, to_pix(img_dtype), plan, t_begin, t_end, tiles_base, reinterpret_cast<cudaStream_t>(stream));
}

int innfer_rrdb_blend_tiles(innfer_rrdb* h, const void* tiles_base, int H, int W, int patch_size, double step, void* dst,
                            int dst_dtype, void* stream) {
  if (!h || !tiles_base || !dst) return fail(INNFER_E_INVALID, "null argument");
  int rc;
  GUARD_DEVICE(h->device);
  TilePlan plan;
  if ((rc = make_plan_checked(h, H, W, patch_size, step, plan))) return rc;
  return blend_tiles(h, tiles_base, plan, dst, to_pix(dst_dtype), reinterpret_cast<cudaStream_t>(stream));
}

int innfer_ipc_export(const void* device_ptr, uint8_t handle[64]) {
  if (!device_ptr || !handle) return fail(INNFER_E_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  cudaIpcMemHandle_t hd;
  CU_TRY(cudaIpcGetMemHandle(&hd, const_cast<void*>(device_ptr)));
  std::memcpy(handle, &hd, 64);
  return 0;
}

int innfer_ipc_open(const uint8_t handle[64], void** device_ptr) {
  if (!handle || !device_ptr) return fail(INNFER_E_INVALID, "null argument");
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle, 64);
  CU_TRY(cudaIpcOpenMemHandle(device_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int innfer_ipc_close(void* device_ptr) {
  if (!device_ptr) return 0;
  CU_TRY(cudaIpcCloseMemHandle(device_ptr));
  return 0;
}

int innfer_device_alloc(int device, uint64_t bytes, void** ptr) {
  if (!ptr) return fail(INNFER_E_INVALID, "null argument");
  GUARD_DEVICE(device);
  CU_TRY(cudaMalloc(ptr, bytes));
  return 0;
}

int innfer_device_upload(void* device_dst, const void* host_src, uint64_t bytes, void* stream) {
  if (!device_dst || !host_src) return fail(INNFER_E_INVALID, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CU_TRY(cudaMemcpyAsync(device_dst, host_src, bytes, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));
  return 0;
}

int innfer_device_memset(void* ptr, int value, uint64_t bytes) {
  if (!ptr) return fail(INNFER_E_INVALID, "null argument");
  CU_TRY(cudaMemset(ptr, value, bytes));
  CU_TRY(cudaDeviceSynchronize());
  return 0;
}

int innfer_memcpy_async(void* dst, const void* src, uint64_t bytes, void* stream) {
  if (!dst || !src) return fail(INNFER_E_INVALID, "null argument");
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

int innfer_stream_signal(void* const* flags, int n, uint32_t value, void* stream) {
  if (!flags || n < 1 || n > kMaxSyncFlags) return fail(INNFER_E_INVALID, "bad flag list");
  int rc = launch_signal(reinterpret_cast<uint32_t* const*>(flags), n, value, reinterpret_cast<cudaStream_t>(stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc) return fail(INNFER_E_CUDA, "signal launch failed");
  return 0;
}

int innfer_stream_wait(void* const* flags, int n, uint32_t value, void* err_flag, uint64_t timeout_ms, void* stream) {
  if (!flags || n < 1 || n > kMaxSyncFlags) return fail(INNFER_E_INVALID, "bad flag list");
  int rc = launch_wait(reinterpret_cast<uint32_t* const*>(flags), n, value, reinterpret_cast<uint32_t*>(err_flag),
                       (unsigned long long)timeout_ms * 1000000ull, reinterpret_cast<cudaStream_t>(stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc) return fail(INNFER_E_CUDA, "wait launch failed");
  return 0;
}

int innfer_device_free(void* ptr) {
  if (ptr) CU_TRY(cudaFree(ptr));
  return 0;
}

int innfer_tiles_plan(int H, int W, int patch_size, double step, innfer_tile* out, int cap, int* n, int* tile_size) {
  TilePlan plan;
  if (make_tile_plan(H, W, patch_size, step, plan)) return fail(INNFER_E_INVALID, "cannot tile this image size");
  const int nt = plan.nty * plan.ntx;
  if (n) *n = nt;
  if (tile_size) *tile_size = plan.p;
  if (out) {
    for (int i = 0; i < nt && i < cap; ++i) {
      out[i].y0 = plan.ys[i / plan.ntx];
      out[i].x0 = plan.xs[i % plan.ntx];
    }
  }
  return 0;
}

int innfer_image_to_tiles(const void* src, int src_dtype, int C, int H, int W, int patch_size, double step,
                          void* dst_tiles, void* stream) {
  if (!src || !dst_tiles) return fail(INNFER_E_INVALID, "null argument");
  TilePlan plan;
  if (make_tile_plan(H, W, patch_size, step, plan)) return fail(INNFER_E_INVALID, "cannot tile this image size");
  const int CT = (C + 15) / 16 * 2;
  int rc = launch_image_to_tiles(src, to_pix(src_dtype), C, plan, 0, plan.nty * plan.ntx,
                                 reinterpret_cast<__half*>(dst_tiles), CT, reinterpret_cast<cudaStream_t>(stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc) return fail(INNFER_E_CUDA, "image_to_tiles launch failed");
  return 0;
}

int innfer_blend(const void* tiles, int H, int W, int patch_size, double step, int scale, int C, void* dst,
                 int dst_dtype, void* stream) {
  if (!tiles || !dst) return fail(INNFER_E_INVALID, "null argument");
  if (!(step >= 0.5 && step <= 1.0)) return fail(INNFER_E_INVALID, "step must be in [0.5, 1.0]");
  TilePlan plan;
  if (make_tile_plan(H, W, patch_size, step, plan)) return fail(INNFER_E_INVALID, "cannot tile this image size");
  int rc = launch_blend(reinterpret_cast<const __half*>(tiles), 1, plan, scale, C, dst, to_pix(dst_dtype),
                        reinterpret_cast<cudaStream_t>(stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc == -2) return fail(INNFER_E_UNSUPPORTED, "blend geometry not reproducible: negative blend core or an HR tile grid that differs from the LR one (utils.py:396-411)");
  if (rc) return fail(INNFER_E_CUDA, "blend launch failed");
  return 0;
}

int innfer_blend_f32(const float* tiles, int H, int W, int patch_size, double step, int scale, int C, void* dst,
                     int dst_dtype, void* stream) {
  if (!tiles || !dst) return fail(INNFER_E_INVALID, "null argument");
  if (!(step >= 0.5 && step <= 1.0)) return fail(INNFER_E_INVALID, "step must be in [0.5, 1.0]");
  TilePlan plan;
  if (make_tile_plan(H, W, patch_size, step, plan)) return fail(INNFER_E_INVALID, "cannot tile this image size");
  int rc = launch_blend_f32(tiles, 1, plan, scale, C, dst, to_pix(dst_dtype), reinterpret_cast<cudaStream_t>(stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (rc == -2) return fail(INNFER_E_UNSUPPORTED, "blend geometry not reproducible: negative blend core or an HR tile grid that differs from the LR one (utils.py:396-411)");
  if (rc) return fail(INNFER_E_CUDA, "blend launch failed");
  return 0;
}

int innfer_conv3x3(const void* x, int n, int Cin, int hgt, int wid, const float* w_oihw, const float* bias,
                   int Cout, int up, int lrelu, const void* res1, float alpha1, void* y, int dtype,
                   int use_fp32_kernel, void* stream) {
  if (!x || !w_oihw || !y) return fail(INNFER_E_INVALID, "null argument");
  g_wide_sep = 1;
  if (dtype != INNFER_F16 && dtype != INNFER_F32) return fail(INNFER_E_INVALID, "dtype must be F16 or F32");
  if (res1 && up != 1) return fail(INNFER_E_INVALID, "residual needs up == 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(INNFER_E_UNSUPPORTED, "device is not compute capability 10.x");
  ConvLayer L;
  std::string err;
  int rc = conv_layer_build(L, w_oihw, bias, Cout, Cin, up, err);
  if (rc) return fail(rc == -2 ? INNFER_E_UNSUPPORTED : INNFER_E_CUDA, err);
  const bool wide_mode = use_fp32_kernel == 2;   // 0: fp16 tiled layout, 1: fp32 direct kernel, 2: fp16 wide layout
  if (wide_mode) use_fp32_kernel = 0;
  const size_t esz = use_fp32_kernel ? 4 : 2;
  const int ict = L.Cin_pad / 8, oct = (Cout + 7) / 8;
  const size_t ipx = wide_mode ? (size_t)hgt * wide_cols(n, wid) : (size_t)n * hgt * wid, opx = ipx * up * up;
  DevBuf bi, bo, br;
  TmapCache cache;
  auto cleanup = [&]() {
    bi.release();
    bo.release();
    br.release();
    conv_layer_free(L);
  };
  if (bi.ensure(ipx * ict * 8 * esz) || bo.ensure(opx * oct * 8 * esz) || (res1 && br.ensure(opx * oct * 8 * esz))) {
    cleanup();
    return fail(INNFER_E_NOMEM, "allocation failed");
  }
  Epilogue ep;
  ep.lrelu = lrelu != 0;
  if (use_fp32_kernel) {
    rc = conv_direct_upload(L);
    rc |= launch_nchw_to_chunks_f32(x, to_pix(dtype), n, Cin, hgt, wid, reinterpret_cast<float*>(bi.p), ict, st);
    if (res1) {
      rc |= launch_nchw_to_chunks_f32(res1, to_pix(dtype), n, Cout, hgt, wid, reinterpret_cast<float*>(br.p), oct, st);
      ep.res1 = view(br, oct, 0);
      ep.alpha1 = alpha1;
    }
    if (!rc) rc = conv_direct_run(L, view(bi, ict, 0), n, hgt, wid, view(bo, oct, 0), oct, ep, st);
    if (!rc) rc = launch_chunks_to_nchw_f32(reinterpret_cast<const float*>(bo.p), oct, n, Cout, hgt * up, wid * up, y, to_pix(dtype), st);
  } else if (wide_mode) {
    // the production layout of the fp16 path: the n images side by side in one wide image
    const int pitch = wide_pitch(wid), cols = wide_cols(n, wid);
    cudaMemsetAsync(bi.p, 0, bi.bytes, st);
    rc = launch_nchw_to_wide(x, to_pix(dtype), n, Cin, hgt, wid, reinterpret_cast<__half*>(bi.p), ict, pitch, cols, st);
    if (res1 == x && Cin >= Cout) {
      // the residual is the conv's own input (its first Cout channels), like conv5 inside a dense block: the view the
      // engine passes there, which lets the CTA-pair kernel take it through identity MMAs (conv_rows.cu, IDT)
      ep.res1 = wview(bi, ict, 0, n, wid, 1);
      ep.alpha1 = alpha1;
    } else if (res1) {
      cudaMemsetAsync(br.p, 0, br.bytes, st);
      rc |= launch_nchw_to_wide(res1, to_pix(dtype), n, Cout, hgt, wid, reinterpret_cast<__half*>(br.p), oct, pitch, cols, st);
      ep.res1 = wview(br, oct, 0, n, wid, 1);
      ep.alpha1 = alpha1;
    }
    if (!rc) rc = conv_layer_run(L, cache, wview(bi, ict, 0, n, wid, 1), n, hgt, wid, wview(bo, oct, 0, n, wid, up), oct, ep,
                                 prop.multiProcessorCount, st);
    if (!rc) rc = launch_wide_to_nchw(reinterpret_cast<const __half*>(bo.p), oct, n, Cout, hgt * up, wid * up, pitch * up,
                                      cols * up, y, to_pix(dtype), st);
  } else {
    rc = launch_nchw_to_chunks(x, to_pix(dtype), n, Cin, hgt, wid, reinterpret_cast<__half*>(bi.p), ict, st);
    if (res1) {
      rc |= launch_nchw_to_chunks(res1, to_pix(dtype), n, Cout, hgt, wid, reinterpret_cast<__half*>(br.p), oct, st);
      ep.res1 = view(br, oct, 0);
      ep.alpha1 = alpha1;
    }
    if (!rc) rc = conv_layer_run(L, cache, view(bi, ict, 0), n, hgt, wid, view(bo, oct, 0), oct, ep, prop.multiProcessorCount, st);
    if (!rc) rc = launch_chunks_to_nchw(reinterpret_cast<const __half*>(bo.p), oct, n, Cout, hgt * up, wid * up, y, to_pix(dtype), st);
  }
  g_launches.fetch_add(3 + (res1 ? 1 : 0), std::memory_order_relaxed);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (rc) return fail(INNFER_E_CUDA, "conv3x3 launch failed (rc=" + std::to_string(rc) + ")");
  if (e != cudaSuccess) return cuda_fail(e, "conv3x3 execution");
  return 0;
}

// Debugging / measurement aid: `iters` back-to-back launches of ONE 3x3 conv (Cin -> Cout, + LeakyReLU,
// optional residual) on a wide batch of B images of H x W random pixels, timed with CUDA events after
// `warm` untimed launches.  Long runs show the kernel's power-limited steady state.
int innfer_gen_conv(const void* x, int n, int Cin, int hgt, int wid, const float* w, const float* bias, int Cout, int k,
                    int stride, int pad, int transposed, int out_pad, int reflect, int norm, const float* norm_weight,
                    const float* norm_bias, int act, int final_path, void* y, int dtype, void* stream) {
  if (!x || !w || !y) return fail(INNFER_E_INVALID, "null argument");
  if (dtype != INNFER_F16 && dtype != INNFER_F32) return fail(INNFER_E_INVALID, "dtype must be F16 or F32");
  if (n < 1 || Cin < 1 || Cout < 1 || hgt < 1 || wid < 1 || norm < 0 || norm > 2 || act < 0 || act > 3)
    return fail(INNFER_E_INVALID, "bad argument");
  if (norm && final_path) return fail(INNFER_E_INVALID, "the last-layer path has no norm");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(INNFER_E_UNSUPPORTED, "device is not compute capability 10.x");
  I2ICfg c;
  c.kind = 2;
  c.in_nc = Cin;
  c.out_nc = Cout;
  c.norm = norm == 1 ? 1 : 0;
  c.train = 1;
  c.fp16 = dtype == INNFER_F16;
  c.sl_cout = Cout; c.sl_k = k; c.sl_stride = stride; c.sl_pad = pad; c.sl_transposed = transposed; c.sl_out_pad = out_pad;
  c.sl_reflect = reflect; c.sl_bias = bias != nullptr; c.sl_norm = norm != 0; c.sl_act = act; c.sl_final = final_path;
  I2INet net(c, prop.multiProcessorCount);
  const std::vector<float> ones((size_t)Cout, 1.f), zeros((size_t)Cout, 0.f);
  auto get = [&](const std::string& key, std::vector<int64_t>& shape) -> const float* {
    if (key == "conv.weight") {
      shape = transposed ? std::vector<int64_t>{Cin, Cout, k, k} : std::vector<int64_t>{Cout, Cin, k, k};
      return w;
    }
    shape = {Cout};
    if (key == "conv.bias") return bias;
    if (key == "norm.weight") return norm_weight ? norm_weight : ones.data();
    if (key == "norm.bias") return norm_bias ? norm_bias : zeros.data();
    if (key == "norm.running_mean") return zeros.data();
    if (key == "norm.running_var") return ones.data();
    return nullptr;
  };
  std::string err;
  size_t consumed = 0;
  int rc = net.build(get, err, &consumed);
  if (rc) return fail(i2i_code(rc), err);
  if ((rc = net.check_size(hgt, wid, err))) return fail(i2i_code(rc), err);
  const int Ho = net.out_size(hgt), Wo = net.out_size(wid);
  const int ict = (Cin + 7) / 8, oct = (Cout + 7) / 8;
  const size_t esz = dtype == INNFER_F16 ? 2 : 4;
  DevBuf in, out;
  if (in.ensure((size_t)n * ict * hgt * wid * 8 * esz) || out.ensure((size_t)n * oct * Ho * Wo * 8 * esz)) {
    in.release();
    out.release();
    return fail(INNFER_E_NOMEM, "temporary allocation failed");
  }
  if (dtype == INNFER_F16)
    rc = launch_nchw_to_chunks(x, kF16, n, Cin, hgt, wid, reinterpret_cast<__half*>(in.p), ict, st);
  else
    rc = launch_nchw_to_chunks_f32(x, kF32, n, Cin, hgt, wid, reinterpret_cast<float*>(in.p), ict, st);
  if (!rc) {
    rc = net.forward(in.p, ict, n, hgt, wid, GenView{out.p, oct, 0}, false, false, st, err);
    if (rc) rc = fail(i2i_code(rc), err);
  } else {
    rc = fail(INNFER_E_CUDA, "nchw_to_chunks launch failed");
  }
  if (!rc) {
    int e2;
    if (dtype == INNFER_F16)
      e2 = launch_chunks_to_nchw(reinterpret_cast<const __half*>(out.p), oct, n, Cout, Ho, Wo, y, kF16, st);
    else
      e2 = launch_chunks_to_nchw_f32(reinterpret_cast<const float*>(out.p), oct, n, Cout, Ho, Wo, y, kF32, st);
    if (e2) rc = fail(INNFER_E_CUDA, "chunks_to_nchw launch failed");
  }
  g_launches.fetch_add(net.launches() + 2, std::memory_order_relaxed);
  const cudaError_t se = cudaStreamSynchronize(st);
  in.release();
  out.release();
  if (!rc && se != cudaSuccess) return cuda_fail(se, "innfer_gen_conv");
  return rc;
}

uint64_t innfer_debug_i2i_halo_launches(void) { return i2i_halo_launches(); }
uint64_t innfer_debug_i2i_graph_replays(void) { return i2i_graph_replays(); }

int innfer_debug_conv_loop(int Cin, int Cout, int B, int H, int W, int with_res, int warm, int iters, float* ms_out) {
  if (!ms_out) return fail(INNFER_E_INVALID, "null argument");
  g_wide_sep = 1;
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  std::vector<float> w((size_t)Cout * Cin * 9), b(Cout);
  uint32_t rng = 12345u;
  auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return ((rng >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : w) v = rnd() * 0.1f;
  for (auto& v : b) v = rnd();
  ConvLayer L;
  std::string err;
  if (conv_layer_build(L, w.data(), b.data(), Cout, Cin, 1, err)) return fail(INNFER_E_CUDA, err);
  const int ict = L.Cin_pad / 8, oct = (Cout + 7) / 8;
  const int Wtot = wide_cols(B, W);
  DevBuf bi, bo, br;
  TmapCache cache;
  const size_t ib = (size_t)ict * H * Wtot * 16, ob = (size_t)oct * H * Wtot * 16;
  if (bi.ensure(ib) || bo.ensure(ob) || br.ensure(ob)) return fail(INNFER_E_NOMEM, "allocation failed");
  {
    std::vector<__half> hbuf(ib / 2);
    for (auto& v : hbuf) v = __float2half_rn(rnd());
    // separators must be zero
    for (int c = 0; c < ict; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < Wtot; ++x)
          if (x % wide_pitch(W) >= W || x / wide_pitch(W) >= B)
            for (int e = 0; e < 8; ++e) hbuf[(((size_t)c * H + y) * Wtot + x) * 8 + e] = __float2half_rn(0.f);
    CU_TRY(cudaMemcpy(bi.p, hbuf.data(), ib, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(br.p, hbuf.data(), ob < ib ? ob : ib, cudaMemcpyHostToDevice));
  }
  Epilogue ep;
  ep.lrelu = !with_res;
  ChunkView vi = wview(bi, ict, 0, B, W, 1), vo = wview(bo, oct, 0, B, W, 1);
  if (with_res) {
    ep.res1 = wview(br, oct, 0, B, W, 1);
    ep.alpha1 = 0.2f;
  }
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  int rc = 0;
  for (int i = 0; i < warm + iters && !rc; ++i) {
    if (i == warm) CU_TRY(cudaEventRecord(e0, nullptr));
    rc = conv_layer_run(L, cache, vi, B, H, W, vo, oct, ep, prop.multiProcessorCount, nullptr);
  }
  CU_TRY(cudaEventRecord(e1, nullptr));
  cudaError_t e = cudaDeviceSynchronize();
  if (!rc && e == cudaSuccess) cudaEventElapsedTime(ms_out, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  bi.release(); bo.release(); br.release();
  conv_layer_free(L);
  if (rc) return fail(INNFER_E_CUDA, "conv launch failed rc=" + std::to_string(rc));
  if (e != cudaSuccess) return cuda_fail(e, "conv loop");
  return 0;
}

int innfer_debug_set_trace(void* device_buf) {
  innfer::g_rows_trace = reinterpret_cast<long long*>(device_buf);
  return 0;
}

int innfer_color_fix(const uint8_t* lr, int h, int w, const uint8_t* sr, int H, int W, uint8_t* out, void* stream) {
  if (!lr || !sr || !out) return fail(INNFER_E_INVALID, "null argument");
  int launches = 0;
  int rc = color_fix_run(lr, h, w, sr, H, W, out, reinterpret_cast<cudaStream_t>(stream), &launches);
  g_launches.fetch_add(launches, std::memory_order_relaxed);
  if (rc == -1) return fail(INNFER_E_INVALID, "color_fix: unsupported image shapes");
  if (rc == -5) return fail(INNFER_E_NOMEM, "color_fix: scratch allocation failed");
  if (rc) return fail(INNFER_E_CUDA, "color_fix launch failed");
  return 0;
}

int innfer_color_fix_host(const uint8_t* lr, int h, int w, const uint8_t* sr, int H, int W, uint8_t* out, int device) {
  if (!lr || !sr || !out) return fail(INNFER_E_INVALID, "null argument");
  GUARD_DEVICE(device);
  DevBuf dl, ds, dout;
  const size_t lb = (size_t)h * w * 3, sb = (size_t)H * W * 3;
  if (dl.ensure(lb) || ds.ensure(sb) || dout.ensure(sb)) {
    dl.release(); ds.release(); dout.release();
    return fail(INNFER_E_NOMEM, "allocation failed");
  }
  int rc = 0;
  cudaError_t e = cudaMemcpy(dl.p, lr, lb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(ds.p, sr, sb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = innfer_color_fix(reinterpret_cast<const uint8_t*>(dl.p), h, w, reinterpret_cast<const uint8_t*>(ds.p), H, W,
                          reinterpret_cast<uint8_t*>(dout.p), nullptr);
    if (!rc) e = cudaMemcpy(out, dout.p, sb, cudaMemcpyDeviceToHost);
  }
  dl.release(); ds.release(); dout.release();
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "color_fix_host");
  return 0;
}

#pragma GCC visibility pop
}  // extern "C"
