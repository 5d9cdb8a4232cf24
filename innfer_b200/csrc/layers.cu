#include "layers.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "conv_rows.cuh"
#include "tmap.cuh"

namespace innfer {

thread_local const char* g_last_conv_kernel = "";   // which kernel the last conv_layer_run picked (profiling)
long long* g_rows_trace = nullptr;  // device buffer of 3072 int64 (innfer_debug_set_trace), debugging only

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

int conv_layer_build(ConvLayer& L, const float* w_in, const float* bias, int Cout, int Cin, int up,
                     std::string& err, int ksize, int dil) {
  // a 1x1 kernel is embedded as the centre tap of a 3x3 one; the tap tables below then keep only
  // taps with a non-zero weight plane, i.e. exactly that centre tap
  std::vector<float> w3;
  const float* w = w_in;
  if (ksize == 1) {
    if (up != 1) {
      err = "1x1 conv with upsample is not supported";
      return -2;
    }
    w3.assign((size_t)Cout * Cin * 9, 0.f);
    for (size_t i = 0; i < (size_t)Cout * Cin; ++i) w3[i * 9 + 4] = w_in[i];
    w = w3.data();
  } else if (ksize != 3) {
    err = "kernel size must be 1 or 3";
    return -2;
  }
  if (dil < 1 || dil > 8 || (dil != 1 && (up != 1 || ksize != 3))) {
    err = "dilation must be 1..8 and needs a plain 3x3 conv";
    return -2;
  }
  L.dil = dil;
  L.centre_only = ksize == 1;
  if (up < 1 || up > 3) {
    err = "unsupported upsample factor (1, 2 or 3 expected)";
    return -2;
  }
  L.Cin = Cin;
  L.Cout = Cout;
  L.Cin_pad = (Cin + 15) / 16 * 16;
  L.N = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
  if (Cout > 64) {
    err = "conv with more than 64 output channels is not supported by the tcgen05 kernel";
    return -2;
  }
  L.up = up;
  L.nphase = up * up;
  const int kslabs = L.Cin_pad / 16;
  const int N = L.N;

  // per-axis folding tables: for phase a, kernel index k -> halo offset h (0..2)
  int hof[3][3];
  for (int a = 0; a < up; ++a)
    for (int k = 0; k < 3; ++k) hof[a][k] = floordiv(a + k - 1, up) + 1;

  std::vector<__half> packed;
  L.max_taps = 0;
  for (int a = 0; a < up; ++a) {
    for (int b = 0; b < up; ++b) {
      const int ph = a * up + b;
      std::vector<int> hys, hxs;
      for (int k = 0; k < 3; ++k) {
        if (std::find(hys.begin(), hys.end(), hof[a][k]) == hys.end()) hys.push_back(hof[a][k]);
        if (std::find(hxs.begin(), hxs.end(), hof[b][k]) == hxs.end()) hxs.push_back(hof[b][k]);
      }
      std::sort(hys.begin(), hys.end());
      std::sort(hxs.begin(), hxs.end());
      if (ksize == 1) {
        hys.assign(1, 1);
        hxs.assign(1, 1);
      }
      const int ntaps = (int)(hys.size() * hxs.size());
      L.ph_ntaps[ph] = (uint8_t)ntaps;
      L.ph_a[ph] = (uint8_t)a;
      L.ph_b[ph] = (uint8_t)b;
      L.ph_woff[ph] = (uint32_t)(packed.size() * sizeof(__half));
      L.max_taps = std::max(L.max_taps, ntaps);
      int t = 0;
      for (int hy : hys)
        for (int hx : hxs) {
          L.tap_hy[ph][t] = (uint8_t)hy;
          L.tap_hx[ph][t] = (uint8_t)hx;
          ++t;
        }
      // folded weights for this phase: wf[tap][co][ci]
      std::vector<float> wf((size_t)ntaps * Cout * Cin, 0.f);
      t = 0;
      for (int hy : hys)
        for (int hx : hxs) {
          for (int ky = 0; ky < 3; ++ky) {
            if (hof[a][ky] != hy) continue;
            for (int kx = 0; kx < 3; ++kx) {
              if (hof[b][kx] != hx) continue;
              for (int co = 0; co < Cout; ++co)
                for (int ci = 0; ci < Cin; ++ci)
                  wf[((size_t)t * Cout + co) * Cin + ci] += w[(((size_t)co * Cin + ci) * 3 + ky) * 3 + kx];
            }
          }
          ++t;
        }
      const size_t base = packed.size();
      packed.resize(base + (size_t)kslabs * ntaps * 2 * N * 8);
      size_t o = base;
      for (int ks = 0; ks < kslabs; ++ks)
        for (int tp = 0; tp < ntaps; ++tp)
          for (int kc = 0; kc < 2; ++kc)
            for (int n = 0; n < N; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = ks * 16 + kc * 8 + e;
                float v = 0.f;
                if (n < Cout && ci < Cin) v = wf[((size_t)tp * Cout + n) * Cin + ci];
                packed[o++] = __float2half_rn(v);
              }
    }
  }
  // row-streaming packing (conv_rows.cu): [kslab][dx][kchunk][dy*Cout + co][8]
  std::vector<__half> packed_rows;
  if (up == 1 && ksize == 3 && (Cout == 32 || Cout == 64 || Cout <= 16)) {   // the packing does not depend on the dilation
    const int CR = Cout <= 16 ? 16 : Cout;   // output channels as the kernel sees them (zero rows beyond Cout)
    const int NR = 3 * CR;
    packed_rows.resize((size_t)kslabs * 3 * 2 * NR * 8);
    size_t o = 0;
    for (int ks = 0; ks < kslabs; ++ks)
      for (int dx = 0; dx < 3; ++dx)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < NR; ++n)
            for (int e = 0; e < 8; ++e) {
              const int ci = ks * 16 + kc * 8 + e, dy = n / CR, co = n % CR;
              const float v = (ci < Cin && co < Cout) ? w[(((size_t)co * Cin + ci) * 3 + dy) * 3 + dx] : 0.f;
              packed_rows[o++] = __float2half_rn(v);
            }
  }
  // CTA-pair packing: [half][kslab][dx][kchunk][row of that half][8]
  std::vector<__half> packed_pair;
  if (!packed_rows.empty() && Cout == 64 && Cin > 64) {
    const int NR = 3 * Cout, NH = NR / 2;
    packed_pair.resize(packed_rows.size() + 2 * (size_t)kRowsIdtBytes / sizeof(__half));
    size_t o = 0;
    for (int half = 0; half < 2; ++half) {
      for (int ks = 0; ks < kslabs; ++ks)
        for (int dx = 0; dx < 3; ++dx)
          for (int kc = 0; kc < 2; ++kc)
            for (int n = 0; n < NH; ++n)
              for (int e = 0; e < 8; ++e)
                packed_pair[o++] = packed_rows[((((size_t)ks * 3 + dx) * 2 + kc) * NR + half * NH + n) * 8 + e];
      // "identity" B tiles behind this half's weights (conv_rows.cu, IDT): four K slabs (input channels 0..63) x
      // [kchunk][32 output channels of this half][8], value 1 / 0.2 where ci == co -- the block's `0.2 * conv + x`
      // residual then comes out of four extra N = 64 MMAs into the centre-row block of the accumulator instead of
      // being loaded by the epilogue
      for (int ks = 0; ks < 4; ++ks)
        for (int kc = 0; kc < 2; ++kc)
          for (int n = 0; n < 32; ++n)
            for (int e = 0; e < 8; ++e)
              packed_pair[o++] = __float2half_rn((ks * 16 + kc * 8 + e == half * 32 + n) ? kRowsIdtScale : 0.f);
    }
  }
  L.w_bytes = packed.size() * sizeof(__half);
  std::vector<float> hb((size_t)L.nphase * N, 0.f);
  if (bias)
    for (int ph = 0; ph < L.nphase; ++ph)
      for (int i = 0; i < Cout; ++i) hb[(size_t)ph * N + i] = bias[i];
  L.h_w32.assign(w, w + (size_t)Cout * Cin * 9);
  L.h_b32.assign(hb.begin(), hb.begin() + Cout);
  L.pixel_shuffle = false;

  cudaError_t e;
  if ((e = cudaMalloc(&L.d_w, L.w_bytes)) != cudaSuccess ||
      (e = cudaMalloc(&L.d_bias, hb.size() * sizeof(float))) != cudaSuccess ||
      (e = cudaMemcpy(L.d_w, packed.data(), L.w_bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(L.d_bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (!packed_pair.empty() &&
       ((e = cudaMalloc(&L.d_wrows_pair, packed_pair.size() * sizeof(__half))) != cudaSuccess ||
        (e = cudaMemcpy(L.d_wrows_pair, packed_pair.data(), packed_pair.size() * sizeof(__half), cudaMemcpyHostToDevice)) !=
            cudaSuccess)) ||
      (!packed_rows.empty() &&
       ((e = cudaMalloc(&L.d_wrows, packed_rows.size() * sizeof(__half))) != cudaSuccess ||
        (e = cudaMemcpy(L.d_wrows, packed_rows.data(), packed_rows.size() * sizeof(__half), cudaMemcpyHostToDevice)) !=
            cudaSuccess))) {
    err = std::string("cuda error while uploading weights: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

int conv_layer_build_ps(ConvLayer& L, const float* w, const float* bias, int Cout, int Cin, int r,
                        std::string& err) {
  if (r < 2 || r > 3 || Cout > 64) {
    err = "pixel-shuffle conv: factor must be 2 or 3 and at most 64 channels per sub-pixel";
    return -2;
  }
  L.Cin = Cin;
  L.Cout = Cout;
  L.Cin_pad = (Cin + 15) / 16 * 16;
  L.N = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
  L.up = r;
  L.nphase = r * r;
  L.pixel_shuffle = true;
  L.max_taps = 9;
  const int N = L.N, kslabs = L.Cin_pad / 16;
  std::vector<__half> packed;
  std::vector<float> hb((size_t)L.nphase * N, 0.f);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < r; ++j) {
      const int ph = i * r + j;
      L.ph_ntaps[ph] = 9;
      L.ph_a[ph] = (uint8_t)i;
      L.ph_b[ph] = (uint8_t)j;
      L.ph_woff[ph] = (uint32_t)(packed.size() * sizeof(__half));
      for (int t = 0; t < 9; ++t) {
        L.tap_hy[ph][t] = (uint8_t)(t / 3);
        L.tap_hx[ph][t] = (uint8_t)(t % 3);
      }
      for (int ks = 0; ks < kslabs; ++ks)
        for (int tp = 0; tp < 9; ++tp)
          for (int kc = 0; kc < 2; ++kc)
            for (int n = 0; n < N; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = ks * 16 + kc * 8 + e;
                float v = 0.f;
                if (n < Cout && ci < Cin) v = w[(((size_t)(n * r * r + ph) * Cin + ci) * 9) + tp];
                packed.push_back(__float2half_rn(v));
              }
      if (bias)
        for (int n = 0; n < Cout; ++n) hb[(size_t)ph * N + n] = bias[n * r * r + ph];
    }
  L.w_bytes = packed.size() * sizeof(__half);
  L.h_w32.assign(w, w + (size_t)Cout * r * r * Cin * 9);
  L.h_b32.assign(r * r * Cout, 0.f);
  if (bias) L.h_b32.assign(bias, bias + (size_t)r * r * Cout);
  cudaError_t e;
  if ((e = cudaMalloc(&L.d_w, L.w_bytes)) != cudaSuccess ||
      (e = cudaMalloc(&L.d_bias, hb.size() * sizeof(float))) != cudaSuccess ||
      (e = cudaMemcpy(L.d_w, packed.data(), L.w_bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(L.d_bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
    err = std::string("cuda error while uploading weights: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

void conv_layer_free(ConvLayer& L) {
  if (L.d_w) cudaFree(L.d_w);
  if (L.d_bias) cudaFree(L.d_bias);
  if (L.d_w32) cudaFree(L.d_w32);
  if (L.d_wrows) cudaFree(L.d_wrows);
  if (L.d_wrows_pair) cudaFree(L.d_wrows_pair);
  L.d_wrows_pair = nullptr;
  L.d_wrows = nullptr;
  L.d_w = nullptr;
  L.d_bias = nullptr;
  L.d_w32 = nullptr;
}

const CUtensorMap* TmapCache::get(const void* base, int B, int CT, int H, int W, int box_w, int box_h, int& rc) {
  auto key = std::make_tuple(base, B, CT, H, W, box_w | (box_h << 16));
  auto it = maps_.find(key);
  rc = 0;
  if (it != maps_.end()) return &it->second->m;
  Slot* s = new Slot();
  rc = encode_act_tmap(&s->m, base, B, CT, H, W, box_w, box_h);
  if (rc != 0) {
    delete s;
    return nullptr;
  }
  maps_[key] = s;
  return &s->m;
}

const CUtensorMap* TmapCache::get_rows(const void* base, int CT, int H, int Wtot, int kc, int& rc) {
  auto key = std::make_tuple(base, -1, CT, H, Wtot, -1000 - kc);
  auto it = maps_.find(key);
  rc = 0;
  if (it != maps_.end()) return &it->second->m;
  Slot* s = new Slot();
  rc = encode_wide_rows_tmap(&s->m, base, CT, H, Wtot, kc);
  if (rc != 0) {
    delete s;
    return nullptr;
  }
  maps_[key] = s;
  return &s->m;
}

// Row-streaming kernel on wide tensors; returns -100 when the conv is not eligible.
static int conv_rows_run(const ConvLayer& L, TmapCache& cache, ChunkView in, int B, int H, int W, ChunkView out,
                         int out_nchunks, const Epilogue& ep, int num_sms, cudaStream_t stream) {
  static const int rows_mode = getenv("INNFER_ROWS") ? atoi(getenv("INNFER_ROWS")) : 7;  // bit 0: on, bit 1: Cout = 64 too, bit 2: CTA pairs (conv5)
  const int CR = L.Cout <= 16 ? 16 : L.Cout;   // kernel instantiation: 16 (the net's last conv), 32 or 64
  // dilated variant (PPON): 64 -> 32, LeakyReLU after the res1 add, optional raw second store; the separator between
  // the images must cover the reach of the taps.  INNFER_ROWS bit 3 (default on) enables it.
  const bool dilv = L.dil != 1 || ep.act_after_res || ep.raw_out.base;
  if (dilv) {
    static const int dil_mode = getenv("INNFER_ROWS_DIL") ? atoi(getenv("INNFER_ROWS_DIL")) : 1;
    if (!dil_mode || !in.wide() || CR != 32 || L.Cin_pad != 64 || !ep.act_after_res || ep.res2.base || ep.compact4 ||
        !out.wide() || in.pitch - W < L.dil || (ep.res1.base && ep.alpha1 != 1.f))
      return -100;
    if (ep.raw_out.base && (!ep.raw_out.wide() || ep.raw_out.pitch != in.pitch || ep.raw_out.Wtot != in.Wtot)) return -100;
  }
  if (ep.res1_unact && !dilv) return -100;
  if (!rows_mode || !in.wide() || L.d_wrows == nullptr || L.up != 1 || ep.gate || ep.self_gate ||
      out_nchunks != (L.Cout + 7) / 8 || (out.wide() && (in.pitch != out.pitch || in.Wtot != out.Wtot)) ||
      (ep.compact4 && (out.wide() || L.Cout > 4)))
    return -100;
  for (const ChunkView* v : {&ep.res1, &ep.res2})
    if (v->base && (!v->wide() || v->pitch != in.pitch || v->Wtot != in.Wtot)) return -100;
  // Cout = 64 (N = 192 MMAs run at ~117 cycles, only two TMEM slots): 1.3x faster than the 9-tap kernel for
  // residual-free convs (HR_conv0: 3275 -> 2520 us), no gain with residual epilogues unless on a CTA pair
  if (CR != 16 && CR != 32 && CR != 64) return -100;
  if (CR == 64 && !(rows_mode & 2)) return -100;
  if (CR == 16 && (!(rows_mode & 2) || L.Cin_pad != 64)) return -100;
  const int nch = L.Cin_pad / 8;
  int nsub = (nch + 15) / 16;
  while (nsub <= nch && (nch % nsub != 0 || ((nch / nsub) & 1))) ++nsub;
  if (nsub > nch) return -100;
  const int kc = nch / nsub;
  int wbytes = conv_rows_weight_bytes(nch, CR);
  int S = (232448 - 1024 - wbytes) / conv_rows_stage_bytes(kc);
  // conv5 of the nf = 64 net (192 -> 64): 221 KB of weights only fit when a CTA pair shares them
  // (pairs for the short row stages of conv1..conv4 were measured 2.5x slower: the cross-CTA signalling per row is not
  // amortised over 12..30 MMAs)
  const bool pair = S < 3 && L.d_wrows_pair != nullptr && (rows_mode & 4) && CR == 64 && kc == 12;
  if (pair) {
    wbytes = wbytes / 2 + kRowsIdtBytes;
    S = (232448 - 1024 - wbytes) / conv_rows_stage_bytes(kc);
  }
  if (S < 3) return -100;
  if (S > 8) S = 8;
  // Cout = 64 with ONE residual input on the row kernel (round 2: SRResNet's block convs 53.3 -> 49.4 ms per 1080p frame
  // against the 9-tap kernel; INNFER_ROWS64_RES=0 restores that); two residuals stay on the CTA pair / 9-tap kernel
  static const int rows64res = getenv("INNFER_ROWS64_RES") ? atoi(getenv("INNFER_ROWS64_RES")) : 1;
  if (CR == 64 && !pair && (ep.res1.base || ep.res2.base) && !(rows64res && !ep.res2.base)) return -100;
  ConvRowsParams p;
  std::memset(&p, 0, sizeof(p));
  p.H = H;
  p.Wtot = in.Wtot;
  p.nimg = B;
  p.pitch = in.pitch;
  p.Wimg = W;
  p.magic = (uint32_t)((1ull << 32) / (unsigned)in.pitch) + 1u;
  p.in_chunk0 = in.chunk0;
  p.nch = nch;
  p.kc = kc;
  p.nsub = nsub;
  p.nstrips = (in.Wtot + 15 + 127) / 128;
  p.stages = S;
  p.out = out.base;
  p.out_px = 8;
  if (out.wide()) {
    p.out_bs = (long long)out.pitch * 8;
    p.out_cs = (long long)H * out.Wtot * 8;
    p.out_ys = out.Wtot * 8;
    p.out_wide = 1;
  } else if (ep.compact4) {
    p.out_bs = (long long)H * W * 4;
    p.out_ys = W * 4;
    p.out_px = 4;
    p.out_compact4 = 1;
  } else {
    p.out_bs = (long long)out.CT * H * W * 8;
    p.out_cs = (long long)H * W * 8;
    p.out_ys = W * 8;
  }
  p.out_chunk0 = out.chunk0;
  p.out_nchunks = out_nchunks;
  p.res_cs = (long long)H * in.Wtot * 8;
  p.res_ys = in.Wtot * 8;
  p.w = pair ? L.d_wrows_pair : L.d_wrows;
  p.pair = pair ? 1 : 0;
  static const int pdl = getenv("INNFER_PDL") ? atoi(getenv("INNFER_PDL")) : 1;   // A/B switch
  p.pdl = pdl;
  p.bias = L.d_bias;
  p.lrelu = ep.lrelu ? 1 : 0;
  p.slope = ep.slope;
  p.res1 = ep.res1.base;
  p.res1_chunk0 = ep.res1.chunk0;
  p.alpha1 = ep.alpha1;
  p.res2 = ep.res2.base;
  p.res2_chunk0 = ep.res2.chunk0;
  p.alpha2 = ep.alpha2;
  // conv5 of a dense block on a CTA pair: res1 is the block's input x = the first 64 channels of the conv's own input,
  // already staged in shared memory for the MMAs -- `alpha1 * conv + x` comes out of four identity MMAs per row
  // (x * (1 / alpha1) into the accumulator) and the epilogue neither loads nor waits for it.  INNFER_ROWS_IDT=0 keeps
  // the loaded residual (A/B switch).
  static const int idt_mode = getenv("INNFER_ROWS_IDT") ? atoi(getenv("INNFER_ROWS_IDT")) : 1;
  if (idt_mode && pair && !ep.lrelu && ep.res1.base == in.base && ep.res1.chunk0 == in.chunk0 && ep.res1.CT == in.CT &&
      L.Cout == 64 && std::fabs(ep.alpha1 * kRowsIdtScale - 1.f) < 1e-6f) {
    p.res1 = nullptr;
    p.idt = 1;
  }
  p.dil = L.dil;
  p.act_after_res = ep.act_after_res ? 1 : 0;
  p.res1_unact = ep.res1_unact ? 1 : 0;
  p.raw = ep.raw_out.base;
  p.raw_chunk0 = ep.raw_out.chunk0;
  static const int trace_nch = getenv("INNFER_TRACE_NCH") ? atoi(getenv("INNFER_TRACE_NCH")) : 0;
  p.trace = (trace_nch == nch) ? g_rows_trace : nullptr;
#ifdef INNFER_EXPERIMENTS   // timing experiment with wrong results: special builds only (INNFER_EXPERIMENTS_BUILD=1)
  static const int dbg_dx0 = getenv("INNFER_ROWS_DX0") ? atoi(getenv("INNFER_ROWS_DX0")) : 0;
  p.dbg_dx0 = dbg_dx0;
#endif
  int rc = 0;
  const CUtensorMap* tm = cache.get_rows(in.base, in.CT, H, in.Wtot, kc, rc);
  if (!tm) return rc ? rc : -5;
  g_last_conv_kernel = pair ? "conv_rows_pair" : (dilv ? "conv_rows_dil" : "conv_rows");
  return launch_conv_rows(tm, p, CR, num_sms, stream);
}

int choose_J(int W, int N) {
  // Alternate tiles use alternate sets of J accumulators (double buffering against the epilogue),
  // so 2*J*N fp32 columns must fit the 512 TMEM columns.  Among the feasible J pick the one that
  // loads the fewest halo columns, charging a few columns per tile for fixed per-tile costs.
  int jmax = 256 / N;
  if (jmax > 5) jmax = 5;
  int best = 1;
  long best_cost = -1;
  for (int J = 1; J <= jmax; ++J) {
    const int cps = (W + 8 * J - 1) / (8 * J);
    const long cost = (long)cps * (8 * J + 2 + 6);
    if (best_cost < 0 || cost <= best_cost) {
      best_cost = cost;
      best = J;
    }
  }
  return best;
}

int conv_layer_run(const ConvLayer& L, TmapCache& cache, ChunkView in, int B, int H, int W,
                   ChunkView out, int out_nchunks, const Epilogue& ep, int num_sms,
                   cudaStream_t stream) {
  ConvTcParams p;
  std::memset(&p, 0, sizeof(p));
  const int N = L.N;
  {
    const int rr = conv_rows_run(L, cache, in, B, H, W, out, out_nchunks, ep, num_sms, stream);
    if (rr != -100) return rr;
  }
  // source geometry as the kernel tiles it: B images of width W, or one wide image
  const int srcB = in.wide() ? 1 : B;
  const int srcW = in.wide() ? in.Wtot : W;
  const int J = choose_J(srcW, N);
  p.B = srcB;
  p.H = H;
  p.W = srcW;
  p.in_chunk0 = in.chunk0;
  p.kslabs = L.Cin_pad / 16;
  p.J = J;
  p.bands = (H + kPatchRows - 1) / kPatchRows;
  p.cps = (srcW + 8 * J - 1) / (8 * J);
  if (in.wide()) {
    p.sep_pitch = in.pitch;
    p.sep_w = W;
    p.sep_nimg = B;
    p.sep_magic = (uint32_t)((1ull << 32) / (unsigned)in.pitch) + 1u;
    p.out_zero_sep = out.wide() ? 1 : 0;
    if (out.wide() && (out.pitch != in.pitch * L.up || out.Wtot != in.Wtot * L.up)) return -6;
  }
  {
    const long long Ho = (long long)H * L.up, Wo = (long long)W * L.up;
    auto strides = [&](const ChunkView& v, long long& bs, long long& cs, int& ys) {
      if (v.wide()) {
        bs = (long long)v.pitch * 8;
        cs = Ho * v.Wtot * 8;
        ys = v.Wtot * 8;
      } else {
        bs = (long long)v.CT * Ho * Wo * 8;
        cs = Ho * Wo * 8;
        ys = (int)(Wo * 8);
      }
    };
    strides(out, p.out_bs, p.out_cs, p.out_ys);
    p.out_px = 8;
    if (ep.compact4) {
      p.out_bs = Ho * Wo * 4;
      p.out_cs = 0;
      p.out_ys = (int)(Wo * 4);
      p.out_px = 4;
    }
    if (ep.raw_out.base) {
      if (!ep.act_after_res || ep.compact4) return -8;
      strides(ep.raw_out, p.raw_bs, p.raw_cs, p.raw_ys);
      p.raw = ep.raw_out.base;
      p.raw_chunk0 = ep.raw_out.chunk0;
    }
    p.act_after_res = ep.act_after_res ? 1 : 0;
    p.gate = ep.gate ? 1 : (ep.self_gate ? 2 : 0);
    p.res1_unact = ep.res1_unact ? 1 : 0;
    if (ep.gate && !ep.res1.base) return -9;
    if (ep.self_gate && (ep.gate || out_nchunks > N / 16 || L.nphase != 1)) return -9;
    if (ep.res1.base) strides(ep.res1, p.res1_bs, p.res1_cs, p.res1_ys);
    if (ep.res2.base) strides(ep.res2, p.res2_bs, p.res2_cs, p.res2_ys);
  }
  p.nphase = L.nphase;
  p.up = L.up;
  p.Hout = H * L.up;
  p.Wout = W * L.up;
  p.nslots = 2 * J;
  int cols = p.nslots * N, pw = 32;
  while (pw < cols) pw <<= 1;
  p.tmem_cols = pw;
  // 1x1 convs (one centre tap): "dilation 0" puts that tap at offset 0 of a tile without halo -- 16 x 8J pixels per
  // stage instead of 18 x (8J + 2) (PPON's 256 -> 64 fusion conv: 2.57 -> 2.0 GB of DRAM traffic per 63 tiles);
  // INNFER_1X1_HALO=1 keeps the halo tile (A/B switch)
  static const int halo_1x1 = getenv("INNFER_1X1_HALO") ? atoi(getenv("INNFER_1X1_HALO")) : 0;
  const int tdil = (L.centre_only && !halo_1x1) ? 0 : L.dil;
  p.dil = tdil;
  static const int pdl = getenv("INNFER_PDL") ? atoi(getenv("INNFER_PDL")) : 1;   // A/B switch, as for conv_rows
  p.pdl = pdl;
  const int stage_bytes = conv_tc_a_bytes(J, tdil) + conv_tc_w_bytes(N, L.max_taps);
  int S = (232448 - kConvTailBytes) / stage_bytes;
  if (S > 8) S = 8;
  if (S < 2) return -4;
  p.stages = S;
  p.out = out.base;
  p.out_CT = out.CT;
  p.out_chunk0 = out.chunk0;
  p.out_nchunks = out_nchunks;
  p.out_compact4 = ep.compact4 ? 1 : 0;
  p.w = L.d_w;
  p.bias = L.d_bias;
  p.lrelu = ep.lrelu ? 1 : 0;
  p.slope = ep.slope;
  p.res1 = ep.res1.base;
  p.res1_CT = ep.res1.CT;
  p.res1_chunk0 = ep.res1.chunk0;
  p.alpha1 = ep.alpha1;
  p.res2 = ep.res2.base;
  p.res2_CT = ep.res2.CT;
  p.res2_chunk0 = ep.res2.chunk0;
  p.alpha2 = ep.alpha2;
  for (int i = 0; i < L.nphase; ++i) {
    p.ph_woff[i] = L.ph_woff[i];
    p.ph_ntaps[i] = L.ph_ntaps[i];
    p.ph_a[i] = L.ph_a[i];
    p.ph_b[i] = L.ph_b[i];
    for (int t = 0; t < kMaxTaps; ++t) {
      p.tap_hy[i][t] = L.tap_hy[i][t];
      p.tap_hx[i][t] = L.tap_hx[i][t];
    }
  }
  int rc = 0;
  {
    // x2 upconv: all four phases per tile visit with resident weights (conv_up.cu) when they fit
    static const int up_mode = getenv("INNFER_UP") ? atoi(getenv("INNFER_UP")) : 1;
    bool four = L.up == 2 && !L.pixel_shuffle && N == 64 && L.nphase == 4;
    for (int i = 0; i < 4 && four; ++i) four = L.ph_ntaps[i] == 4;
    const int wb = conv_up_weight_bytes(p.kslabs);
    int Su = (232448 - kConvTailBytes - wb) / conv_up_stage_bytes();
    if (up_mode && four && !ep.res1.base && !ep.res2.base && !ep.compact4 && !ep.raw_out.base && Su >= 4) {
      p.J = 1;
      p.trace = g_rows_trace;
      p.cps = (srcW + 7) / 8;
      p.stages = Su > 12 ? 12 : Su;
      const CUtensorMap* tmu = cache.get(in.base, srcB, in.CT, H, srcW, 10, kPatchRows + 2, rc);
      if (!tmu) return rc ? rc : -5;
      g_last_conv_kernel = "conv_up";
      return launch_conv_up(tmu, p, num_sms, stream);
    }
  }
  const CUtensorMap* tm = cache.get(in.base, srcB, in.CT, H, srcW, 8 * J + 2 * tdil, kPatchRows + 2 * tdil, rc);
  if (!tm) return rc ? rc : -5;
  g_last_conv_kernel = "conv_tc";
  return launch_conv_tc(tm, p, N, num_sms, stream);
}

}  // namespace innfer
