// Image-to-image generators of the reference (BASELINE configs[4]): pix2pix's UnetGenerator
// (architectures/UNet_arch.py:11-165) and CycleGAN's ResnetGenerator (architectures/ResNet_arch.py:11-151) on planar-chunk
// tensors [B][C/8][H][W][8].  New op families compared with the SR nets: k x k convolutions with stride, transposed
// convolutions, reflection padding, BatchNorm with batch statistics / InstanceNorm, tanh.
//
// Every convolution is one implicit GEMM (csrc/i2i.cu: gen_conv_tc_kernel): M = 128 output pixels of one output-parity
// class ("phase": a transposed convolution with stride s is s*s ordinary convolutions with sub-sampled taps), N = up to
// 128 output channels, K = (tap, 8-channel chunk) pairs in steps of 64.  The A operand cannot be a TMA box (stride-2
// taps, reflected borders), so four warps gather it with 16-byte cp.async copies straight from the planar-chunk tensor
// into the SWIZZLE_NONE K-major layout tcgen05.mma reads -- one 16-byte pixel chunk IS one core-matrix row -- and the
// packed weights of the K-step arrive by one bulk copy; accumulators live in TMEM.  Layers with little M (the inner
// levels of the UNet: 1x1 .. 8x8 pixels, 8192 x 512 weights) split K across CTAs into fp32 partial tensors that the
// normalisation kernels sum in a fixed order, so results stay bit-reproducible.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace innfer {

constexpr int kGenMaxTaps = 49;   // 7 x 7
constexpr int kGenMaxPhases = 4;  // stride-2 transposed convolutions: one phase per output parity class
constexpr int kGenMaxSplit = 16;  // split-K partial tensors

enum GenAct { kActNone = 0, kActRelu = 1, kActLrelu = 2, kActTanh = 3, kActTanh01 = 4 /* (tanh(v) + 1) / 2 */ };

// One convolution or transposed convolution with its packed weights.
struct GenConv {
  int Cin = 0, Cout = 0, k = 0, stride = 1, pad = 0, out_pad = 0;
  bool transposed = false, reflect = false;
  int cin_chunks = 0;   // ceil(Cin / 8)
  int cout_chunks = 0;  // ceil(Cout / 8)
  int NT = 0;           // output channels per CTA tile: 16, 32, 64 or 128
  int ntiles = 0;       // ceil(Cout / NT)
  // phases: output pixel (oy' * ostep + py, ox' * ostep + px) reads input (oy' * istep + offy[t], ox' * istep + offx[t])
  int nphase = 1, ostep = 1, istep = 1;
  int ph_py[kGenMaxPhases] = {}, ph_px[kGenMaxPhases] = {}, ph_ntaps[kGenMaxPhases] = {};
  int8_t offy[kGenMaxPhases][kGenMaxTaps] = {}, offx[kGenMaxPhases][kGenMaxTaps] = {};
  int ph_ksteps[kGenMaxPhases] = {};        // ceil(ntaps * cin_chunks / 8)
  size_t ph_woff16[kGenMaxPhases] = {};     // element offset of the phase inside d_w16
  size_t ph_woff32[kGenMaxPhases] = {};     // element offset of the phase inside d_w32
  __half* d_w16 = nullptr;   // [phase][ntile][kstep][8 K-chunks][NT][8] fp16 (tensor-core kernel)
  float* d_w32 = nullptr;    // [phase][tap][cin_chunks * 8][cout_chunks * 8] fp32 (direct kernel, -no_fp16 mode)
  float* d_bias = nullptr;   // [ntiles * NT] (zeros when the layer has no bias)
  // halo-tile kernel (gen_conv_halo_kernel): per phase the taps address up to four parity planes of one gathered input
  // tile; plane p holds input pixels (istep * (y0 + r + pl_r0) + pl_py, istep * (x0 + c + pl_c0) + pl_px)
  int NTh = 0, ntiles_h = 0, nslabs = 0;    // N tile (bounded by the shared-memory stage), ceil(cin_chunks / 2)
  int h_nplanes[kGenMaxPhases] = {}, h_rext[kGenMaxPhases] = {}, h_cext[kGenMaxPhases] = {};
  int8_t h_pl_py[kGenMaxPhases][4] = {}, h_pl_px[kGenMaxPhases][4] = {}, h_pl_r0[kGenMaxPhases][4] = {}, h_pl_c0[kGenMaxPhases][4] = {};
  int8_t h_tap_pl[kGenMaxPhases][kGenMaxTaps] = {}, h_tap_dr[kGenMaxPhases][kGenMaxTaps] = {}, h_tap_dc[kGenMaxPhases][kGenMaxTaps] = {};
  size_t ph_woffh[kGenMaxPhases] = {};
  __half* d_w16h = nullptr;  // [phase][ntile][16-channel slab][tap][2 K-chunks][NTh][8] fp16
  int out_h(int Hin) const {
    return transposed ? (Hin - 1) * stride - 2 * pad + k + out_pad : (Hin + 2 * pad - k) / stride + 1;
  }
};

struct GenView {   // chunks starting at chunk0 of a [B][CT][H][W][8] tensor (fp16 or fp32)
  void* base = nullptr;
  int CT = 0, chunk0 = 0;
};

struct I2ICfg {
  int kind = 0;       // 0: UnetGenerator, 1: ResnetGenerator
  int in_nc = 3, out_nc = 3, ngf = 64;
  int depth = 8;      // num_downs (UNet) or n_blocks (ResNet)
  int norm = 0;       // 0: BatchNorm2d, 1: InstanceNorm2d
  int train = 0;      // BatchNorm: 1 = statistics of the batch (module in training mode, run.py:297), 0 = running statistics
  int fp16 = 1;
  // images in [0, 1] at both ends: the [-1, 1] normalisation run.py applies around these networks (np2tensor(normalize),
  // tensor2np(denormalize); utils.py:136-161) folded into the network -- first conv on 2x - 1 == conv with doubled
  // weights and bias - sum(w) (exact with reflection padding: every tap sees a real pixel), last layer (tanh + 1) / 2.
  // ResnetGenerator only (the UNet's first conv pads with zeros, which are 0.5 in image units).
  int unit_io = 0;
  // kind 2: ONE convolution [+ norm] [+ activation] under the keys "conv" / "norm" -- the operator-level entry point
  // innfer_gen_conv of the parity tests (same kernels and code paths as the networks)
  int sl_cout = 0, sl_k = 3, sl_stride = 1, sl_pad = 1, sl_transposed = 0, sl_out_pad = 0, sl_reflect = 0, sl_bias = 1;
  int sl_norm = 0;    // 1: followed by the norm layer `norm` selects
  int sl_act = 0;     // GenAct
  int sl_final = 0;   // 1: bias + activation in the conv epilogue (the path of a network's last layer), no norm
};

// key lookup into the loaded state dict: returns the data and fills `shape`, or nullptr
using ParamLookup = std::function<const float*(const std::string& key, std::vector<int64_t>& shape)>;

uint64_t i2i_graph_replays();   // forwards of the generators served by replaying a recorded CUDA graph (this process)
uint64_t i2i_halo_launches();   // launches of the halo-tile kernel by this process (tests: which kernel served a layer)

class I2INet {
 public:
  I2INet(const I2ICfg& cfg, int num_sms) : cfg_(cfg), num_sms_(num_sms) {}
  ~I2INet();
  // builds every layer from the reference-named parameters; returns 0 or -1 invalid / -2 unsupported / -3 cuda /
  // -4 missing key / -5 out of memory; `consumed` receives the number of state-dict entries used
  int build(const ParamLookup& get, std::string& err, size_t* consumed);
  // in: [B][in_CT][H][W][8]; out: chunks at out.chunk0 of [B][out.CT][Ho][Wo][8], or compact [B][Ho][Wo][4] fp16.
  // per_sample: every image of the batch is a separate call of the reference (the tile loop of run.py:187-197), so
  // train-mode BatchNorm takes its statistics per image instead of over the batch.
  int forward(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample, cudaStream_t st,
              std::string& err);
  // 0 when an H x W input gives an H x W output (UNet: multiples of 2^depth; ResNet: multiples of 4)
  int check_size(int H, int W, std::string& err) const;
  int out_size(int n) const { return cfg_.kind == 2 ? down_[0].out_h(n) : n; }
  uint64_t launches() const { return launches_; }
  uint64_t graph_replays() const { return graph_replays_; }   // forwards served by replaying a recorded CUDA graph

 private:
  struct Norm {
    int C = 0;
    bool affine = false, running = false;
    float* d_gamma = nullptr;   // affine (BatchNorm)
    float* d_beta = nullptr;
    float* d_ss = nullptr;      // eval-mode BatchNorm: precomputed (scale, shift) per channel
  };
  struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
  };
  struct Raw {   // fp32 output of a convolution before normalisation: nsplit partial tensors [B][CT][H][W][8]
    float* p = nullptr;
    int nsplit = 1;
    size_t split_stride = 0;   // elements
    int stat_slices = 0;       // > 0: the conv kernel left per-tile partial sums of the norm statistics in stats_
  };
  int build_conv(const ParamLookup& get, const std::string& name, GenConv& L, int Cout, int Cin, int k, int stride, int pad,
                 bool transposed, int out_pad, bool reflect, bool has_bias, std::string& err, size_t* consumed);
  int build_norm(const ParamLookup& get, const std::string& name, Norm& n, int C, std::string& err, size_t* consumed);
  int need(Buf& b, size_t bytes);
  // convolution into the raw fp32 scratch (split-K allowed)
  int conv_raw(const GenConv& L, GenView in, int B, int Hin, int Win, Raw& raw, cudaStream_t st, bool want_stats = false);
  // convolution with bias + activation straight into a tensor of the network's precision (the last layer)
  int conv_final(const GenConv& L, GenView in, int B, int Hin, int Win, GenView out, int act, bool compact4, cudaStream_t st);
  // out_a = act_a(norm(raw) [+ res]), optionally out_b = act_b(same); norm == nullptr: identity
  int norm_apply(const Norm* n, const Raw& raw, int B, int C, int H, int W, bool per_sample, int act_a, GenView out_a,
                 int act_b, const GenView* out_b, const GenView* res, cudaStream_t st);
  int forward_unet(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample, cudaStream_t st);
  int forward_resnet(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample, cudaStream_t st);
  int forward_eager(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample, cudaStream_t st);
  void drop_graphs();
  size_t esz() const { return cfg_.fp16 ? 2 : 4; }
  GenView view(Buf& b, int CT, int chunk0) const { return GenView{b.p, CT, chunk0}; }
  // CUDA graphs of whole forwards, keyed by everything the recorded launches depend on
  struct GraphKey {
    const void* in;
    void* out;
    int in_CT, B, H, W, out_CT, out_chunk0;
    bool compact4, per_sample;
    bool operator<(const GraphKey& o) const {
      return std::tie(in, out, in_CT, B, H, W, out_CT, out_chunk0, compact4, per_sample) <
             std::tie(o.in, o.out, o.in_CT, o.B, o.H, o.W, o.out_CT, o.out_chunk0, o.compact4, o.per_sample);
    }
  };
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;
    bool failed = false;
  };
  std::map<GraphKey, GraphEntry> graphs_;
  cudaStream_t cap_ = nullptr;
  bool capturing_ = false;
  uint64_t graph_replays_ = 0;

  I2ICfg cfg_;
  int num_sms_ = 148;
  std::vector<GenConv> down_, up_;      // UNet: per level, outermost first; ResNet: all convolutions in order in down_
  std::vector<Norm> dnorm_, unorm_;     // UNet; ResNet: norms in order in dnorm_
  std::vector<Buf> dbuf_, cat_;         // UNet: lrelu(x_i) per level, concat buffer per level
  Buf raw_, inner_, a_, b_;             // raw conv outputs / UNet bottleneck / ResNet full- and half-resolution tensors
  std::vector<Buf> rb_;                 // ResNet: residual stream ping-pong + block scratch
  Buf stats_, ss_;                      // norm partial sums and per-(image, channel) scale / shift
  uint64_t launches_ = 0;
  std::vector<void*> owned_;            // device allocations of the layers
  std::string err_;
};

}  // namespace innfer
