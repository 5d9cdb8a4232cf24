// Image-to-image generators of the reference (BASELINE configs[4]): pix2pix's UnetGenerator
// (architectures/UNet_arch.py:11-165) and CycleGAN's ResnetGenerator (architectures/ResNet_arch.py:11-151) on planar-chunk
// tensors [B][C/8][H][W][8].  New op families compared with the SR nets: k x k convolutions with stride, transposed
// convolutions, reflection padding, BatchNorm with batch statistics / InstanceNorm, tanh.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

namespace innfer {

constexpr int kGenMaxTaps = 49;   // 7 x 7
constexpr int kGenMaxPhases = 4;  // stride-2 transposed convolutions: one phase per output parity class

enum GenAct { kActNone = 0, kActRelu = 1, kActLrelu = 2, kActTanh = 3 };

// One convolution or transposed convolution with its packed weights.
struct GenConv {
  int Cin = 0, Cout = 0, k = 0, stride = 1, pad = 0, out_pad = 0;
  bool transposed = false, reflect = false, has_bias = false;
  int cin_chunks = 0;   // ceil(Cin / 8)
  int NT = 0;           // output channels per CTA tile: 16, 32, 64 or 128
  int ntiles = 0;       // ceil(Cout / NT)
  // phases: output pixels (oy' * ostep + py, ox' * ostep + px) read input (oy' * istep + offy[t], ox' * istep + offx[t])
  int nphase = 1, ostep = 1, istep = 1;
  int ph_py[kGenMaxPhases] = {}, ph_px[kGenMaxPhases] = {}, ph_ntaps[kGenMaxPhases] = {};
  int8_t offy[kGenMaxPhases][kGenMaxTaps] = {}, offx[kGenMaxPhases][kGenMaxTaps] = {};
  size_t ph_woff[kGenMaxPhases] = {};   // byte offset of the phase's packed fp16 weights
  int ph_ksteps[kGenMaxPhases] = {};
  __half* d_w16 = nullptr;   // [phase][ntile][kstep][8 chunks][NT][8] fp16
  float* d_w32 = nullptr;    // [ky][kx][Cin][Cout_pad8] fp32 (direct kernel)
  float* d_bias = nullptr;   // [ntiles * NT] (zeros when the layer has no bias)
  int cout_pad8 = 0;
};

struct GenView {   // `nch` chunks starting at chunk0 of a [B][CT][H][W][8] tensor
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
};

// key lookup into the loaded state dict: returns the data and fills `shape`, or nullptr
using ParamLookup = std::function<const float*(const std::string& key, std::vector<int64_t>& shape)>;

class I2INet {
 public:
  explicit I2INet(const I2ICfg& cfg) : cfg_(cfg) {}
  ~I2INet();
  // builds every layer from the reference-named parameters; returns 0 or a negative INNFER_E_* style code
  // (-1 invalid, -2 unsupported, -3 cuda, -4 missing key); `expected` receives the number of keys consumed
  int build(const ParamLookup& get, std::string& err, size_t* expected);
  int ensure(int B, int H, int W, std::string& err);
  // in: [B][in_CT][H][W][8]; out: `out_nchunks` chunks at out.chunk0 of [B][out.CT][Ho][Wo][8], or compact [B][Ho][Wo][4]
  int forward(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, cudaStream_t st, int* launches,
              std::string& err);
  int out_h(int H) const;
  int out_w(int W) const;

 private:
  struct Norm {
    int C = 0;
    float* d_gamma = nullptr;   // affine (BatchNorm) or null
    float* d_beta = nullptr;
    float* d_scale = nullptr;   // eval-mode BatchNorm: precomputed scale / shift
    float* d_shift = nullptr;
  };
  struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
  };
  int build_conv(const ParamLookup& get, const std::string& name, GenConv& L, int Cout, int Cin, int k, int stride, int pad,
                 bool transposed, int out_pad, bool reflect, bool has_bias, std::string& err);
  int build_norm(const ParamLookup& get, const std::string& name, Norm& n, int C, std::string& err);
  int need(Buf& b, size_t bytes);
  int run_conv(const GenConv& L, GenView in, int B, int Hin, int Win, GenView out, int Hout, int Wout, int pre_act, int post_act,
               bool compact4, cudaStream_t st);
  // y = act((x - mean) * rstd * gamma + beta) [+ res], statistics per (b, c) (instance) or per c (batch)
  int run_norm(const Norm& n, GenView x, int B, int H, int W, int act, GenView res, GenView out, cudaStream_t st);
  int forward_unet(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, cudaStream_t st);
  int forward_resnet(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, cudaStream_t st);
  size_t esz() const { return cfg_.fp16 ? 2 : 4; }

  I2ICfg cfg_;
  std::vector<GenConv> down_, up_;      // UNet: per level, outermost first; ResNet: all convs in order in down_
  std::vector<Norm> dnorm_, unorm_;     // UNet; ResNet: norms in order in dnorm_
  std::vector<int> chan_;               // UNet: channels of x_i per level
  std::vector<Buf> cat_;                // UNet: concat buffer per level
  Buf tmp_, inner_, a_, b_, c_;         // raw conv outputs / ResNet ping-pong
  Buf stats_, scale_;                   // norm partial sums and per-(b, c) scale / shift
  int launches_ = 0;
  std::vector<void*> owned_;            // device allocations of the layers
};

}  // namespace innfer
