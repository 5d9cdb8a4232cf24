// Host-side description of one fused convolution "layer" of the RRDB path and the helper that
// launches it on planar-chunk tensors.  One ConvLayer corresponds to one reference conv_block
// (architectures/block.py:213-254) optionally preceded by the nearest-neighbour Upsample of
// upconv_block (block.py:286-361), which is folded into per-output-phase 2x2 weight sets.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "conv_tc.cuh"

namespace innfer {

// A view of `nchunks` channel chunks inside a planar-chunk buffer [B][CT][H][W][8] fp16.
struct ChunkView {
  __half* base = nullptr;  // buffer start
  int CT = 0;              // chunks per tile in the buffer
  int chunk0 = 0;          // first chunk of the view
  // wide layout (fp16 engine path): the buffer is ONE image [CT][H][Wtot][8] in which tile image b
  // occupies columns [b*pitch, b*pitch + W) and all other columns (separators, padding of Wtot to a
  // multiple of 16) hold zeros.  pitch == 0: tiled layout [B][CT][H][W][8].
  int pitch = 0;
  int Wtot = 0;
  bool wide() const { return pitch != 0; }
};

struct ConvLayer {
  int Cin = 0, Cout = 0;   // real channel counts
  int Cin_pad = 0;         // multiple of 16
  int N = 0;               // padded Cout: 16, 32 or 64
  int up = 1;              // nearest-upsample factor folded in front of the conv (or PixelShuffle factor)
  int dil = 1;             // dilation (plain 3x3 convs only)
  bool centre_only = false;  // 1x1 conv: the only tap is the centre one, the tensor-core kernel needs no halo
  bool pixel_shuffle = false;  // conv -> PixelShuffle(up) (block.py:333-346): phase = output sub-pixel
  int nphase = 1;
  uint32_t ph_woff[kMaxPhases] = {};
  uint8_t ph_ntaps[kMaxPhases] = {};
  uint8_t ph_a[kMaxPhases] = {}, ph_b[kMaxPhases] = {};
  uint8_t tap_hy[kMaxPhases][kMaxTaps] = {};
  uint8_t tap_hx[kMaxPhases][kMaxTaps] = {};
  int max_taps = 0;
  __half* d_w = nullptr;   // device, packed [phase][kslab][tap][2][N][8]
  __half* d_wrows = nullptr; // device, row-streaming packing [kslab][dx][2][dy*Cout+co][8] (Cout 32/64, plain 3x3)
  __half* d_wrows_pair = nullptr;  // same, split in two halves of the N columns for the CTA-pair mode (Cout 64, Cin > 64)
  float* d_bias = nullptr; // device, [nphase][N]
  size_t w_bytes = 0;
  // fp32 copies (OIHW + bias) kept on the device for the fp32-mode direct kernel
  float* d_w32 = nullptr;
  std::vector<float> h_w32, h_b32;
};

struct Epilogue {
  bool lrelu = false;
  float slope = 0.2f;
  ChunkView res1, res2;   // base == nullptr -> unused
  float alpha1 = 1.f, alpha2 = 1.f;
  bool compact4 = false;  // store out channels 0..3 as [tile][H][W][4] (8 bytes per pixel), see ConvTcParams
  bool act_after_res = false;  // LeakyReLU after the residual adds (PPON running sums) instead of before
  ChunkView raw_out;      // optional second destination for the pre-activation value (needs act_after_res)
  bool res1_unact = false; // res1 holds LeakyReLU_slope(v): the add uses v itself (PPON fp16: the running sum is read
                           // back from its activated copy in the concat buffer, which saves the raw second store)
  bool gate = false;      // pixel attention (PAN_arch.py:22-36,48-57): v = res1 * sigmoid(conv + bias) instead of the add
  bool self_gate = false; // PACnv with k3 and k2 merged into one conv (rows [0, N/2) = k3, [N/2, N) = k2 at the centre
                          // tap): v[c] = conv[c] * sigmoid(conv[N/2 + c] + bias[N/2 + c]); out_nchunks <= N/16
};

// Build the packed fp16 weights (and phase tables) from OIHW fp32 weights.  `bias` may be null.
// Returns 0 or a negative error code; `err` receives a message.
// ksize 3 (default) or 1: a 1x1 conv (ESRGAN+ conv1x1, block.py:390-391) runs as a single centre tap.
int conv_layer_build(ConvLayer& L, const float* w_oihw, const float* bias, int Cout, int Cin,
                     int up, std::string& err, int ksize = 3, int dil = 1);
// conv (Cin -> Cout*r*r) followed by PixelShuffle(r): one 9-tap phase per output sub-pixel (i, j), whose
// Cout filters are rows c*r*r + i*r + j of the weight tensor; the shuffle becomes output addressing.
int conv_layer_build_ps(ConvLayer& L, const float* w_oihw, const float* bias, int Cout, int Cin, int r,
                        std::string& err);
void conv_layer_free(ConvLayer& L);

// Cache of encoded TMA descriptors keyed by (base, B, CT, H, W, box width in pixels).
class TmapCache {
 public:
  // 5-D map of a [B][CT][H][W][8] tensor (or one wide image, B = 1) with a box of box_w pixels
  const CUtensorMap* get(const void* base, int B, int CT, int H, int W, int box_w, int box_h, int& rc);
  // wide-layout row-segment map of the row-streaming kernel (box of `kc` chunks)
  const CUtensorMap* get_rows(const void* base, int CT, int H, int Wtot, int kc, int& rc);
  void clear() { maps_.clear(); }

 private:
  struct alignas(64) Slot { CUtensorMap m; };
  std::map<std::tuple<const void*, int, int, int, int, int>, Slot*> maps_;
};

extern long long* g_rows_trace;
extern thread_local const char* g_last_conv_kernel;   // "conv_rows", "conv_rows_pair", "conv_rows_dil", "conv_up" or "conv_tc"

// Pick the number of 8-pixel sub-patches per CTA for an image of width W and accumulator width N.
int choose_J(int W, int N);

// Launch one fused conv.  in: `L.Cin_pad/8` chunks starting at in.chunk0 of a [B][in.CT][H][W][8]
// buffer; out: `out_nchunks` chunks at out.chunk0 of a [B][out.CT][up*H][up*W][8] buffer.
int conv_layer_run(const ConvLayer& L, TmapCache& cache, ChunkView in, int B, int H, int W,
                   ChunkView out, int out_nchunks, const Epilogue& ep, int num_sms,
                   cudaStream_t stream);

}  // namespace innfer
