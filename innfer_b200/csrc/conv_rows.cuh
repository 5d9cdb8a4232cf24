// Row-streaming dy-taps-as-N convolution on the wide activation layout (conv_rows.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace innfer {

struct ConvRowsParams {
  // wide source image [chunk][H][Wtot][8]; Wtot is a multiple of 16
  int H, Wtot;
  int nimg, pitch, Wimg;   // image b occupies columns [b*pitch, b*pitch + Wimg)
  uint32_t magic;          // floor(2^32 / pitch) + 1: column / pitch == umulhi(column, magic)
  int in_chunk0;           // first input chunk inside the TMA-mapped buffer
  int nch;                 // input chunks (Cin_pad / 8), even
  int kc;                  // chunks per pipeline stage (= box depth of the tensor map), even, divides nch
  int nsub;                // nch / kc stages per row
  int nstrips;             // ceil((Wtot + 15) / 128); strip m covers columns [128m - 15, 128m + 113)
  int stages;              // shared-memory ring depth
  // destination: address = out + image * out_bs + chunk * out_cs + row * out_ys + column_in_image * out_px (elements).
  // Wide (same geometry as the source): bs = pitch*8, cs = H*Wtot*8, ys = Wtot*8, px = 8, separator columns are stored
  // as zeros (out_wide).  Tiled [B][CT][H][W][8] and compact [B][H][W][4] (out_compact4, channels 0..3) destinations
  // receive real pixels only -- the last conv of a net, possibly into a peer GPU's tile buffer.
  __half* out;
  long long out_bs, out_cs;
  int out_ys, out_px;
  int out_chunk0, out_nchunks;
  int out_wide, out_compact4;
  // residual tensors are wide with the source's geometry
  long long res_cs;
  int res_ys;
  // optional residuals, wide tensors of the destination's geometry (same row / chunk strides):
  //   t = lrelu(acc + bias);  t = t*alpha1 + res1;  t = t*alpha2 + res2
  const __half* res1;
  const __half* res2;
  int res1_chunk0, res2_chunk0;
  float alpha1, alpha2;
  // weights [slab][dx][2][dy*COUT + co][8] fp16 (pair: [half][slab][dx][2][rows of that half][8]), bias [COUT]
  const __half* w;
  const float* bias;
  int lrelu;
  float slope;
  int pair;              // 1: clusters of two CTAs, M = 256 MMAs, half of the weight columns per CTA (conv_rows.cu)
  int idt;               // pair only: t = (acc + bias) * alpha1 where the accumulator already holds x / alpha1 for the
                         // residual x = input channels 0..63 (identity MMAs, see kRowsIdtBytes); res1 is null then
  // dilated variant (PPON's 64 -> 32 convs, conv_rows.cu DILV): taps at distance `dil` (1..8; the separator between the
  // images must be at least that wide), LeakyReLU AFTER the res1 add, optional second store of the pre-activation value
  int dil;
  int act_after_res;
  int res1_unact;        // res1 holds LeakyReLU(v) of the value to add: undo it (v > 0 ? v : v / slope)
  __half* raw;           // wide tensor with the geometry of res1 / res2 (res_cs, res_ys), or null
  int raw_chunk0;
  int pdl;               // 1: programmatic dependent launch -- the prologue overlaps the previous kernel's tail
  long long* trace;      // debugging: clock64 samples of CTA 0 (see tests/gpu_bringup.py --stage trace), or null
  int dbg_dx0;           // timing experiment (INNFER_ROWS_DX0=1, wrong results): all three horizontal taps read the 128-byte
                         // aligned window, to measure what the 16/32-byte shifted A descriptors cost
};

// CTA-pair weights: every half is followed by four identity B tiles ([kchunk][32 rows][8] fp16, 1 KB each)
constexpr int kRowsIdtBytes = 4096;
constexpr float kRowsIdtScale = 5.0f;   // 1 / 0.2, exact in fp16; 0.2f * 5.0f == 1.0f in fp32

int launch_conv_rows(const CUtensorMap* tmap_in, const ConvRowsParams& p, int cout, int num_sms, cudaStream_t stream);
int conv_rows_stage_bytes(int kc);
int conv_rows_weight_bytes(int nch, int cout);

}  // namespace innfer
