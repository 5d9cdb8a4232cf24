// Implicit-GEMM 3x3 (or phase-folded 2x2) convolution on tcgen05 tensor cores for sm_100a.
//
// Replaces, on the GPU path, what the reference runs as nn.Conv2d(k=3,s=1,p=1)+LeakyReLU(0.2)
// (+ torch.cat, + "*0.2 + x") per call of conv_block / ResidualDenseBlock_5C / RRDB / upconv_block
// (reference: architectures/block.py:213-254,348-361, architectures/RRDBNet_arch.py:91-98,152-165).
//
// Data layout in HBM ("planar chunk", NC/8HW8): activations are [tile][chunk][H][W][8] fp16, one
// chunk = 8 channels = 16 bytes per pixel.  A dense block's concat buffer is just a tensor with 24
// chunks; each conv writes its own chunk range, so torch.cat never runs.
//
// GEMM mapping: M = 128 output pixels (a 16-row x 8-col sub-patch), N = Cout, K = 16 input
// channels per MMA, one MMA per (sub-patch, tap, 16-channel slab).  A CTA owns a patch of
// 16 x (8*J) pixels: the (16+2) x (8J+2) halo tile of one 16-channel slab is brought in ONCE by a
// 5-D TMA box load (OOB -> zero = the conv's zero padding) and all taps / sub-patches address it
// through shifted SWIZZLE_NONE K-major matrix descriptors:
//   addr(pixel row m, kchunk c) = base + ((hy + m/8) * Wh + hx + 8j + m%8) * 16 + c * (Rh*Wh*16)
// i.e. SBO = Wh*16, LBO = Rh*Wh*16, start shifted by (hy*Wh + hx + 8j) * 16 bytes.
// Accumulators (J per tile, 128 lanes x N fp32 columns each) live in two TMEM sets used by
// alternate tiles, so the MMA warp works on the next tile while the epilogue drains the previous
// one.  The epilogue (bias, LeakyReLU, up to two scaled residual adds, fp16 pack) reads them with
// tcgen05.ld and stores 16-byte channel chunks straight into the destination chunk slice.
//
// Warp roles (608 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..17 = epilogue (TMEM lane quarter = warp_id % 4; four warps per quarter = sub-tile parity x chunk parity:
// the 128-pixel sub-tiles of a tile alternate between two groups, each of which splits the accumulator's 8-channel
// chunks between two warps), warp 18 = scout: it waits on the stage and
// accumulator-set barriers for the issuer and publishes the number of ready stages in shared memory
// (an mbarrier wait on the issuing thread costs 170-260 cycles even when already satisfied).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace innfer {

constexpr int kConvThreads = 608;         // producer, issuer, 16 epilogue warps, scout (does the issuer's barrier waits)
constexpr int kConvScoutWarp = 18;
constexpr int kPatchRows = 16;            // rows per CTA patch (= rows of one M=128 sub-patch)
constexpr int kMaxPhases = 9;
constexpr int kMaxTaps = 9;
constexpr int kConvTailBytes = 4096;       // barriers + TMEM slot + per-phase bias behind the stage ring

struct ConvTcParams {
  // source geometry
  int B, H, W;
  int in_chunk0;   // first input chunk inside the TMA-mapped buffer
  int kslabs;      // Cin / 16
  // CTA tiling
  int J;           // sub-patches per CTA (patch width = 8*J)
  int bands, cps;  // ceil(H/16), ceil(W/(8J))
  int nphase;      // 1 for a plain 3x3, up*up for a nearest-upsample-folded conv
  int up;          // output = up * source resolution
  int Hout, Wout;
  int stages;      // smem pipeline depth
  int nslots;      // TMEM accumulator slots of N columns: 2*J (two sets of J, alternate tiles)
  int tmem_cols;   // power of two >= nslots*N
  long long* trace;  // debugging: clock64 samples of CTA 0 (conv_up.cu, -DINNFER_ROWS_TRACE builds), or null
  int dil;         // dilation of a plain 3x3 conv (PPON's d1..d8, block.py:364-366): taps at (hy*dil, hx*dil) of
                   // a (16 + 2*dil) x (8J + 2*dil) halo tile; 1 for every other conv, 0 for 1x1 convs (no halo)
  // destination
  __half* out;
  int out_CT, out_chunk0, out_nchunks;
  int out_compact4;  // final conv only: store channels 0..3 as [tile][Hout][Wout][4] fp16 (8 B / pixel)
  // element strides of the destination / residual tensors: address = base + image * bs + chunk * cs +
  // row * ys + column * px.  Tiled [B][CT][H][W][8]: bs = CT*H*W*8, cs = H*W*8, ys = W*8, px = 8.
  // Wide [CT][H][Wtot][8] (images side by side, `pitch` columns apart): bs = pitch*8, cs = H*Wtot*8,
  // ys = Wtot*8.  Compact [B][H][W][4]: bs = H*W*4, ys = W*4, px = 4.
  long long out_bs, out_cs;
  int out_ys, out_px;
  long long res1_bs, res1_cs, res2_bs, res2_cs;
  int res1_ys, res2_ys;
  // second destination (or null) with the layout class of `out`: receives the value BEFORE the activation
  // when act_after_res is set (PPON's running sums add_k feed both the next dilated conv and, activated, c2)
  __half* raw;
  long long raw_bs, raw_cs;
  int raw_ys, raw_chunk0;
  int act_after_res;   // apply the LeakyReLU after the residual adds instead of before
  int pdl;             // programmatic dependent launch (ptx.cuh: pdl_wait): the prologue overlaps the previous kernel's tail
  int res1_unact;      // res1 holds LeakyReLU(v) of the value to add: undo it (v > 0 ? v : v / slope)
  int gate;            // 1: res1 multiplies sigmoid(conv + bias) instead of being added (PAN pixel attention);
                       // 2: column c is multiplied by sigmoid(column N/2 + c) of the same accumulator (merged PACnv)
  // wide SOURCE: B == 1, W == Wtot and a column decomposes as image * sep_pitch + x; columns with
  // x >= sep_w or image >= sep_nimg are separators: never computed, stored as zeros when the
  // destination is wide too (out_zero_sep), skipped otherwise.  sep_pitch == 0: tiled source.
  int sep_pitch, sep_w, sep_nimg, out_zero_sep;
  uint32_t sep_magic;    // floor(2^32 / sep_pitch) + 1
  // weights / bias
  const __half* w;       // packed [phase][kslab][tap][2][N][8]
  const float* bias;     // [nphase][N]
  // epilogue
  int lrelu;
  float slope;
  const __half* res1;
  int res1_CT, res1_chunk0;
  float alpha1;  // t = t*alpha1 + res1
  const __half* res2;
  int res2_CT, res2_chunk0;
  float alpha2;  // t = t*alpha2 + res2
  // phase tables
  uint32_t ph_woff[kMaxPhases];           // byte offset of the phase's weights inside w
  uint8_t ph_ntaps[kMaxPhases];
  uint8_t ph_a[kMaxPhases], ph_b[kMaxPhases];      // output sub-pixel offsets
  uint8_t tap_hy[kMaxPhases][kMaxTaps];   // halo-tile offsets of each tap (0..2)
  uint8_t tap_hx[kMaxPhases][kMaxTaps];
};

// Host-side launcher (conv_tc.cu). Returns cudaError_t as int.
int launch_conv_tc(const CUtensorMap* tmap_in, const ConvTcParams& p, int N, int num_sms,
                   cudaStream_t stream);
// x2 upsample-folded conv, all four phases per source-tile visit, resident weights (conv_up.cu); N = 64 only.
int launch_conv_up(const CUtensorMap* tmap_in, const ConvTcParams& p, int num_sms, cudaStream_t stream);
int conv_up_weight_bytes(int kslabs);
int conv_up_stage_bytes();
// Bytes of dynamic smem / stage geometry helpers shared by host and device.
__host__ __device__ inline int conv_tc_a_bytes(int J, int dil) {
  int b = 2 * (kPatchRows + 2 * dil) * (8 * J + 2 * dil) * 16;
  return (b + 127) & ~127;
}
__host__ __device__ inline int conv_tc_w_bytes(int N, int max_taps) { return max_taps * 2 * N * 16; }

}  // namespace innfer
