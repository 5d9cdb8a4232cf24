// HBM-bound pixel kernels around the conv trunk: image -> tile conversion (np2tensor +
// extract_patches_2d fused), tile blending (recompose_tensor) with optional uint8 quantisation
// (tensor2np), and layout conversions between NCHW tensors and the planar-chunk layout.
// Reference: utils/utils.py:164-248 (np2tensor / tensor2np), 318-445 (extract / recompose).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace innfer {

constexpr int kMaxTilesPerAxis = 192;

// Tile geometry of chop_forward for one image (all coordinates in LOW-RES pixels).
struct TilePlan {
  int H, W;        // image size
  int p;           // tile size = min(H, W, patch)
  int step;        // int(p * step)
  double stepf;    // the step factor itself (0.5 .. 1.0): recompose_tensor derives its overlap from it
  int nty, ntx;    // tiles per axis
  int ys[kMaxTilesPerAxis];
  int xs[kMaxTilesPerAxis];
};
// Fills the plan following extract_patches_2d (utils.py:349-362). Returns 0 or negative error.
int make_tile_plan(int H, int W, int patch, double step, TilePlan& plan);

enum PixelDType { kF16 = 0, kF32 = 1, kU8 = 2 };

// src: NCHW image [1][C][H][W] (fp16/fp32, already RGB in [0,1]) or uint8 HWC BGR (C==3, /255 and
// channel flip applied).  dst: tiles [nt][CT][p][p][8] fp16, channels >= C zero-filled; tiles
// [t0, t0+nt) of the row-major plan are produced.
// skip_pad (uint8 sources with C <= 8 only; ignored otherwise): write chunk 0 only -- the caller guarantees that the
// other chunks of dst already hold zeros (the engine zero-fills its tile buffer once per tile geometry).
int launch_image_to_tiles(const void* src, PixelDType st, int C, const TilePlan& plan, int t0, int nt,
                          __half* dst, int CT, cudaStream_t stream, bool skip_pad = false);

// tiles: [ntiles][CT][P][P][8] fp16 (P = scale*p), channel c of chunk 0 is output channel c.
// dst: NCHW [1][C][scale*H][scale*W] fp16/fp32, or uint8 HWC BGR with clip(255x).round().
// CT == 0 selects the compact tile layout [ntiles][P][P][4] fp16 (C <= 4) written by the last conv.
int launch_blend(const __half* tiles, int CT, const TilePlan& plan, int scale, int C, void* dst,
                 PixelDType dt, cudaStream_t stream);

// Plain layout conversions for the un-chopped forward.
int launch_nchw_to_chunks(const void* src, PixelDType st, int n, int C, int H, int W, __half* dst,
                          int CT, cudaStream_t stream);
int launch_chunks_to_nchw(const __half* src, int CT, int n, int C, int H, int W, void* dst,
                          PixelDType dt, cudaStream_t stream);

// fp32-mode flavours: identical semantics on [..][8] float chunks (32 bytes per chunk pixel).
int launch_image_to_tiles_f32(const void* src, PixelDType st, int C, const TilePlan& plan, int t0,
                              int nt, float* dst, int CT, cudaStream_t stream, bool skip_pad = false);
int launch_blend_f32(const float* tiles, int CT, const TilePlan& plan, int scale, int C, void* dst,
                     PixelDType dt, cudaStream_t stream);
int launch_nchw_to_chunks_f32(const void* src, PixelDType st, int n, int C, int H, int W, float* dst,
                              int CT, cudaStream_t stream);
int launch_chunks_to_nchw_f32(const float* src, int CT, int n, int C, int H, int W, void* dst,
                              PixelDType dt, cudaStream_t stream);

// dst = a + alpha * b over n16 16-byte groups of fp16 (whole wide / tiled tensors of equal geometry; dst may alias a)
int launch_axpy_f16(__half* dst, const __half* a, const __half* b, float alpha, size_t n16, cudaStream_t stream);

// NCHW <-> wide layout [CT][H][cols][8] (image b in columns [b*pitch, b*pitch + W)); separator columns are not
// touched by the first (the caller zeroes the buffer) and skipped by the second.
int launch_nchw_to_wide(const void* src, PixelDType st, int n, int C, int H, int W, __half* dst, int CT, int pitch,
                        int cols, cudaStream_t stream);
int launch_wide_to_nchw(const __half* src, int CT, int n, int C, int H, int W, int pitch, int cols, void* dst,
                        PixelDType dt, cudaStream_t stream);

}  // namespace innfer
