// -cf colour correction on the GPU. Follows color_fix (utils/utils.py:278-315):
//   A = srgb2linear(LR), B = srgb2linear(SR)                      (utils/colors.py:29-46)
//   Bd = cv2.resize(B, LR size, INTER_CUBIC)                      (a = -0.75, half-pixel centres,
//                                                                  replicate border, no antialias)
//   blurred = cv2.GaussianBlur(A - Bd, (3,3), 0)                  ([.25 .5 .25] separable, REFLECT_101)
//   out = linear2srgb(cv2.resize(blurred, SR size, INTER_CUBIC) + B)   (colors.py:49-60, truncating)
// Three HBM-bound kernels: (1) downscale+diff (reads SR once), (2) 3x3 blur on the small LR-sized
// difference, (3) upscale+add+encode (reads SR once more, writes the result once).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace innfer {
// lr [h][w][3], sr [H][W][3], out [H][W][3], all device uint8. Returns 0, -1 (shapes), -5 (memory)
// or a positive cudaError. *launches receives the number of kernels launched.
int color_fix_run(const uint8_t* lr, int h, int w, const uint8_t* sr, int H, int W, uint8_t* out,
                  cudaStream_t stream, int* launches);
}  // namespace innfer
