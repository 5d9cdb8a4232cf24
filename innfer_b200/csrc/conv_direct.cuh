// fp32 CUDA-core direct 3x3 convolution on planar-chunk float tensors [B][CT][H][W][8].
// This is the -no_fp16 path (run.py:333,345): fp32 storage, fp32 FFMA accumulation, the nearest
// upsample of upconv_block evaluated by source addressing (src = dst / up) with the ORIGINAL 3x3
// weights, so results follow the reference's fp32 forward to rounding-order differences.  It also
// serves as an independent on-device cross-check of the tcgen05 kernel.  It is not a fallback:
// a handle runs either this (cfg.fp16 == 0) or the tensor-core kernel (cfg.fp16 == 1).
#pragma once
#include "layers.cuh"

namespace innfer {

// Upload the fp32 weights of `L` in the [ci][tap][N] layout the kernel reads. Returns 0 / cudaError.
int conv_direct_upload(ConvLayer& L);

// Same contract as conv_layer_run; ChunkView::base points at float data in this mode.
int conv_direct_run(const ConvLayer& L, ChunkView in, int B, int H, int W, ChunkView out,
                    int out_nchunks, const Epilogue& ep, cudaStream_t stream);

}  // namespace innfer
