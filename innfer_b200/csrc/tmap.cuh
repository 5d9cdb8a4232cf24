// Host helper: encode the 5-D TMA descriptor of a planar-chunk activation tensor
// [B][CT][H][W][8] fp16 with a (8, box_w, 18, 2, 1) box. cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so the library never links libcuda at build time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace innfer {
// Returns 0 on success, a CUresult/cudaError-style non-zero code otherwise.
int encode_act_tmap(CUtensorMap* out, const void* base, int B, int CT, int H, int W, int box_w);
// Same tensor seen as 4-D [B][CT][H][W*8]: the 8 channels of a chunk and the W pixels are merged
// into one contiguous inner dimension so that a box row is box_w*16 bytes (TMA moves whole rows;
// with the 5-D form every 16-byte pixel chunk is its own request and the TMA unit becomes the
// bottleneck).  box_w * 8 must be <= 256 elements.  Box = (box_w*8, 18, box_chunks, 1), OOB -> zero.
int encode_act_tmap_merged(CUtensorMap* out, const void* base, int B, int CT, int H, int W, int box_w,
                           int box_chunks);
// Wide layout [CT][H][Wtot][8] (Wtot a multiple of 16) seen as [CT][H][Wtot/16][128]: the box is one
// row segment of 9 groups of 16 pixels (144 pixels, 256-byte inner rows) of `box_chunks` chunks;
// groups outside [0, Wtot/16) are zero-filled.  Used by the row-streaming kernel (conv_rows.cu).
int encode_wide_rows_tmap(CUtensorMap* out, const void* base, int CT, int H, int Wtot, int box_chunks);
}  // namespace innfer
