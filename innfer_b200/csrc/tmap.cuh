// Host helper: encode the 5-D TMA descriptor of a planar-chunk activation tensor
// [B][CT][H][W][8] fp16 with a (8, box_w, box_h, 2, 1) box. cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so the library never links libcuda at build time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace innfer {
// Returns 0 on success, a CUresult/cudaError-style non-zero code otherwise.
int encode_act_tmap(CUtensorMap* out, const void* base, int B, int CT, int H, int W, int box_w, int box_h);
// Wide layout [CT][H][Wtot][8] (Wtot a multiple of 16) seen as [CT][H][Wtot/16][128]: the box is one
// row segment of 9 groups of 16 pixels (144 pixels, 256-byte inner rows) of `box_chunks` chunks;
// groups outside [0, Wtot/16) are zero-filled.  Used by the row-streaming kernel (conv_rows.cu).
int encode_wide_rows_tmap(CUtensorMap* out, const void* base, int CT, int H, int Wtot, int box_chunks);
}  // namespace innfer
