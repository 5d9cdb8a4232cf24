// Host helper: encode the 5-D TMA descriptor of a planar-chunk activation tensor
// [B][CT][H][W][8] fp16 with a (8, 8J+2, 18, 2, 1) box. cuTensorMapEncodeTiled is resolved through
// cudaGetDriverEntryPoint so the library never links libcuda at build time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace innfer {
// Returns 0 on success, a CUresult/cudaError-style non-zero code otherwise.
int encode_act_tmap(CUtensorMap* out, const void* base, int B, int CT, int H, int W, int J);
}  // namespace innfer
