#include "tmap.cuh"

#include "conv_tc.cuh"

namespace innfer {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_act_tmap(CUtensorMap* out, const void* base, int B, int CT, int H, int W, int box_w, int box_h) {
  EncodeTiledFn enc = resolve_encode();
  if (!enc) return -1;
  const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)CT, (cuuint64_t)B};
  const cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16,
                                 (cuuint64_t)CT * H * W * 16};
  const cuuint32_t box[5] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, 2, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

int encode_wide_rows_tmap(CUtensorMap* out, const void* base, int CT, int H, int Wtot, int box_chunks) {
  EncodeTiledFn enc = resolve_encode();
  if (!enc) return -1;
  if (Wtot % 16 != 0 || box_chunks < 1 || box_chunks > 256) return -2;
  const cuuint64_t dims[4] = {128, (cuuint64_t)(Wtot / 16), (cuuint64_t)H, (cuuint64_t)CT};
  const cuuint64_t strides[3] = {256, (cuuint64_t)Wtot * 16, (cuuint64_t)H * Wtot * 16};
  const cuuint32_t box[4] = {128, 9, 1, (cuuint32_t)box_chunks};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

}  // namespace innfer
