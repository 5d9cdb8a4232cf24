// Non-convolution pieces of PAN (architectures/PAN_arch.py:171-222): the max-pooled self-attention block
// (block.py:398-473), its bicubic resize back to the feature map, and the bilinear "ILR" skip of the input image.
// All tensors are tiled planar-chunk [B][CT][H][W][8] of T (__half in fp16 mode, float in fp32 mode); the attention
// itself runs in fp32 on channel-last scratch arrays whatever the mode.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace innfer {

constexpr int kPanRow = 64;   // floats per pooled pixel in the channel-last scratch arrays (nf <= 64)
constexpr int kPanQK = 8;     // floats per pooled pixel in the query / key arrays (nf / 8 <= 8)

// MaxPool2d(pool, pool) (block.py:418-419,443-444): x [B][CT][H][W][8] (first `nchunks` chunks) ->
// pooled [B][hp*wp][kPanRow] fp32, hp = H / pool, wp = W / pool.
template <typename T>
int launch_pan_maxpool(const T* x, int CT, int nchunks, int B, int H, int W, int pool, float* pooled, cudaStream_t st);

// The three 1x1 Conv1d projections (block.py:452-454).  wcat: the [2 * kPanQK + kPanRow] x [kPanRow] fp32 matrix with
// rows 0..7 = conv_f, 8..15 = conv_g, 16.. = conv_h (zero rows / columns beyond the real sizes), stored TRANSPOSED
// ([input channel][row]); bcat the matching biases.
// nfp = channels rounded up to 8 (columns of `pooled` that the max-pool wrote).
int launch_pan_proj(const float* pooled, long long npix, int nfp, const float* wcat, const float* bcat, float* f, float* g,
                    float* hv, cudaStream_t st);

// out[b][i][c] = sum_j softmax_j(<f_i, g_j>) * hv[b][j][c]  (block.py:456-461); n = pooled pixels per image,
// nfp = channels rounded up to 8.
int launch_pan_attention(const float* f, const float* g, const float* hv, int B, int n, int nfp, float* out,
                         cudaStream_t st);

// The same on the tensor cores (fp16 engine mode): S = Q K^T and O = P V as tcgen05 MMAs with fp32 accumulators in TMEM,
// online softmax in fp32 registers between them (pan_ops.cu).
int launch_pan_attention_tc(const float* f, const float* g, const float* hv, int B, int n, int nfp, float* out,
                            cudaStream_t st);

// y = gamma * bicubic_resize(att, (H, W), align_corners=False) + x  (block.py:463-468): att [B][hp*wp][kPanRow].
template <typename T>
int launch_pan_bicubic_add(const float* att, int hp, int wp, const T* x, int x_CT, T* y, int y_CT, int nchunks,
                           int B, int H, int W, float gamma, cudaStream_t st);

// ILR = F.interpolate(x, scale_factor=s, mode='bilinear', align_corners=True) (PAN_arch.py:215-219) of chunk 0 of
// the input tiles x [B][x_CT][h][w][8] -> y [B][1][s*h][s*w][8].
template <typename T>
int launch_pan_bilinear(const T* x, int x_CT, int B, int h, int w, int s, T* y, cudaStream_t st);

}  // namespace innfer
