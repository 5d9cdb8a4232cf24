#include "color_fix.cuh"

#include <cmath>
#include <mutex>

namespace innfer {

namespace {

struct Scratch {
  float* diff = nullptr;
  float* blur = nullptr;
  float* lut = nullptr;
  size_t cap = 0;
  int device = -1;
};
thread_local Scratch g_scratch;

// cv::interpolateCubic with A = -0.75 (OpenCV imgproc/resize.cpp)
__device__ __forceinline__ void cubic_coeffs(float x, float (&c)[4]) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
  c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  c[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
  c[3] = 1.f - c[0] - c[1] - c[2];
}

// source index / fraction of destination coordinate d: fx = (d + 0.5) * scale - 0.5
__device__ __forceinline__ void src_coord(int d, double scale, int& s, float& f) {
  const float fx = (float)(((double)d + 0.5) * scale - 0.5);
  const float fl = floorf(fx);
  s = (int)fl;
  f = fx - fl;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect101(int v, int n) {
  if (n == 1) return 0;
  if (v < 0) v = -v;
  if (v >= n) v = 2 * n - 2 - v;
  return v;
}

// (1) diff[y][x][c] = lin(lr) - cubic_down(lin(sr))   (or lin(lr) - lin(sr) when not scaling)
__global__ void cf_down_diff_kernel(const uint8_t* __restrict__ lr, int h, int w, const uint8_t* __restrict__ sr,
                                    int H, int W, const float* __restrict__ lut_g, double sy_scale, double sx_scale,
                                    int scaling, float* __restrict__ diff) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  float bd[3];
  if (scaling) {
    int sx, sy;
    float fx, fy;
    src_coord(x, sx_scale, sx, fx);
    src_coord(y, sy_scale, sy, fy);
    float cx[4], cy[4];
    cubic_coeffs(fx, cx);
    cubic_coeffs(fy, cy);
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = clampi(sx + j - 1, 0, W - 1);
    float rows[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = clampi(sy + k - 1, 0, H - 1);
      const uint8_t* row = sr + (size_t)yy * W * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        rows[k][c] = lut[row[xs[0] * 3 + c]] * cx[0] + lut[row[xs[1] * 3 + c]] * cx[1] +
                     lut[row[xs[2] * 3 + c]] * cx[2] + lut[row[xs[3] * 3 + c]] * cx[3];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      bd[c] = rows[0][c] * cy[0] + rows[1][c] * cy[1] + rows[2][c] * cy[2] + rows[3][c] * cy[3];
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) bd[c] = lut[sr[((size_t)y * W + x) * 3 + c]];
  }
  const size_t o = ((size_t)y * w + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) diff[o + c] = lut[lr[o + c]] - bd[c];
}

// (2) separable [0.25 0.5 0.25] with BORDER_REFLECT_101, row pass then column pass
__global__ void cf_blur_kernel(const float* __restrict__ diff, int h, int w, float* __restrict__ blur) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
  const int yu = reflect101(y - 1, h), yd = reflect101(y + 1, h);
  const int ys[3] = {yu, y, yd};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* row = diff + (size_t)ys[k] * w * 3;
      r[k] = row[x * 3 + c] * 0.5f + (row[xl * 3 + c] + row[xr * 3 + c]) * 0.25f;
    }
    blur[((size_t)y * w + x) * 3 + c] = r[1] * 0.5f + (r[0] + r[2]) * 0.25f;
  }
}

__device__ __forceinline__ uint8_t encode_srgb(float v) {
  // linear2srgb (colors.py:49-60): clip, piecewise gamma, *255, clip, truncating cast.
  // pow(v, 1/2.4) = exp2(log2(v)/2.4) with the SFU approximations (abs. error ~1e-6 on [0,1]):
  // a value that sits within that distance of an integer boundary can truncate to the neighbour,
  // which the <= 1 LSB tolerance of the path allows (measured flip rate < 1e-3).
  // The exponent lies in (-3.5, 0], so the bare ex2.approx needs no range scaling; the truncating cast is an add of
  // 2^23 rounded towards -inf (FP32 pipe) instead of a float-to-int conversion (which shares the SFU with lg2 / ex2).
  v = fminf(fmaxf(v, 0.f), 1.f);
  const float inv_gamma = (float)(1.0 / 2.4);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(inv_gamma * __log2f(v)));
  const float s = v <= 0.0031308f ? v * 12.92f : 1.055f * e - 0.055f;
  const float q = fminf(fmaxf(s * 255.0f, 0.f), 255.f);
  return (uint8_t)(__float_as_uint(__fadd_rd(q, 8388608.0f)) & 0xFFu);
}

// (3) out = encode(cubic_up(blur) + lin(sr)); each thread produces PX consecutive pixels of a row.
// Tap indices / weights are recomputed per thread (a lookup table of them costs more memory
// traffic than the image itself: measured 2.3x slower).
template <int PX>
__global__ void cf_up_apply_kernel(const float* __restrict__ blur, int h, int w, const uint8_t* __restrict__ sr,
                                   int H, int W, const float* __restrict__ lut_g, double sy_scale, double sx_scale,
                                   int scaling, uint8_t* __restrict__ out) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
  __syncthreads();
  const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
  const int Y = blockIdx.y;
  if (X0 >= W) return;
  int sy = Y;
  float cy[4] = {0.f, 1.f, 0.f, 0.f};
  if (scaling) {
    float fy;
    src_coord(Y, sy_scale, sy, fy);
    cubic_coeffs(fy, cy);
  }
  const float* brow[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) brow[k] = blur + (size_t)clampi(sy + k - 1, 0, h - 1) * w * 3;
  const size_t base = ((size_t)Y * W + X0) * 3;
  uint8_t in[PX * 3], res[PX * 3];
  if (PX == 4) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sr + base);
    uint32_t* i32 = reinterpret_cast<uint32_t*>(in);
    i32[0] = s32[0];
    i32[1] = s32[1];
    i32[2] = s32[2];
  } else {
    for (int i = 0; i < 3; ++i) in[i] = sr[base + i];
  }
#pragma unroll
  for (int px = 0; px < PX; ++px) {
    const int X = X0 + px;
    float up[3];
    if (scaling) {
      int sx;
      float fx;
      src_coord(X, sx_scale, sx, fx);
      float cx[4];
      cubic_coeffs(fx, cx);
      int xs[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xs[j] = clampi(sx + j - 1, 0, w - 1) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          r[k] = brow[k][xs[0] + c] * cx[0] + brow[k][xs[1] + c] * cx[1] + brow[k][xs[2] + c] * cx[2] +
                 brow[k][xs[3] + c] * cx[3];
        up[c] = r[0] * cy[0] + r[1] * cy[1] + r[2] * cy[2] + r[3] * cy[3];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) up[c] = blur[((size_t)Y * w + X) * 3 + c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) res[px * 3 + c] = encode_srgb(up[c] + lut[in[px * 3 + c]]);
  }
  if (PX == 4) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(out + base);
    const uint32_t* r32 = reinterpret_cast<const uint32_t*>(res);
    o32[0] = r32[0];
    o32[1] = r32[1];
    o32[2] = r32[2];
  } else {
    for (int i = 0; i < 3; ++i) out[base + i] = res[i];
  }
}

// (3') tiled form of cf_up_apply for rows of whole 4-pixel groups (W % 4 == 0): a block produces a kUpTY x kUpTX output
// tile.  The blur rows / columns the tile's taps touch are staged in shared memory once and cv::resize's two passes run
// separably through shared memory (horizontal pass for every staged source row, then the vertical pass per output
// pixel -- the order and the expressions of the per-pixel kernel above, so the values are the same floats).  A thread
// owns 4 consecutive pixels of a row: their 12 SR bytes are loaded as three words before anything else, the vertical
// pass reads 12 consecutive floats per tap row as three 16-byte shared-memory loads, the 12 result bytes leave as
// three words.  Per output value: 4 + 4 * (staged rows / tile rows) multiply-adds and the two SFU operations of the
// sRGB encode; the per-pixel kernel spent 48 cached global loads per pixel (164 us per 720p -> 2880p frame; this
// one: profiles/r02e_pixel_kernels.md).
constexpr int kUpTX = 128, kUpTY = 32;
constexpr int kUpItems = kUpTY * (kUpTX / 4) / 256;   // (row, 4-pixel group) items per thread
struct UpCoef {
  float c[4];
  int s[4];   // tap positions relative to the staged tile, already clamped to the image
};
__global__ void __launch_bounds__(256)
cf_up_apply_tile_kernel(const float* __restrict__ blur, int h, int w, const uint8_t* __restrict__ sr, int H, int W,
                        const float* __restrict__ lut_g, double sy_scale, double sx_scale, int rows_cap, int cols_cap,
                        uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t cf_smem[];
  __shared__ float lut[256];
  __shared__ __align__(16) UpCoef cxs[kUpTX];
  __shared__ __align__(16) UpCoef cys[kUpTY];
  __shared__ int s_org[4];   // first staged source column / row, staged column / row count
  float* sT = reinterpret_cast<float*>(cf_smem);                       // [rows_cap][kUpTX * 3] after the horizontal pass
  float* sB = sT + (size_t)rows_cap * kUpTX * 3;                       // [rows_cap][cols_cap * 3] blur tile
  const int tid = threadIdx.x;
  const int X0 = blockIdx.x * kUpTX, Y0 = blockIdx.y * kUpTY;
  const int nx = min(kUpTX, W - X0), ny = min(kUpTY, H - Y0);
  // the SR bytes of this thread's items (4 pixels = 12 bytes = three words each) start their trip first
  uint32_t srw[kUpItems][3];
#pragma unroll
  for (int it = 0; it < kUpItems; ++it) {
    const int i = tid + it * 256, y = i >> 5, q = i & 31;
    if (y < ny && 4 * q < nx) {
      const uint32_t* p32 = reinterpret_cast<const uint32_t*>(sr + ((size_t)(Y0 + y) * W + X0 + 4 * q) * 3);
      srw[it][0] = p32[0];
      srw[it][1] = p32[1];
      srw[it][2] = p32[2];
    }
  }
  lut[tid] = lut_g[tid];
  if (tid == 0) {
    int s0, s1;
    float f;
    src_coord(X0, sx_scale, s0, f);
    src_coord(X0 + nx - 1, sx_scale, s1, f);
    const int xa = clampi(s0 - 1, 0, w - 1), xb = clampi(s1 + 2, 0, w - 1);
    s_org[0] = xa;
    s_org[2] = xb - xa + 1;
    src_coord(Y0, sy_scale, s0, f);
    src_coord(Y0 + ny - 1, sy_scale, s1, f);
    const int ya = clampi(s0 - 1, 0, h - 1), yb = clampi(s1 + 2, 0, h - 1);
    s_org[1] = ya;
    s_org[3] = yb - ya + 1;
  }
  __syncthreads();
  const int xa = s_org[0], ya = s_org[1], ncol = s_org[2], nrow = s_org[3];
  if (tid < kUpTX) {
    if (tid < nx) {
      int sx;
      float fx;
      src_coord(X0 + tid, sx_scale, sx, fx);
      cubic_coeffs(fx, cxs[tid].c);
#pragma unroll
      for (int j = 0; j < 4; ++j) cxs[tid].s[j] = (clampi(sx + j - 1, 0, w - 1) - xa) * 3;
    }
  } else if (tid < kUpTX + kUpTY) {
    const int i = tid - kUpTX;
    if (i < ny) {
      int sy;
      float fy;
      src_coord(Y0 + i, sy_scale, sy, fy);
      cubic_coeffs(fy, cys[i].c);
#pragma unroll
      for (int k = 0; k < 4; ++k) cys[i].s[k] = (clampi(sy + k - 1, 0, h - 1) - ya) * (kUpTX * 3);
    }
  }
  // stage the blur tile (rows ya.., columns xa..; consecutive threads read consecutive floats)
  const int rowf = ncol * 3;
  for (int i = tid; i < nrow * rowf; i += 256) {
    const int r = i / rowf, c = i - r * rowf;
    sB[r * (cols_cap * 3) + c] = blur[((size_t)(ya + r) * w + xa) * 3 + c];
  }
  __syncthreads();
  // horizontal pass: every staged row, every output column of the tile
  for (int i = tid; i < nrow * kUpTX; i += 256) {
    const int r = i >> 7, x = i & (kUpTX - 1);
    if (x >= nx) continue;
    const float4 kc = *reinterpret_cast<const float4*>(cxs[x].c);
    const int4 ks = *reinterpret_cast<const int4*>(cxs[x].s);
    const float* b = sB + r * (cols_cap * 3);
    float* t = sT + (r * kUpTX + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      t[c] = b[ks.x + c] * kc.x + b[ks.y + c] * kc.y + b[ks.z + c] * kc.z + b[ks.w + c] * kc.w;
  }
  __syncthreads();
  // vertical pass + add + encode: 12 consecutive floats (4 pixels) per tap row as three 16-byte reads
#pragma unroll
  for (int it = 0; it < kUpItems; ++it) {
    const int i = tid + it * 256, y = i >> 5, q = i & 31;
    if (y >= ny || 4 * q >= nx) continue;
    const float4 kc = *reinterpret_cast<const float4*>(cys[y].c);
    const int4 ks = *reinterpret_cast<const int4*>(cys[y].s);
    const float* t = sT + q * 12;
    float up[12];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const float4 a0 = *reinterpret_cast<const float4*>(t + ks.x + 4 * v);
      const float4 a1 = *reinterpret_cast<const float4*>(t + ks.y + 4 * v);
      const float4 a2 = *reinterpret_cast<const float4*>(t + ks.z + 4 * v);
      const float4 a3 = *reinterpret_cast<const float4*>(t + ks.w + 4 * v);
      up[4 * v + 0] = a0.x * kc.x + a1.x * kc.y + a2.x * kc.z + a3.x * kc.w;
      up[4 * v + 1] = a0.y * kc.x + a1.y * kc.y + a2.y * kc.z + a3.y * kc.w;
      up[4 * v + 2] = a0.z * kc.x + a1.z * kc.y + a2.z * kc.z + a3.z * kc.w;
      up[4 * v + 3] = a0.w * kc.x + a1.w * kc.y + a2.w * kc.z + a3.w * kc.w;
    }
    uint32_t res[3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      uint32_t o = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t byte = (srw[it][v] >> (8 * e)) & 0xFFu;
        o |= (uint32_t)encode_srgb(up[4 * v + e] + lut[byte]) << (8 * e);
      }
      res[v] = o;
    }
    uint32_t* o32 = reinterpret_cast<uint32_t*>(out + ((size_t)(Y0 + y) * W + X0 + 4 * q) * 3);
    o32[0] = res[0];
    o32[1] = res[1];
    o32[2] = res[2];
  }
}

// (1'+2) exact 4x case, difference and blur in one pass.  Every LR pixel owns a private 4x4 block of SR (sx = 4x + 1,
// fraction 0.5), read with three 4-byte loads per row.  A block computes the difference image of a kDbTY x kDbTX LR
// tile plus a one-pixel ring (reflected at the image border like BORDER_REFLECT_101 does) into shared memory and blurs
// it from there -- the same expressions as cf_down_diff_kernel followed by cf_blur_kernel, so the same floats, without
// the round trip of the difference image and with one launch less.
constexpr int kDbTX = 64, kDbTY = 16;
__global__ void __launch_bounds__(256)
cf_down_diff4_blur_kernel(const uint8_t* __restrict__ lr, int h, int w, const uint8_t* __restrict__ sr, int W,
                          const float* __restrict__ lut_g, float* __restrict__ blur) {
  __shared__ float lut[256];
  __shared__ float sD[(kDbTY + 2) * (kDbTX + 2) * 3];
  const int tid = threadIdx.x;
  lut[tid] = lut_g[tid];
  __syncthreads();
  const int x0 = blockIdx.x * kDbTX, y0 = blockIdx.y * kDbTY;
  float c[4];
  cubic_coeffs(0.5f, c);
  for (int i = tid; i < (kDbTY + 2) * (kDbTX + 2); i += 256) {
    const int ry = i / (kDbTX + 2), rx = i - ry * (kDbTX + 2);
    const int yy = y0 + ry - 1, xx = x0 + rx - 1;
    if (yy > h || xx > w) continue;                 // ring positions past the reflected border are never read
    const int y = reflect101(yy, h), x = reflect101(xx, w);
    float rows[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(sr + ((size_t)(4 * y + k) * W + 4 * x) * 3);
      uint32_t raw[3] = {p[0], p[1], p[2]};
      const uint8_t* b = reinterpret_cast<const uint8_t*>(raw);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        rows[k][ch] = lut[b[ch]] * c[0] + lut[b[3 + ch]] * c[1] + lut[b[6 + ch]] * c[2] + lut[b[9 + ch]] * c[3];
    }
    const size_t o = ((size_t)y * w + x) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float bd = rows[0][ch] * c[0] + rows[1][ch] * c[1] + rows[2][ch] * c[2] + rows[3][ch] * c[3];
      sD[i * 3 + ch] = lut[lr[o + ch]] - bd;
    }
  }
  __syncthreads();
  for (int i = tid; i < kDbTY * kDbTX; i += 256) {
    const int ry = i / kDbTX, rx = i - ry * kDbTX;
    const int y = y0 + ry, x = x0 + rx;
    if (y >= h || x >= w) continue;
    const float* d = sD + ((ry + 1) * (kDbTX + 2) + rx + 1) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float r[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* row = d + (k - 1) * (kDbTX + 2) * 3;
        r[k] = row[ch] * 0.5f + (row[ch - 3] + row[ch + 3]) * 0.25f;
      }
      blur[((size_t)y * w + x) * 3 + ch] = r[1] * 0.5f + (r[0] + r[2]) * 0.25f;
    }
  }
}

}  // namespace

int color_fix_run(const uint8_t* lr, int h, int w, const uint8_t* sr, int H, int W, uint8_t* out,
                  cudaStream_t stream, int* launches) {
  if (launches) *launches = 0;
  if (h < 1 || w < 1 || H < 1 || W < 1) return -1;
  const int scaling = (h < H && w < W) ? 1 : 0;
  if (!scaling && !(h == H && w == W)) return -1;  // numpy would fail to broadcast imgA - imgB
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  Scratch& s = g_scratch;
  const size_t need = (size_t)h * w * 3 * sizeof(float);
  if (s.device != dev || s.cap < need) {
    if (s.diff) cudaFree(s.diff);
    if (s.blur) cudaFree(s.blur);
    s.diff = s.blur = nullptr;
    s.cap = 0;
    if (cudaMalloc(&s.diff, need) != cudaSuccess || cudaMalloc(&s.blur, need) != cudaSuccess) return -5;
    s.cap = need;
    if (s.device != dev || !s.lut) {
      if (cudaMalloc(&s.lut, 256 * sizeof(float)) != cudaSuccess) return -5;
      // srgb2linear on the 256 possible uint8 values, float32 arithmetic like numpy (colors.py:43-46)
      float lut[256];
      for (int i = 0; i < 256; ++i) {
        const float v = (float)i / 255.0f;
        lut[i] = v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
      }
      if (cudaMemcpy(s.lut, lut, sizeof lut, cudaMemcpyHostToDevice) != cudaSuccess) return -5;
    }
    s.device = dev;
  }
  // cv::resize: scale = 1 / (dst / src)
  const double down_x = 1.0 / ((double)w / (double)W), down_y = 1.0 / ((double)h / (double)H);
  const double up_x = 1.0 / ((double)W / (double)w), up_y = 1.0 / ((double)H / (double)h);
  int nl = 0;
  dim3 block(128);
  dim3 g1((w + 127) / 128, h);
  const bool exact4 = scaling && H == 4 * h && W == 4 * w && (reinterpret_cast<uintptr_t>(sr) % 4 == 0);
  if (exact4) {
    dim3 gdb((w + kDbTX - 1) / kDbTX, (h + kDbTY - 1) / kDbTY);
    cf_down_diff4_blur_kernel<<<gdb, 256, 0, stream>>>(lr, h, w, sr, W, s.lut, s.blur);
    nl = -1;
  } else {
    cf_down_diff_kernel<<<g1, block, 0, stream>>>(lr, h, w, sr, H, W, s.lut, down_y, down_x, scaling, s.diff);
    cf_blur_kernel<<<g1, block, 0, stream>>>(s.diff, h, w, s.blur);
  }
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(sr) | reinterpret_cast<uintptr_t>(out)) % 4 == 0);
  // staged source rows / columns of a tile: the span of the first and last pixel's taps (ratio < 1 when scaling)
  const int rows_cap = (int)std::ceil((kUpTY - 1) * up_y) + 5, cols_cap = (int)std::ceil((kUpTX - 1) * up_x) + 5;
  const size_t tile_smem = ((size_t)rows_cap * cols_cap * 3 + (size_t)rows_cap * kUpTX * 3) * sizeof(float);
  const bool tiled = scaling && vec && tile_smem <= 160 * 1024;
  if (tiled) {
    static thread_local int attr_dev = -1;
    if (attr_dev != dev) {   // once per (thread, device): opt in to the worst-case tile (ratios close to 1)
      if (cudaFuncSetAttribute(cf_up_apply_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) !=
          cudaSuccess)
        return -5;
      attr_dev = dev;
    }
    dim3 g3((W + kUpTX - 1) / kUpTX, (H + kUpTY - 1) / kUpTY);
    cf_up_apply_tile_kernel<<<g3, 256, tile_smem, stream>>>(s.blur, h, w, sr, H, W, s.lut, up_y, up_x, rows_cap,
                                                            cols_cap, out);
  } else if (vec) {
    dim3 g3((W / 4 + 127) / 128, H);
    cf_up_apply_kernel<4><<<g3, block, 0, stream>>>(s.blur, h, w, sr, H, W, s.lut, up_y, up_x, scaling, out);
  } else {
    dim3 g3((W + 127) / 128, H);
    cf_up_apply_kernel<1><<<g3, block, 0, stream>>>(s.blur, h, w, sr, H, W, s.lut, up_y, up_x, scaling, out);
  }
  if (launches) *launches = nl + 3;
  return (int)cudaGetLastError();
}

}  // namespace innfer
