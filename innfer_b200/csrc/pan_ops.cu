#include "pan_ops.cuh"

#include <math_constants.h>
#include <stdint.h>

namespace innfer {

namespace {

__device__ __forceinline__ void load8(const __half* p, float v[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h[e]);
    v[2 * e] = t.x;
    v[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float v[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__half* p, const float v[8]) {
  uint4 o;
  __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// one thread per (image, chunk, pooled pixel)
template <typename T>
__global__ void pan_maxpool_kernel(const T* __restrict__ x, int CT, int nchunks, int H, int W, int pool, int hp, int wp,
                                   long long total, float* __restrict__ pooled) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = (int)(idx % wp);
  long long r = idx / wp;
  const int py = (int)(r % hp);
  r /= hp;
  const int ch = (int)(r % nchunks), b = (int)(r / nchunks);
  const T* src = x + (((size_t)b * CT + ch) * H * W) * 8;
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -CUDART_INF_F;
  for (int dy = 0; dy < pool; ++dy)
    for (int dx = 0; dx < pool; ++dx) {
      float v[8];
      load8(src + ((size_t)(py * pool + dy) * W + (px * pool + dx)) * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    }
  store8(pooled + ((size_t)b * hp * wp + (size_t)py * wp + px) * kPanRow + ch * 8, m);
}

// one thread per (pooled pixel, projection row)
__global__ void pan_proj_kernel(const float* __restrict__ pooled, long long npix, int nfp, const float* __restrict__ wcat,
                                const float* __restrict__ bcat, float* __restrict__ f, float* __restrict__ g,
                                float* __restrict__ hv) {
  constexpr int R = 2 * kPanQK + kPanRow;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * R) return;
  const long long i = idx / R;
  const int j = (int)(idx % R);
  const float* xr = pooled + i * kPanRow;   // the same address across most of a warp (broadcast)
  const float* wr = wcat + j;               // [channel][row]: consecutive rows -> consecutive addresses
  float s = 0.f;
  for (int c = 0; c < nfp; ++c) s = fmaf(xr[c], wr[(size_t)c * R], s);
  s += bcat[j];
  if (j < kPanQK) f[i * kPanQK + j] = s;
  else if (j < 2 * kPanQK) g[i * kPanQK + j - kPanQK] = s;
  else hv[i * kPanRow + j - 2 * kPanQK] = s;
}

// 128 queries of one image per block; keys / values stream through shared memory in tiles of 64.  Two passes over
// the keys: the row maximum first (softmax subtracts it, as torch does), then the weighted sum.
template <int NFP>
__global__ void __launch_bounds__(128)
pan_attention_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ hv, int n,
                     float* __restrict__ out) {
  constexpr int TJ = 64;
  __shared__ __align__(16) float s_g[TJ][kPanQK];
  __shared__ __align__(16) float s_h[TJ][NFP];
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool live = i < n;
  const float* fb = f + (size_t)b * n * kPanQK;
  const float* gb = g + (size_t)b * n * kPanQK;
  const float* hb = hv + (size_t)b * n * kPanRow;
  float q[kPanQK];
#pragma unroll
  for (int e = 0; e < kPanQK; ++e) q[e] = live ? fb[(size_t)i * kPanQK + e] : 0.f;
  float m = -CUDART_INF_F;
  for (int j0 = 0; j0 < n; j0 += TJ) {
    __syncthreads();
    for (int t = threadIdx.x; t < TJ * kPanQK; t += 128) {
      const int j = j0 + t / kPanQK;
      s_g[t / kPanQK][t % kPanQK] = j < n ? gb[(size_t)j * kPanQK + t % kPanQK] : 0.f;
    }
    __syncthreads();
    const int jn = (n - j0) < TJ ? (n - j0) : TJ;
    for (int j = 0; j < jn; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < kPanQK; ++e) s = fmaf(q[e], s_g[j][e], s);
      m = fmaxf(m, s);
    }
  }
  float l = 0.f, acc[NFP];
#pragma unroll
  for (int c = 0; c < NFP; ++c) acc[c] = 0.f;
  for (int j0 = 0; j0 < n; j0 += TJ) {
    __syncthreads();
    for (int t = threadIdx.x; t < TJ * kPanQK; t += 128) {
      const int j = j0 + t / kPanQK;
      s_g[t / kPanQK][t % kPanQK] = j < n ? gb[(size_t)j * kPanQK + t % kPanQK] : 0.f;
    }
    for (int t = threadIdx.x; t < TJ * NFP; t += 128) {
      const int j = j0 + t / NFP;
      s_h[t / NFP][t % NFP] = j < n ? hb[(size_t)j * kPanRow + t % NFP] : 0.f;
    }
    __syncthreads();
    const int jn = (n - j0) < TJ ? (n - j0) : TJ;
    for (int j = 0; j < jn; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < kPanQK; ++e) s = fmaf(q[e], s_g[j][e], s);
      const float p = expf(s - m);
      l += p;
      const float4* hrow = reinterpret_cast<const float4*>(&s_h[j][0]);   // 16-byte broadcast loads
#pragma unroll
      for (int c4 = 0; c4 < NFP / 4; ++c4) {
        const float4 v = hrow[c4];
        acc[4 * c4 + 0] = fmaf(p, v.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(p, v.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(p, v.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(p, v.w, acc[4 * c4 + 3]);
      }
    }
  }
  if (!live) return;
  const float inv = 1.f / l;
  float* o = out + ((size_t)b * n + i) * kPanRow;
#pragma unroll
  for (int c = 0; c < NFP; ++c) o[c] = acc[c] * inv;
}

// torch's cubic convolution coefficients (A = -0.75) for the taps at floor(src) - 1 .. floor(src) + 2
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = 2.f - t, x2 = 1.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// one thread per (image, chunk, pixel)
template <typename T>
__global__ void pan_bicubic_add_kernel(const float* __restrict__ att, int hp, int wp, const T* __restrict__ x, int x_CT,
                                       T* __restrict__ y, int y_CT, int nchunks, int H, int W, float gamma,
                                       long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % W);
  long long r = idx / W;
  const int oy = (int)(r % H);
  r /= H;
  const int ch = (int)(r % nchunks), b = (int)(r / nchunks);
  // align_corners=False: src = scale * (dst + 0.5) - 0.5 with scale = in / out, not clamped for cubic
  const float sy = ((float)hp / (float)H) * ((float)oy + 0.5f) - 0.5f;
  const float sx = ((float)wp / (float)W) * ((float)ox + 0.5f) - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float cy[4], cx[4];
  cubic_coeffs(sy - fy, cy);
  cubic_coeffs(sx - fx, cx);
  const int iy = (int)fy, ix = (int)fx;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const float* ab = att + (size_t)b * hp * wp * kPanRow + ch * 8;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int yy = min(max(iy - 1 + ky, 0), hp - 1);
    float row[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) row[e] = 0.f;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int xx = min(max(ix - 1 + kx, 0), wp - 1);
      float v[8];
      load8(ab + ((size_t)yy * wp + xx) * kPanRow, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) row[e] = fmaf(cx[kx], v[e], row[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(cy[ky], row[e], acc[e]);
  }
  const size_t pix = (size_t)oy * W + ox;
  float xv[8];
  load8(x + (((size_t)b * x_CT + ch) * H * W + pix) * 8, xv);
#pragma unroll
  for (int e = 0; e < 8; ++e) xv[e] = fmaf(gamma, acc[e], xv[e]);
  store8(y + (((size_t)b * y_CT + ch) * H * W + pix) * 8, xv);
}

// one thread per (image, output pixel): the 8 channels of chunk 0
template <typename T>
__global__ void pan_bilinear_kernel(const T* __restrict__ x, int x_CT, int h, int w, int s, T* __restrict__ y,
                                    long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Ho = h * s, Wo = w * s;
  const int ox = (int)(idx % Wo);
  const long long r = idx / Wo;
  const int oy = (int)(r % Ho), b = (int)(r / Ho);
  // align_corners=True: src = dst * (in - 1) / (out - 1)
  const float ry = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float rx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float sy = ry * (float)oy, sx = rx * (float)ox;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const T* src = x + ((size_t)b * x_CT) * h * w * 8;
  float a[8], bb[8], c[8], d[8], o[8];
  load8(src + ((size_t)y0 * w + x0) * 8, a);
  load8(src + ((size_t)y0 * w + x1) * 8, bb);
  load8(src + ((size_t)y1 * w + x0) * 8, c);
  load8(src + ((size_t)y1 * w + x1) * 8, d);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = ly0 * (lx0 * a[e] + lx1 * bb[e]) + ly1 * (lx0 * c[e] + lx1 * d[e]);
  store8(y + ((size_t)b * Ho * Wo + (size_t)oy * Wo + ox) * 8, o);
}

inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace

template <typename T>
int launch_pan_maxpool(const T* x, int CT, int nchunks, int B, int H, int W, int pool, float* pooled, cudaStream_t st) {
  const int hp = H / pool, wp = W / pool;
  if (hp < 1 || wp < 1 || nchunks * 8 > kPanRow) return -2;
  const long long total = (long long)B * nchunks * hp * wp;
  pan_maxpool_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>(x, CT, nchunks, H, W, pool, hp, wp, total, pooled);
  return (int)cudaGetLastError();
}
template int launch_pan_maxpool<__half>(const __half*, int, int, int, int, int, int, float*, cudaStream_t);
template int launch_pan_maxpool<float>(const float*, int, int, int, int, int, int, float*, cudaStream_t);

int launch_pan_proj(const float* pooled, long long npix, int nfp, const float* wcat, const float* bcat, float* f, float* g,
                    float* hv, cudaStream_t st) {
  const long long total = npix * (2 * kPanQK + kPanRow);
  pan_proj_kernel<<<blocks_for(total, 256), 256, 0, st>>>(pooled, npix, nfp, wcat, bcat, f, g, hv);
  return (int)cudaGetLastError();
}

int launch_pan_attention(const float* f, const float* g, const float* hv, int B, int n, int nfp, float* out,
                         cudaStream_t st) {
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)B);
  switch (nfp) {
    case 8: pan_attention_kernel<8><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 16: pan_attention_kernel<16><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 24: pan_attention_kernel<24><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 32: pan_attention_kernel<32><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 40: pan_attention_kernel<40><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 48: pan_attention_kernel<48><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 56: pan_attention_kernel<56><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    case 64: pan_attention_kernel<64><<<grid, 128, 0, st>>>(f, g, hv, n, out); break;
    default: return -2;
  }
  return (int)cudaGetLastError();
}

template <typename T>
int launch_pan_bicubic_add(const float* att, int hp, int wp, const T* x, int x_CT, T* y, int y_CT, int nchunks,
                           int B, int H, int W, float gamma, cudaStream_t st) {
  const long long total = (long long)B * nchunks * H * W;
  pan_bicubic_add_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>(att, hp, wp, x, x_CT, y, y_CT, nchunks, H, W, gamma,
                                                                    total);
  return (int)cudaGetLastError();
}
template int launch_pan_bicubic_add<__half>(const float*, int, int, const __half*, int, __half*, int, int, int, int, int,
                                            float, cudaStream_t);
template int launch_pan_bicubic_add<float>(const float*, int, int, const float*, int, float*, int, int, int, int, int,
                                           float, cudaStream_t);

template <typename T>
int launch_pan_bilinear(const T* x, int x_CT, int B, int h, int w, int s, T* y, cudaStream_t st) {
  const long long total = (long long)B * h * s * w * s;
  pan_bilinear_kernel<T><<<blocks_for(total, 256), 256, 0, st>>>(x, x_CT, h, w, s, y, total);
  return (int)cudaGetLastError();
}
template int launch_pan_bilinear<__half>(const __half*, int, int, int, int, int, __half*, cudaStream_t);
template int launch_pan_bilinear<float>(const float*, int, int, int, int, int, float*, cudaStream_t);

}  // namespace innfer
