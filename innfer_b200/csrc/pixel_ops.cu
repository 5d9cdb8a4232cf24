#include "pixel_ops.cuh"

#include <algorithm>
#include <vector>

#include <cmath>

namespace innfer {

int make_tile_plan(int H, int W, int patch, double step, TilePlan& plan) {
  if (H <= 0 || W <= 0 || patch <= 0) return -1;
  int p = H < W ? H : W;
  if (patch < p) p = patch;
  plan.H = H;
  plan.W = W;
  plan.p = p;
  // int(patch_H * step) with a Python float product (utils.py:351-352)
  int s = (int)((double)p * (double)step);
  if (s < 1) return -1;
  plan.step = s;
  plan.stepf = (double)step;
  auto fill = [&](int L, int* o, int& n) -> int {
    n = 0;
    const int cnt = (L - p) / s + 1;  // tensor.unfold window count
    for (int i = 0; i < cnt; ++i) {
      if (n >= kMaxTilesPerAxis) return -1;
      o[n++] = i * s;
    }
    if ((L - p) % s != 0) {            // extra window anchored at the far edge (utils.py:355-362)
      if (n >= kMaxTilesPerAxis) return -1;
      o[n++] = L - p;
    }
    return 0;
  };
  if (fill(H, plan.ys, plan.nty)) return -2;
  if (fill(W, plan.xs, plan.ntx)) return -2;
  return 0;
}

namespace {

struct TileGeom {
  int H, W, p, nty, ntx;
  int ys[kMaxTilesPerAxis];
  int xs[kMaxTilesPerAxis];
};

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, size_t i) {
  return __half2float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, size_t i) {
  return p[i];
}

// one thread = one (tile, chunk, y, x): 16-byte store, reads are coalesced along x per channel plane
template <typename E>
__device__ __forceinline__ void store_chunk(E* dst, size_t chunk_index, const float (&f)[8]);
template <>
__device__ __forceinline__ void store_chunk<__half>(__half* dst, size_t i, const float (&f)[8]) {
  __align__(16) __half v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = __float2half_rn(f[e]);
  *reinterpret_cast<uint4*>(dst + i * 8) = *reinterpret_cast<const uint4*>(v);
}
template <>
__device__ __forceinline__ void store_chunk<float>(float* dst, size_t i, const float (&f)[8]) {
  float4* d = reinterpret_cast<float4*>(dst + i * 8);
  d[0] = make_float4(f[0], f[1], f[2], f[3]);
  d[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <typename E>
__device__ __forceinline__ void load_chunk(const E* src, size_t chunk_index, float (&f)[8]);
template <>
__device__ __forceinline__ void load_chunk<__half>(const __half* src, size_t i, float (&f)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(src + i * 8);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = __half22float2(h[e]);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}
template <>
__device__ __forceinline__ void load_chunk<float>(const float* src, size_t i, float (&f)[8]) {
  const float4* s = reinterpret_cast<const float4*>(src + i * 8);
  const float4 a = s[0], b = s[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <typename T, bool U8, typename E>
__global__ void image_to_tiles_kernel(const void* __restrict__ src_, int C, const __grid_constant__ TileGeom g,
                                      int t0, int nt, E* __restrict__ dst, int CT) {
  const int p = g.p;
  const size_t total = (size_t)nt * CT * p * p;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p);
    size_t r = i / p;
    const int y = (int)(r % p);
    r /= p;
    const int ch = (int)(r % CT);
    const int t = (int)(r / CT) + t0;
    const int sy = g.ys[t / g.ntx] + y;
    const int sx = g.xs[t % g.ntx] + x;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = ch * 8 + e;
      float f = 0.f;
      if (c < C) {
        if (U8) {
          // HWC BGR uint8 -> RGB /255 (np2tensor + bgr_to_rgb, utils.py:177-187, colors.py:5-11)
          const uint8_t* s8 = reinterpret_cast<const uint8_t*>(src_);
          f = (float)s8[((size_t)sy * g.W + sx) * C + (C - 1 - c)] / 255.0f;
        } else {
          f = load_as_float<T>(reinterpret_cast<const T*>(src_), ((size_t)c * g.H + sy) * g.W + sx);
        }
      }
      v[e] = f;
    }
    store_chunk<E>(dst, i, v);
  }
}

// uint8 HWC source with at most 8 channels (the CLI's path: np2tensor + bgr_to_rgb + extract_patches_2d + .half(),
// utils.py:164-194,318-369): a block is (4 tile rows) x (64 pixel columns), the 256 possible `v / 255` values come from
// a shared-memory table filled with the same expression the generic kernel evaluates (bit-identical), every thread
// converts one pixel at a time and stores its 16-byte chunk; chunks past the image's channels are zero fill.
// No 64-bit index arithmetic, no per-pixel division: 62 -> 35 us per 95-tile batch with both chunks written, 23 us per
// 63 tiles when the engine skips the pad chunk (profiles/r02e_pixel_kernels.md).
template <typename E>
__global__ void __launch_bounds__(256)
image_to_tiles_u8_kernel(const uint8_t* __restrict__ src, int C, const __grid_constant__ TileGeom g, int t0, int nt,
                         E* __restrict__ dst, int CT, int CTW) {
  __shared__ E lut[256];
  {
    const float f = (float)threadIdx.x / 255.0f;
    if constexpr (sizeof(E) == 2) lut[threadIdx.x] = __float2half_rn(f);
    else lut[threadIdx.x] = f;
  }
  __syncthreads();
  const int p = g.p;
  const size_t plane = (size_t)p * p;                  // pixels per (tile, chunk) plane
  const int nrows = nt * p;
  // persistent blocks: the table is built once per block, every quarter of the block walks (tile, row) pairs
  for (int r = blockIdx.x * 4 + (threadIdx.x >> 6); r < nrows; r += gridDim.x * 4) {
    const int tl = r / p, y = r - tl * p;
    const int t = tl + t0;
    const int ty = t / g.ntx, tx = t - ty * g.ntx;
    const uint8_t* srow = src + ((size_t)(g.ys[ty] + y) * g.W + g.xs[tx]) * C;
    E* drow = dst + (((size_t)tl * CT) * plane + (size_t)y * p) * 8;
    for (int x0 = threadIdx.x & 63; x0 < p; x0 += 256) {
      // up to four pixels per thread with all their byte loads in flight before the first table lookup
      uint8_t raw[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x0 + 64 * k;
#pragma unroll
        for (int e = 0; e < 8; ++e) raw[k][e] = (e < C && x < p) ? srow[x * C + (C - 1 - e)] : (uint8_t)0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x0 + 64 * k;
        if (x >= p) break;
        __align__(16) E v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = lut[raw[k][e]];   // lut[0] == +0 for the padding channels
        E* d = drow + (size_t)x * 8;
        if constexpr (sizeof(E) == 2) {
          *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(v);
          for (int ch = 1; ch < CTW; ++ch) *reinterpret_cast<uint4*>(d + (size_t)ch * plane * 8) = make_uint4(0u, 0u, 0u, 0u);
        } else {
          reinterpret_cast<uint4*>(d)[0] = reinterpret_cast<const uint4*>(v)[0];
          reinterpret_cast<uint4*>(d)[1] = reinterpret_cast<const uint4*>(v)[1];
          for (int ch = 1; ch < CTW; ++ch) {
            uint4* z = reinterpret_cast<uint4*>(d + (size_t)ch * plane * 8);
            z[0] = z[1] = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
    }
  }
}

struct BlendGeom {
  int Hs, Ws;      // full output size
  int P;           // HR tile size
  int eff;         // int(0.5 * P)
  int overlap;
  int nty, ntx;
  int oys[kMaxTilesPerAxis];
  int oxs[kMaxTilesPerAxis];
};

// torch.linspace(0.1, 1.0, n) in fp32 (symmetric evaluation), then ones, then linspace(1.0, 0.1, n)
__device__ __forceinline__ float blend_profile(int i, int P, int n) {
  if (n == 1) return i == 0 ? 0.1f : 1.0f;   // linspace(a, b, 1) == [a]: ramp-in [0.1], ramp-out [1.0]
  const float step = n > 1 ? (1.0f - 0.1f) / (float)(n - 1) : 0.f;
  if (i < n) {
    return (i < n / 2) ? 0.1f + step * (float)i : 1.0f - step * (float)(n - 1 - i);
  }
  if (i >= P - n) {
    const int k = i - (P - n);
    const float dstep = n > 1 ? (0.1f - 1.0f) / (float)(n - 1) : 0.f;
    return (k < n / 2) ? 1.0f + dstep * (float)k : 0.1f - dstep * (float)(n - 1 - k);
  }
  return 1.0f;
}

__device__ __forceinline__ float quant_u8(float v) {
  // np.clip(255 * x, 0, 255).round() -- round half to even (utils.py:245)
  v = fminf(fmaxf(v * 255.0f, 0.f), 255.0f);
  return rintf(v);
}

// Compact tile pixel: 4 fp16 channels in 8 bytes.
__device__ __forceinline__ void load_compact4(const __half* src, size_t pixel_index, float (&f)[8]) {
  const uint2 raw = *reinterpret_cast<const uint2*>(src + pixel_index * 4);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
  const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  f[4] = f[5] = f[6] = f[7] = 0.f;
}

// Visits the tiles covering output position (Y, X) in the reference's row-major accumulation order
// (utils.py:436-443): regular tiles ascending, then the edge-anchored last tile of each axis.
template <typename F>
__device__ __forceinline__ void for_each_covering_tile(const BlendGeom& g, int Y, int X, F&& fn) {
  const int ty_hi = min(Y / g.eff, g.nty - 1);
  const int ty_lo = max(0, (Y - g.P) / g.eff);
  const int tx_hi = min(X / g.eff, g.ntx - 1);
  const int tx_lo = max(0, (X - g.P) / g.eff);
  for (int pass_y = 0; pass_y < 2; ++pass_y) {
    const int ya = pass_y == 0 ? ty_lo : g.nty - 1;
    const int yb = pass_y == 0 ? ty_hi : (ty_hi < g.nty - 1 ? g.nty - 1 : g.nty - 2);
    for (int ty = ya; ty <= yb; ++ty) {
      const int ly = Y - g.oys[ty];
      if (ly < 0 || ly >= g.P) continue;
      for (int pass_x = 0; pass_x < 2; ++pass_x) {
        const int xa = pass_x == 0 ? tx_lo : g.ntx - 1;
        const int xb = pass_x == 0 ? tx_hi : (tx_hi < g.ntx - 1 ? g.ntx - 1 : g.ntx - 2);
        for (int tx = xa; tx <= xb; ++tx) {
          const int lx = X - g.oxs[tx];
          if (lx < 0 || lx >= g.P) continue;
          fn(ty, tx, ly, lx);
        }
      }
    }
  }
}

// Scalar gather-blend: one thread per output pixel (any geometry).
template <int DT, typename E, bool COMPACT>
__global__ void blend_kernel(const E* __restrict__ tiles, int CT, const __grid_constant__ BlendGeom g, int C,
                             void* __restrict__ dst) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y;
  if (X >= g.Ws) return;
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  float wsum = 0.f;
  for_each_covering_tile(g, Y, X, [&](int ty, int tx, int ly, int lx) {
    const float w = blend_profile(lx, g.P, g.overlap) * blend_profile(ly, g.P, g.overlap);
    const size_t t = (size_t)ty * g.ntx + tx;
    float f[8];
    if (COMPACT) load_compact4(reinterpret_cast<const __half*>(tiles), (t * g.P + ly) * g.P + lx, f);
    else load_chunk<E>(tiles, (t * CT * g.P + ly) * g.P + lx, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e] * w;
    wsum += w;
  });
  const size_t plane = (size_t)g.Hs * g.Ws;
  const size_t pix = (size_t)Y * g.Ws + X;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c >= C) break;
    const float v = acc[c] / wsum;
    if (DT == kF16) reinterpret_cast<__half*>(dst)[c * plane + pix] = __float2half_rn(v);
    if (DT == kF32) reinterpret_cast<float*>(dst)[c * plane + pix] = v;
    if (DT == kU8) reinterpret_cast<uint8_t*>(dst)[pix * C + (C - 1 - c)] = (uint8_t)quant_u8(v);
  }
}

// Vector gather-blend for 3-channel images whose tile origins, tile size and width are multiples of
// 4 (every reference geometry with an even tile size and scale 4; checked on the host): one thread
// produces 4 consecutive pixels, which share their covering tiles, and stores 12 bytes (uint8 HWC)
// or 8 bytes per channel plane (fp16 NCHW) at once.  Same accumulation order as the scalar kernel.
// The profile of one tile axis, evaluated once per geometry by the same device function the scalar kernel calls per
// pixel (so both kernels see identical floats): prof[i] = blend_profile(i, P, overlap).
__global__ void blend_profile_table_kernel(int P, int overlap, float* __restrict__ prof) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) prof[i] = blend_profile(i, P, overlap);
}

// The tiles covering position v of one axis, in the reference's accumulation order (regular tiles ascending, then the
// edge-anchored last tile): indices into t[], local coordinates into l[]; returns the count (<= 4).
__device__ __forceinline__ int covering_axis(int v, int P, int eff, int nt, const int* origins, int (&t)[4], int (&l)[4]) {
  const int hi = min(v / eff, nt - 1);
  const int lo = max(0, (v - P) / eff);
  int n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int a = pass == 0 ? lo : nt - 1;
    const int b = pass == 0 ? hi : (hi < nt - 1 ? nt - 1 : nt - 2);
    for (int i = a; i <= b; ++i) {
      const int loc = v - origins[i];
      if (loc < 0 || loc >= P || n >= 3) continue;
      t[n] = i;
      l[n] = loc;
      ++n;
    }
  }
  return n;
}

template <int DT, typename E, bool COMPACT>
__global__ void blend_vec4_kernel(const E* __restrict__ tiles, int CT, const __grid_constant__ BlendGeom g,
                                  const float* __restrict__ prof, void* __restrict__ dst) {
  const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int Y = blockIdx.y;
  // the row's covering tiles and their vertical weights are the same for the whole block
  __shared__ int s_ny, s_ty[4], s_ly[4];
  __shared__ float s_wy[4];
  if (threadIdx.x == 0) {
    int t[4], l[4];
    const int n = covering_axis(Y, g.P, g.eff, g.nty, g.oys, t, l);
    s_ny = n;
    for (int i = 0; i < n; ++i) {
      s_ty[i] = t[i];
      s_ly[i] = l[i];
      s_wy[i] = prof[l[i]];
    }
  }
  __syncthreads();
  if (X0 >= g.Ws) return;
  int tx[4], lx[4];
  const int nx = covering_axis(X0, g.P, g.eff, g.ntx, g.oxs, tx, lx);   // 4 consecutive pixels share their tiles
  float acc[4][3];
  float wsum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    wsum[k] = 0.f;
    acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
  }
  const int ny = s_ny;
  // at most three tiles cover a position per axis (two regular ones + the edge-anchored last one); fixed trip counts
  // with predicates keep tx / lx in registers
#pragma unroll
  for (int iy = 0; iy < 3; ++iy) {
    if (iy >= ny) break;
    const float wy = s_wy[iy];
    const int ly = s_ly[iy];
    const size_t trow = (size_t)s_ty[iy] * g.ntx;
#pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
      if (ix >= nx) break;
      const size_t t = trow + tx[ix];
      const float4 wx = *reinterpret_cast<const float4*>(prof + lx[ix]);   // lx and the table are 16-byte aligned
      const float w4[4] = {wx.x * wy, wx.y * wy, wx.z * wy, wx.w * wy};
      if (COMPACT) {
        // 4 pixels x 4 halves = 32 contiguous, 32-byte aligned bytes (lx and P are multiples of 4)
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(tiles) +
                                                          ((t * g.P + ly) * g.P + lx[ix]) * 4);
        const uint4 a = src[0], b = src[1];
        const uint32_t raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float w = w4[k];
          const float2 rg = __half22float2(*reinterpret_cast<const __half2*>(&raw[2 * k]));
          const float2 bx = __half22float2(*reinterpret_cast<const __half2*>(&raw[2 * k + 1]));
          acc[k][0] += rg.x * w;
          acc[k][1] += rg.y * w;
          acc[k][2] += bx.x * w;
          wsum[k] += w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float w = w4[k];
          float f[8];
          load_chunk<E>(tiles, (t * CT * g.P + ly) * g.P + lx[ix] + k, f);
          acc[k][0] += f[0] * w;
          acc[k][1] += f[1] * w;
          acc[k][2] += f[2] * w;
          wsum[k] += w;
        }
      }
    }
  }
  const size_t plane = (size_t)g.Hs * g.Ws;
  const size_t pix = (size_t)Y * g.Ws + X0;
  if (DT == kU8) {
    __align__(4) uint8_t o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[k * 3 + (2 - c)] = (uint8_t)quant_u8(acc[k][c] / wsum[k]);
    uint32_t* out = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(dst) + pix * 3);
    const uint32_t* o32 = reinterpret_cast<const uint32_t*>(o);
    out[0] = o32[0];
    out[1] = o32[1];
    out[2] = o32[2];
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (DT == kF16) {
        __align__(8) __half hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) hv[k] = __float2half_rn(acc[k][c] / wsum[k]);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(dst) + c * plane + pix) = *reinterpret_cast<const uint2*>(hv);
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + c * plane + pix) =
            make_float4(acc[0][c] / wsum[0], acc[1][c] / wsum[1], acc[2][c] / wsum[2], acc[3][c] / wsum[3]);
      }
    }
  }
}

// Chunk-pixel strides of a planar-chunk tensor, in 16-byte units (one 8-channel pixel chunk each): image, chunk, row.
// Tiled [n][CT][H][W][8]: {CT*H*W, H*W, W}; wide [CT][H][Wtot][8] with images `pitch` columns apart: {pitch, H*Wtot, Wtot}.
struct ChunkStrides {
  size_t bs, cs, ys;
};

template <typename T, typename E>
__global__ void nchw_to_chunks_kernel(const T* __restrict__ src, int n, int C, int H, int W,
                                      E* __restrict__ dst, int CT, ChunkStrides st) {
  const size_t plane = (size_t)H * W;
  const size_t total = (size_t)n * CT * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % plane;
    const size_t r = i / plane;
    const int ch = (int)(r % CT);
    const size_t b = r / CT;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = ch * 8 + e;
      v[e] = c < C ? load_as_float<T>(src, (b * C + c) * plane + pix) : 0.f;
    }
    store_chunk<E>(dst, b * st.bs + ch * st.cs + (pix / W) * st.ys + pix % W, v);
  }
}

template <typename T, typename E>
__global__ void chunks_to_nchw_kernel(const E* __restrict__ src, int CT, int n, int C, int H, int W,
                                      T* __restrict__ dst, ChunkStrides st) {
  const size_t plane = (size_t)H * W;
  const size_t total = (size_t)n * C * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % plane;
    const size_t r = i / plane;
    const int c = (int)(r % C);
    const size_t b = r / C;
    const float v = load_as_float<E>(src, (b * st.bs + (c / 8) * st.cs + (pix / W) * st.ys + pix % W) * 8 + (c & 7));
    if (sizeof(T) == 2) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(v);
    else reinterpret_cast<float*>(dst)[i] = v;
  }
}

int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

template <typename E>
int image_to_tiles_impl(const void* src, PixelDType st, int C, const TilePlan& plan, int t0, int nt, E* dst,
                        int CT, cudaStream_t stream, bool skip_pad) {
  TileGeom g;
  g.H = plan.H;
  g.W = plan.W;
  g.p = plan.p;
  g.nty = plan.nty;
  g.ntx = plan.ntx;
  for (int i = 0; i < kMaxTilesPerAxis; ++i) {
    g.ys[i] = plan.ys[i];
    g.xs[i] = plan.xs[i];
  }
  const size_t total = (size_t)nt * CT * plan.p * plan.p;
  const int block = 256, grid = grid_for(total, block);
  if (st == kF16)
    image_to_tiles_kernel<__half, false, E><<<grid, block, 0, stream>>>(src, C, g, t0, nt, dst, CT);
  else if (st == kF32)
    image_to_tiles_kernel<float, false, E><<<grid, block, 0, stream>>>(src, C, g, t0, nt, dst, CT);
  else if (C <= 8 && (long long)nt * plan.p < (1ll << 31) - 8)
    image_to_tiles_u8_kernel<E><<<(unsigned)std::min<long long>(((long long)nt * plan.p + 3) / 4, 148 * 8), 256, 0, stream>>>(
        reinterpret_cast<const uint8_t*>(src), C, g, t0, nt, dst, CT, skip_pad ? 1 : CT);
  else
    image_to_tiles_kernel<float, true, E><<<grid, block, 0, stream>>>(src, C, g, t0, nt, dst, CT);
  return (int)cudaGetLastError();
}

// Per (device, tile size, overlap) profile table of the vector blend kernel; built on first use on the caller's stream
// (later launches on other streams of the same device find it complete: the build is tiny and the first launch that
// needs it is ordered behind it; the cache is per thread like the rest of the per-call scratch).
const float* blend_profile_table(int P, int overlap, cudaStream_t stream) {
  struct Entry {
    int dev, P, overlap;
    float* p;
  };
  static thread_local std::vector<Entry> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (const Entry& e : cache)
    if (e.dev == dev && e.P == P && e.overlap == overlap) return e.p;
  float* p = nullptr;
  if (cudaMalloc(&p, ((size_t)P + 4) * sizeof(float)) != cudaSuccess) return nullptr;
  blend_profile_table_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, overlap, p);
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
    cudaFree(p);
    return nullptr;
  }
  cache.push_back(Entry{dev, P, overlap, p});
  return p;
}

template <typename E>
int blend_impl(const E* tiles, int CT, const TilePlan& plan, int scale, int C, void* dst, PixelDType dt,
               cudaStream_t stream) {
  if (C > 8) return -1;
  BlendGeom g;
  g.Hs = plan.H * scale;
  g.Ws = plan.W * scale;
  g.P = plan.p * scale;
  // recompose_tensor (utils.py:396-399): overlap = scale*int(round((1-step)*(P/scale))), eff = int(step*P), all in
  // Python float (double) arithmetic; Python's round() is round-half-to-even, as is nearbyint in the default mode.
  const double sf = plan.stepf;
  g.overlap = scale * (int)std::nearbyint((1.0 - sf) * ((double)g.P / (double)scale));
  g.eff = (int)(sf * (double)g.P);
  if (g.P - 2 * g.overlap < 0 || g.eff < 1) return -2;
  g.nty = plan.nty;
  g.ntx = plan.ntx;
  // the reference re-derives the tile grid from the HR sizes (utils.py:405-411); when that disagrees with the grid
  // the tiles were cut on (possible for step != 0.5 with odd products) its patch index runs off: refuse loudly
  {
    const int stride = (int)((double)g.P * sf);
    if (stride < 1) return -2;
    auto count = [&](int full) {
      const int span = (full > g.P ? full : g.P) - g.P;
      return 1 + span / stride + (span % stride != 0 ? 1 : 0);
    };
    if (count(g.Hs) != plan.nty || count(g.Ws) != plan.ntx) return -2;
  }
  for (int i = 0; i < plan.nty; ++i) {
    const int o = i * g.eff;
    g.oys[i] = o < g.Hs - g.P ? o : g.Hs - g.P;
  }
  for (int i = 0; i < plan.ntx; ++i) {
    const int o = i * g.eff;
    g.oxs[i] = o < g.Ws - g.P ? o : g.Ws - g.P;
  }
  if (CT == 0 && (C > 4 || sizeof(E) != 2)) return -1;
  bool vec = (C == 3) && (g.P % 4 == 0) && (g.Ws % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  for (int i = 0; i < g.ntx && vec; ++i) vec = (g.oxs[i] % 4 == 0);
  dim3 block(128), grid(vec ? (g.Ws / 4 + 127) / 128 : (g.Ws + 127) / 128, g.Hs);
  const float* prof = nullptr;
  if (vec) {
    prof = blend_profile_table(g.P, g.overlap, stream);
    if (!prof) return (int)cudaErrorMemoryAllocation;
  }
#define INNFER_BLEND_LAUNCH(DT, COMPACT)                                                        \
  do {                                                                                          \
    if (vec) blend_vec4_kernel<DT, E, COMPACT><<<grid, block, 0, stream>>>(tiles, CT, g, prof, dst);  \
    else blend_kernel<DT, E, COMPACT><<<grid, block, 0, stream>>>(tiles, CT, g, C, dst);        \
  } while (0)
  if (CT == 0) {
    if (dt == kF16) INNFER_BLEND_LAUNCH(kF16, true);
    else if (dt == kF32) INNFER_BLEND_LAUNCH(kF32, true);
    else INNFER_BLEND_LAUNCH(kU8, true);
  } else {
    if (dt == kF16) INNFER_BLEND_LAUNCH(kF16, false);
    else if (dt == kF32) INNFER_BLEND_LAUNCH(kF32, false);
    else INNFER_BLEND_LAUNCH(kU8, false);
  }
#undef INNFER_BLEND_LAUNCH
  return (int)cudaGetLastError();
}

ChunkStrides chunk_strides(int CT, int H, int W, int wide_pitch, int wide_cols) {
  ChunkStrides st;
  if (wide_pitch > 0) {
    st.bs = (size_t)wide_pitch;
    st.cs = (size_t)H * wide_cols;
    st.ys = (size_t)wide_cols;
  } else {
    st.bs = (size_t)CT * H * W;
    st.cs = (size_t)H * W;
    st.ys = (size_t)W;
  }
  return st;
}

template <typename E>
int nchw_to_chunks_impl(const void* src, PixelDType st, int n, int C, int H, int W, E* dst, int CT,
                        cudaStream_t stream, int wide_pitch = 0, int wide_cols = 0) {
  const size_t total = (size_t)n * CT * H * W;
  const int block = 256, grid = grid_for(total, block);
  const ChunkStrides cs = chunk_strides(CT, H, W, wide_pitch, wide_cols);
  if (st == kF16)
    nchw_to_chunks_kernel<__half, E><<<grid, block, 0, stream>>>(reinterpret_cast<const __half*>(src), n, C, H, W, dst, CT, cs);
  else if (st == kF32)
    nchw_to_chunks_kernel<float, E><<<grid, block, 0, stream>>>(reinterpret_cast<const float*>(src), n, C, H, W, dst, CT, cs);
  else
    return -1;
  return (int)cudaGetLastError();
}

template <typename E>
int chunks_to_nchw_impl(const E* src, int CT, int n, int C, int H, int W, void* dst, PixelDType dt,
                        cudaStream_t stream, int wide_pitch = 0, int wide_cols = 0) {
  const size_t total = (size_t)n * C * H * W;
  const int block = 256, grid = grid_for(total, block);
  const ChunkStrides cs = chunk_strides(CT, H, W, wide_pitch, wide_cols);
  if (dt == kF16)
    chunks_to_nchw_kernel<__half, E><<<grid, block, 0, stream>>>(src, CT, n, C, H, W, reinterpret_cast<__half*>(dst), cs);
  else if (dt == kF32)
    chunks_to_nchw_kernel<float, E><<<grid, block, 0, stream>>>(src, CT, n, C, H, W, reinterpret_cast<float*>(dst), cs);
  else
    return -1;
  return (int)cudaGetLastError();
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) axpy_f16_kernel(uint4* __restrict__ dst, const uint4* __restrict__ a,
                                                       const uint4* __restrict__ b, float alpha, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 ua = a[i], ub = b[i];
    const __half2* ha = reinterpret_cast<const __half2*>(&ua);
    const __half2* hb = reinterpret_cast<const __half2*>(&ub);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __half22float2(ha[e]), y = __half22float2(hb[e]);
      ho[e] = __floats2half2_rn(fmaf(alpha, y.x, x.x), fmaf(alpha, y.y, x.y));
    }
    dst[i] = o;
  }
}
}  // namespace

int launch_axpy_f16(__half* dst, const __half* a, const __half* b, float alpha, size_t n16, cudaStream_t stream) {
  const int block = 256;
  const unsigned grid = (unsigned)std::min<size_t>((n16 + block - 1) / block, 148u * 16u);
  axpy_f16_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(a),
                                              reinterpret_cast<const uint4*>(b), alpha, n16);
  return (int)cudaGetLastError();
}

int launch_image_to_tiles(const void* src, PixelDType st, int C, const TilePlan& plan, int t0, int nt,
                          __half* dst, int CT, cudaStream_t stream, bool skip_pad) {
  return image_to_tiles_impl<__half>(src, st, C, plan, t0, nt, dst, CT, stream, skip_pad);
}
int launch_image_to_tiles_f32(const void* src, PixelDType st, int C, const TilePlan& plan, int t0, int nt,
                              float* dst, int CT, cudaStream_t stream, bool skip_pad) {
  return image_to_tiles_impl<float>(src, st, C, plan, t0, nt, dst, CT, stream, skip_pad);
}
int launch_blend(const __half* tiles, int CT, const TilePlan& plan, int scale, int C, void* dst,
                 PixelDType dt, cudaStream_t stream) {
  return blend_impl<__half>(tiles, CT, plan, scale, C, dst, dt, stream);
}
int launch_blend_f32(const float* tiles, int CT, const TilePlan& plan, int scale, int C, void* dst,
                     PixelDType dt, cudaStream_t stream) {
  return blend_impl<float>(tiles, CT, plan, scale, C, dst, dt, stream);
}
int launch_nchw_to_chunks(const void* src, PixelDType st, int n, int C, int H, int W, __half* dst,
                          int CT, cudaStream_t stream) {
  return nchw_to_chunks_impl<__half>(src, st, n, C, H, W, dst, CT, stream);
}
int launch_nchw_to_chunks_f32(const void* src, PixelDType st, int n, int C, int H, int W, float* dst,
                              int CT, cudaStream_t stream) {
  return nchw_to_chunks_impl<float>(src, st, n, C, H, W, dst, CT, stream);
}
int launch_chunks_to_nchw(const __half* src, int CT, int n, int C, int H, int W, void* dst,
                          PixelDType dt, cudaStream_t stream) {
  return chunks_to_nchw_impl<__half>(src, CT, n, C, H, W, dst, dt, stream);
}
int launch_chunks_to_nchw_f32(const float* src, int CT, int n, int C, int H, int W, void* dst,
                              PixelDType dt, cudaStream_t stream) {
  return chunks_to_nchw_impl<float>(src, CT, n, C, H, W, dst, dt, stream);
}
int launch_nchw_to_wide(const void* src, PixelDType st, int n, int C, int H, int W, __half* dst, int CT, int pitch,
                        int cols, cudaStream_t stream) {
  return nchw_to_chunks_impl<__half>(src, st, n, C, H, W, dst, CT, stream, pitch, cols);
}
int launch_wide_to_nchw(const __half* src, int CT, int n, int C, int H, int W, int pitch, int cols, void* dst,
                        PixelDType dt, cudaStream_t stream) {
  return chunks_to_nchw_impl<__half>(src, CT, n, C, H, W, dst, dt, stream, pitch, cols);
}

}  // namespace innfer
