#include "pan_ops.cuh"

#include <cmath>
#include <cstdlib>

#include "ptx.cuh"

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


// ------------------------------------------------------------------------------------------------ attention on tcgen05
// The same contraction (block.py:456-461) as two tensor-core GEMMs per block of 128 keys with an online softmax between
// them (fp16 operands, fp32 accumulation in TMEM, fp32 softmax in registers):
//     S[128 q x 128 k] = Q[128 x 16] K[128 x 16]^T            one M = 128, N = 128, K = 16 MMA (channels 8..15 are zero)
//     P = exp(S - rowmax), running rowmax / rowsum, previous output rescaled
//     O[128 q x NV]   += P[128 x 128] V[128 k x NV]             eight M = 128, N = NV, K = 16 MMAs
// 128 threads own one query row each (= TMEM lane): they stage K / V^T of the block into the SWIZZLE_NONE K-major operand
// layouts (fp32 -> fp16), read S, write P as the A operand of the second GEMM, and fold the block's O into their
// registers.  One more warp issues the MMAs.  The phases of a block are sequential (two CTAs per SM overlap each other).
constexpr int kAttThreads = 160;

template <int NV>   // value channels rounded up to 16
__global__ void __launch_bounds__(kAttThreads, 2)
pan_attention_tc_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ hv, int n, int nfp,
                        float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [Q: 2 x 128 x 16 B][K: 2 x 128 x 16 B][P: 16 x 128 x 16 B][Vt: 16 x NV x 16 B][barriers]
  constexpr uint32_t kQ = 0, kK = 4096, kP = 8192, kV = 8192 + 32768, kBar = kV + 16u * NV * 16u;
  constexpr uint32_t TCOLS = 256;   // S: columns [0, 128), O block: [128, 128 + NV)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * 128;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kBar);
  uint64_t* o_bar = s_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_bar + 1);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(s_bar), 1);
    mbar_init(smem_u32(o_bar), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), TCOLS);
    tmem_relinquish();
  }
  const float* fb = f + (size_t)b * n * kPanQK;
  const float* gb = g + (size_t)b * n * kPanQK;
  const float* hb = hv + (size_t)b * n * kPanRow;
  const int tid = threadIdx.x;
  auto pack8 = [](const float* v) {
    uint4 o;
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2);
    o.w = *reinterpret_cast<uint32_t*>(&h3);
    return o;
  };
  if (warp < 4) {
    // Q: row tid, K-chunk 0 = its 8 projection values, K-chunk 1 = zeros
    float v[8];
    const int i = q0 + tid;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = i < n ? fb[(size_t)i * kPanQK + e] : 0.f;
    *reinterpret_cast<uint4*>(smem + kQ + tid * 16) = pack8(v);
    *reinterpret_cast<uint4*>(smem + kQ + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(smem + kK + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);   // K-chunk 1 of every key block
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int nblocks = (n + 127) / 128;
  float m_run = -CUDART_INF_F, l_run = 0.f;
  float o_acc[NV];
#pragma unroll
  for (int c = 0; c < NV; ++c) o_acc[c] = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;

  for (int kb = 0; kb < nblocks; ++kb) {
    const int j0 = kb * 128;
    if (warp < 4) {
      // ---- stage the key block: K rows (B operand [K-chunk][key][8]) and V transposed (B operand [key chunk][channel][8 keys])
      {
        float v[8];
        const int j = j0 + tid;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = j < n ? gb[(size_t)j * kPanQK + e] : 0.f;
        *reinterpret_cast<uint4*>(smem + kK + tid * 16) = pack8(v);
      }
      // thread -> (key chunk kc = 8 keys, channel c): NV x 16 items, coalesced over channels for a fixed key
      for (int it = tid; it < 16 * NV; it += 128) {
        const int kc = it / NV, c = it - kc * NV;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = j0 + kc * 8 + e;
          v[e] = (j < n && c < nfp) ? hb[(size_t)j * kPanRow + c] : 0.f;
        }
        *reinterpret_cast<uint4*>(smem + kV + ((uint32_t)kc * NV + c) * 16) = pack8(v);
      }
      fence_proxy_async_smem();
    }
    __syncthreads();
    if (warp == 4) {
      // ---- S = Q K^T
      if (lane == 0) {
        tc_fence_after();
        const uint64_t ad = make_smem_desc(sbase + kQ, 2048u, 128u);
        const uint64_t bd = make_smem_desc(sbase + kK, 2048u, 128u);
        umma_f16_ss(tmem_base, ad, bd, make_idesc_f16(128), 0u);
        umma_commit(smem_u32(s_bar));
      }
      __syncwarp();
    } else {
      mbar_wait(smem_u32(s_bar), (uint32_t)kb & 1u);
      tc_fence_after();
      // ---- online softmax over this block's 128 keys; P goes to shared memory as the next A operand
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      float sv[128];
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 16) tmem_ld16(trow + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&sv[c0]));
      tmem_ld_wait();
      float mx = m_run;
#pragma unroll
      for (int j = 0; j < 128; ++j) {
        if (j0 + j >= n) sv[j] = -CUDART_INF_F;
        mx = fmaxf(mx, sv[j]);
      }
      const float alpha = exp2f((m_run - mx) * kLog2e);   // 0 for the first block (m_run = -inf)
      m_run = mx;
      float psum = 0.f;
#pragma unroll
      for (int kc = 0; kc < 16; ++kc) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = exp2f((sv[kc * 8 + e] - mx) * kLog2e);
          psum += pv[e];
        }
        *reinterpret_cast<uint4*>(smem + kP + ((uint32_t)kc * 128 + tid) * 16) = pack8(pv);
      }
      l_run = l_run * alpha + psum;
#pragma unroll
      for (int c = 0; c < NV; ++c) o_acc[c] *= alpha;
      fence_proxy_async_smem();
      tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
      // ---- O_block = P V
      if (lane == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(NV);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_smem_desc(sbase + kP + (uint32_t)kk * 4096u, 2048u, 128u);
          const uint64_t bd = make_smem_desc(sbase + kV + (uint32_t)kk * 2u * NV * 16u, (uint32_t)NV * 16u, 128u);
          umma_f16_ss(tmem_base + 128u, ad, bd, idesc, kk != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(o_bar));
      }
      __syncwarp();
    } else {
      mbar_wait(smem_u32(o_bar), (uint32_t)kb & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + 128u;
      uint32_t ov[NV];
#pragma unroll
      for (int c0 = 0; c0 < NV; c0 += 16) tmem_ld16(trow + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&ov[c0]));
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < NV; ++c) o_acc[c] += __uint_as_float(ov[c]);
      tc_fence_before();
    }
    __syncthreads();   // the operand buffers and both accumulators are free again
  }
  if (warp < 4) {
    const int i = q0 + tid;
    if (i < n) {
      const float inv = 1.f / l_run;
      float* o = out + ((size_t)b * n + i) * kPanRow;
#pragma unroll
      for (int c = 0; c < NV; ++c)
        if (c < nfp) o[c] = o_acc[c] * inv;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

template <int NV>
int launch_att_tc(const float* f, const float* g, const float* hv, int B, int n, int nfp, float* out, cudaStream_t st) {
  constexpr size_t smem = 8192 + 32768 + 16 * NV * 16 + 64;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(pan_attention_tc_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)   // two CTAs per SM hide each other's sequential phases: ask for the shared-memory carve-out they need
      e = cudaFuncSetAttribute(pan_attention_tc_kernel<NV>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev & 63] = true;
  }
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)B);
  pan_attention_tc_kernel<NV><<<grid, kAttThreads, smem, st>>>(f, g, hv, n, nfp, out);
  return (int)cudaGetLastError();
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

// Tiled form of pan_bicubic_add_kernel: a block produces a kBcTY x kBcTX output tile of one (image, chunk).  The
// attention rows / columns its taps touch are staged in shared memory once and the two passes run separably through
// shared memory -- the per-pixel kernel fetched its 16 taps (32 bytes each) from L2 for every output pixel, 6.4 GB of
// L2 -> SM traffic per 63-tile batch for a 0.4 GB result (552 us).  Same expressions in the same order (row = fma chain
// over kx from 0, acc = fma chain over ky), so the output is bit-identical to the per-pixel kernel.
constexpr int kBcTX = 64, kBcTY = 16;
struct BcCoef {
  float c[4];
  int s[4];   // tap positions relative to the staged tile (clamped to the attention map like the per-pixel kernel)
};
template <typename T>
__global__ void __launch_bounds__(256)
pan_bicubic_add_tile_kernel(const float* __restrict__ att, int hp, int wp, const T* __restrict__ x, int x_CT,
                            T* __restrict__ y, int y_CT, int nchunks, int H, int W, float gamma, int rows_cap,
                            int cols_cap) {
  extern __shared__ __align__(16) float bc_smem[];
  __shared__ __align__(16) BcCoef cxs[kBcTX];
  __shared__ __align__(16) BcCoef cys[kBcTY];
  __shared__ int s_org[4];
  float* sT = bc_smem;                                        // [rows_cap][kBcTX][8] after the horizontal pass
  float* sS = sT + (size_t)rows_cap * kBcTX * 8;              // [rows_cap][cols_cap][8] staged attention pixels
  const int tid = threadIdx.x;
  const int ch = blockIdx.z % nchunks, b = blockIdx.z / nchunks;
  const int ox0 = blockIdx.x * kBcTX, oy0 = blockIdx.y * kBcTY;
  const int nx = min(kBcTX, W - ox0), ny = min(kBcTY, H - oy0);
  // align_corners=False: src = scale * (dst + 0.5) - 0.5 with scale = in / out, not clamped for cubic
  auto src_of = [](int in, int out, int o, int& i, float& t) {
    const float sf = ((float)in / (float)out) * ((float)o + 0.5f) - 0.5f;
    const float fl = floorf(sf);
    i = (int)fl;
    t = sf - fl;
  };
  if (tid == 0) {
    int i0, i1;
    float t;
    src_of(wp, W, ox0, i0, t);
    src_of(wp, W, ox0 + nx - 1, i1, t);
    const int xa = min(max(i0 - 1, 0), wp - 1), xb = min(max(i1 + 2, 0), wp - 1);
    s_org[0] = xa;
    s_org[2] = xb - xa + 1;
    src_of(hp, H, oy0, i0, t);
    src_of(hp, H, oy0 + ny - 1, i1, t);
    const int ya = min(max(i0 - 1, 0), hp - 1), yb = min(max(i1 + 2, 0), hp - 1);
    s_org[1] = ya;
    s_org[3] = yb - ya + 1;
  }
  __syncthreads();
  const int xa = s_org[0], ya = s_org[1], ncol = s_org[2], nrow = s_org[3];
  if (tid < kBcTX) {
    if (tid < nx) {
      int ix;
      float t;
      src_of(wp, W, ox0 + tid, ix, t);
      cubic_coeffs(t, cxs[tid].c);
#pragma unroll
      for (int k = 0; k < 4; ++k) cxs[tid].s[k] = (min(max(ix - 1 + k, 0), wp - 1) - xa) * 8;
    }
  } else if (tid < kBcTX + kBcTY) {
    const int i = tid - kBcTX;
    if (i < ny) {
      int iy;
      float t;
      src_of(hp, H, oy0 + i, iy, t);
      cubic_coeffs(t, cys[i].c);
#pragma unroll
      for (int k = 0; k < 4; ++k) cys[i].s[k] = (min(max(iy - 1 + k, 0), hp - 1) - ya) * (kBcTX * 8);
    }
  }
  const float* ab = att + (size_t)b * hp * wp * kPanRow + ch * 8;
  for (int i = tid; i < nrow * ncol; i += 256) {
    const int r = i / ncol, c = i - r * ncol;
    float v[8];
    load8(ab + ((size_t)(ya + r) * wp + xa + c) * kPanRow, v);
    float4* d = reinterpret_cast<float4*>(sS + ((size_t)r * cols_cap + c) * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
  for (int i = tid; i < nrow * kBcTX; i += 256) {
    const int r = i / kBcTX, xx = i - r * kBcTX;
    if (xx >= nx) continue;
    const float4 kc = *reinterpret_cast<const float4*>(cxs[xx].c);
    const int4 ks = *reinterpret_cast<const int4*>(cxs[xx].s);
    const float cx[4] = {kc.x, kc.y, kc.z, kc.w};
    const int so[4] = {ks.x, ks.y, ks.z, ks.w};
    const float* srow = sS + (size_t)r * cols_cap * 8;
    float row[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) row[e] = 0.f;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const float4 a = *reinterpret_cast<const float4*>(srow + so[kx]), c2 = *reinterpret_cast<const float4*>(srow + so[kx] + 4);
      const float v[8] = {a.x, a.y, a.z, a.w, c2.x, c2.y, c2.z, c2.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) row[e] = fmaf(cx[kx], v[e], row[e]);
    }
    float4* d = reinterpret_cast<float4*>(sT + ((size_t)r * kBcTX + xx) * 8);
    d[0] = make_float4(row[0], row[1], row[2], row[3]);
    d[1] = make_float4(row[4], row[5], row[6], row[7]);
  }
  __syncthreads();
  for (int i = tid; i < ny * kBcTX; i += 256) {
    const int yy = i / kBcTX, xx = i - yy * kBcTX;
    if (xx >= nx) continue;
    const float4 kc = *reinterpret_cast<const float4*>(cys[yy].c);
    const int4 ks = *reinterpret_cast<const int4*>(cys[yy].s);
    const float cy[4] = {kc.x, kc.y, kc.z, kc.w};
    const int so[4] = {ks.x, ks.y, ks.z, ks.w};
    const float* t = sT + (size_t)xx * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const float4 a = *reinterpret_cast<const float4*>(t + so[ky]), c2 = *reinterpret_cast<const float4*>(t + so[ky] + 4);
      const float v[8] = {a.x, a.y, a.z, a.w, c2.x, c2.y, c2.z, c2.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(cy[ky], v[e], acc[e]);
    }
    const size_t pix = (size_t)(oy0 + yy) * W + ox0 + xx;
    float xv[8];
    load8(x + (((size_t)b * x_CT + ch) * H * W + pix) * 8, xv);
#pragma unroll
    for (int e = 0; e < 8; ++e) xv[e] = fmaf(gamma, acc[e], xv[e]);
    store8(y + (((size_t)b * y_CT + ch) * H * W + pix) * 8, xv);
  }
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

int launch_pan_attention_tc(const float* f, const float* g, const float* hv, int B, int n, int nfp, float* out,
                            cudaStream_t st) {
  if (nfp <= 16) return launch_att_tc<16>(f, g, hv, B, n, nfp, out, st);
  if (nfp <= 32) return launch_att_tc<32>(f, g, hv, B, n, nfp, out, st);
  if (nfp <= 48) return launch_att_tc<48>(f, g, hv, B, n, nfp, out, st);
  if (nfp <= 64) return launch_att_tc<64>(f, g, hv, B, n, nfp, out, st);
  return -2;
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
  // staged attention rows / columns of a tile: the span of the first and last pixel's taps
  const int rows_cap = (int)std::ceil((kBcTY - 1) * (double)hp / H) + 6, cols_cap = (int)std::ceil((kBcTX - 1) * (double)wp / W) + 6;
  const size_t smem = ((size_t)rows_cap * kBcTX + (size_t)rows_cap * cols_cap) * 8 * sizeof(float);
  static const int tiled = getenv("INNFER_PAN_BICUBIC_TILED") ? atoi(getenv("INNFER_PAN_BICUBIC_TILED")) : 1;   // A/B switch
  if (tiled && hp <= H && wp <= W && smem <= 48 * 1024 && (long long)B * nchunks <= 65535) {
    dim3 grid((unsigned)((W + kBcTX - 1) / kBcTX), (unsigned)((H + kBcTY - 1) / kBcTY), (unsigned)(B * nchunks));
    pan_bicubic_add_tile_kernel<T><<<grid, 256, smem, st>>>(att, hp, wp, x, x_CT, y, y_CT, nchunks, H, W, gamma, rows_cap,
                                                            cols_cap);
    return (int)cudaGetLastError();
  }
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
