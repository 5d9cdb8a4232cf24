// pix2pix UnetGenerator / CycleGAN ResnetGenerator on sm_100a.  See i2i.cuh for the design notes.
//
// Reference semantics restated here (file:line relative to the reference tree):
//   architectures/UNet_arch.py:98-165   UnetSkipConnectionBlock: [LeakyReLU(0.2, inplace), Conv2d(4, 2, 1), norm,
//                                       submodule, ReLU(inplace), ConvTranspose2d(4, 2, 1), norm], cat([x, model(x)])
//   architectures/ResNet_arch.py:55-91  ReflectionPad2d(3) + Conv2d(7), two Conv2d(3, 2, 1), ResnetBlocks,
//                                       two ConvTranspose2d(3, 2, 1, output_padding 1), ReflectionPad2d(3) + Conv2d(7), Tanh
//   architectures/ResNet_arch.py:113-151 ResnetBlock: x + [pad, conv, norm, ReLU, pad, conv, norm](x)
//   torch.nn.BatchNorm2d / InstanceNorm2d: eps 1e-5, biased variance; BatchNorm in training mode (run.py:297,
//   pix2pix_extras meval False) normalises with the statistics of the batch it is given.
#include "i2i.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "ptx.cuh"

namespace innfer {

namespace {

constexpr int kGenThreads = 160;            // warps 0..3: operand gather, then epilogue; warp 4: TMEM owner + MMA issuer
constexpr int kGenABytes = 8 * 128 * 16;    // one K-step of A: 8 K-chunks x 128 pixels x 16 bytes
constexpr int kGenTailBytes = 256;          // barriers + TMEM slot behind the stage ring
constexpr float kNormEps = 1e-5f;
constexpr float kLreluSlope = 0.2f;

__host__ __device__ constexpr int gen_stages(int NT) { return NT >= 128 ? 3 : 4; }

struct GenConvParams {
  const void* in;            // [B][*][Hin][Win][8] fp16 (tensor-core kernel) or fp32 (direct kernel)
  long long in_bs, in_cs;    // elements per image / per chunk plane
  int Hin, Win;
  void* out;
  long long out_bs, out_cs;
  long long split_stride;    // elements between split-K partial tensors (raw fp32 only)
  int Hout, Wout;
  int out_mode;              // 0: chunks of the kernel's element type, 1: raw fp32 chunks, 2: compact [B][H][W][4] fp16
  int out_chunk0, out_nchunks;
  int B, nphase, nsplit;
  int Hp[kGenMaxPhases], Wp[kGenMaxPhases], Mp[kGenMaxPhases];
  int py[kGenMaxPhases], px[kGenMaxPhases], ntaps[kGenMaxPhases], ksteps[kGenMaxPhases];
  long long woff[kGenMaxPhases];   // element offset of the phase inside w
  int ostep, istep;
  int cin_chunks, in_chunk0, cout_pad;
  int reflect, act;
  const void* w;
  const float* bias;
  int8_t offy[kGenMaxPhases][kGenMaxTaps], offx[kGenMaxPhases][kGenMaxTaps];
  // halo-tile kernel only
  int J, nslabs, tmem_cols;
  unsigned stage_bytes;
  int bands[kGenMaxPhases], cps[kGenMaxPhases], nplanes[kGenMaxPhases], Rpl[kGenMaxPhases], Wpl[kGenMaxPhases];
  int8_t pl_py[kGenMaxPhases][4], pl_px[kGenMaxPhases][4], pl_r0[kGenMaxPhases][4], pl_c0[kGenMaxPhases][4];
  uint16_t tap_aoff[kGenMaxPhases][kGenMaxTaps];   // 16-byte units inside one K-chunk block of the gathered tile
  // fused normalisation statistics (raw outputs only): per CTA tile the per-channel sum and sum of squares of its valid
  // pixels go to stats[((b * out_nchunks + chunk) * stat_slices + slice) * 16 + {e, 8 + e}], slice = slice0[phase] + tile
  double* stats;
  int stat_slices, slice0[kGenMaxPhases];
  int split_for_halo;   // nsplit was chosen for the halo-tile kernel (else a split launch belongs to the im2col kernel)
  int debug;   // INNFER_I2I_DEBUG bit mask for timing experiments (wrong results): 1 no MMA, 2 no A gather, 4 no weight copy
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case kActRelu: return fmaxf(v, 0.f);
    case kActLrelu: return v > 0.f ? v : v * kLreluSlope;
    case kActTanh: return tanhf(v);
    case kActTanh01: return 0.5f * tanhf(v) + 0.5f;
    default: return v;
  }
}

// the activations a norm layer can be followed by (no tanh: 16 inlined copies of it were most of the apply kernel's code)
__device__ __forceinline__ float apply_act_light(float v, int act) {
  return act == kActRelu ? fmaxf(v, 0.f) : (act == kActLrelu ? (v > 0.f ? v : v * kLreluSlope) : v);
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival (counted in its init value) when all cp.async copies this thread has issued so far
// have landed: the gather threads never block on their own copies
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ int reflect_index(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// ------------------------------------------------------------------------------------------------ tensor-core conv
template <int NT, bool RAW>
__global__ void __launch_bounds__(kGenThreads) gen_conv_tc_kernel(const __grid_constant__ GenConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int S = gen_stages(NT);
  constexpr uint32_t B_BYTES = 8u * NT * 16u;
  constexpr uint32_t STAGE = kGenABytes + B_BYTES;
  constexpr uint32_t TCOLS = NT < 32 ? 32u : (uint32_t)NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ph = (int)blockIdx.z / p.nsplit, split = (int)blockIdx.z - ph * p.nsplit;
  const int Mp = p.Mp[ph];
  if ((int)blockIdx.x * 128 >= Mp) return;       // phases of a transposed conv differ by at most one row / column
  const int KS = p.ksteps[ph];
  const int k0 = (int)((long long)KS * split / p.nsplit), k1 = (int)((long long)KS * (split + 1) / p.nsplit);
  const int nk = k1 - k0;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 128 + 1);   // every gather thread + the weight copy's expect_tx
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tfull_bar), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), TCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp < 4) {
    // ------------------------------------------------------------ gather: thread <-> output pixel (row of the M tile)
    const int tid = threadIdx.x;
    const int m = (int)blockIdx.x * 128 + tid;
    const bool valid = m < Mp;
    const int Wp = p.Wp[ph], HWp = p.Hp[ph] * Wp;
    int b = 0, oyp = 0, oxp = 0;
    if (valid) {
      b = m / HWp;
      const int r = m - b * HWp;
      oyp = r / Wp;
      oxp = r - oyp * Wp;
    }
    const int iy0 = oyp * p.istep, ix0 = oxp * p.istep;
    const __half* inb = reinterpret_cast<const __half*>(p.in) + (size_t)b * p.in_bs + (size_t)p.in_chunk0 * p.in_cs;
    const int ntaps = p.ntaps[ph];
    const int8_t* offy = p.offy[ph];
    const int8_t* offx = p.offx[ph];
    int tap = (k0 * 8) / p.cin_chunks;
    int chunk = k0 * 8 - tap * p.cin_chunks;
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(reinterpret_cast<const __half*>(p.w) + p.woff[ph]) +
                          ((size_t)blockIdx.y * KS + k0) * B_BYTES;
    for (int i = 0; i < nk; ++i) {
      const int s = i % S;
      mbar_wait(smem_u32(&empty_bar[s]), (((uint32_t)(i / S)) & 1u) ^ 1u);
      const uint32_t a_dst = smem_base + (uint32_t)s * STAGE;
      if (tid == 0) {
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, B_BYTES);
        bulk_load(a_dst + kGenABytes, wsrc + (size_t)i * B_BYTES, B_BYTES, fb);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const void* src = p.in;
        uint32_t sz = 0;
        if (valid && tap < ntaps) {
          int iy = iy0 + offy[tap], ix = ix0 + offx[tap];
          if (p.reflect) {
            iy = reflect_index(iy, p.Hin);
            ix = reflect_index(ix, p.Win);
          }
          if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) {
            src = inb + (size_t)chunk * p.in_cs + ((size_t)iy * p.Win + ix) * 8;
            sz = 16;
          }
        }
        cp_async_16(a_dst + (uint32_t)j * 2048u + (uint32_t)tid * 16u, src, sz);   // sz 0: zero fill (padding, K tail)
        if (++chunk == p.cin_chunks) {
          chunk = 0;
          ++tap;
        }
      }
      cp_async_mbar_arrive(smem_u32(&full_bar[s]));   // arrives when this thread's copies of the K-step have landed
    }
    cp_async_wait_all();

    // ------------------------------------------------------------ epilogue: TMEM lane = pixel row = this thread
    mbar_wait(smem_u32(tfull_bar), 0u);
    tc_fence_after();
    const int oy = oyp * p.ostep + p.py[ph], ox = oxp * p.ostep + p.px[ph];
    const size_t pix = (size_t)oy * p.Wout + ox;
    const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tacc + (uint32_t)c0, v);
      tmem_ld_wait();
      if (!valid) continue;
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int gc = (int)blockIdx.y * (NT / 8) + c0 / 8 + hh;
        if (gc >= p.out_nchunks) continue;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = __uint_as_float(hh ? v[8 + e] : v[e]);
          if (split == 0) t += p.bias[gc * 8 + e];
          f[e] = RAW ? t : apply_act(t, p.act);
        }
        if constexpr (RAW) {
          float* op = reinterpret_cast<float*>(p.out) + (size_t)split * p.split_stride + (size_t)b * p.out_bs +
                      (size_t)(p.out_chunk0 + gc) * p.out_cs + pix * 8;
          *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
        } else {
          __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
          __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
          if (p.out_mode == 0) {
            __half* op = reinterpret_cast<__half*>(p.out) + (size_t)b * p.out_bs + (size_t)(p.out_chunk0 + gc) * p.out_cs + pix * 8;
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0);
            o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2);
            o.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(op) = o;
          } else if (gc == 0) {
            __half* op = reinterpret_cast<__half*>(p.out) + ((size_t)b * p.Hout * p.Wout + pix) * 4;
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0);
            o.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(op) = o;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(NT);
    for (int i = 0; i < nk; ++i) {
      const int s = i % S;
      mbar_wait(smem_u32(&full_bar[s]), ((uint32_t)(i / S)) & 1u);
      fence_proxy_async_smem();   // the gather's generic-proxy writes (visible through the barrier) -> async-proxy reads
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_base + (uint32_t)s * STAGE;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // A: [K-chunk][pixel][8]: 8-row groups 128 B apart (SBO), K-chunks 2048 B apart (LBO); B: [K-chunk][n][8]
          const uint64_t adesc = make_smem_desc(sa + (uint32_t)kk * 4096u, 2048u, 128u);
          const uint64_t bdesc = make_smem_desc(sa + kGenABytes + (uint32_t)kk * 2u * NT * 16u, (uint32_t)NT * 16u, 128u);
          umma_f16_ss(tmem_base, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));   // the stage is free once these MMAs have read it
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(smem_u32(tfull_bar));
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}


// ------------------------------------------------------------------------------------------------ halo-tile conv
// The same implicit GEMM with the A operand staged ONCE per 16-channel slab: the CTA's input tile (16 + k - 1 rows of
// 8J + k - 1 pixels; for stride 2 four parity planes of it, so that every tap reads unit-stride pixels) is gathered into
// shared memory -- reflection / zero padding resolved by a per-CTA table of source offsets -- and every tap is a shifted
// SWIZZLE_NONE descriptor into it (rows of a core matrix = 8 consecutive pixels, SBO = tile row pitch, LBO = K-chunk
// block).  M = 128 is a 16 x 8 pixel sub-patch; J sub-patches side by side share the tile and the weights.  Against the
// im2col gather above this moves k*k/1.4 times fewer activation bytes (3x3: 6.4x, 7x7: 20x).
// RAW: fp32 output for a norm layer (no activation, optional fused statistics); !RAW: a network's last layer (bias +
// activation, fp16 chunks or compact tile pixels).  Two kernels because tanh inlined 64 times into one shared epilogue
// made the function 214 KB of code -- an instruction-cache problem, not an arithmetic one.
template <int NT, int S, bool RAW>
__global__ void __launch_bounds__(kGenThreads) gen_conv_halo_kernel(const __grid_constant__ GenConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-K (raw outputs only): the 16-channel slabs [i0, i0 + nslabs) of this CTA go to partial tensor `split`
  const int ph = (int)blockIdx.z / p.nsplit, split = (int)blockIdx.z - ph * p.nsplit;
  const int bands = p.bands[ph], cps = p.cps[ph];
  int t = blockIdx.x;
  if (t >= p.B * bands * cps) return;
  const int cp = t % cps;
  t /= cps;
  const int band = t % bands, b = t / bands;
  const int y0 = band * 16, x0 = cp * 8 * p.J;
  const int Hp = p.Hp[ph], Wp = p.Wp[ph];
  const int rem = (Wp - x0 + 7) >> 3;
  const int jeff = rem < p.J ? rem : p.J;
  const int Rpl = p.Rpl[ph], Wpl = p.Wpl[ph];
  const int Npos = p.nplanes[ph] * Rpl * Wpl;
  const int ntaps = p.ntaps[ph];
  const uint32_t A_BYTES = (2u * (uint32_t)Npos * 16u + 127u) & ~127u;
  const uint32_t B_BYTES = (uint32_t)ntaps * 2u * NT * 16u;
  const uint32_t STAGE = p.stage_bytes;
  const int i0 = (int)((long long)p.nslabs * split / p.nsplit);
  const int nslabs = (int)((long long)p.nslabs * (split + 1) / p.nsplit) - i0;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  float* red = reinterpret_cast<float*>(smem + (size_t)S * STAGE + 128);            // [2][4 warps][32]: statistics exchange
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)S * STAGE + 128 + 1024);   // [NT] bias of this N tile
  int32_t* tbl = reinterpret_cast<int32_t*>(smem + (size_t)S * STAGE + 128 + 1024 + 512);
  if (threadIdx.x < NT) s_bias[threadIdx.x] = (p.debug & 32) ? 0.f : p.bias[blockIdx.y * NT + threadIdx.x];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 128 + 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tfull_bar), 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  // source offset of every position of the gathered tile (elements from the image's chunk plane; -1: zero)
  for (int pos = threadIdx.x; pos < Npos; pos += kGenThreads) {
    const int pl = pos / (Rpl * Wpl);
    const int rr = pos - pl * Rpl * Wpl;
    const int r = rr / Wpl, c = rr - r * Wpl;
    int iy = p.istep * (y0 + r + p.pl_r0[ph][pl]) + p.pl_py[ph][pl];
    int ix = p.istep * (x0 + c + p.pl_c0[ph][pl]) + p.pl_px[ph][pl];
    if (p.reflect) {
      iy = reflect_index(iy, p.Hin);
      ix = reflect_index(ix, p.Win);
    }
    tbl[pos] = ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) ? (iy * p.Win + ix) * 8 : -1;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (warp < 4) {
    // ------------------------------------------------------------ gather
    const int tid = threadIdx.x;
    const __half* inb = reinterpret_cast<const __half*>(p.in) + (size_t)b * p.in_bs + (size_t)p.in_chunk0 * p.in_cs;
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(reinterpret_cast<const __half*>(p.w) + p.woff[ph]) +
                          ((size_t)blockIdx.y * p.nslabs + i0) * B_BYTES;
    for (int i = 0; i < nslabs; ++i) {
      const int s = i % S;
      mbar_wait(smem_u32(&empty_bar[s]), (((uint32_t)(i / S)) & 1u) ^ 1u);
      const uint32_t a_dst = smem_base + (uint32_t)s * STAGE;
      if (tid == 0) {
        const uint32_t fb = smem_u32(&full_bar[s]);
        if (p.debug & 4) {
          mbar_arrive(fb);
        } else {
          mbar_expect_tx(fb, B_BYTES);
          bulk_load(a_dst + A_BYTES, wsrc + (size_t)i * B_BYTES, B_BYTES, fb);
        }
      }
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        const int chunk = 2 * (i0 + i) + kc;
        const bool cv = chunk < p.cin_chunks;   // odd chunk counts: the second K-chunk of the last slab is zeros
        const __half* cb = inb + (size_t)chunk * p.in_cs;
        const uint32_t dst = a_dst + (uint32_t)kc * (uint32_t)Npos * 16u;
        for (int pos = tid; pos < ((p.debug & 2) ? 0 : Npos); pos += 128) {
          const int off = tbl[pos];
          const bool ok = cv && off >= 0;
          cp_async_16(dst + (uint32_t)pos * 16u, ok ? (const void*)(cb + off) : p.in, ok ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive(smem_u32(&full_bar[s]));   // arrives when this thread's copies have landed
    }
    cp_async_wait_all();

    // ------------------------------------------------------------ epilogue: lane m = pixel (m / 8, m % 8) of a sub-patch
    mbar_wait(smem_u32(tfull_bar), 0u);
    tc_fence_after();
    const int r = tid >> 3, c = tid & 7;
    const int oyp = y0 + r;
    const int oy = oyp * p.ostep + p.py[ph];
    const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16);
    if constexpr (RAW) {
      const bool want_stats = p.stats != nullptr && !(p.debug & 8);
      int grp = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16, ++grp) {
        float sq[32];   // [0, 16): sums of this thread's pixels (one per sub-patch), [16, 32): sums of squares
#pragma unroll
        for (int e = 0; e < 32; ++e) sq[e] = 0.f;
        uint32_t vv[4][16];   // the group's 16 columns of every sub-patch: one TMEM round trip instead of J
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < jeff) tmem_ld16(tacc + (uint32_t)(j * NT + c0), vv[j]);
        tmem_ld_wait();
        float bias16[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) bias16[e] = split == 0 ? s_bias[c0 + e] : 0.f;
        const int gc0 = (int)blockIdx.y * (NT / 8) + c0 / 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= jeff) continue;
          const int oxp = x0 + 8 * j + c;
          if (!(oyp < Hp && oxp < Wp)) continue;
          const size_t pix = (size_t)oy * p.Wout + (size_t)(oxp * p.ostep + p.px[ph]);
          float* op = reinterpret_cast<float*>(p.out) + (size_t)split * p.split_stride + (size_t)b * p.out_bs +
                      (size_t)(p.out_chunk0 + gc0) * p.out_cs + pix * 8;
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            f[e] = __uint_as_float(vv[j][e]) + bias16[e];
            sq[e] += f[e];
            sq[16 + e] = fmaf(f[e], f[e], sq[16 + e]);
          }
          if (!(p.debug & 16)) {
            if (gc0 < p.out_nchunks) {
              *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
            }
            if (gc0 + 1 < p.out_nchunks) {
              *reinterpret_cast<float4*>(op + p.out_cs) = make_float4(f[8], f[9], f[10], f[11]);
              *reinterpret_cast<float4*>(op + p.out_cs + 4) = make_float4(f[12], f[13], f[14], f[15]);
            }
          }
        }
        if (want_stats) {
          // transpose-reduce: 32 values per lane -> lane l holds the warp's total of value l (31 shuffles, fixed order)
#pragma unroll
          for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
              const float send = upper ? sq[i] : sq[i + n];
              const float keep = upper ? sq[i + n] : sq[i];
              sq[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          float* rb = red + (grp & 1) * 128;
          rb[warp * 32 + lane] = sq[0];
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (warp == 0) {
            const float tot = ((rb[lane] + rb[32 + lane]) + rb[64 + lane]) + rb[96 + lane];
            const int ch = (int)blockIdx.y * NT + c0 + (lane & 15);
            const int chunk = ch >> 3;
            if (chunk < p.out_nchunks) {
              const int slice = p.slice0[ph] + band * cps + cp;
              p.stats[(((size_t)b * p.out_nchunks + chunk) * p.stat_slices + slice) * 16 + (lane < 16 ? 0 : 8) + (ch & 7)] = (double)tot;
            }
          }
        }
      }
    } else {
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 16) {
#pragma unroll 1
        for (int j = 0; j < jeff; ++j) {
          uint32_t v[16];
          tmem_ld16(tacc + (uint32_t)(j * NT + c0), v);
          tmem_ld_wait();
          const int oxp = x0 + 8 * j + c;
          if (!(oyp < Hp && oxp < Wp)) continue;
          const size_t pix = (size_t)oy * p.Wout + (size_t)(oxp * p.ostep + p.px[ph]);
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            const int gc = (int)blockIdx.y * (NT / 8) + c0 / 8 + hh;
            if (gc >= p.out_nchunks) continue;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              f[e] = apply_act(__uint_as_float(hh ? v[8 + e] : v[e]) + s_bias[c0 + hh * 8 + e], p.act);
            __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
            __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
            if (p.out_mode == 0) {
              __half* op = reinterpret_cast<__half*>(p.out) + (size_t)b * p.out_bs + (size_t)(p.out_chunk0 + gc) * p.out_cs + pix * 8;
              uint4 o;
              o.x = *reinterpret_cast<uint32_t*>(&h0);
              o.y = *reinterpret_cast<uint32_t*>(&h1);
              o.z = *reinterpret_cast<uint32_t*>(&h2);
              o.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(op) = o;
            } else if (gc == 0) {
              __half* op = reinterpret_cast<__half*>(p.out) + ((size_t)b * p.Hout * p.Wout + pix) * 4;
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&h0);
              o.y = *reinterpret_cast<uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(op) = o;
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ MMA issuer: taps x sub-patches per 16-channel slab
    const uint32_t idesc = make_idesc_f16(NT);
    const uint32_t lbo = (uint32_t)Npos * 16u, sbo = (uint32_t)Wpl * 16u;
    for (int i = 0; i < nslabs; ++i) {
      const int s = i % S;
      mbar_wait(smem_u32(&full_bar[s]), ((uint32_t)(i / S)) & 1u);
      fence_proxy_async_smem();   // the gather's generic-proxy writes (visible through the barrier) -> async-proxy reads
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_base + (uint32_t)s * STAGE;
        const uint32_t sb = sa + A_BYTES;
        for (int tp = 0; tp < ntaps; ++tp) {
          const uint32_t aoff = (uint32_t)p.tap_aoff[ph][tp] * 16u;
          const uint64_t bdesc = make_smem_desc(sb + (uint32_t)tp * 2u * NT * 16u, (uint32_t)NT * 16u, 128u);
          for (int j = 0; j < ((p.debug & 1) ? 0 : jeff); ++j) {
            const uint64_t adesc = make_smem_desc(sa + aoff + (uint32_t)j * 128u, lbo, sbo);
            umma_f16_ss(tmem_base + (uint32_t)(j * NT), adesc, bdesc, idesc, (i | tp) != 0 ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(smem_u32(tfull_bar));
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ fp32 direct conv
// -no_fp16 mode (correctness mode, CUDA cores): thread <-> (output pixel, 8 output channels), FMAs in a fixed order.
__global__ void __launch_bounds__(128) gen_conv_direct_kernel(const __grid_constant__ GenConvParams p) {
  const int ph = blockIdx.z, oc = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.Mp[ph]) return;
  const int Wp = p.Wp[ph], HWp = p.Hp[ph] * Wp;
  const int b = m / HWp;
  const int r = m - b * HWp;
  const int oyp = r / Wp, oxp = r - oyp * Wp;
  const float* inb = reinterpret_cast<const float*>(p.in) + (size_t)b * p.in_bs + (size_t)p.in_chunk0 * p.in_cs;
  const float* wbase = reinterpret_cast<const float*>(p.w) + p.woff[ph] + oc * 8;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = p.bias[oc * 8 + o];
  for (int t = 0; t < p.ntaps[ph]; ++t) {
    int iy = oyp * p.istep + p.offy[ph][t], ix = oxp * p.istep + p.offx[ph][t];
    if (p.reflect) {
      iy = reflect_index(iy, p.Hin);
      ix = reflect_index(ix, p.Win);
    }
    if ((unsigned)iy >= (unsigned)p.Hin || (unsigned)ix >= (unsigned)p.Win) continue;
    const float* ip = inb + ((size_t)iy * p.Win + ix) * 8;
    const float* wt = wbase + (size_t)t * p.cin_chunks * 8 * p.cout_pad;
    for (int c = 0; c < p.cin_chunks; ++c) {
      const float4 a0 = *reinterpret_cast<const float4*>(ip + (size_t)c * p.in_cs);
      const float4 a1 = *reinterpret_cast<const float4*>(ip + (size_t)c * p.in_cs + 4);
      const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float* wr = wt + (size_t)(c * 8 + e) * p.cout_pad;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + 4));
        acc[0] = fmaf(x[e], w0.x, acc[0]);
        acc[1] = fmaf(x[e], w0.y, acc[1]);
        acc[2] = fmaf(x[e], w0.z, acc[2]);
        acc[3] = fmaf(x[e], w0.w, acc[3]);
        acc[4] = fmaf(x[e], w1.x, acc[4]);
        acc[5] = fmaf(x[e], w1.y, acc[5]);
        acc[6] = fmaf(x[e], w1.z, acc[6]);
        acc[7] = fmaf(x[e], w1.w, acc[7]);
      }
    }
  }
  const int oy = oyp * p.ostep + p.py[ph], ox = oxp * p.ostep + p.px[ph];
  float* op = reinterpret_cast<float*>(p.out) + (size_t)b * p.out_bs + (size_t)(p.out_chunk0 + oc) * p.out_cs +
              ((size_t)oy * p.Wout + ox) * 8;
  *reinterpret_cast<float4*>(op) = make_float4(apply_act(acc[0], p.act), apply_act(acc[1], p.act), apply_act(acc[2], p.act),
                                               apply_act(acc[3], p.act));
  *reinterpret_cast<float4*>(op + 4) = make_float4(apply_act(acc[4], p.act), apply_act(acc[5], p.act),
                                                   apply_act(acc[6], p.act), apply_act(acc[7], p.act));
}

// ------------------------------------------------------------------------------------------------ normalisation
// The value of a raw pixel chunk: split-K partial tensors summed in index order (the same order everywhere).
__device__ __forceinline__ void load_raw8(const float* p, int nsplit, long long split_stride, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 c = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  for (int s = 1; s < nsplit; ++s) {
    const float4 d = *reinterpret_cast<const float4*>(p + (size_t)s * split_stride);
    const float4 e = *reinterpret_cast<const float4*>(p + (size_t)s * split_stride + 4);
    v[0] += d.x; v[1] += d.y; v[2] += d.z; v[3] += d.w;
    v[4] += e.x; v[5] += e.y; v[6] += e.z; v[7] += e.w;
  }
}

// grid (slices, chunks, B), 256 threads: per-channel sum and sum of squares of one slice of pixels, in double, reduced
// in a fixed order.  part: [B][chunks][slices][16] (8 sums, 8 sums of squares).
__global__ void __launch_bounds__(256) norm_stats_kernel(const float* __restrict__ raw, int nsplit, long long split_stride,
                                                         long long bs, long long cs, int HW, double* __restrict__ part) {
  const int slice = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z, nslices = gridDim.x;
  const float* base = raw + (size_t)b * bs + (size_t)chunk * cs;
  const int len = (HW + nslices - 1) / nslices;
  const int start = slice * len, end = min(HW, start + len);
  double s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.0;
  for (int px = start + (int)threadIdx.x; px < end; px += 256) {
    float v[8];
    load_raw8(base + (size_t)px * 8, nsplit, split_stride, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += (double)v[e];
      q[e] += (double)v[e] * (double)v[e];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += __shfl_xor_sync(0xffffffffu, s[e], off);
      q[e] += __shfl_xor_sync(0xffffffffu, q[e], off);
    }
  __shared__ double sh[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sh[warp][e] = s[e];
      sh[warp][8 + e] = q[e];
    }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    part[(((size_t)b * gridDim.y + chunk) * nslices + slice) * 16 + threadIdx.x] = t;
  }
}

// one 128-thread block per (image, channel) [per_sample] or per channel [batch statistics]: threads take the partial sums
// in a strided order, then a shuffle tree and a 4-term sum -- a fixed summation order, so the statistics are
// bit-reproducible.  (scale, shift) of y = x * scale + shift.
__global__ void __launch_bounds__(128) norm_finalize_kernel(const double* __restrict__ part, int B, int chunks, int nslices, int HW,
                                                            int per_sample, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float2* __restrict__ ss) {
  const int C8 = chunks * 8;
  const int idx = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = idx % C8, b0 = per_sample ? idx / C8 : 0, b1 = per_sample ? b0 + 1 : B;
  double S = 0.0, Q = 0.0;
  const int n_part = (b1 - b0) * nslices;
  for (int i = threadIdx.x; i < n_part; i += 128) {
    const int b = b0 + i / nslices, sl = i - (i / nslices) * nslices;
    const double* pp = part + (((size_t)b * chunks + c / 8) * nslices + sl) * 16;
    S += pp[c % 8];
    Q += pp[8 + c % 8];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    S += __shfl_xor_sync(0xffffffffu, S, off);
    Q += __shfl_xor_sync(0xffffffffu, Q, off);
  }
  __shared__ double sh[4][2];
  if (lane == 0) {
    sh[warp][0] = S;
    sh[warp][1] = Q;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  S = ((sh[0][0] + sh[1][0]) + sh[2][0]) + sh[3][0];
  Q = ((sh[0][1] + sh[1][1]) + sh[2][1]) + sh[3][1];
  const double n = (double)(b1 - b0) * (double)HW;
  const double mean = S / n;
  double var = Q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const double rstd = 1.0 / sqrt(var + (double)kNormEps);
  const double g = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
  const float2 o = make_float2((float)(g * rstd), (float)(be - mean * g * rstd));
  for (int b = b0; b < b1; ++b) ss[(size_t)b * C8 + c] = o;
}

struct ApplyParams {
  const float* raw;
  int nsplit;
  long long split_stride, raw_bs, raw_cs;
  int B, chunks, HW;
  const float2* ss;    // null: identity
  int ss_per_sample;   // 1: [B][C], 0: [C]
  void* out_a;
  long long a_bs, a_cs;
  int act_a;
  void* out_b;         // optional second copy with its own activation
  long long b_bs, b_cs;
  int act_b;
  const void* res;     // optional residual, added before the activation
  long long r_bs, r_cs;
};

template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8], int act);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&f)[8], int act) {
  *reinterpret_cast<float4*>(p) = make_float4(apply_act_light(f[0], act), apply_act_light(f[1], act), apply_act_light(f[2], act), apply_act_light(f[3], act));
  *reinterpret_cast<float4*>(p + 4) = make_float4(apply_act_light(f[4], act), apply_act_light(f[5], act), apply_act_light(f[6], act), apply_act_light(f[7], act));
}
template <>
__device__ __forceinline__ void store8<__half>(__half* p, const float (&f)[8], int act) {
  __half2 h0 = __floats2half2_rn(apply_act_light(f[0], act), apply_act_light(f[1], act));
  __half2 h1 = __floats2half2_rn(apply_act_light(f[2], act), apply_act_light(f[3], act));
  __half2 h2 = __floats2half2_rn(apply_act_light(f[4], act), apply_act_light(f[5], act));
  __half2 h3 = __floats2half2_rn(apply_act_light(f[6], act), apply_act_light(f[7], act));
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&h0);
  o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2);
  o.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), c = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = c.x; f[5] = c.y; f[6] = c.z; f[7] = c.w;
}
__device__ __forceinline__ void load8(const __half* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) norm_apply_kernel(const __grid_constant__ ApplyParams p) {
  const long long total = (long long)p.B * p.chunks * p.HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(idx % p.HW);
    const long long r = idx / p.HW;
    const int chunk = (int)(r % p.chunks), b = (int)(r / p.chunks);
    float v[8];
    load_raw8(p.raw + (size_t)b * p.raw_bs + (size_t)chunk * p.raw_cs + (size_t)px * 8, p.nsplit, p.split_stride, v);
    if (p.ss) {
      const float2* ss = p.ss + ((size_t)(p.ss_per_sample ? b : 0) * p.chunks + chunk) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 t = ss[e];
        v[e] = fmaf(v[e], t.x, t.y);
      }
    }
    if (p.res) {
      float rv[8];
      load8(reinterpret_cast<const T*>(p.res) + (size_t)b * p.r_bs + (size_t)chunk * p.r_cs + (size_t)px * 8, rv);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += rv[e];
    }
    store8<T>(reinterpret_cast<T*>(p.out_a) + (size_t)b * p.a_bs + (size_t)chunk * p.a_cs + (size_t)px * 8, v, p.act_a);
    if (p.out_b)
      store8<T>(reinterpret_cast<T*>(p.out_b) + (size_t)b * p.b_bs + (size_t)chunk * p.b_cs + (size_t)px * 8, v, p.act_b);
  }
}

// eval-mode BatchNorm: (scale, shift) from the running statistics
std::vector<float> running_scale_shift(const float* gamma, const float* beta, const float* mean, const float* var, int C) {
  std::vector<float> ss((size_t)2 * C);
  for (int c = 0; c < C; ++c) {
    const double rstd = 1.0 / std::sqrt((double)var[c] + (double)kNormEps);
    const double sc = (double)gamma[c] * rstd;
    ss[2 * c] = (float)sc;
    ss[2 * c + 1] = (float)((double)beta[c] - (double)mean[c] * sc);
  }
  return ss;
}

template <int NT, bool RAW>
cudaError_t launch_tc_r(const GenConvParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)gen_stages(NT) * (kGenABytes + 8 * NT * 16) + kGenTailBytes;
  static bool attr_set[64] = {};   // once per kernel and device (not per launch: launches may be recorded into a CUDA graph)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gen_conv_tc_kernel<NT, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  gen_conv_tc_kernel<NT, RAW><<<grid, kGenThreads, smem, st>>>(p);
  return cudaGetLastError();
}
template <int NT>
cudaError_t launch_tc(const GenConvParams& p, dim3 grid, cudaStream_t st) {
  return p.out_mode == 1 ? launch_tc_r<NT, true>(p, grid, st) : launch_tc_r<NT, false>(p, grid, st);
}

template <int NT, int S, bool RAW>
cudaError_t launch_halo_r(const GenConvParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  static bool attr_set[64] = {};   // once per kernel and device, at the largest size any launch asks for
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gen_conv_halo_kernel<NT, S, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(227 * 1024));
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  gen_conv_halo_kernel<NT, S, RAW><<<grid, kGenThreads, smem, st>>>(p);
  return cudaGetLastError();
}
template <int NT, int S>
cudaError_t launch_halo_s(const GenConvParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  return p.out_mode == 1 ? launch_halo_r<NT, S, true>(p, grid, smem, st) : launch_halo_r<NT, S, false>(p, grid, smem, st);
}
template <int NT>
cudaError_t launch_halo(const GenConvParams& p, int stages, dim3 grid, size_t smem, cudaStream_t st) {
  switch (stages) {
    case 1: return launch_halo_s<NT, 1>(p, grid, smem, st);
    case 2: return launch_halo_s<NT, 2>(p, grid, smem, st);
    case 3: return launch_halo_s<NT, 3>(p, grid, smem, st);
    default: return launch_halo_s<NT, 4>(p, grid, smem, st);
  }
}

std::atomic<uint64_t> g_halo_launches{0};
std::atomic<uint64_t> g_graph_replays{0};
constexpr size_t kSmemMax = 227 * 1024;
constexpr size_t kSmemHalf = 113 * 1024;   // two CTAs per SM: one CTA's epilogue overlaps the other's main loop

}  // namespace

uint64_t i2i_halo_launches() { return g_halo_launches.load(); }
uint64_t i2i_graph_replays() { return g_graph_replays.load(); }

// ==================================================================================================== host side
I2INet::~I2INet() {
  drop_graphs();
  if (cap_) cudaStreamDestroy(cap_);
  for (void* p : owned_) cudaFree(p);
  for (auto* v : {&dbuf_, &cat_, &rb_})
    for (auto& b : *v)
      if (b.p) cudaFree(b.p);
  for (Buf* b : {&raw_, &inner_, &a_, &b_, &stats_, &ss_})
    if (b->p) cudaFree(b->p);
}

int I2INet::need(Buf& b, size_t bytes) {
  if (bytes <= b.bytes) return 0;
  if (capturing_) {   // cannot happen: a shape is captured only after an eager run of the same shape sized every buffer
    err_ = "internal: workspace growth during graph capture";
    return -3;
  }
  drop_graphs();      // recorded launches point into the buffers that are about to move
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  const size_t want = (bytes + 255) & ~(size_t)255;
  if (cudaMalloc(&b.p, want) != cudaSuccess) {
    cudaGetLastError();
    err_ = "workspace allocation of " + std::to_string(want) + " bytes failed";
    return -5;
  }
  b.bytes = want;
  return 0;
}

int I2INet::check_size(int H, int W, std::string& err) const {
  if (cfg_.kind == 2) {
    if (down_[0].out_h(H) < 1 || down_[0].out_h(W) < 1 || (down_[0].reflect && (H <= down_[0].pad || W <= down_[0].pad))) {
      err = "input too small for this convolution";
      return -1;
    }
    return 0;
  }
  if (cfg_.kind == 0) {
    const int g = 1 << cfg_.depth;
    if (H % g || W % g) {
      err = "UnetGenerator with num_downs " + std::to_string(cfg_.depth) + " needs image sizes that are multiples of " +
            std::to_string(g) + " (run.py:412-413 resizes to multiples of the -a size for pix2pix)";
      return -1;
    }
  } else if (H % 4 || W % 4 || H < 8 || W < 8) {
    err = "ResnetGenerator needs image (tile) sizes that are multiples of 4, at least 8";
    return -1;
  }
  return 0;
}

int I2INet::build_conv(const ParamLookup& get, const std::string& name, GenConv& L, int Cout, int Cin, int k, int stride,
                       int pad, bool transposed, int out_pad, bool reflect, bool has_bias, std::string& err, size_t* consumed) {
  std::vector<int64_t> ws, bshape;
  const float* w = get(name + ".weight", ws);
  if (!w) {
    err = "missing key " + name + ".weight";
    return -4;
  }
  const std::vector<int64_t> want = transposed ? std::vector<int64_t>{Cin, Cout, k, k} : std::vector<int64_t>{Cout, Cin, k, k};
  if (ws != want) {
    err = "size mismatch for " + name + ".weight";
    return -1;
  }
  const float* bias = nullptr;
  if (has_bias) {
    bias = get(name + ".bias", bshape);
    if (!bias) {
      err = "missing key " + name + ".bias";
      return -4;
    }
    if (bshape != std::vector<int64_t>{Cout}) {
      err = "size mismatch for " + name + ".bias";
      return -1;
    }
  }
  *consumed += has_bias ? 2 : 1;
  if (k * k > kGenMaxTaps || stride < 1 || stride > 2 || (transposed && stride * stride > kGenMaxPhases)) {
    err = name + ": unsupported kernel size / stride";
    return -2;
  }
  L.Cin = Cin; L.Cout = Cout; L.k = k; L.stride = stride; L.pad = pad; L.out_pad = out_pad;
  L.transposed = transposed; L.reflect = reflect;
  L.cin_chunks = (Cin + 7) / 8;
  L.cout_chunks = (Cout + 7) / 8;
  L.NT = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128));
  L.ntiles = (Cout + L.NT - 1) / L.NT;
  // taps of every phase; tapk = ky * k + kx of each tap (host only)
  std::vector<int> tapk[kGenMaxPhases];
  if (!transposed) {
    L.nphase = 1; L.ostep = 1; L.istep = stride;
    L.ph_py[0] = L.ph_px[0] = 0;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) {
        const int t = (int)tapk[0].size();
        L.offy[0][t] = (int8_t)(ky - pad);
        L.offx[0][t] = (int8_t)(kx - pad);
        tapk[0].push_back(ky * k + kx);
      }
  } else {
    // out[iy * s - pad + ky] += in[iy] * w[ky]: the outputs of parity class (py, px) use the taps with
    // ky = (py + pad) mod s (+ s, + 2s, ...) and read in[oy' + (py + pad - ky) / s]
    L.nphase = stride * stride; L.ostep = stride; L.istep = 1;
    for (int py = 0; py < stride; ++py)
      for (int px = 0; px < stride; ++px) {
        const int ph = py * stride + px;
        L.ph_py[ph] = py;
        L.ph_px[ph] = px;
        for (int ky = (py + pad) % stride; ky < k; ky += stride)
          for (int kx = (px + pad) % stride; kx < k; kx += stride) {
            const int t = (int)tapk[ph].size();
            L.offy[ph][t] = (int8_t)((py + pad - ky) / stride);
            L.offx[ph][t] = (int8_t)((px + pad - kx) / stride);
            tapk[ph].push_back(ky * k + kx);
          }
      }
  }
  auto weight = [&](int co, int ci, int kk) -> float {
    return transposed ? w[((size_t)ci * Cout + co) * k * k + kk] : w[((size_t)co * Cin + ci) * k * k + kk];
  };
  size_t n16 = 0, n32 = 0;
  for (int ph = 0; ph < L.nphase; ++ph) {
    L.ph_ntaps[ph] = (int)tapk[ph].size();
    L.ph_ksteps[ph] = std::max(1, (L.ph_ntaps[ph] * L.cin_chunks + 7) / 8);
    L.ph_woff16[ph] = n16;
    L.ph_woff32[ph] = n32;
    n16 += (size_t)L.ntiles * L.ph_ksteps[ph] * 8 * L.NT * 8;
    n32 += (size_t)L.ph_ntaps[ph] * L.cin_chunks * 8 * L.cout_chunks * 8;
  }
  cudaError_t e;
  if (cfg_.fp16) {
    std::vector<__half> pk(n16);
    for (int ph = 0; ph < L.nphase; ++ph)
      for (int nt = 0; nt < L.ntiles; ++nt)
        for (int ks = 0; ks < L.ph_ksteps[ph]; ++ks)
          for (int j = 0; j < 8; ++j) {
            const int q = ks * 8 + j, tap = q / L.cin_chunks, chunk = q % L.cin_chunks;
            __half* dst = pk.data() + L.ph_woff16[ph] + ((((size_t)nt * L.ph_ksteps[ph] + ks) * 8 + j) * L.NT) * 8;
            for (int n = 0; n < L.NT; ++n)
              for (int ee = 0; ee < 8; ++ee) {
                const int co = nt * L.NT + n, ci = chunk * 8 + ee;
                const float v = (tap < L.ph_ntaps[ph] && co < Cout && ci < Cin) ? weight(co, ci, tapk[ph][tap]) : 0.f;
                dst[(size_t)n * 8 + ee] = __float2half_rn(v);
              }
          }
    e = cudaMalloc(&L.d_w16, n16 * sizeof(__half));
    if (e == cudaSuccess) {
      owned_.push_back(L.d_w16);
      e = cudaMemcpy(L.d_w16, pk.data(), n16 * sizeof(__half), cudaMemcpyHostToDevice);
    }
  } else {
    const int cip = L.cin_chunks * 8, cop = L.cout_chunks * 8;
    std::vector<float> pk(n32, 0.f);
    for (int ph = 0; ph < L.nphase; ++ph)
      for (int t = 0; t < L.ph_ntaps[ph]; ++t)
        for (int ci = 0; ci < Cin; ++ci)
          for (int co = 0; co < Cout; ++co)
            pk[L.ph_woff32[ph] + ((size_t)t * cip + ci) * cop + co] = weight(co, ci, tapk[ph][t]);
    e = cudaMalloc(&L.d_w32, n32 * sizeof(float));
    if (e == cudaSuccess) {
      owned_.push_back(L.d_w32);
      e = cudaMemcpy(L.d_w32, pk.data(), n32 * sizeof(float), cudaMemcpyHostToDevice);
    }
  }
  if (e != cudaSuccess) {
    err = name + ": weight upload failed: " + cudaGetErrorString(e);
    return -3;
  }
  const int nb = std::max(L.ntiles * L.NT, L.cout_chunks * 8);
  std::vector<float> hb((size_t)nb, 0.f);
  if (bias) std::copy(bias, bias + Cout, hb.begin());
  e = cudaMalloc(&L.d_bias, hb.size() * sizeof(float));
  if (e == cudaSuccess) {
    owned_.push_back(L.d_bias);
    e = cudaMemcpy(L.d_bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    err = name + ": bias upload failed: " + cudaGetErrorString(e);
    return -3;
  }
  if (!cfg_.fp16) return 0;
  // ---- halo-tile kernel: parity planes of the gathered tile and the taps' positions inside them
  L.nslabs = (L.cin_chunks + 1) / 2;
  int max_taps = 0;
  for (int ph = 0; ph < L.nphase; ++ph) {
    max_taps = std::max(max_taps, L.ph_ntaps[ph]);
    int qy[kGenMaxTaps], qx[kGenMaxTaps], pary[kGenMaxTaps], parx[kGenMaxTaps];
    int r0[2] = {1 << 20, 1 << 20}, r1[2] = {-(1 << 20), -(1 << 20)}, c0[2] = {1 << 20, 1 << 20}, c1[2] = {-(1 << 20), -(1 << 20)};
    for (int t = 0; t < L.ph_ntaps[ph]; ++t) {
      const int oy = L.offy[ph][t], ox = L.offx[ph][t];
      qy[t] = (int)std::floor((double)oy / L.istep);
      qx[t] = (int)std::floor((double)ox / L.istep);
      pary[t] = oy - qy[t] * L.istep;
      parx[t] = ox - qx[t] * L.istep;
      r0[pary[t]] = std::min(r0[pary[t]], qy[t]);
      r1[pary[t]] = std::max(r1[pary[t]], qy[t]);
      c0[parx[t]] = std::min(c0[parx[t]], qx[t]);
      c1[parx[t]] = std::max(c1[parx[t]], qx[t]);
    }
    int plane_of[2][2] = {{-1, -1}, {-1, -1}};
    L.h_nplanes[ph] = 0;
    L.h_rext[ph] = L.h_cext[ph] = 0;
    for (int t = 0; t < L.ph_ntaps[ph]; ++t) {
      int& pl = plane_of[pary[t]][parx[t]];
      if (pl < 0) {
        pl = L.h_nplanes[ph]++;
        L.h_pl_py[ph][pl] = (int8_t)pary[t];
        L.h_pl_px[ph][pl] = (int8_t)parx[t];
        L.h_pl_r0[ph][pl] = (int8_t)r0[pary[t]];
        L.h_pl_c0[ph][pl] = (int8_t)c0[parx[t]];
      }
      L.h_tap_pl[ph][t] = (int8_t)pl;
      L.h_tap_dr[ph][t] = (int8_t)(qy[t] - r0[pary[t]]);
      L.h_tap_dc[ph][t] = (int8_t)(qx[t] - c0[parx[t]]);
      L.h_rext[ph] = std::max(L.h_rext[ph], r1[pary[t]] - r0[pary[t]]);
      L.h_cext[ph] = std::max(L.h_cext[ph], c1[parx[t]] - c0[parx[t]]);
    }
  }
  // N tile: as wide as two stages of (tile + weights of all taps for 16 channels) allow
  L.NTh = L.NT;
  auto stage_bytes = [&](int nt) {
    size_t worst = 0;
    for (int ph = 0; ph < L.nphase; ++ph) {
      const int J = nt >= 128 ? 2 : 4;
      const size_t npos = (size_t)L.h_nplanes[ph] * (16 + L.h_rext[ph]) * (8 * J + L.h_cext[ph]);
      worst = std::max(worst, ((2 * npos * 16 + 127) & ~(size_t)127) + (size_t)L.ph_ntaps[ph] * 2 * nt * 16);
    }
    return worst;
  };
  while (L.NTh > 16 && 2 * stage_bytes(L.NTh) + 16384 > kSmemMax) L.NTh /= 2;
  if (2 * stage_bytes(L.NTh) + 16384 > kSmemMax) {
    L.NTh = 0;   // does not fit: the im2col kernel serves this layer
    return 0;
  }
  L.ntiles_h = (Cout + L.NTh - 1) / L.NTh;
  size_t nh = 0;
  for (int ph = 0; ph < L.nphase; ++ph) {
    L.ph_woffh[ph] = nh;
    nh += (size_t)L.ntiles_h * L.nslabs * L.ph_ntaps[ph] * 2 * L.NTh * 8;
  }
  {
    std::vector<__half> pk(nh);
    for (int ph = 0; ph < L.nphase; ++ph)
      for (int nt = 0; nt < L.ntiles_h; ++nt)
        for (int sl = 0; sl < L.nslabs; ++sl)
          for (int t = 0; t < L.ph_ntaps[ph]; ++t)
            for (int kc = 0; kc < 2; ++kc) {
              __half* dst = pk.data() + L.ph_woffh[ph] +
                            (((((size_t)nt * L.nslabs + sl) * L.ph_ntaps[ph] + t) * 2 + kc) * L.NTh) * 8;
              for (int n = 0; n < L.NTh; ++n)
                for (int ee = 0; ee < 8; ++ee) {
                  const int co = nt * L.NTh + n, ci = (2 * sl + kc) * 8 + ee;
                  dst[(size_t)n * 8 + ee] = __float2half_rn((co < Cout && ci < Cin) ? weight(co, ci, tapk[ph][t]) : 0.f);
                }
            }
    e = cudaMalloc(&L.d_w16h, nh * sizeof(__half));
    if (e == cudaSuccess) {
      owned_.push_back(L.d_w16h);
      e = cudaMemcpy(L.d_w16h, pk.data(), nh * sizeof(__half), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
      err = name + ": weight upload failed: " + cudaGetErrorString(e);
      return -3;
    }
  }
  // the bias array must cover the halo kernel's N tiles as well
  if (L.ntiles_h * L.NTh > nb) {
    std::vector<float> hb2((size_t)L.ntiles_h * L.NTh, 0.f);
    if (bias) std::copy(bias, bias + Cout, hb2.begin());
    e = cudaMalloc(&L.d_bias, hb2.size() * sizeof(float));
    if (e == cudaSuccess) {
      owned_.push_back(L.d_bias);
      e = cudaMemcpy(L.d_bias, hb2.data(), hb2.size() * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
      err = name + ": bias upload failed";
      return -3;
    }
  }
  return 0;
}

int I2INet::build_norm(const ParamLookup& get, const std::string& name, Norm& n, int C, std::string& err, size_t* consumed) {
  n.C = C;
  if (C % 8) {
    err = name + ": normalised channel counts must be multiples of 8 (ngf a multiple of 8)";
    return -2;
  }
  if (cfg_.norm == 1) return 0;   // InstanceNorm2d: no parameters, statistics per image always
  std::vector<int64_t> sh;
  const float* g = get(name + ".weight", sh);
  if (!g || sh != std::vector<int64_t>{C}) {
    err = (g ? "size mismatch for " : "missing key ") + name + ".weight";
    return g ? -1 : -4;
  }
  const float* b = get(name + ".bias", sh);
  if (!b || sh != std::vector<int64_t>{C}) {
    err = (b ? "size mismatch for " : "missing key ") + name + ".bias";
    return b ? -1 : -4;
  }
  const float* rm = get(name + ".running_mean", sh);
  if (!rm || sh != std::vector<int64_t>{C}) {
    err = (rm ? "size mismatch for " : "missing key ") + name + ".running_mean";
    return rm ? -1 : -4;
  }
  const float* rv = get(name + ".running_var", sh);
  if (!rv || sh != std::vector<int64_t>{C}) {
    err = (rv ? "size mismatch for " : "missing key ") + name + ".running_var";
    return rv ? -1 : -4;
  }
  *consumed += 4;
  n.affine = true;
  n.running = !cfg_.train;
  auto up = [&](const float* src, size_t count, float** dst) -> bool {
    if (cudaMalloc(dst, count * sizeof(float)) != cudaSuccess) return false;
    owned_.push_back(*dst);
    return cudaMemcpy(*dst, src, count * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  bool ok = up(g, C, &n.d_gamma) && up(b, C, &n.d_beta);
  if (ok && n.running) {
    const std::vector<float> ss = running_scale_shift(g, b, rm, rv, C);
    ok = up(ss.data(), ss.size(), &n.d_ss);
  }
  if (!ok) {
    cudaGetLastError();
    err = name + ": upload failed";
    return -3;
  }
  return 0;
}

int I2INet::build(const ParamLookup& get, std::string& err, size_t* consumed) {
  *consumed = 0;
  if (cfg_.kind == 2) {
    down_.resize(1);
    dnorm_.resize(1);
    int rc = build_conv(get, "conv", down_[0], cfg_.sl_cout, cfg_.in_nc, cfg_.sl_k, cfg_.sl_stride, cfg_.sl_pad,
                        cfg_.sl_transposed != 0, cfg_.sl_out_pad, cfg_.sl_reflect != 0, cfg_.sl_bias != 0, err, consumed);
    if (!rc && cfg_.sl_norm) rc = build_norm(get, "norm", dnorm_[0], cfg_.sl_cout, err, consumed);
    return rc;
  }
  const int ngf = cfg_.ngf;
  if (ngf % 8 || ngf < 8) {
    err = "ngf must be a multiple of 8";
    return -2;
  }
  if (cfg_.in_nc < 1 || cfg_.in_nc > 16 || cfg_.out_nc < 1 || cfg_.out_nc > 8) {
    err = "in_nc must be <= 16 and out_nc <= 8";
    return -2;
  }
  const bool bias = cfg_.norm == 1;   // use_bias = norm_layer == InstanceNorm2d (UNet_arch.py:100-103, ResNet_arch.py:47-50)
  int rc;
  if (cfg_.unit_io && cfg_.kind != 1) {
    err = "unit_io (normalisation folded into the network) needs reflection padding: ResnetGenerator only";
    return -2;
  }
  if (cfg_.kind == 0) {
    const int D = cfg_.depth;
    if (D < 5 || D > 12) {
      err = "num_downs must be in [5, 12]";
      return -2;
    }
    down_.resize(D);
    up_.resize(D);
    dnorm_.resize(D);
    unorm_.resize(D);
    auto inner_nc = [&](int i) { return ngf * std::min(1 << i, 8); };
    std::string prefix = "model.model";
    for (int i = 0; i < D; ++i) {
      const bool outer = i == 0, innermost = i == D - 1;
      const int inner = inner_nc(i), outer_nc = outer ? cfg_.out_nc : inner_nc(i - 1);
      const int cin = outer ? cfg_.in_nc : outer_nc;
      // Sequential indices (UNet_arch.py:120-158): outermost [down, sub, relu, up, tanh]; middle [lrelu, down, norm, sub,
      // relu, up, norm]; innermost [lrelu, down, relu, up, norm]
      const std::string dn = prefix + (outer ? ".0" : ".1");
      const std::string un = prefix + (outer ? ".3" : (innermost ? ".3" : ".5"));
      if ((rc = build_conv(get, dn, down_[i], inner, cin, 4, 2, 1, false, 0, false, bias, err, consumed))) return rc;
      if ((rc = build_conv(get, un, up_[i], outer_nc, innermost ? inner : 2 * inner, 4, 2, 1, true, 0, false,
                           outer ? true : bias, err, consumed)))
        return rc;
      if (!outer && !innermost && (rc = build_norm(get, prefix + ".2", dnorm_[i], inner, err, consumed))) return rc;
      if (!outer && (rc = build_norm(get, prefix + (innermost ? ".4" : ".6"), unorm_[i], outer_nc, err, consumed))) return rc;
      prefix += outer ? ".1.model" : ".3.model";
    }
  } else {
    const int nb = cfg_.depth;
    if (nb < 0 || nb > 64) {
      err = "n_blocks must be in [0, 64]";
      return -2;
    }
    const int nconv = 3 + 2 * nb + 3;
    down_.resize(nconv);
    dnorm_.resize(nconv);
    int li = 0;
    auto key = [](int i) { return "model." + std::to_string(i); };
    if (cfg_.unit_io) {
      // conv(w, 2x - 1) + b == conv(2w, x) + (b - sum over taps and input channels of w)
      std::vector<int64_t> ws, bs;
      const float* w = get(key(1) + ".weight", ws);
      if (!w || ws != std::vector<int64_t>{ngf, cfg_.in_nc, 7, 7}) {
        err = (w ? "size mismatch for " : "missing key ") + key(1) + ".weight";
        return w ? -1 : -4;
      }
      const float* b0 = bias ? get(key(1) + ".bias", bs) : nullptr;
      if (bias && (!b0 || bs != std::vector<int64_t>{ngf})) {
        err = (b0 ? "size mismatch for " : "missing key ") + key(1) + ".bias";
        return b0 ? -1 : -4;
      }
      const size_t per = (size_t)cfg_.in_nc * 49;
      std::vector<float> w2((size_t)ngf * per), b2((size_t)ngf);
      for (int co = 0; co < ngf; ++co) {
        double sum = 0.0;
        for (size_t i = 0; i < per; ++i) {
          sum += (double)w[co * per + i];
          w2[co * per + i] = 2.f * w[co * per + i];
        }
        b2[co] = (float)((b0 ? (double)b0[co] : 0.0) - sum);
      }
      const ParamLookup folded = [&](const std::string& k, std::vector<int64_t>& shape) -> const float* {
        if (k == key(1) + ".weight") {
          shape = {ngf, cfg_.in_nc, 7, 7};
          return w2.data();
        }
        shape = {ngf};
        return b2.data();
      };
      size_t dummy = 0;
      if ((rc = build_conv(folded, key(1), down_[li], ngf, cfg_.in_nc, 7, 1, 3, false, 0, true, true, err, &dummy))) return rc;
      *consumed += bias ? 2 : 1;
    } else if ((rc = build_conv(get, key(1), down_[li], ngf, cfg_.in_nc, 7, 1, 3, false, 0, true, bias, err, consumed))) {
      return rc;
    }
    if ((rc = build_norm(get, key(2), dnorm_[li], ngf, err, consumed))) return rc;
    ++li;
    int ch = ngf;
    for (int d = 0; d < 2; ++d, ++li) {
      if ((rc = build_conv(get, key(4 + 3 * d), down_[li], ch * 2, ch, 3, 2, 1, false, 0, false, bias, err, consumed))) return rc;
      if ((rc = build_norm(get, key(5 + 3 * d), dnorm_[li], ch * 2, err, consumed))) return rc;
      ch *= 2;
    }
    for (int b = 0; b < nb; ++b) {
      const std::string p = key(10 + b) + ".conv_block";
      if ((rc = build_conv(get, p + ".1", down_[li], ch, ch, 3, 1, 1, false, 0, true, bias, err, consumed))) return rc;
      if ((rc = build_norm(get, p + ".2", dnorm_[li], ch, err, consumed))) return rc;
      ++li;
      if ((rc = build_conv(get, p + ".5", down_[li], ch, ch, 3, 1, 1, false, 0, true, bias, err, consumed))) return rc;
      if ((rc = build_norm(get, p + ".6", dnorm_[li], ch, err, consumed))) return rc;
      ++li;
    }
    int idx = 10 + nb;
    for (int u = 0; u < 2; ++u, ++li, idx += 3) {
      if ((rc = build_conv(get, key(idx), down_[li], ch / 2, ch, 3, 2, 1, true, 1, false, bias, err, consumed))) return rc;
      if ((rc = build_norm(get, key(idx + 1), dnorm_[li], ch / 2, err, consumed))) return rc;
      ch /= 2;
    }
    if ((rc = build_conv(get, key(idx + 1), down_[li], cfg_.out_nc, ngf, 7, 1, 3, false, 0, true, true, err, consumed))) return rc;
  }
  return 0;
}

namespace {

void fill_params(GenConvParams& p, const GenConv& L, GenView in, int B, int Hin, int Win, int Hout, int Wout, size_t esz_in) {
  (void)esz_in;
  memset(&p, 0, sizeof p);
  p.in = in.base;
  p.in_cs = (long long)Hin * Win * 8;
  p.in_bs = (long long)in.CT * p.in_cs;
  p.in_chunk0 = in.chunk0;
  p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout;
  p.B = B;
  p.nphase = L.nphase;
  p.nsplit = 1;
  p.ostep = L.ostep; p.istep = L.istep;
  p.cin_chunks = L.cin_chunks;
  p.cout_pad = L.cout_chunks * 8;
  p.reflect = L.reflect ? 1 : 0;
  p.out_nchunks = L.cout_chunks;
  p.bias = L.d_bias;
  for (int ph = 0; ph < L.nphase; ++ph) {
    p.py[ph] = L.ph_py[ph];
    p.px[ph] = L.ph_px[ph];
    p.Hp[ph] = (Hout - L.ph_py[ph] + L.ostep - 1) / L.ostep;
    p.Wp[ph] = (Wout - L.ph_px[ph] + L.ostep - 1) / L.ostep;
    p.Mp[ph] = B * p.Hp[ph] * p.Wp[ph];
    p.ntaps[ph] = L.ph_ntaps[ph];
    p.ksteps[ph] = L.ph_ksteps[ph];
    memcpy(p.offy[ph], L.offy[ph], kGenMaxTaps);
    memcpy(p.offx[ph], L.offx[ph], kGenMaxTaps);
  }
}

int max_m(const GenConvParams& p) {
  int m = 0;
  for (int ph = 0; ph < p.nphase; ++ph) m = std::max(m, p.Mp[ph]);
  return m;
}

int halo_mode() {
  // INNFER_I2I_HALO (read per launch; tests and A/B measurements): 0 = im2col kernel only, 2 = halo kernel whenever the
  // geometry allows, even for layers with too few CTAs to pay; default 1
  const char* env = getenv("INNFER_I2I_HALO");
  return env ? atoi(env) : 1;
}

bool halo_geometry_ok(const GenConv& L, const GenConvParams& p) {
  if (!L.NTh) return false;
  for (int ph = 0; ph < L.nphase; ++ph)
    if (p.Hp[ph] < 8 || p.Wp[ph] < 8) return false;
  return true;
}

// The halo-tile kernel when the layer has enough pixels to fill 16 x 8J patches and CTAs; fills p's halo fields.
bool halo_setup(const GenConv& L, GenConvParams& p, int num_sms, dim3& grid, size_t& smem, int& stages) {
  const int mode = halo_mode();
  if (mode == 0 || (p.nsplit != 1 && !(p.split_for_halo && p.out_mode == 1)) || !halo_geometry_ok(L, p)) return false;
#ifdef INNFER_EXPERIMENTS   // timing experiments that switch parts of the kernel off (wrong results): special builds only
  p.debug = getenv("INNFER_I2I_DEBUG") ? atoi(getenv("INNFER_I2I_DEBUG")) : 0;
#endif
  // sub-patches per CTA: 4 x NT accumulator columns must fit TMEM's 512.  With NT = 128 that is the whole TMEM (one CTA per
  // SM, no second CTA to hide the epilogue) but it halves the weight bytes per MMA, which is what bounds the wide layers
  // (every CTA re-reads the slab's weights from L2): taken when there is more than a wave of such tiles and it fits.
  int J = 4, max_tiles = 0;
  if (L.NTh >= 128) {
    long long tiles4 = 0;
    for (int ph = 0; ph < L.nphase; ++ph) tiles4 += (long long)p.B * ((p.Hp[ph] + 15) / 16) * ((p.Wp[ph] + 31) / 32) * L.ntiles_h;
    const char* ej = getenv("INNFER_I2I_J");
    J = ej ? atoi(ej) : 2;   // measured (resnet_9blocks 1024x1024): J = 4 is 1.0 ms per forward SLOWER than J = 2
    (void)tiles4;
    if (J != 4) J = 2;
  }
  for (int ph = 0; ph < L.nphase; ++ph) J = std::min(J, (p.Wp[ph] + 7) / 8);
  size_t stage = 0, tail = 0;
  for (;;) {
    stage = 0;
    size_t tbl = 0;
    max_tiles = 0;
    bool ok = true;
    for (int ph = 0; ph < L.nphase; ++ph) {
      p.bands[ph] = (p.Hp[ph] + 15) / 16;
      p.cps[ph] = (p.Wp[ph] + 8 * J - 1) / (8 * J);
      max_tiles = std::max(max_tiles, p.B * p.bands[ph] * p.cps[ph]);
      p.nplanes[ph] = L.h_nplanes[ph];
      p.Rpl[ph] = 16 + L.h_rext[ph];
      p.Wpl[ph] = 8 * J + L.h_cext[ph];
      const size_t npos = (size_t)p.nplanes[ph] * p.Rpl[ph] * p.Wpl[ph];
      if (npos > 16000) ok = false;
      stage = std::max(stage, ((2 * npos * 16 + 127) & ~(size_t)127) + (size_t)L.ph_ntaps[ph] * 2 * L.NTh * 16);
      tbl = std::max(tbl, npos * 4);
    }
    stage = (stage + 127) & ~(size_t)127;
    tail = 128 + 1024 + 512 + ((tbl + 127) & ~(size_t)127);
    if (ok && (L.nslabs == 1 ? 1 : 2) * stage + tail <= kSmemMax) break;
    if (J <= 1) return false;
    J /= 2;   // a narrower patch needs a smaller tile
  }
  for (int ph = 0; ph < L.nphase; ++ph) {
    for (int pl = 0; pl < 4; ++pl) {
      p.pl_py[ph][pl] = L.h_pl_py[ph][pl];
      p.pl_px[ph][pl] = L.h_pl_px[ph][pl];
      p.pl_r0[ph][pl] = L.h_pl_r0[ph][pl];
      p.pl_c0[ph][pl] = L.h_pl_c0[ph][pl];
    }
    for (int t = 0; t < L.ph_ntaps[ph]; ++t)
      p.tap_aoff[ph][t] = (uint16_t)((size_t)L.h_tap_pl[ph][t] * p.Rpl[ph] * p.Wpl[ph] + (size_t)L.h_tap_dr[ph][t] * p.Wpl[ph] +
                                     L.h_tap_dc[ph][t]);
  }
  if (mode != 2 && max_tiles * L.ntiles_h * L.nphase * p.nsplit * 4 < num_sms) return false;   // too few CTAs: im2col kernel
  p.stat_slices = 0;
  for (int ph = 0; ph < L.nphase; ++ph) {
    p.slice0[ph] = p.stat_slices;
    p.stat_slices += p.bands[ph] * p.cps[ph];
  }
  if (L.nslabs == 1) {
    stages = 1;       // a single 16-channel slab (the 3-channel first layer): nothing to pipeline, keep the CTA small
  } else {
    if (2 * stage + tail <= kSmemHalf && J * L.NTh <= 256)
      stages = 2;     // two CTAs per SM: one CTA's epilogue overlaps the other's main loop
    else
      stages = (int)std::min<size_t>(4, (kSmemMax - tail) / stage);
    if (getenv("INNFER_I2I_STAGES")) stages = std::min<int>(atoi(getenv("INNFER_I2I_STAGES")), (int)((kSmemMax - tail) / stage));
    stages = std::max(2, std::min(stages, L.nslabs));
  }
  p.J = J;
  p.nslabs = L.nslabs;
  p.stage_bytes = (unsigned)stage;
  int cols = 32;
  while (cols < J * L.NTh) cols *= 2;
  p.tmem_cols = cols;
  smem = (size_t)stages * stage + tail;
  grid = dim3((unsigned)max_tiles, (unsigned)L.ntiles_h, (unsigned)(L.nphase * p.nsplit));
  p.w = L.d_w16h;
  for (int ph = 0; ph < L.nphase; ++ph) p.woff[ph] = (long long)L.ph_woffh[ph];
  return true;
}

// `stats_alloc` (optional): called with the number of doubles the fused statistics of this launch need; returns the
// buffer (or null: no fused statistics).  *stat_slices receives the slice count when the statistics were fused.
cudaError_t launch_conv(const GenConv& L, GenConvParams& p, bool fp16, int num_sms, cudaStream_t st,
                        const std::function<double*(size_t)>* stats_alloc = nullptr, int* stat_slices = nullptr) {
  if (stat_slices) *stat_slices = 0;
  p.stats = nullptr;
  if (fp16) {
    dim3 hgrid;
    size_t hsmem = 0;
    int hstages = 0;
    if (halo_setup(L, p, num_sms, hgrid, hsmem, hstages)) {
      g_halo_launches.fetch_add(1, std::memory_order_relaxed);
      if (stats_alloc && p.out_mode == 1 && p.nsplit == 1) {
        p.stats = (*stats_alloc)((size_t)p.B * p.out_nchunks * p.stat_slices * 16);
        if (p.stats && stat_slices) *stat_slices = p.stat_slices;
      }
      switch (L.NTh) {
        case 16: return launch_halo<16>(p, hstages, hgrid, hsmem, st);
        case 32: return launch_halo<32>(p, hstages, hgrid, hsmem, st);
        case 64: return launch_halo<64>(p, hstages, hgrid, hsmem, st);
        default: return launch_halo<128>(p, hstages, hgrid, hsmem, st);
      }
    }
    p.w = L.d_w16;
    for (int ph = 0; ph < L.nphase; ++ph) p.woff[ph] = (long long)L.ph_woff16[ph];
    const dim3 grid((unsigned)((max_m(p) + 127) / 128), (unsigned)L.ntiles, (unsigned)(L.nphase * p.nsplit));
    switch (L.NT) {
      case 16: return launch_tc<16>(p, grid, st);
      case 32: return launch_tc<32>(p, grid, st);
      case 64: return launch_tc<64>(p, grid, st);
      default: return launch_tc<128>(p, grid, st);
    }
  }
  p.w = L.d_w32;
  for (int ph = 0; ph < L.nphase; ++ph) p.woff[ph] = (long long)L.ph_woff32[ph];
  const dim3 grid((unsigned)((max_m(p) + 127) / 128), (unsigned)L.cout_chunks, (unsigned)L.nphase);
  gen_conv_direct_kernel<<<grid, 128, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

int I2INet::conv_raw(const GenConv& L, GenView in, int B, int Hin, int Win, Raw& raw, cudaStream_t st, bool want_stats) {
  const int Hout = L.out_h(Hin), Wout = L.out_h(Win);
  GenConvParams p;
  fill_params(p, L, in, B, Hin, Win, Hout, Wout, esz());
  // split K across CTAs when the layer has too few output tiles to fill the GPU (inner levels of the UNet)
  int nsplit = 1;
  bool halo_split = false;
  if (cfg_.fp16 && halo_mode() != 0 && halo_geometry_ok(L, p)) {
    // mid-size layers (UNet 256 -> 512 at 64 x 64: 64 tiles for 148 SMs, 16 slabs of 100 KB each): the halo kernel with
    // the slabs of a tile split over up to four CTAs
    int J = L.NTh >= 128 ? 2 : 4, tiles = 0;
    for (int ph = 0; ph < L.nphase; ++ph) J = std::min(J, (p.Wp[ph] + 7) / 8);
    for (int ph = 0; ph < L.nphase; ++ph) tiles += p.B * ((p.Hp[ph] + 15) / 16) * ((p.Wp[ph] + 8 * J - 1) / (8 * J));
    const int ctas = tiles * L.ntiles_h;
    static const int hs = getenv("INNFER_I2I_HALO_SPLIT") ? atoi(getenv("INNFER_I2I_HALO_SPLIT")) : 1;
    if (hs && ctas * 4 >= num_sms_ && ctas < num_sms_ && L.nslabs >= 8) {
      nsplit = std::max(1, std::min({4, num_sms_ / ctas, L.nslabs / 4}));
      halo_split = nsplit > 1;
    }
  }
  if (cfg_.fp16 && !halo_split) {
    const int ctas = ((max_m(p) + 127) / 128) * L.ntiles * L.nphase;
    int ksmin = 1 << 30;
    for (int ph = 0; ph < L.nphase; ++ph) ksmin = std::min(ksmin, L.ph_ksteps[ph]);
    if (ctas * 2 <= num_sms_ && ksmin >= 4 && !(halo_mode() == 2 && halo_geometry_ok(L, p)))
      nsplit = std::max(1, std::min({ksmin / 2, kGenMaxSplit, (num_sms_ + ctas - 1) / ctas}));
  }
  const size_t per = (size_t)B * L.cout_chunks * Hout * Wout * 8;
  int rc = need(raw_, per * nsplit * sizeof(float));
  if (rc) return rc;
  raw.p = reinterpret_cast<float*>(raw_.p);
  raw.nsplit = nsplit;
  raw.split_stride = per;
  p.out = raw.p;
  p.out_cs = (long long)Hout * Wout * 8;
  p.out_bs = (long long)L.cout_chunks * p.out_cs;
  p.out_chunk0 = 0;
  p.out_mode = 1;
  p.nsplit = nsplit;
  p.split_for_halo = halo_split ? 1 : 0;
  p.split_stride = (long long)per;
  p.act = kActNone;
  // statistics of the norm layer that follows, fused into the halo kernel's epilogue when that kernel takes the layer
  static const bool fuse = !(getenv("INNFER_I2I_FUSED_STATS") && atoi(getenv("INNFER_I2I_FUSED_STATS")) == 0);
  int alloc_rc = 0;
  const std::function<double*(size_t)> alloc = [&](size_t n) -> double* {
    alloc_rc = need(stats_, n * sizeof(double));
    return alloc_rc ? nullptr : reinterpret_cast<double*>(stats_.p);
  };
  raw.stat_slices = 0;
  const cudaError_t e = launch_conv(L, p, cfg_.fp16 != 0, num_sms_, st, want_stats && fuse ? &alloc : nullptr, &raw.stat_slices);
  if (alloc_rc) return alloc_rc;
  ++launches_;
  if (e != cudaSuccess) {
    err_ = std::string("conv launch failed: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

int I2INet::conv_final(const GenConv& L, GenView in, int B, int Hin, int Win, GenView out, int act, bool compact4,
                       cudaStream_t st) {
  const int Hout = L.out_h(Hin), Wout = L.out_h(Win);
  GenConvParams p;
  fill_params(p, L, in, B, Hin, Win, Hout, Wout, esz());
  p.out = out.base;
  p.out_cs = (long long)Hout * Wout * 8;
  p.out_bs = (long long)out.CT * p.out_cs;
  p.out_chunk0 = out.chunk0;
  p.out_mode = compact4 ? 2 : 0;
  p.act = act;
  if (compact4 && !cfg_.fp16) {
    err_ = "compact tiles exist in fp16 mode only";
    return -1;
  }
  const cudaError_t e = launch_conv(L, p, cfg_.fp16 != 0, num_sms_, st);
  ++launches_;
  if (e != cudaSuccess) {
    err_ = std::string("conv launch failed: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

int I2INet::norm_apply(const Norm* n, const Raw& raw, int B, int C, int H, int W, bool per_sample, int act_a, GenView out_a,
                       int act_b, const GenView* out_b, const GenView* res, cudaStream_t st) {
  const int chunks = (C + 7) / 8, HW = H * W;
  const long long cs = (long long)HW * 8;
  ApplyParams p;
  memset(&p, 0, sizeof p);
  p.raw = raw.p;
  p.nsplit = raw.nsplit;
  p.split_stride = (long long)raw.split_stride;
  p.raw_cs = cs;
  p.raw_bs = (long long)chunks * cs;
  p.B = B; p.chunks = chunks; p.HW = HW;
  if (n) {
    if (n->running) {
      p.ss = reinterpret_cast<const float2*>(n->d_ss);
      p.ss_per_sample = 0;
    } else {
      // InstanceNorm2d, or BatchNorm2d in training mode (per image when every image is its own reference call)
      const bool ps = cfg_.norm == 1 || per_sample;
      const int nslices = raw.stat_slices ? raw.stat_slices : std::max(1, std::min(64, (HW + 4095) / 4096));
      int rc = raw.stat_slices ? 0 : need(stats_, (size_t)B * chunks * nslices * 16 * sizeof(double));
      if (!rc) rc = need(ss_, (size_t)B * chunks * 8 * sizeof(float2));
      if (rc) return rc;
      if (!raw.stat_slices) {   // the conv kernel did not leave per-tile partial sums: one pass over the raw tensor
        norm_stats_kernel<<<dim3((unsigned)nslices, (unsigned)chunks, (unsigned)B), 256, 0, st>>>(
            raw.p, raw.nsplit, (long long)raw.split_stride, p.raw_bs, p.raw_cs, HW, reinterpret_cast<double*>(stats_.p));
        ++launches_;
      }
      norm_finalize_kernel<<<(ps ? B : 1) * chunks * 8, 128, 0, st>>>(reinterpret_cast<const double*>(stats_.p), B, chunks, nslices,
                                                                  HW, ps ? 1 : 0, n->d_gamma, n->d_beta,
                                                                  reinterpret_cast<float2*>(ss_.p));
      ++launches_;
      p.ss = reinterpret_cast<const float2*>(ss_.p);
      p.ss_per_sample = 1;
    }
  }
  p.out_a = reinterpret_cast<uint8_t*>(out_a.base) + (size_t)out_a.chunk0 * cs * esz();
  p.a_cs = cs;
  p.a_bs = (long long)out_a.CT * cs;
  p.act_a = act_a;
  if (out_b) {
    p.out_b = reinterpret_cast<uint8_t*>(out_b->base) + (size_t)out_b->chunk0 * cs * esz();
    p.b_cs = cs;
    p.b_bs = (long long)out_b->CT * cs;
    p.act_b = act_b;
  }
  if (res) {
    p.res = reinterpret_cast<const uint8_t*>(res->base) + (size_t)res->chunk0 * cs * esz();
    p.r_cs = cs;
    p.r_bs = (long long)res->CT * cs;
  }
  const long long total = (long long)B * chunks * HW;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)num_sms_ * 16);
  if (cfg_.fp16)
    norm_apply_kernel<__half><<<blocks, 256, 0, st>>>(p);
  else
    norm_apply_kernel<float><<<blocks, 256, 0, st>>>(p);
  ++launches_;
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    err_ = std::string("norm launch failed: ") + cudaGetErrorString(e);
    return -3;
  }
  return 0;
}

int I2INet::forward_unet(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample,
                         cudaStream_t st) {
  const int D = cfg_.depth;
  auto inner_nc = [&](int i) { return cfg_.ngf * std::min(1 << i, 8); };
  int rc;
  dbuf_.resize(D);
  cat_.resize(D);
  for (int i = 1; i < D; ++i) {   // x_i has inner_nc(i - 1) channels at (H >> i) x (W >> i)
    const size_t n = (size_t)B * inner_nc(i - 1) * (H >> i) * (W >> i) * esz();
    if ((rc = need(dbuf_[i], n)) || (rc = need(cat_[i], 2 * n))) return rc;
  }
  GenView cur{const_cast<void*>(in), in_CT, 0};
  Raw raw;
  // down: d_i = downconv_i(lrelu(x_i)) [+ norm]; lrelu(d_i) feeds the next level, relu(d_i) = relu(lrelu(d_i)) is the skip
  // half of what the next level returns after the in-place ReLU in front of this level's transposed conv
  for (int i = 0; i < D; ++i) {
    const int hi = H >> i, wi = W >> i, C = inner_nc(i);
    if ((rc = conv_raw(down_[i], cur, B, hi, wi, raw, st, i > 0 && i < D - 1 && !dnorm_[i].running))) return rc;
    if (i < D - 1) {
      const GenView a = view(dbuf_[i + 1], C / 8, 0), skip = view(cat_[i + 1], 2 * C / 8, 0);
      if ((rc = norm_apply(i > 0 ? &dnorm_[i] : nullptr, raw, B, C, hi / 2, wi / 2, per_sample, kActLrelu, a, kActRelu, &skip,
                           nullptr, st)))
        return rc;
      cur = a;
    } else {
      if ((rc = need(inner_, (size_t)B * C * (hi / 2) * (wi / 2) * esz()))) return rc;
      const GenView o = view(inner_, C / 8, 0);
      if ((rc = norm_apply(nullptr, raw, B, C, hi / 2, wi / 2, per_sample, kActRelu, o, 0, nullptr, nullptr, st))) return rc;
      cur = o;
    }
  }
  // up: y_i = norm(upconv_i(relu(inner))) goes, already passed through the next ReLU, behind the skip half
  for (int i = D - 1; i >= 1; --i) {
    const int hin = H >> (i + 1), win = W >> (i + 1), X = inner_nc(i - 1);
    if ((rc = conv_raw(up_[i], cur, B, hin, win, raw, st, !unorm_[i].running))) return rc;
    const GenView o = view(cat_[i], 2 * X / 8, X / 8);
    if ((rc = norm_apply(&unorm_[i], raw, B, X, H >> i, W >> i, per_sample, kActRelu, o, 0, nullptr, nullptr, st))) return rc;
    cur = view(cat_[i], 2 * X / 8, 0);
  }
  return conv_final(up_[0], cur, B, H >> 1, W >> 1, out, kActTanh, compact4, st);
}

int I2INet::forward_resnet(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample,
                           cudaStream_t st) {
  const int ngf = cfg_.ngf, nb = cfg_.depth;
  const int H1 = H / 2, W1 = W / 2, H2 = H / 4, W2 = W / 4;
  int rc;
  rb_.resize(3);
  if ((rc = need(a_, (size_t)B * ngf * H * W * esz())) || (rc = need(b_, (size_t)B * 2 * ngf * H1 * W1 * esz()))) return rc;
  for (auto& b : rb_)
    if ((rc = need(b, (size_t)B * 4 * ngf * H2 * W2 * esz()))) return rc;
  Raw raw;
  int li = 0;
  GenView x{const_cast<void*>(in), in_CT, 0};
  const GenView va = view(a_, ngf / 8, 0), vb = view(b_, 2 * ngf / 8, 0);
  if ((rc = conv_raw(down_[li], x, B, H, W, raw, st, !dnorm_[li].running))) return rc;
  if ((rc = norm_apply(&dnorm_[li], raw, B, ngf, H, W, per_sample, kActRelu, va, 0, nullptr, nullptr, st))) return rc;
  ++li;
  if ((rc = conv_raw(down_[li], va, B, H, W, raw, st, !dnorm_[li].running))) return rc;
  if ((rc = norm_apply(&dnorm_[li], raw, B, 2 * ngf, H1, W1, per_sample, kActRelu, vb, 0, nullptr, nullptr, st))) return rc;
  ++li;
  int y = 0, t = 1, y2 = 2;
  const int C = 4 * ngf;
  if ((rc = conv_raw(down_[li], vb, B, H1, W1, raw, st, !dnorm_[li].running))) return rc;
  if ((rc = norm_apply(&dnorm_[li], raw, B, C, H2, W2, per_sample, kActRelu, view(rb_[y], C / 8, 0), 0, nullptr, nullptr, st)))
    return rc;
  ++li;
  for (int b = 0; b < nb; ++b) {
    const GenView vy = view(rb_[y], C / 8, 0), vt = view(rb_[t], C / 8, 0), vy2 = view(rb_[y2], C / 8, 0);
    if ((rc = conv_raw(down_[li], vy, B, H2, W2, raw, st, !dnorm_[li].running))) return rc;
    if ((rc = norm_apply(&dnorm_[li], raw, B, C, H2, W2, per_sample, kActRelu, vt, 0, nullptr, nullptr, st))) return rc;
    ++li;
    if ((rc = conv_raw(down_[li], vt, B, H2, W2, raw, st, !dnorm_[li].running))) return rc;
    if ((rc = norm_apply(&dnorm_[li], raw, B, C, H2, W2, per_sample, kActNone, vy2, 0, nullptr, &vy, st))) return rc;
    ++li;
    std::swap(y, y2);
  }
  if ((rc = conv_raw(down_[li], view(rb_[y], C / 8, 0), B, H2, W2, raw, st, !dnorm_[li].running))) return rc;
  if ((rc = norm_apply(&dnorm_[li], raw, B, 2 * ngf, H1, W1, per_sample, kActRelu, vb, 0, nullptr, nullptr, st))) return rc;
  ++li;
  if ((rc = conv_raw(down_[li], vb, B, H1, W1, raw, st, !dnorm_[li].running))) return rc;
  if ((rc = norm_apply(&dnorm_[li], raw, B, ngf, H, W, per_sample, kActRelu, va, 0, nullptr, nullptr, st))) return rc;
  ++li;
  return conv_final(down_[li], va, B, H, W, out, cfg_.unit_io ? kActTanh01 : kActTanh, compact4, st);
}

int I2INet::forward_eager(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample,
                          cudaStream_t st) {
  if (cfg_.kind == 2) {
    int rc;
    GenView x{const_cast<void*>(in), in_CT, 0};
    if (cfg_.sl_final) {
      rc = conv_final(down_[0], x, B, H, W, out, cfg_.sl_act, compact4, st);
    } else {
      Raw raw;
      rc = conv_raw(down_[0], x, B, H, W, raw, st, cfg_.sl_norm && !dnorm_[0].running);
      if (!rc)
        rc = norm_apply(cfg_.sl_norm ? &dnorm_[0] : nullptr, raw, B, cfg_.sl_cout, down_[0].out_h(H), down_[0].out_h(W),
                        per_sample, cfg_.sl_act, out, 0, nullptr, nullptr, st);
    }
    return rc;
  }
  return cfg_.kind == 0 ? forward_unet(in, in_CT, B, H, W, out, compact4, per_sample, st)
                        : forward_resnet(in, in_CT, B, H, W, out, compact4, per_sample, st);
}

void I2INet::drop_graphs() {
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
}

// Small problems are launch-latency bound (resnet_9blocks at 256x256: 91 launches in 1.1 ms): the launch sequence of a
// (pointers, shape) combination is recorded into a CUDA graph the second time it is seen and replayed from then on.
// The first call runs eagerly and sizes every buffer; a later buffer growth drops all graphs.  INNFER_I2I_GRAPH=0 disables.
int I2INet::forward(const void* in, int in_CT, int B, int H, int W, GenView out, bool compact4, bool per_sample, cudaStream_t st,
                    std::string& err) {
  int rc = check_size(H, W, err);
  if (rc) return rc;
  err_.clear();
  static const int graph_mode = getenv("INNFER_I2I_GRAPH") ? atoi(getenv("INNFER_I2I_GRAPH")) : 1;
  const bool small = (long long)B * H * W <= (graph_mode == 2 ? (1ll << 40) : 512ll * 512);
  if (!graph_mode || cfg_.kind == 2 || !small) {
    rc = forward_eager(in, in_CT, B, H, W, out, compact4, per_sample, st);
    if (rc) err = err_;
    return rc;
  }
  const GraphKey key{in, out.base, in_CT, B, H, W, out.CT, out.chunk0, compact4, per_sample};
  auto it = graphs_.find(key);
  if (it == graphs_.end()) {   // first sight: eager (allocations, kernel attributes), remember the shape
    rc = forward_eager(in, in_CT, B, H, W, out, compact4, per_sample, st);
    if (rc) {
      err = err_;
      return rc;
    }
    if (graphs_.size() >= 64) drop_graphs();   // a caller that keeps changing buffers: do not hoard recordings
    graphs_[key] = GraphEntry{};
    return 0;
  }
  if (!it->second.exec && !it->second.failed) {
    // record on an internal stream (the caller's may be the legacy default stream, which cannot be captured)
    if (!cap_ && cudaStreamCreateWithFlags(&cap_, cudaStreamNonBlocking) != cudaSuccess) cap_ = nullptr;
    cudaGraph_t graph = nullptr;
    const uint64_t l0 = launches_;
    bool ok = cap_ && cudaStreamBeginCapture(cap_, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      capturing_ = true;
      rc = forward_eager(in, in_CT, B, H, W, out, compact4, per_sample, cap_);
      capturing_ = false;
      ok = cudaStreamEndCapture(cap_, &graph) == cudaSuccess && rc == 0 && graph != nullptr;
    }
    const uint64_t recorded = launches_ - l0;
    launches_ = l0;   // nothing ran yet
    if (ok) ok = cudaGraphInstantiate(&it->second.exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      it->second.exec = nullptr;
      it->second.failed = true;   // keep serving this shape eagerly
    } else {
      it->second.launches = recorded;
    }
  }
  if (it->second.exec) {
    const cudaError_t e = cudaGraphLaunch(it->second.exec, st);
    if (e != cudaSuccess) {
      err = std::string("graph launch failed: ") + cudaGetErrorString(e);
      return -3;
    }
    launches_ += it->second.launches;
    ++graph_replays_;
    g_graph_replays.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  rc = forward_eager(in, in_CT, B, H, W, out, compact4, per_sample, st);
  if (rc) err = err_;
  return rc;
}

}  // namespace innfer
