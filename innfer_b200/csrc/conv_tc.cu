// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel. See conv_tc.cuh for the design notes.
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace innfer {

namespace {

struct TileCoord {
  int b, y0, x0, phase, jeff;
};

// Walks the tile list t0, t0+stride, ... as a mixed-radix counter (phase, col-patch, band, image)
// so that the per-tile decode on the MMA warp's critical path needs no integer division.
struct TileIter {
  int phase, cp, band, b;
  int dphase, dcp, dband, db;
  int nphase, cps, bands, B, J, W;
  __device__ __forceinline__ void init(const ConvTcParams& p, int t0, int stride) {
    nphase = p.nphase; cps = p.cps; bands = p.bands; B = p.B; J = p.J; W = p.W;
    phase = t0 % nphase; int r = t0 / nphase;
    cp = r % cps; r /= cps;
    band = r % bands; b = r / bands;
    dphase = stride % nphase; r = stride / nphase;
    dcp = r % cps; r /= cps;
    dband = r % bands; db = r / bands;
  }
  __device__ __forceinline__ bool valid() const { return b < B; }
  __device__ __forceinline__ void advance() {
    phase += dphase;
    int c = phase >= nphase ? 1 : 0;
    phase -= c ? nphase : 0;
    cp += dcp + c;
    c = cp >= cps ? 1 : 0;
    cp -= c ? cps : 0;
    band += dband + c;
    c = band >= bands ? 1 : 0;
    band -= c ? bands : 0;
    b += db + c;
  }
  __device__ __forceinline__ TileCoord coord() const {
    TileCoord c;
    c.b = b;
    c.phase = phase;
    c.y0 = band * kPatchRows;
    c.x0 = cp * 8 * J;
    const int rem = (W - c.x0 + 7) >> 3;  // sub-patches that still touch the image
    c.jeff = rem < J ? rem : J;
    return c;
  }
};

__device__ __forceinline__ float lrelu_f(float v, float slope) { return v > 0.f ? v : v * slope; }

template <int N>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel of the stream may take this SM as soon as this CTA leaves it

  const int Wh = 8 * p.J + 2 * p.dil;              // halo tile: Rh x Wh pixels
  const int Rh = kPatchRows + 2 * p.dil;
  const int a_bytes = conv_tc_a_bytes(p.J, p.dil);
  int max_taps = 0;
  for (int i = 0; i < p.nphase; ++i) max_taps = max_taps > p.ph_ntaps[i] ? max_taps : p.ph_ntaps[i];
  const int w_bytes = conv_tc_w_bytes(N, max_taps);
  const int stage_bytes = a_bytes + w_bytes;
  const int S = p.stages;

  uint8_t* bar_base = smem + (size_t)S * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* set_bar = tfull_bar + 2;       // "accumulator set is drained" (16 epilogue warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(set_bar + 2);
  volatile uint32_t* ready_cnt = tmem_slot + 1;   // stages whose barriers the scout warp has seen complete
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&tfull_bar[i]), 1);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&set_bar[i]), 16);
    *ready_cnt = 0u;
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.nphase * N; i += blockDim.x) s_bias[i] = p.bias[i];  // bias is [phase][N]
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t smem_base = smem_u32(smem);

  // programmatic dependent launch: everything above touched only constants (bias), shared memory and TMEM; the warps
  // that read or write activations wait here for the previous kernel of the stream
  if (warp != 1 && warp != kConvScoutWarp) pdl_wait();
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      TileIter ti;
      ti.init(p, blockIdx.x, gridDim.x);
      for (; ti.valid(); ti.advance()) {
        const TileCoord c = ti.coord();
        const uint32_t wb = (uint32_t)p.ph_ntaps[c.phase] * 2u * N * 16u;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) + p.ph_woff[c.phase];
        for (int ks = 0; ks < p.kslabs; ++ks) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          const uint32_t dstA = smem_base + (uint32_t)s * stage_bytes;
          mbar_expect_tx(fb, (uint32_t)(2 * Rh * Wh * 16) + wb);
          tma_load_5d(dstA, &tmap_in, fb, 0, c.x0 - p.dil, c.y0 - p.dil, p.in_chunk0 + 2 * ks, c.b);
          bulk_load(dstA + a_bytes, wsrc + (size_t)ks * wb, wb, fb);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // tcgen05.mma is issued by one thread and the tensor pipe buffers only a few instructions:
    // with M=128 x N=32..64 x K=16 MMAs (16-32 tensor cycles each) every instruction on this
    // warp's path shows up as a pipe bubble, so the loop is kept as lean as possible (division-free
    // tile iteration, one accumulator-set barrier per tile, taps fully unrolled).  Measured on
    // B200: a second issuing warp does NOT help -- MMAs from two threads interleave worse than the
    // same MMAs from one thread.  The whole warp runs the (warp-uniform) control flow so that
    // descriptor arithmetic stays in the uniform datapath; only the elected lane issues.
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t a_lbo = (uint32_t)(Rh * Wh);          // in 16-byte units
    const uint32_t a_sbo = (uint32_t)Wh;
    const uint32_t a_hi = a_sbo | (1u << 14);            // SBO [32,46) + version=1 [46,48)
    const uint32_t dil = (uint32_t)p.dil, dil_row = dil * a_sbo;   // tap spacing in pixels / in tile rows
    const uint32_t b_hi = 8u | (1u << 14);               // SBO = 128 B
    const uint32_t b_lbo = (uint32_t)N;                  // N * 16 B
    const uint32_t tap_stride = 2u * N;                  // 16-byte units per tap in the weight stage
    const int J = p.J;
    const bool plain3x3 = (p.nphase == 1) && (p.ph_ntaps[0] == 9);
    int s = 0;
    uint32_t it = 0;
    uint32_t stage_i = 0, ready = 0;
    const uint32_t ready_addr = smem_u32((const void*)ready_cnt);
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (; ti.valid(); ti.advance(), ++it) {
      const TileCoord c = ti.coord();
      const int ntaps = plain3x3 ? 9 : (int)p.ph_ntaps[c.phase];
      const uint32_t set = it & 1u;
      const uint32_t acc0 = tmem_base + set * (uint32_t)(J * N);
      // halo-tile offsets of this phase's taps (16-byte units), once per tile
      uint32_t aoff[kMaxTaps];
      if (!plain3x3) {
#pragma unroll
        for (int tp = 0; tp < kMaxTaps; ++tp)
          aoff[tp] = (uint32_t)p.tap_hy[c.phase][tp] * dil_row + p.tap_hx[c.phase][tp] * dil;   // dil = 0: halo-free 1x1
      }
      for (int ks = 0; ks < p.kslabs; ++ks, ++stage_i) {
        // stage stage_i is ready when the scout has seen its TMA data (and, for the first stage of a
        // tile, the accumulator set drained by the epilogue); the count is re-read only when needed
        if (ready <= stage_i) {
          uint32_t spins = 0;
          do {
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ready) : "r"(ready_addr) : "memory");
            if (++spins > (1u << 26)) __trap();
          } while (ready <= stage_i);
          tc_fence_after();
        }
        const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t a_lo = ((sa & 0x3FFFFu) >> 4) | (a_lbo << 16);
        const uint32_t b_lo = (((sa + a_bytes) & 0x3FFFFu) >> 4) | (b_lbo << 16);
        const uint32_t first = ks != 0 ? 1u : 0u;
        if (N == 64 && ntaps == 9) {
          // weight-stationary order: tap outer, sub-patch inner; the tap's B tile is read from shared
          // memory once (collector fill) and reused for the other sub-patches
          if (leader) {
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
              const int hy = tp / 3, hx = tp % 3;
              const uint64_t bd = make_desc64(b_lo + tp * tap_stride, b_hi);
              const uint32_t at = a_lo + hy * dil_row + hx * dil;
              const uint32_t accf = tp == 0 ? first : 1u;
              if (tp & 1) {
                umma_f16_ws<1, true>(acc0, make_desc64(at, a_hi), bd, idesc, accf);
                for (int j = 1; j < c.jeff; ++j)
                  umma_f16_ws<1, false>(acc0 + (uint32_t)(j * N), make_desc64(at + 8u * j, a_hi), bd, idesc, accf);
              } else {
                umma_f16_ws<0, true>(acc0, make_desc64(at, a_hi), bd, idesc, accf);
                for (int j = 1; j < c.jeff; ++j)
                  umma_f16_ws<0, false>(acc0 + (uint32_t)(j * N), make_desc64(at + 8u * j, a_hi), bd, idesc, accf);
              }
            }
          }
        } else if (ntaps == 9) {
          // plain 3x3: taps are (hy, hx) in row-major order, fully unrolled
          for (int j = 0; j < c.jeff; ++j) {
            const uint32_t acc = acc0 + (uint32_t)(j * N);
            const uint32_t aj = a_lo + 8u * j;
            if (leader) {
#pragma unroll
              for (int hy = 0; hy < 3; ++hy) {
#pragma unroll
                for (int hx = 0; hx < 3; ++hx) {
                  const int tp = hy * 3 + hx;
                  umma_f16_ss(acc, make_desc64(aj + hy * dil_row + hx * dil, a_hi),
                              make_desc64(b_lo + tp * tap_stride, b_hi), idesc, tp == 0 ? first : 1u);
                }
              }
            }
          }
        } else if (N == 64) {
          // phase-folded / pixel-shuffle / 1x1 convs (1..9 taps from the phase's table), weight-stationary order
          if (leader) {
#pragma unroll
            for (int tp = 0; tp < kMaxTaps; ++tp) {
              if (tp < ntaps) {
                const uint64_t bd = make_desc64(b_lo + tp * tap_stride, b_hi);
                const uint32_t at = a_lo + aoff[tp];
                const uint32_t accf = tp == 0 ? first : 1u;
                if (tp & 1) {
                  umma_f16_ws<1, true>(acc0, make_desc64(at, a_hi), bd, idesc, accf);
                  for (int j = 1; j < c.jeff; ++j)
                    umma_f16_ws<1, false>(acc0 + (uint32_t)(j * N), make_desc64(at + 8u * j, a_hi), bd, idesc, accf);
                } else {
                  umma_f16_ws<0, true>(acc0, make_desc64(at, a_hi), bd, idesc, accf);
                  for (int j = 1; j < c.jeff; ++j)
                    umma_f16_ws<0, false>(acc0 + (uint32_t)(j * N), make_desc64(at + 8u * j, a_hi), bd, idesc, accf);
                }
              }
            }
          }
        } else {
          for (int j = 0; j < c.jeff; ++j) {
            const uint32_t acc = acc0 + (uint32_t)(j * N);
            const uint32_t aj = a_lo + 8u * j;
            if (leader) {
#pragma unroll
              for (int tp = 0; tp < kMaxTaps; ++tp)
                if (tp < ntaps)
                  umma_f16_ss(acc, make_desc64(aj + aoff[tp], a_hi), make_desc64(b_lo + tp * tap_stride, b_hi), idesc,
                              tp == 0 ? first : 1u);
            }
          }
        }
        if (leader) umma_commit(smem_u32(&empty_bar[s]));
        __syncwarp();
        if (++s == S) s = 0;
      }
      if (leader) umma_commit(smem_u32(&tfull_bar[set]));
      __syncwarp();
    }
  } else if (warp == kConvScoutWarp) {
    // ------------------------------------------------------------ scout
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, it = 0, done = 0;
      const uint32_t ready_addr = smem_u32((const void*)ready_cnt);
      TileIter ti;
      ti.init(p, blockIdx.x, gridDim.x);
      for (; ti.valid(); ti.advance(), ++it) {
        // the epilogue must have drained the tile that used this accumulator set two tiles ago
        mbar_wait(smem_u32(&set_bar[it & 1u]), ((it >> 1) & 1u) ^ 1u);
        for (int ks = 0; ks < p.kslabs; ++ks) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
          ++done;
          asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(ready_addr), "r"(done) : "memory");
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    // Sixteen epilogue warps = four per TMEM lane quarter: `grp` takes the sub-tiles j = grp, grp + 2, ... of every
    // tile, `hh` the 8-channel chunks ch = hh, hh + 2, ... of the accumulator (N = 16: both chunks, the hh = 1 warps
    // idle).  The epilogue of these small-N convs is a chain of latencies (accumulator wait, TMEM read, residual
    // loads, stores): eight warps -- two per scheduler -- left the SM waiting on it (round 2e).  A self-gated conv
    // pairs chunk ch with chunk ch + NCH / 2, which has the same parity for N = 32 and 64.
    const int q = warp & 3;
    const int grp = ((warp - 2) >> 2) & 1;
    const int hh = (warp - 2) >> 3;
    constexpr int NCH = N / 8;
    constexpr int KC = N >= 32 ? NCH / 2 : NCH;   // chunks per thread
    constexpr int CSTEP = N >= 32 ? 2 : 1;
    const bool active = N >= 32 || hh == 0;
    const int ch0 = N >= 32 ? hh : 0;
    const int m = q * 32 + lane;
    const int r = m >> 3, cc = m & 7;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    TileIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (; ti.valid(); ti.advance(), ++it) {
      const TileCoord c = ti.coord();
      mbar_wait(smem_u32(&tfull_bar[it & 1]), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const int y = c.y0 + r;
      const int oy = y * p.up + p.ph_a[c.phase];
      const uint32_t sb = smem_u32(&set_bar[it & 1]);
      if (!active || grp >= c.jeff) {   // nothing of this tile for this warp: its share of the set is "drained"
        __syncwarp();
        if (lane == 0) mbar_arrive(sb);
        continue;
      }
      for (int j = grp; j < c.jeff; j += 2) {
        const uint32_t tacc = tmem_base + lane_base + (uint32_t)(((it & 1) * p.J + j) * N);
        const int x = c.x0 + 8 * j + cc;
        const bool inside = (y < p.H) && (x < p.W);
        // wide source: column -> (image, column inside the image); separator columns are not real pixels
        int img = c.b, xi = x;
        bool valid = inside;
        if (p.sep_pitch) {
          img = (int)__umulhi((uint32_t)x, p.sep_magic);
          xi = x - img * p.sep_pitch;
          valid = inside && (img < p.sep_nimg) && (xi < p.sep_w);
        }
        const int ox = xi * p.up + p.ph_b[c.phase];
        // 1. residual loads first: their latency overlaps the TMEM read and nothing below the
        //    slot release depends on the tensor pipe any more
        uint4 r1[KC], r2[KC];
        if (p.res1 != nullptr) {
          const __half* rp = p.res1 + (size_t)img * p.res1_bs + (size_t)p.res1_chunk0 * p.res1_cs + (size_t)oy * p.res1_ys + (size_t)ox * 8;
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int ch = ch0 + CSTEP * k;
            if (valid && ch < p.out_nchunks) r1[k] = *reinterpret_cast<const uint4*>(rp + (size_t)ch * p.res1_cs);
          }
        }
        if (p.res2 != nullptr) {
          const __half* rp = p.res2 + (size_t)img * p.res2_bs + (size_t)p.res2_chunk0 * p.res2_cs + (size_t)oy * p.res2_ys + (size_t)ox * 8;
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int ch = ch0 + CSTEP * k;
            if (valid && ch < p.out_nchunks) r2[k] = *reinterpret_cast<const uint4*>(rp + (size_t)ch * p.res2_cs);
          }
        }
        // 2. drain this warp's chunks of the accumulator into registers (the set goes back to the MMA warp after the
        //    last one)
        uint32_t v[KC * 8];
#pragma unroll
        for (int k = 0; k < KC; ++k)
          tmem_ld8(tacc + (uint32_t)((ch0 + CSTEP * k) * 8), *reinterpret_cast<uint32_t(*)[8]>(&v[k * 8]));
        tmem_ld_wait();
        if (j + 2 >= c.jeff) {  // this warp's last accumulator of the tile is in registers: release its share of the set
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sb);
        }
        // 3. bias, activation, residuals, fp16 pack, 16-byte stores into the destination chunk slice
        __half* op = p.out + (size_t)img * p.out_bs + (size_t)p.out_chunk0 * p.out_cs + (size_t)oy * p.out_ys + (size_t)ox * p.out_px;
        if (valid) {
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int ch = ch0 + CSTEP * k;
            if (ch < p.out_nchunks) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float tv = __uint_as_float(v[k * 8 + e]) + s_bias[c.phase * N + ch * 8 + e];
                if (p.lrelu && !p.act_after_res) tv = lrelu_f(tv, p.slope);
                f[e] = tv;
              }
              if (p.gate == 2) {
                constexpr int PC = (NCH / 2 > 0 ? NCH / 2 : 1);   // chunks per half of a self-gated conv
                constexpr int PK = (PC / CSTEP > 0 ? PC / CSTEP : 1);   // ... in this thread's chunk list
                if (ch < PC && k + PK < KC) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float gv = __uint_as_float(v[(k + PK) * 8 + e]) + s_bias[c.phase * N + (ch + PC) * 8 + e];
                    f[e] = __fdividef(f[e], 1.f + __expf(-gv));
                  }
                }
              }
              if (p.res1 != nullptr) {
                const __half2* hp = reinterpret_cast<const __half2*>(&r1[k]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 rv = __half22float2(hp[e]);
                  if (p.res1_unact) {
                    const float inv = 1.f / p.slope;
                    rv.x = rv.x > 0.f ? rv.x : rv.x * inv;
                    rv.y = rv.y > 0.f ? rv.y : rv.y * inv;
                  }
                  if (p.gate == 1) {
                    // approximate division: the IEEE one takes its slow path (a call) for the zero padding channels
                    f[2 * e] = __fdividef(rv.x, 1.f + __expf(-f[2 * e]));
                    f[2 * e + 1] = __fdividef(rv.y, 1.f + __expf(-f[2 * e + 1]));
                  } else {
                    f[2 * e] = f[2 * e] * p.alpha1 + rv.x;
                    f[2 * e + 1] = f[2 * e + 1] * p.alpha1 + rv.y;
                  }
                }
              }
              if (p.res2 != nullptr) {
                const __half2* hp = reinterpret_cast<const __half2*>(&r2[k]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 rv = __half22float2(hp[e]);
                  f[2 * e] = f[2 * e] * p.alpha2 + rv.x;
                  f[2 * e + 1] = f[2 * e + 1] * p.alpha2 + rv.y;
                }
              }
              if (p.raw != nullptr) {   // pre-activation copy
                uint4 o;
                const __half2 h0 = __floats2half2_rn(f[0], f[1]);
                const __half2 h1 = __floats2half2_rn(f[2], f[3]);
                const __half2 h2 = __floats2half2_rn(f[4], f[5]);
                const __half2 h3 = __floats2half2_rn(f[6], f[7]);
                o.x = *reinterpret_cast<const uint32_t*>(&h0);
                o.y = *reinterpret_cast<const uint32_t*>(&h1);
                o.z = *reinterpret_cast<const uint32_t*>(&h2);
                o.w = *reinterpret_cast<const uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(p.raw + (size_t)img * p.raw_bs + (size_t)(p.raw_chunk0 + ch) * p.raw_cs +
                                          (size_t)oy * p.raw_ys + (size_t)ox * 8) = o;
              }
              if (p.lrelu && p.act_after_res) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = lrelu_f(f[e], p.slope);
              }
              uint4 o;
              const __half2 h0 = __floats2half2_rn(f[0], f[1]);
              const __half2 h1 = __floats2half2_rn(f[2], f[3]);
              const __half2 h2 = __floats2half2_rn(f[4], f[5]);
              const __half2 h3 = __floats2half2_rn(f[6], f[7]);
              o.x = *reinterpret_cast<const uint32_t*>(&h0);
              o.y = *reinterpret_cast<const uint32_t*>(&h1);
              o.z = *reinterpret_cast<const uint32_t*>(&h2);
              o.w = *reinterpret_cast<const uint32_t*>(&h3);
              if (p.out_compact4) {
                *reinterpret_cast<uint2*>(op) = make_uint2(o.x, o.y);
              } else {
                *reinterpret_cast<uint4*>(op + (size_t)ch * p.out_cs) = o;
              }
            }
          }
        } else if (inside && p.out_zero_sep) {
          // separator column of a wide destination: keep the zero padding between images intact
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int ch = ch0 + CSTEP * k;
            if (ch < p.out_nchunks) {
              *reinterpret_cast<uint4*>(op + (size_t)ch * p.out_cs) = make_uint4(0u, 0u, 0u, 0u);
              if (p.raw != nullptr)
                *reinterpret_cast<uint4*>(p.raw + (size_t)img * p.raw_bs + (size_t)(p.raw_chunk0 + ch) * p.raw_cs +
                                          (size_t)oy * p.raw_ys + (size_t)ox * 8) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int N>
int launch_impl(const CUtensorMap* tmap_in, const ConvTcParams& p, int num_sms, cudaStream_t stream) {
  int max_taps = 0;
  for (int i = 0; i < p.nphase; ++i) max_taps = max_taps > p.ph_ntaps[i] ? max_taps : p.ph_ntaps[i];
  const int stage_bytes = conv_tc_a_bytes(p.J, p.dil) + conv_tc_w_bytes(N, max_taps);
  const size_t smem_bytes = (size_t)p.stages * stage_bytes + kConvTailBytes;
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int num_tiles = p.B * p.bands * p.cps * p.nphase;
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, conv_tc_kernel<N>, *tmap_in, p);
}

}  // namespace

int launch_conv_tc(const CUtensorMap* tmap_in, const ConvTcParams& p, int N, int num_sms,
                   cudaStream_t stream) {
  switch (N) {
    case 16: return launch_impl<16>(tmap_in, p, num_sms, stream);
    case 32: return launch_impl<32>(tmap_in, p, num_sms, stream);
    case 64: return launch_impl<64>(tmap_in, p, num_sms, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace innfer
