// upconv_block (block.py:348-361: nearest Upsample(x2) + 3x3 conv + LeakyReLU) with all four output phases
// computed in ONE visit of a source tile.
//
// The generic 9-tap kernel (conv_tc.cu) treats every output phase (a, b) of the folded conv as a tile of its own:
// the source halo tile is fetched four times and the phase's weights travel with every pipeline stage, 11x the
// source bytes in L2 -> shared-memory traffic (ncu: 15 GB for the second upconv of a 63-tile batch) and 20 % tensor
// pipe.  Here a CTA tile is 16 x 8 source pixels; per 16-channel slab the (18 x 10) halo tile is loaded once and the
// 4 phases x 4 folded taps issue 16 MMAs (M = 128, N = 64) into four TMEM accumulators (one per phase); the folded
// weights of the whole conv (4 phases x Cin/16 slabs x 4 taps x 2 KB = 128 KB for 64 -> 64) stay resident in shared
// memory.  Two accumulator sets (2 x 4 x 64 = 512 TMEM columns) alternate between tiles so that the epilogue -- phase
// (a, b) of source pixel (y, x) goes to output pixel (2y + a, 2x + b) -- overlaps the next tile's MMAs.
//
// Same conventions as conv_tc.cu: planar-chunk fp16 tensors, tiled or wide source (separator columns skipped /
// stored as zeros), generic destination strides, producer / issuer / 16 epilogue warps / scout.
#include "conv_tc.cuh"
#include "ptx.cuh"

#ifdef INNFER_ROWS_TRACE
#define UP_TRACE(...) __VA_ARGS__
#else
#define UP_TRACE(...)
#endif

namespace innfer {

namespace {

constexpr int kUpN = 64;
constexpr int kUpPhases = 4, kUpTaps = 4;
constexpr int kUpWh = 10, kUpRh = kPatchRows + 2;            // halo tile of a 16 x 8 patch
constexpr int kUpABytes = 2 * kUpRh * kUpWh * 16;            // 5760 B per 16-channel slab (multiple of 128)
constexpr uint32_t kUpTapUnits = 2 * kUpN;                   // 16-byte units per (phase, slab, tap) weight block
constexpr int kUpThreads = 608;                              // producer, issuer, 16 epilogue warps, scout
constexpr int kUpScoutWarp = 18;

struct UpTile {
  int b, y0, x0;
};

// tiles in (image, band, column-patch) order, column-patch fastest
struct UpIter {
  long long t, stride, total;
  int cps, bands;
  __device__ __forceinline__ void init(const ConvTcParams& p) {
    cps = p.cps;
    bands = p.bands;
    total = (long long)p.B * bands * cps;
    t = blockIdx.x;
    stride = gridDim.x;
  }
  __device__ __forceinline__ bool valid() const { return t < total; }
  __device__ __forceinline__ void advance() { t += stride; }
  __device__ __forceinline__ UpTile coord() const {
    UpTile c;
    const long long r = t / cps;
    c.x0 = (int)(t - r * cps) * 8;
    c.b = (int)(r / bands);
    c.y0 = (int)(r - (long long)c.b * bands) * kPatchRows;
    return c;
  }
};

__global__ void __launch_bounds__(kUpThreads, 1)
conv_up_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel of the stream may take this SM as soon as this CTA leaves it
  const int S = p.stages;
  const uint32_t w_total = (uint32_t)kUpPhases * p.kslabs * kUpTaps * kUpTapUnits * 16u;

  uint8_t* bar_base = smem + w_total + (size_t)S * kUpABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* set_bar = tfull_bar + 2;
  uint64_t* wfull_bar = set_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  volatile uint32_t* ready_cnt = tmem_slot + 1;
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&tfull_bar[i]), 1);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&set_bar[i]), 16);
    mbar_init(smem_u32(wfull_bar), 1);
    *ready_cnt = 0u;
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < kUpPhases * kUpN; i += blockDim.x) s_bias[i] = p.bias[i];
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t w_base = smem_u32(smem);
  const uint32_t ring_base = w_base + w_total;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t wb = smem_u32(wfull_bar);
      mbar_expect_tx(wb, w_total);
      for (uint32_t off = 0; off < w_total; off += 32768u) {
        const uint32_t n = (w_total - off) < 32768u ? (w_total - off) : 32768u;
        bulk_load(w_base + off, reinterpret_cast<const uint8_t*>(p.w) + off, n, wb);
      }
      pdl_wait();   // weights are constants; the activations below belong to the previous kernel of the stream
      int s = 0;
      uint32_t ph = 0;
      UpIter ti;
      ti.init(p);
      for (; ti.valid(); ti.advance()) {
        const UpTile c = ti.coord();
        for (int ks = 0; ks < p.kslabs; ++ks) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, (uint32_t)kUpABytes);
          tma_load_5d(ring_base + (uint32_t)s * kUpABytes, &tmap_in, fb, 0, c.x0 - 1, c.y0 - 1, p.in_chunk0 + 2 * ks, c.b);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (no mbarrier waits: see the scout)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(kUpN);
    constexpr uint32_t a_lbo = kUpRh * kUpWh;
    constexpr uint32_t a_hi = (uint32_t)kUpWh | (1u << 14);      // SBO = one halo-tile row
    constexpr uint32_t b_hi = 8u | (1u << 14);
    const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | ((uint32_t)kUpN << 16);
    const uint32_t ph_units = (uint32_t)p.kslabs * kUpTaps * kUpTapUnits;   // weight units per phase
    uint32_t aoff[kUpPhases][kUpTaps];
#pragma unroll
    for (int ph = 0; ph < kUpPhases; ++ph)
#pragma unroll
      for (int tp = 0; tp < kUpTaps; ++tp) aoff[ph][tp] = (uint32_t)p.tap_hy[ph][tp] * kUpWh + p.tap_hx[ph][tp];
    const uint32_t ready_addr = smem_u32((const void*)ready_cnt);
    uint32_t ready = 0, stage_i = 0, it = 0;
    int s = 0;
    mbar_wait(smem_u32(wfull_bar), 0u);
    UpIter ti;
    ti.init(p);
    for (; ti.valid(); ti.advance(), ++it) {
      const uint32_t set = it & 1u;
      const uint32_t acc0 = tmem_base + set * (uint32_t)(kUpPhases * kUpN);
      for (int ks = 0; ks < p.kslabs; ++ks, ++stage_i) {
        UP_TRACE(if (p.trace && blockIdx.x == 0 && lane == 0 && stage_i < 256) p.trace[stage_i * 4 + 0] = clock64());
        if (ready <= stage_i) {
          uint32_t spins = 0;
          do {
            asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ready) : "r"(ready_addr) : "memory");
            if (++spins > (1u << 26)) __trap();
          } while (ready <= stage_i);
          tc_fence_after();
        }
        UP_TRACE(if (p.trace && blockIdx.x == 0 && lane == 0 && stage_i < 256) p.trace[stage_i * 4 + 1] = clock64());
        const uint32_t sa = ring_base + (uint32_t)s * kUpABytes;
        const uint32_t a_lo = ((sa & 0x3FFFFu) >> 4) | (a_lbo << 16);
        const uint32_t b_lo = b_lo0 + (uint32_t)ks * (kUpTaps * kUpTapUnits);
        const uint32_t first = ks != 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int ph = 0; ph < kUpPhases; ++ph)
#pragma unroll
            for (int tp = 0; tp < kUpTaps; ++tp)
              umma_f16_ss(acc0 + (uint32_t)(ph * kUpN), make_desc64(a_lo + aoff[ph][tp], a_hi),
                          make_desc64(b_lo + (uint32_t)ph * ph_units + (uint32_t)tp * kUpTapUnits, b_hi), idesc,
                          tp == 0 ? first : 1u);
          umma_commit(smem_u32(&empty_bar[s]));
        }
        __syncwarp();
        UP_TRACE(if (p.trace && blockIdx.x == 0 && lane == 0 && stage_i < 256) p.trace[stage_i * 4 + 2] = clock64());
        if (++s == S) s = 0;
      }
      if (leader) umma_commit(smem_u32(&tfull_bar[set]));
      __syncwarp();
    }
  } else if (warp == kUpScoutWarp) {
    // ------------------------------------------------------------ scout: waits for the issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, it = 0, done = 0;
      const uint32_t ready_addr = smem_u32((const void*)ready_cnt);
      UpIter ti;
      ti.init(p);
      for (; ti.valid(); ti.advance(), ++it) {
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
    pdl_wait();
    const int q = warp & 3;
    const int pa = ((warp - 2) >> 2) & 1;    // output row parity of this warp: phases (pa, 0) and (pa, 1)
    const int hc = (warp - 2) >> 3;          // channel half: channels [32 hc, 32 hc + 32)
    const int m = q * 32 + lane;
    const int r = m >> 3, cc = m & 7;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool lrelu = p.lrelu != 0;
    const float slope = p.slope;
    const int nchunks = p.out_nchunks;
    uint32_t it = 0;
    UpIter ti;
    ti.init(p);
    for (; ti.valid(); ti.advance(), ++it) {
      const UpTile c = ti.coord();
      const int y = c.y0 + r, x = c.x0 + cc;
      const bool inside = (y < p.H) && (x < p.W);
      int img = c.b, xi = x;
      bool valid = inside;
      if (p.sep_pitch) {
        img = (int)__umulhi((uint32_t)x, p.sep_magic);
        xi = x - img * p.sep_pitch;
        valid = inside && (img < p.sep_nimg) && (xi < p.sep_w);
      }
      mbar_wait(smem_u32(&tfull_bar[it & 1u]), (it >> 1) & 1u);
      tc_fence_after();
      UP_TRACE(if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < 64) p.trace[2048 + it * 2] = clock64());
      // Sixteen epilogue warps: four per TMEM lane quarter = output row parity x channel half.  A thread drains 32
      // channels of BOTH horizontal phases of its source pixel (two rounds of 16 channels x 2 phases) and stores them
      // to the output pixels (2y + pa, 2x) and (2y + pa, 2x + 1), which are adjacent: one 32-byte store per chunk, so a
      // warp writes 256 contiguous bytes per output row and every 32-byte sector is written whole (with one 16-byte
      // store per phase each sector was filled by two warps at different times).
      // (The epilogue is latency / instruction bound: a tile has 9/4 of the outputs per MMA of a plain 3x3 conv; four
      // warps took 7300 cycles per tile, eight 3600, against ~2100-3000 cycles of MMAs.)
      const uint32_t tacc = tmem_base + lane_base + (uint32_t)(((it & 1u) * kUpPhases + 2 * pa) * kUpN + 32 * hc);
      const int oy = 2 * y + pa, ox = 2 * xi;
      __half* op = p.out + (size_t)img * p.out_bs + (size_t)p.out_chunk0 * p.out_cs + (size_t)oy * p.out_ys + (size_t)ox * 8;
#pragma unroll
      for (int rd = 0; rd < 2; ++rd) {
        uint32_t v0[16], v1[16];
        tmem_ld16(tacc + rd * 16, v0);
        tmem_ld16(tacc + kUpN + rd * 16, v1);
        tmem_ld_wait();
        if (rd == 1) {   // this warp's part of the set is in registers: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&set_bar[it & 1u]));
          UP_TRACE(if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < 64) p.trace[2048 + it * 2 + 1] = clock64());
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int ch = hc * 4 + rd * 2 + k;
          if (ch >= nchunks) continue;
          __half* dst = op + (size_t)ch * p.out_cs;
          if (valid) {
            // (the bias table sits 8 bytes past a 16-byte boundary: 8-byte reads, the same address for the warp)
            const float2* bp = reinterpret_cast<const float2*>(s_bias + ch * 8);
            const float2 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
            const float bb[8] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y, b3.x, b3.y};
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float f0 = __uint_as_float(v0[k * 8 + 2 * e]) + bb[2 * e], f1 = __uint_as_float(v0[k * 8 + 2 * e + 1]) + bb[2 * e + 1];
              float g0 = __uint_as_float(v1[k * 8 + 2 * e]) + bb[2 * e], g1 = __uint_as_float(v1[k * 8 + 2 * e + 1]) + bb[2 * e + 1];
              if (lrelu) {
                f0 = f0 > 0.f ? f0 : f0 * slope;
                f1 = f1 > 0.f ? f1 : f1 * slope;
                g0 = g0 > 0.f ? g0 : g0 * slope;
                g1 = g1 > 0.f ? g1 : g1 * slope;
              }
              const __half2 hf = __floats2half2_rn(f0, f1), hg = __floats2half2_rn(g0, g1);
              w[e] = *reinterpret_cast<const uint32_t*>(&hf);
              w[4 + e] = *reinterpret_cast<const uint32_t*>(&hg);
            }
            st_global_256(dst, w);
          } else if (inside && p.out_zero_sep) {
            const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            st_global_256(dst, z);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace

int conv_up_weight_bytes(int kslabs) { return kUpPhases * kslabs * kUpTaps * (int)kUpTapUnits * 16; }
int conv_up_stage_bytes() { return kUpABytes; }

int launch_conv_up(const CUtensorMap* tmap_in, const ConvTcParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem_bytes = (size_t)conv_up_weight_bytes(p.kslabs) + (size_t)p.stages * kUpABytes + kConvTailBytes;
  cudaError_t e = cudaFuncSetAttribute(conv_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (long long)p.B * p.bands * p.cps;
  const int grid = tiles < num_sms ? (int)tiles : num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kUpThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, conv_up_kernel, *tmap_in, p);
}

}  // namespace innfer
