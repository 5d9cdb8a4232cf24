// "dx-taps-as-N" 3x3 convolution for Cout = 32 (conv1..conv4 of every residual dense block).
//
// Why: with M=128 x N=32 MMAs the SS-mode A operand (4 KB of shared-memory reads per MMA) bounds
// conv_tc_kernel<32> at ~44 cycles per 16-cycle MMA.  Here the three horizontal taps are folded
// into the N dimension instead of shifting A:
//     P[y, q, dx, co] = sum_{dy, ci} W[dy, dx, ci, co] * in[y + dy - 1, q, ci]        (one GEMM, N = 96)
//     out[y, x, co]   = P[y, x-1, 0, co] + P[y, x, 1, co] + P[y, x+1, 2, co]
// so A is read 3x (once per dy) instead of 9x and each MMA is N=96 (48 tensor cycles).  The
// horizontal recombination runs in the epilogue with two warp shuffles per channel: the thread
// that owns accumulator position q produces output x = q - 1 from its own dx=2 column block, its
// left neighbour's dx=1 block and its second-left neighbour's dx=0 block (lanes of one 8-pixel row
// group; the two leftmost positions take them from the previous sub-patch, kept in registers).
//
// Geometry: a CTA tile covers 16 rows x (8J - 2) output columns; its accumulator positions
// q = x0-1 .. x0-2+8J are exactly J sub-patches of 8 columns, and the TMA box (8J x 18 pixels, two
// chunks) needs no horizontal halo beyond them.  Accumulators (96 fp32 columns) rotate through five
// TMEM slots with one mbarrier each, so the MMA warp can run ahead of the epilogue by one slot.
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace innfer {

namespace {

constexpr int kDxN = 96;      // 3 dx x 32 output channels
constexpr int kDxSlots = 5;   // 5 x 96 = 480 TMEM columns
constexpr int kDxChunks = 4;  // 8-channel chunks per pipeline stage (32 channels)
constexpr int kDxThreads = 320;  // producer, issuer, 8 epilogue warps (2 per TMEM lane quarter)

struct DxIter {
  int cp, band, b;
  int dcp, dband, db;
  int cps, bands, B, J, W, OW;
  __device__ __forceinline__ void init(const ConvTcParams& p, int t0, int stride) {
    cps = p.cps; bands = p.bands; B = p.B; J = p.J; W = p.W; OW = 8 * p.J - 2;
    cp = t0 % cps; int r = t0 / cps;
    band = r % bands; b = r / bands;
    dcp = stride % cps; r = stride / cps;
    dband = r % bands; db = r / bands;
  }
  __device__ __forceinline__ bool valid() const { return b < B; }
  __device__ __forceinline__ void advance() {
    cp += dcp;
    int c = cp >= cps ? 1 : 0;
    cp -= c ? cps : 0;
    band += dband + c;
    c = band >= bands ? 1 : 0;
    band -= c ? bands : 0;
    b += db + c;
  }
  __device__ __forceinline__ int x0() const { return cp * OW; }
  __device__ __forceinline__ int y0() const { return band * kPatchRows; }
  __device__ __forceinline__ int nout() const {
    const int rem = W - cp * OW;
    return rem < OW ? rem : OW;
  }
  __device__ __forceinline__ int jeff() const { return (nout() + 2 + 7) >> 3; }
};

__global__ void __launch_bounds__(kDxThreads, 1)
conv_dx_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int Wh = 8 * p.J;                                  // smem tile width in pixels (no x halo)
  // One pipeline stage = 32 input channels (4 chunks, two K=16 MMA slabs): with N=96 MMAs a
  // 16-channel stage is only ~900 tensor-pipe cycles and the issuing warp's per-stage barrier
  // round trip (~360 cycles, measured) would be a 28 % bubble.
  const int a_bytes = kDxChunks * kHaloRows * Wh * 16;     // multiple of 128
  const int w_bytes = (kDxChunks / 2) * 3 * 2 * kDxN * 16; // 18432 per 32-channel stage
  // debug bit 3: all weight slabs of this conv (kslabs x 9216 B <= 92 KB) stay resident in shared
  // memory for the whole persistent kernel instead of travelling with every stage;
  // bit 4: the resident block sits in front of the stage ring instead of behind it.
  const bool resident = (p.debug & 8) != 0;
  const int stage_bytes = resident ? a_bytes : a_bytes + w_bytes;
  const int S = p.stages;
  const uint32_t w_total = resident ? (uint32_t)p.kslabs * w_bytes : 0u;
  const uint32_t ring_off = (resident && (p.debug & 16)) ? w_total : 0u;

  uint8_t* bar_base = smem + (size_t)S * stage_bytes + w_total;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;                     // one per slot: accumulator complete
  uint64_t* slot_bar = tfull_bar + kDxSlots;               // one per slot: accumulator drained
  uint64_t* wfull_bar = slot_bar + kDxSlots;               // resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(wfull_bar), 1);
    for (int i = 0; i < kDxSlots; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&slot_bar[i]), 8);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem) + ring_off;
  const uint32_t w_base = ring_off ? smem_u32(smem) : smem_base + (uint32_t)S * stage_bytes;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (resident) {
        const uint32_t wb = smem_u32(wfull_bar);
        mbar_expect_tx(wb, w_total);
        for (int ks = 0; ks < p.kslabs; ++ks)
          bulk_load(w_base + (uint32_t)ks * w_bytes, reinterpret_cast<const uint8_t*>(p.w) + (size_t)ks * w_bytes,
                    w_bytes, wb);
      }
      int s = 0;
      uint32_t ph = 0;
      DxIter ti;
      ti.init(p, blockIdx.x, gridDim.x);
      for (; ti.valid(); ti.advance()) {
        const int x0 = ti.x0(), y0 = ti.y0();
        for (int ks = 0; ks < p.kslabs; ++ks) {
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[s]);
          const uint32_t dstA = smem_base + (uint32_t)s * stage_bytes;
          mbar_expect_tx(fb, (uint32_t)(resident ? a_bytes : a_bytes + w_bytes));
          if (p.debug & 4)
            tma_load_4d(dstA, &tmap_in, fb, (x0 - 1) * 8, y0 - 1, p.in_chunk0 + kDxChunks * ks, ti.b);
          else
            tma_load_5d(dstA, &tmap_in, fb, 0, x0 - 1, y0 - 1, p.in_chunk0 + kDxChunks * ks, ti.b);
          if (!resident)
            bulk_load(dstA + a_bytes, reinterpret_cast<const uint8_t*>(p.w) + (size_t)ks * w_bytes, w_bytes, fb);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one lean thread)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(kDxN);
    const uint32_t a_lbo = (uint32_t)(kHaloRows * Wh);   // 16-byte units
    const uint32_t a_sbo = (uint32_t)Wh;
    const uint32_t a_hi = a_sbo | (1u << 14);
    const uint32_t b_hi = 8u | (1u << 14);               // SBO = 128 B
    const uint32_t b_lbo = (uint32_t)kDxN;               // 96 rows * 16 B
    const uint32_t dy_stride = 2u * kDxN;                // 16-byte units per dy block of a weight slab
    const uint32_t slab_stride = 3u * dy_stride;         // 16-byte units per 16-channel weight slab
    int s = 0;
    uint32_t ph = 0;
    int slot0 = 0;          // slot of sub-patch 0 of the current tile (accumulators rotate over 5 slots)
    uint32_t use0 = 0;      // how many times slot0's ring position has wrapped
    DxIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    if (resident) mbar_wait(smem_u32(wfull_bar), 0u);
    for (; ti.valid(); ti.advance()) {
      const int jeff = ti.jeff();
      for (int ks = 0; ks < p.kslabs; ++ks) {
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t a_lo = ((sa & 0x3FFFFu) >> 4) | (a_lbo << 16);
        const uint32_t wsrc = resident ? w_base + (uint32_t)ks * w_bytes : sa + a_bytes;
        const uint32_t b_lo = ((wsrc & 0x3FFFFu) >> 4) | (b_lbo << 16);
        const uint32_t first = ks != 0 ? 1u : 0u;
        const bool last = ks == p.kslabs - 1;
        for (int j = 0; j < jeff; ++j) {
          int slot = slot0 + j;
          uint32_t use = use0;
          if (slot >= kDxSlots) {
            slot -= kDxSlots;
            ++use;
          }
          if (ks == 0 && !(p.debug & 32)) {
            mbar_wait(smem_u32(&slot_bar[slot]), (use & 1u) ^ 1u);  // previous user of the slot drained
            tc_fence_after();
          }
          const uint32_t acc = tmem_base + (uint32_t)(slot * kDxN);
          const uint32_t aj = a_lo + 8u * j;
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < kDxChunks / 2; ++kk) {
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
                umma_f16_ss(acc, make_desc64(aj + kk * 2u * a_lbo + dy * a_sbo, a_hi),
                            make_desc64(b_lo + kk * slab_stride + dy * dy_stride, b_hi), idesc,
                            (kk | dy) == 0 ? first : 1u);
            }
            if (last) umma_commit(smem_u32(&tfull_bar[slot]));  // this accumulator is complete
          }
        }
        if (leader) umma_commit(smem_u32(&empty_bar[s]));
        __syncwarp();
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
      }
      slot0 += jeff;
      if (slot0 >= kDxSlots) {
        slot0 -= kDxSlots;
        ++use0;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    // 8 warps: warp & 3 selects the TMEM lane quarter, (warp - 2) / 4 the half of the 32 output
    // channels this warp recombines (16 channels = 2 chunks, 3 x 16 accumulator columns).
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q4 * 32 + lane;
    const int r = m >> 3, cc = m & 7;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const size_t oplane = (size_t)p.Hout * p.Wout;
    // shuffle plan: position q needs dx=1 of q-1 and dx=0 of q-2 (see file header)
    const int src1 = cc >= 1 ? lane - 1 : lane + 7;
    const int src0 = cc >= 2 ? lane - 2 : lane + 6;
    const bool send_prev1 = cc == 7, send_prev0 = cc >= 6;
    float bias[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) bias[c] = s_bias[half * 16 + c];
    const bool noshfl = (p.debug & 64) != 0;   // timing experiments only (wrong results)
    const bool nostore = (p.debug & 128) != 0;
    int slot = 0;
    uint32_t use = 0;
    DxIter ti;
    ti.init(p, blockIdx.x, gridDim.x);
    for (; ti.valid(); ti.advance()) {
      const int jeff = ti.jeff();
      const int x0 = ti.x0(), nout = ti.nout();
      const int y = ti.y0() + r;
      float prev0[16], prev1[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) prev0[c] = prev1[c] = 0.f;
      for (int j = 0; j < jeff; ++j) {
        mbar_wait(smem_u32(&tfull_bar[slot]), use & 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_base + (uint32_t)(slot * kDxN + half * 16);
        uint32_t v0[16], v1[16], v2[16];
        tmem_ld16(tacc, v0);
        tmem_ld16(tacc + 32, v1);
        tmem_ld16(tacc + 64, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&slot_bar[slot]));
        if (++slot == kDxSlots) {
          slot = 0;
          ++use;
        }
        const int xo = 8 * j + cc - 2;   // output column relative to x0
        const bool valid = (xo >= 0) && (xo < nout) && (y < p.H);
        __half* op = p.out + (((size_t)ti.b * p.out_CT + p.out_chunk0 + half * 2) * oplane + (size_t)y * p.Wout + (x0 + xo)) * 8;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = ch * 8 + e;
            const float cur0 = __uint_as_float(v0[c]);
            const float cur1 = __uint_as_float(v1[c]);
            const float cur2 = __uint_as_float(v2[c]);
            const float l1 = noshfl ? cur1 : __shfl_sync(0xffffffffu, send_prev1 ? prev1[c] : cur1, src1);
            const float l0 = noshfl ? cur0 : __shfl_sync(0xffffffffu, send_prev0 ? prev0[c] : cur0, src0);
            prev0[c] = cur0;
            prev1[c] = cur1;
            float t = (l0 + l1) + cur2 + bias[c];
            if (p.lrelu) t = t > 0.f ? t : t * p.slope;
            f[e] = t;
          }
          if (valid && !nostore) {
            uint4 o;
            const __half2 h0 = __floats2half2_rn(f[0], f[1]);
            const __half2 h1 = __floats2half2_rn(f[2], f[3]);
            const __half2 h2 = __floats2half2_rn(f[4], f[5]);
            const __half2 h3 = __floats2half2_rn(f[6], f[7]);
            o.x = *reinterpret_cast<const uint32_t*>(&h0);
            o.y = *reinterpret_cast<const uint32_t*>(&h1);
            o.z = *reinterpret_cast<const uint32_t*>(&h2);
            o.w = *reinterpret_cast<const uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(op + (size_t)ch * oplane * 8) = o;
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

int conv_dx_stage_bytes(int J) { return kDxChunks * kHaloRows * 8 * J * 16; }
int conv_dx_weight_bytes(int kstages) { return kstages * (kDxChunks / 2) * 3 * 2 * kDxN * 16; }

int launch_conv_dx(const CUtensorMap* tmap_in, const ConvTcParams& p, int num_sms, cudaStream_t stream) {
  const bool resident = (p.debug & 8) != 0;
  const size_t smem_bytes = resident
                                ? (size_t)p.stages * conv_dx_stage_bytes(p.J) + conv_dx_weight_bytes(p.kslabs) + 1024
                                : (size_t)p.stages * (conv_dx_stage_bytes(p.J) + conv_dx_weight_bytes(1)) + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int num_tiles = p.B * p.bands * p.cps;
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  conv_dx_kernel<<<grid, kDxThreads, smem_bytes, stream>>>(*tmap_in, p);
  return (int)cudaGetLastError();
}

}  // namespace innfer
