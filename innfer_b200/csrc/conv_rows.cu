// Row-streaming 3x3 convolution on the "wide" activation layout: the three VERTICAL taps are folded
// into the N dimension of the MMA and recombined by the epilogue thread that owns the pixel column.
//
// Replaces (fp16 engine path) conv1..conv4 of ResidualDenseBlock_5C (RRDBNet_arch.py:152-165:
// Conv2d(k=3,p=1) + LeakyReLU(0.2), Cout = 32) and, with COUT = 64, the 64->64 convs of the tail.
//
// Wide layout: the B tile images of a batch stand side by side in one image [chunk][H][Wtot][8],
// image b in columns [b*pitch, b*pitch + Wimg), the `pitch - Wimg` separator columns hold zeros (they
// are the conv's zero padding between neighbouring images; every conv writes zeros there again).
//
// GEMM view, per image row r and 128-pixel strip m (an "M-tile" = 128 consecutive pixels of ONE row):
//     Q[r, x, dy, co] = sum_{dx, ci} W[co, ci, dy, dx] * in[r, x + dx - 1, ci]          N = 3 * COUT
//     out[y, x, co]   = Q[y-1, x, 0, co] + Q[y, x, 1, co] + Q[y+1, x, 2, co] + bias
// The dx shift is a 16-byte shift of the A descriptor inside the row's halo tile; the dy recombination
// adds accumulators of three different rows at the SAME TMEM lane, so the epilogue thread of pixel
// column x keeps two running sums in registers while the rows stream by: no shuffles, no shared
// memory, no halo columns.  Every input row is loaded once (one TMA box per row and channel group,
// all channels of the row accumulate into one TMEM slot), the conv's weights stay resident in shared
// memory, and a CTA walks down a strip so that consecutive row accumulators rotate through the TMEM
// slots while the epilogue drains them.
//
// Work split: the nstrips * H (strip, row) units are cut into 148 equal contiguous ranges; a range is
// a few "pieces" (strip, [ya, yb)), each of which reads input rows ya-1 .. yb.
#include "conv_rows.cuh"
#include "ptx.cuh"

namespace innfer {

namespace {

constexpr int kRowsThreads = 320;  // producer, issuer, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kRowPx = 144;        // pixels per staged row segment: 9 groups of 16 (strip of 128 + halo)

struct Piece {
  int m, ya, yb, r0, r1;
};

// Walks the pieces of this CTA's unit range [u, u1).
struct PieceIter {
  long long u, u1;
  int H;
  __device__ __forceinline__ void init(const ConvRowsParams& p) {
    const long long T = (long long)p.nstrips * p.H;
    u = T * blockIdx.x / gridDim.x;
    u1 = T * (blockIdx.x + 1) / gridDim.x;
    H = p.H;
  }
  __device__ __forceinline__ bool next(Piece& pc) {
    if (u >= u1) return false;
    pc.m = (int)(u / H);
    pc.ya = (int)(u - (long long)pc.m * H);
    const long long left = u1 - u;
    pc.yb = (H - pc.ya) < left ? H : pc.ya + (int)left;
    pc.r0 = pc.ya > 0 ? pc.ya - 1 : 0;
    pc.r1 = pc.yb < H ? pc.yb : H - 1;
    u += pc.yb - pc.ya;
    return true;
  }
};

template <int COUT>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ ConvRowsParams p) {
  constexpr int N = 3 * COUT;                 // dy-major: column dy*COUT + co
  constexpr int NSLOT = 512 / N;              // 5 (COUT=32) or 2 (COUT=64)
  constexpr int CH = COUT / 2;                // channels per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int S = p.stages;
  const uint32_t stage_bytes = (uint32_t)p.kc * kRowPx * 16;
  const uint32_t w_total = (uint32_t)(p.nch / 2) * 3u * 2u * N * 16u;
  uint8_t* bar_base = smem + w_total + (size_t)S * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;            // accumulator of a row is complete
  uint64_t* slot_bar = tfull_bar + NSLOT;         // accumulator slot is drained
  uint64_t* wfull_bar = slot_bar + NSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&slot_bar[i]), 8);
    }
    mbar_init(smem_u32(wfull_bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < COUT) s_bias[threadIdx.x] = p.bias[threadIdx.x];
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
      // bulk copies of at most 32 KB each
      for (uint32_t off = 0; off < w_total; off += 32768u) {
        const uint32_t n = (w_total - off) < 32768u ? (w_total - off) : 32768u;
        bulk_load(w_base + off, reinterpret_cast<const uint8_t*>(p.w) + off, n, wb);
      }
      int s = 0;
      uint32_t ph = 0;
      PieceIter it;
      it.init(p);
      Piece pc;
      while (it.next(pc)) {
        const int gx = 8 * pc.m - 1;  // first 16-pixel group of the strip's halo tile (-1 -> zero fill)
        for (int r = pc.r0; r <= pc.r1; ++r) {
          for (int sub = 0; sub < p.nsub; ++sub) {
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, stage_bytes);
            tma_load_4d(ring_base + (uint32_t)s * stage_bytes, &tmap_in, fb, 0, gx, r, p.in_chunk0 + sub * p.kc);
            if (++s == S) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(N);
    constexpr uint32_t a_lbo = kRowPx;                   // 16-byte units between the two K chunks
    const uint32_t a_hi = 8u | (1u << 14);               // SBO = 128 B (8 consecutive pixels)
    const uint32_t b_hi = 8u | (1u << 14);
    const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | ((uint32_t)N << 16);
    constexpr uint32_t b_blk = 2u * N;                   // 16-byte units per (slab, dx) weight block
    const int kslabs = p.kc >> 1;                        // K=16 slabs per sub-stage
    int s = 0;
    uint32_t ph = 0;
    int slot = 0;
    uint32_t use = 0;
    mbar_wait(smem_u32(wfull_bar), 0u);
    PieceIter it;
    it.init(p);
    Piece pc;
    while (it.next(pc)) {
      for (int r = pc.r0; r <= pc.r1; ++r) {
        mbar_wait(smem_u32(&slot_bar[slot]), (use & 1u) ^ 1u);   // drained NSLOT rows ago
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(slot * N);
        for (int sub = 0; sub < p.nsub; ++sub) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint32_t sa = ring_base + (uint32_t)s * stage_bytes;
          const uint32_t a_lo = ((sa & 0x3FFFFu) >> 4) | (a_lbo << 16);
          const uint32_t b_lo = b_lo0 + (uint32_t)(sub * kslabs) * 3u * b_blk;
          if (leader) {
            // TMEM lane l is output column 128m - 15 + l; its tap dx reads pixel l + dx of the staged
            // segment, which starts at column 128m - 16
            umma_f16_ss(acc, make_desc64(a_lo, a_hi), make_desc64(b_lo, b_hi), idesc, sub != 0 ? 1u : 0u);
            umma_f16_ss(acc, make_desc64(a_lo + 1u, a_hi), make_desc64(b_lo + b_blk, b_hi), idesc, 1u);
            umma_f16_ss(acc, make_desc64(a_lo + 2u, a_hi), make_desc64(b_lo + 2u * b_blk, b_hi), idesc, 1u);
#pragma unroll 2
            for (int kk = 1; kk < kslabs; ++kk) {
              const uint32_t ak = a_lo + (uint32_t)kk * 2u * a_lbo;
              const uint32_t bk = b_lo + (uint32_t)kk * 3u * b_blk;
              umma_f16_ss(acc, make_desc64(ak, a_hi), make_desc64(bk, b_hi), idesc, 1u);
              umma_f16_ss(acc, make_desc64(ak + 1u, a_hi), make_desc64(bk + b_blk, b_hi), idesc, 1u);
              umma_f16_ss(acc, make_desc64(ak + 2u, a_hi), make_desc64(bk + 2u * b_blk, b_hi), idesc, 1u);
            }
            umma_commit(smem_u32(&empty_bar[s]));
          }
          __syncwarp();
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
        if (leader) umma_commit(smem_u32(&tfull_bar[slot]));
        __syncwarp();
        if (++slot == NSLOT) {
          slot = 0;
          ++use;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const float* bias = s_bias + half * CH;
    const bool nostore = (p.debug & 128) != 0;
    int slot = 0;
    uint32_t use = 0;
    PieceIter it;
    it.init(p);
    Piece pc;
    while (it.next(pc)) {
      const int xw = 128 * pc.m - 15 + q4 * 32 + lane;   // wide column of this thread's TMEM lane
      const bool in_range = xw >= 0 && xw < p.Wtot;
      const uint32_t b = __umulhi((uint32_t)(xw < 0 ? 0 : xw), p.magic);
      const int xi = xw - (int)b * p.pitch;
      const bool real = in_range && (int)b < p.nimg && xi < p.Wimg;   // else separator column: zeros
      __half* const obase = p.out + (size_t)(p.out_chunk0 + half * (CH / 8)) * p.out_cs + (size_t)(xw < 0 ? 0 : xw) * 8;
      float accA[CH], accB[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) accA[c] = accB[c] = 0.f;
      auto store_row = [&](int y, const float (&o)[CH]) {
        if (!in_range || nostore) return;
        __half* op = obase + (size_t)y * p.out_ys;
#pragma unroll
        for (int ch = 0; ch < CH / 8; ++ch) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (real) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float t = o[ch * 8 + e] + bias[ch * 8 + e];
              if (p.lrelu) t = t > 0.f ? t : t * p.slope;
              f[e] = t;
            }
            const __half2 h0 = __floats2half2_rn(f[0], f[1]);
            const __half2 h1 = __floats2half2_rn(f[2], f[3]);
            const __half2 h2 = __floats2half2_rn(f[4], f[5]);
            const __half2 h3 = __floats2half2_rn(f[6], f[7]);
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            pk.z = *reinterpret_cast<const uint32_t*>(&h2);
            pk.w = *reinterpret_cast<const uint32_t*>(&h3);
          }
          *reinterpret_cast<uint4*>(op + (size_t)ch * p.out_cs) = pk;
        }
      };
      for (int r = pc.r0; r <= pc.r1; ++r) {
        mbar_wait(smem_u32(&tfull_bar[slot]), use & 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_base + (uint32_t)(slot * N + half * CH);
        uint32_t v0[CH], v1[CH], v2[CH];
#pragma unroll
        for (int g = 0; g < CH / 16; ++g) {
          tmem_ld16(tacc + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&v0[g * 16]));
          tmem_ld16(tacc + COUT + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&v1[g * 16]));
          tmem_ld16(tacc + 2 * COUT + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&v2[g * 16]));
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&slot_bar[slot]));
        if (++slot == NSLOT) {
          slot = 0;
          ++use;
        }
        if (r - 1 >= pc.ya) {
          float o[CH];
#pragma unroll
          for (int c = 0; c < CH; ++c) o[c] = accA[c] + __uint_as_float(v2[c]);
          store_row(r - 1, o);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          accA[c] = accB[c] + __uint_as_float(v1[c]);
          accB[c] = __uint_as_float(v0[c]);
        }
      }
      if (pc.yb == p.H) store_row(p.H - 1, accA);   // bottom row: the row below is zero padding
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

template <int COUT>
int launch_rows_impl(const CUtensorMap* tmap_in, const ConvRowsParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem_bytes = conv_rows_weight_bytes(p.nch, COUT) + (size_t)p.stages * conv_rows_stage_bytes(p.kc) + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv_rows_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const long long T = (long long)p.nstrips * p.H;
  const int grid = T < num_sms ? (int)T : num_sms;
  conv_rows_kernel<COUT><<<grid, kRowsThreads, smem_bytes, stream>>>(*tmap_in, p);
  return (int)cudaGetLastError();
}

}  // namespace

int conv_rows_stage_bytes(int kc) { return kc * kRowPx * 16; }
int conv_rows_weight_bytes(int nch, int cout) { return (nch / 2) * 3 * 2 * (3 * cout) * 16; }

int launch_conv_rows(const CUtensorMap* tmap_in, const ConvRowsParams& p, int cout, int num_sms, cudaStream_t stream) {
  if (cout == 32) return launch_rows_impl<32>(tmap_in, p, num_sms, stream);
  if (cout == 64) return launch_rows_impl<64>(tmap_in, p, num_sms, stream);
  return (int)cudaErrorInvalidValue;
}

}  // namespace innfer
