// Row-streaming 3x3 convolution on the "wide" activation layout: the three VERTICAL taps are folded
// into the N dimension of the MMA and recombined by the epilogue thread that owns the pixel column.
//
// Replaces (fp16 engine path) conv1..conv4 of ResidualDenseBlock_5C (RRDBNet_arch.py:152-165:
// Conv2d(k=3,p=1) + LeakyReLU(0.2), Cout = 32), incl. the ESRGAN+ residual adds and, for nf = 32 nets,
// conv5 with its "*0.2 + x" epilogues; with COUT = 64 (N = 192, two TMEM slots) the residual-free
// 64 -> 64 convs whose weights fit in shared memory (HR_conv0, SRResNet's first block convs) and, as
// clusters of two CTAs (PAIR, tcgen05 cta_group::2), conv5 of the nf = 64 net with both residual
// epilogues (RRDBNet_arch.py:98,165), whose weights only fit when the pair shares them.
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
#include <cstdio>
#include <cstdlib>

#include "conv_rows.cuh"
#include "ptx.cuh"

// clock64 tracing of CTA 0 (tests/gpu_bringup.py --stage trace) is compiled in only with
// -DINNFER_ROWS_TRACE: the samples themselves cost ~50 cycles each on the issuing thread.
#ifdef INNFER_ROWS_TRACE
#define ROWS_TRACE(...) __VA_ARGS__
#else
#define ROWS_TRACE(...)
#endif

namespace innfer {

namespace {

// warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue (two per TMEM lane quarter, each half of the
// output channels), 10 = scout.  352 threads leave the epilogue threads 168 registers (COUT = 64 keeps
// 3 x 32 running sums per thread).  Since round 2e only the Cout = 16 kernel and the INNFER_ROWS_WEPI=0 fallbacks
// run this narrow form.
// The wide epilogue (WEPI: the CTA-pair kernel, the dilated kernels and the WE instantiations of the plain ones) runs 16
// epilogue warps -- four per TMEM lane quarter, 8 (Cout = 32) or 16 (Cout = 64) channels per thread, 608 threads,
// 96 registers: with 12..36 MMAs per row the epilogue is a chain of latencies (accumulator wait, TMEM read, residual
// load, convert, store) and two warps per scheduler could not hide it (ncu on PPON's dilated convs: 1.4 instructions
// per cycle and SM, 2 250 cycles per row for 660 cycles of MMAs; profiles/r02e_epilogue_warps.md).
__host__ __device__ constexpr int rows_threads(int /*cout*/, bool wide_epi) { return wide_epi ? 608 : 352; }
constexpr int kRowPx = 144;        // pixels per staged row segment: 9 groups of 16 (strip of 128 + halo)

struct Piece {
  int m, ya, yb, r0, r1;
};

// Walks the pieces of this CTA's unit range [u, u1).
struct PieceIter {
  long long u, u1;
  int H;
  // In pair mode (p.pair) the unit of work is a PAIR of neighbouring strips handled by the two CTAs of a cluster;
  // `m` then counts strip pairs and CTA `rank` works on strip 2m + rank.
  // Dilation d > 1 (DILV kernels): the rows of one residue class c = y mod d form a sub-image of H = ceil(p.H / d)
  // "virtual" rows in which the vertical taps are neighbours again; `m` then counts (strip, class) pairs,
  // m = strip * d + c, and virtual row k is image row c + k * d (rows >= p.H read as zeros and are not stored).
  __device__ __forceinline__ void init(const ConvRowsParams& p, int d = 1) {
    const long long worker = p.pair ? (blockIdx.x >> 1) : blockIdx.x;
    const long long nworkers = p.pair ? (gridDim.x >> 1) : gridDim.x;
    H = (p.H + d - 1) / d;
    const long long T = (long long)(p.pair ? (p.nstrips + 1) / 2 : p.nstrips * d) * H;
    u = T * worker / nworkers;
    u1 = T * (worker + 1) / nworkers;
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

template <int COUT, int KSLABS, bool RES, bool PAIR, bool DILV = false, bool WE = false, bool NR1 = false>
__global__ void __launch_bounds__(rows_threads(COUT, PAIR || DILV || WE), 1)
conv_rows_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ ConvRowsParams p) {
  constexpr int N = 3 * COUT;                 // dy-major: column dy*COUT + co
  // PAIR: the two CTAs of a cluster issue M = 256 MMAs (tcgen05 cta_group::2) over two neighbouring strips; each CTA
  // stages its own strip and keeps only HALF of the weight columns (N / 2 rows of every B tile), which is what lets
  // conv5's 221 KB of row-streaming weights be resident.  CTA 0 issues, both drain their own accumulators.
  constexpr int NB = PAIR ? N / 2 : N;        // B-tile rows held by this CTA
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int NSLOT = 512 / N;              // 5 (COUT=32) or 2 (COUT=64)
  // channels per epilogue thread: two warps per lane quarter share COUT; COUT = 16 (the net's last conv, N = 48) is
  // drained by ONE warp per quarter, the other four epilogue warps idle
  constexpr bool WEPI = PAIR || DILV || WE;    // 16 epilogue warps (608 threads)
  static_assert(!WE || COUT == 32 || COUT == 64, "the wide epilogue splits 32 or 64 channels over four warp groups");
  constexpr int CH = WEPI ? COUT / 4 : (COUT == 16 ? 16 : COUT / 2);
  constexpr int NGRP = COUT / CH;              // active epilogue warps per lane quarter
  constexpr int SCOUT_WARP = WEPI ? 2 + 4 * NGRP : 10;
  static_assert(!DILV || COUT == 32, "the dilated variant exists for 64 -> 32 convs");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  ROWS_TRACE(if (p.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[3072 + blockIdx.x * 8 + 0] = (long long)gt;
    p.trace[3072 + blockIdx.x * 8 + 1] = clock64();
  });

  // Programmatic dependent launch (p.pdl): let the next conv of the stream take this SM as soon as this CTA leaves it;
  // everything below up to pdl_wait() touches only this kernel's own constants (weights, bias), shared memory and TMEM.
  pdl_launch_dependents();
  const int S = p.stages;
  const int dl = DILV ? p.dil : 1;   // dilation (compile-time 1 for every kernel of the plain nets)
  const uint32_t stage_bytes = (uint32_t)p.kc * kRowPx * 16;
  const uint32_t w_main = (uint32_t)(p.nch / 2) * 3u * 2u * NB * 16u;
  const uint32_t w_total = w_main + (PAIR ? (uint32_t)kRowsIdtBytes : 0u);   // PAIR: + the identity tiles (IDT)
  uint8_t* bar_base = smem + w_total + (size_t)S * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;            // accumulator of a row is complete
  uint64_t* slot_bar = tfull_bar + NSLOT;         // accumulator slot is drained
  uint64_t* wfull_bar = slot_bar + NSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 2);   // wfull_bar[1]: PAIR, peer's weights landed
  volatile uint32_t* ready_cnt = tmem_slot + 1;    // stages whose barriers the scout warp has seen complete
  volatile uint32_t* peer_cnt = tmem_slot + 2;     // PAIR, leader CTA: the same count published by the peer's scout
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&slot_bar[i]), (PAIR ? 8 : 4) * NGRP);   // PAIR: the leader's barrier counts both CTAs' warps
    }
    mbar_init(smem_u32(wfull_bar), 1);
    mbar_init(smem_u32(wfull_bar + 1), 1);
    *ready_cnt = 0u;
    *peer_cnt = 0u;
    fence_mbar_init();
  }
  if (threadIdx.x < COUT) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_pair(smem_u32(tmem_slot), 512u);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(tmem_slot), 512u);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must exist before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t w_base = smem_u32(smem);
  const uint32_t ring_base = w_base + w_total;
  // activations (TMA loads, residual reads, output stores) belong to the previous kernels of the stream: every warp but
  // the producer, which first starts the weight copies, waits for them here
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t wb = smem_u32(wfull_bar);
      mbar_expect_tx(wb, w_total);
      // bulk copies of at most 32 KB each
      for (uint32_t off = 0; off < w_total; off += 32768u) {
        const uint32_t n = (w_total - off) < 32768u ? (w_total - off) : 32768u;
        bulk_load(w_base + off, reinterpret_cast<const uint8_t*>(p.w) + (size_t)rank * w_total + off, n, wb);
      }
      pdl_wait();
      int s = 0;
      uint32_t ph = 0;
      ROWS_TRACE(int tcount = 0);
      PieceIter it;
      it.init(p, dl);
      Piece pc;
      while (it.next(pc)) {
        const int strip = PAIR ? 2 * pc.m + (int)rank : (DILV ? pc.m / dl : pc.m);
        const int cres = DILV ? pc.m % dl : 0;
        const int gx = 8 * strip - 1;  // first 16-pixel group of the strip's halo tile (-1 -> zero fill)
        for (int r = pc.r0; r <= pc.r1; ++r) {
          for (int sub = 0; sub < p.nsub; ++sub) {
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, stage_bytes);
            tma_load_4d(ring_base + (uint32_t)s * stage_bytes, &tmap_in, fb, 0, gx, DILV ? cres + r * dl : r,
                        p.in_chunk0 + sub * p.kc);
            ROWS_TRACE(if (p.trace && blockIdx.x == 0 && tcount < 256) p.trace[1024 + tcount++] = clock64());
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
    // Everything on this warp's path is a potential tensor-pipe bubble (the pipe buffers very few
    // MMAs and these are 48..96-cycle MMAs), so the loop is specialised on the number of K slabs per
    // stage, keeps all of its state as running values (no multiplies, no 64-bit counters) and never
    // touches an mbarrier: an mbarrier wait on this thread costs 170-260 cycles even when the phase
    // completed long ago (measured with clock64: the SYNCS instruction queues behind the MMAs already
    // handed to the pipe).  The scout warp does the waiting and publishes the number of ready stages
    // in shared memory; this thread re-reads that word only when it runs out of known-ready stages,
    // in the middle of a stage, while the first half of the stage's MMAs is still queued.
    const bool leader = elect_one();
#ifdef INNFER_EXPERIMENTS   // INNFER_ROWS_DX0=3 (wrong results): MMAs of half the N extent -- half the B operand per MMA
    const uint32_t idesc = PAIR ? make_idesc_f16_m256(p.dbg_dx0 == 3 ? N / 2 : N) : make_idesc_f16(p.dbg_dx0 == 3 ? N / 2 : N);
#else
    const uint32_t idesc = PAIR ? make_idesc_f16_m256(N) : make_idesc_f16(N);
#endif
    const uint32_t dxu = p.dbg_dx0 == 1 ? 0u : (DILV ? (uint32_t)p.dil : 1u);    // pixels (16-byte units) between the horizontal taps
    constexpr uint32_t a_lbo = kRowPx;                   // 16-byte units between the two K chunks
    constexpr uint32_t a_hi = 8u | (1u << 14);           // SBO = 128 B (8 consecutive pixels)
    constexpr uint32_t b_hi = 8u | (1u << 14);
    constexpr uint32_t b_blk = 2u * NB;                  // 16-byte units per (slab, dx) weight block (of this CTA)
    constexpr uint32_t b_sub_step = (uint32_t)KSLABS * 3u * b_blk;
    constexpr int KH = KSLABS > 1 ? KSLABS / 2 : 1;
    const uint32_t b_lo0 = ((w_base & 0x3FFFFu) >> 4) | ((uint32_t)NB << 16);
    const bool idt = PAIR && p.idt != 0;
    const uint32_t idt_lo = (((w_base + w_main) & 0x3FFFFu) >> 4) | (32u << 16);   // 32 rows per K chunk
    const uint32_t idesc_idt = PAIR ? make_idesc_f16_m256(COUT) : 0u;
    const uint32_t a_step = stage_bytes >> 4;
    const uint32_t a_first = ((ring_base & 0x3FFFFu) >> 4) | (a_lbo << 16);
    const uint32_t a_last = a_first + (uint32_t)(S - 1) * a_step;
    const uint32_t ebar0 = smem_u32(&empty_bar[0]);
    const uint32_t tbar0 = smem_u32(&tfull_bar[0]);
    const uint32_t acc_last = tmem_base + (uint32_t)((NSLOT - 1) * N);
    const uint32_t ready_addr = smem_u32((const void*)ready_cnt);
    const uint32_t peer_addr = smem_u32((const void*)peer_cnt);
    uint32_t nstage = 0;
    {
      PieceIter it;
      it.init(p, dl);
      Piece pc;
      while (it.next(pc)) nstage += (uint32_t)(pc.r1 - pc.r0 + 1);
      nstage *= (uint32_t)p.nsub;
    }
    const int nsub = p.nsub;
    if (PAIR && rank != 0) nstage = 0;   // only the leader CTA issues MMAs
    // number of stages known to be ready: in PAIR mode the minimum over both CTAs' scouts
    auto read_ready = [&]() -> uint32_t {
      uint32_t v;
      if constexpr (PAIR) {
        const uint32_t a = ld_acquire_cluster(ready_addr), b = ld_acquire_cluster(peer_addr);
        v = a < b ? a : b;
      } else {
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(ready_addr) : "memory");
      }
      return v;
    };
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t accum) {
      if constexpr (PAIR) umma_f16_ss_pair(d, ad, bd, idesc, accum);
      else umma_f16_ss(d, ad, bd, idesc, accum);
    };
    auto commit = [&](uint32_t bar) {
      if constexpr (PAIR) umma_commit_pair(bar);   // arrives on the barrier at this offset in BOTH CTAs
      else umma_commit(bar);
    };
    uint32_t ready = 0;
    uint32_t a_lo = a_first, ebar = ebar0, tbar = tbar0, acc = tmem_base, b_lo = b_lo0;
    int sub = 0;
    mbar_wait(smem_u32(wfull_bar), 0u);
    if constexpr (PAIR) {
      // the leader must not issue before the PEER's weights have landed: the peer's issuer warp forwards that event
      if (rank != 0) {
        if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(wfull_bar + 1), 0));
      } else {
        mbar_wait_cluster(smem_u32(wfull_bar + 1), 0u);
      }
    }
    ROWS_TRACE(if (p.trace && lane == 0) p.trace[3072 + blockIdx.x * 8 + 2] = clock64());
    for (uint32_t i = 0; i < nstage; ++i) {
      if (ready <= i) {   // only at the very start, or when the producer is behind
        uint32_t spins = 0;
        do {
          ready = read_ready();
          if (++spins > (1u << 26)) __trap();
        } while (ready <= i);
        tc_fence_after();
      }
      if (leader) {
        // TMEM lane l is output column 128m - 16 + d + l (d = 1: 128m - 15 + l); its tap dx reads pixel l + dx * d of
        // the staged segment, which starts at column 128m - 16
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            mma(acc, make_desc64(a_lo + (uint32_t)(kk * 2 * a_lbo) + (uint32_t)dx * dxu, a_hi),
                make_desc64(b_lo + (uint32_t)((kk * 3 + dx) * b_blk), b_hi), (kk | dx) == 0 ? (sub != 0 ? 1u : 0u) : 1u);
        }
      }
      if (ready <= i + 1 && i + 1 < nstage) {   // look ahead while the pipe is busy with the first half
        ready = read_ready();
        tc_fence_after();
      }
      const bool row_end = sub == nsub - 1;
      if (leader) {
#pragma unroll
        for (int kk = KH; kk < KSLABS; ++kk) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            mma(acc, make_desc64(a_lo + (uint32_t)(kk * 2 * a_lbo) + (uint32_t)dx * dxu, a_hi),
                make_desc64(b_lo + (uint32_t)((kk * 3 + dx) * b_blk), b_hi), 1u);
        }
        if constexpr (PAIR) {
          // IDT: the residual x (input channels 0..63, centre tap) enters the centre-row block of the accumulator
          // through four N = 64 MMAs against identity tiles: D[:, COUT + co] += x[:, co] / alpha1
          if (idt && sub == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_ss_pair(acc + (uint32_t)COUT, make_desc64(a_lo + (uint32_t)(kk * 2 * a_lbo) + dxu, a_hi),
                               make_desc64(idt_lo + (uint32_t)(kk * 64), b_hi), idesc_idt, 1u);
          }
        }
        commit(ebar);
        if (row_end) commit(tbar);
      }
      __syncwarp();
      if (a_lo == a_last) {
        a_lo = a_first;
        ebar = ebar0;
      } else {
        a_lo += a_step;
        ebar += 8u;
      }
      if (row_end) {
        sub = 0;
        b_lo = b_lo0;
        if (acc == acc_last) {
          acc = tmem_base;
          tbar = tbar0;
        } else {
          acc += (uint32_t)N;
          tbar += 8u;
        }
      } else {
        ++sub;
        b_lo += b_sub_step;
      }
    }
    ROWS_TRACE(if (p.trace && lane == 0) p.trace[3072 + blockIdx.x * 8 + 3] = clock64());
    ROWS_TRACE(if (p.trace && lane == 0) p.trace[3072 + blockIdx.x * 8 + 5] = nstage);
  } else if (warp == SCOUT_WARP) {
    // ------------------------------------------------------------ scout: waits for the issuer
    if (lane == 0) {
      long long nrows = 0;
      {
        PieceIter it;
        it.init(p, dl);
        Piece pc;
        while (it.next(pc)) nrows += pc.r1 - pc.r0 + 1;
      }
      int s = 0;
      uint32_t ph = 0;
      int slot = 0;
      uint32_t use = 0;
      uint32_t done = 0;
      // PAIR: the leader's scout also waits for the accumulator slots (its barriers count the epilogue warps of both
      // CTAs); the peer's scout only follows its own stage ring and publishes the count into the LEADER's memory
      const uint32_t pub_addr = (PAIR && rank != 0) ? map_to_cta(smem_u32((const void*)peer_cnt), 0)
                                                    : smem_u32((const void*)ready_cnt);
      for (long long r = 0; r < nrows; ++r) {
        if (!PAIR) mbar_wait(smem_u32(&slot_bar[slot]), (use & 1u) ^ 1u);
        else if (rank == 0) mbar_wait_cluster(smem_u32(&slot_bar[slot]), (use & 1u) ^ 1u);
        if (++slot == NSLOT) {
          slot = 0;
          ++use;
        }
        for (int sub = 0; sub < p.nsub; ++sub) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
          ++done;
          if constexpr (PAIR) st_release_cluster(pub_addr, done);
          else asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(pub_addr), "r"(done) : "memory");
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q4 = warp & 3;
    const int grp = (warp - 2) >> 2;                     // which half of the output channels
    if (grp < NGRP) {
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const float* bias = s_bias + grp * CH;
    int slot = 0;
    uint32_t use = 0;
    ROWS_TRACE(int ecount = 0);
    PieceIter it;
    it.init(p, dl);
    Piece pc;
    while (it.next(pc)) {
      const int strip = PAIR ? 2 * pc.m + (int)rank : (DILV ? pc.m / dl : pc.m);
      const int cres = DILV ? pc.m % dl : 0;
      const int Hv = DILV ? (p.H + dl - 1) / dl : p.H;    // rows of the (virtual) image the pieces walk
      const int xw = 128 * strip - 16 + dl + q4 * 32 + lane;   // wide column of this thread's TMEM lane
      const bool in_range = xw >= 0 && xw < p.Wtot;
      const uint32_t b = __umulhi((uint32_t)(xw < 0 ? 0 : xw), p.magic);
      const int xi = xw - (int)b * p.pitch;
      const bool real = in_range && (int)b < p.nimg && xi < p.Wimg;   // else separator column: zeros
      const size_t col = (size_t)(xw < 0 ? 0 : xw) * 8;
      // wide: image * pitch + xi == xw, so this is xw * 8 for separator columns too
      __half* const obase = p.out + (size_t)b * p.out_bs + (size_t)(p.out_chunk0 + grp * (CH / 8)) * p.out_cs +
                            (size_t)(xw < 0 ? 0 : xi) * p.out_px;
      // NR1: the launch has no first residual (conv5 of RDB3 with the block residual in the accumulator, IDT) -- known
      // at compile time so that its registers go to the one-row-ahead prefetch of the second one
      const __half* const r1base = (RES && !NR1 && p.res1) ? p.res1 + (size_t)(p.res1_chunk0 + grp * (CH / 8)) * p.res_cs + col : nullptr;
      // (the dilated kernels never get a second residual: layers.cu refuses it)
      const __half* const r2base = (RES && !DILV && p.res2) ? p.res2 + (size_t)(p.res2_chunk0 + grp * (CH / 8)) * p.res_cs + col : nullptr;
      const int nchunks = p.out_nchunks - grp * (CH / 8);   // chunks of this thread that exist in the destination
      float accA[CH], accB[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) accA[c] = accB[c] = 0.f;
      // residual inputs of one output row: prefetched into L2 one row stage ahead, loaded chunk by chunk in
      // store_row (keeps the live register set small)
      auto load_side = [&](int y) {
        if (!RES || !real || y < pc.ya) return;
        const int yr = DILV ? cres + y * dl : y;
        if (DILV && yr >= p.H) return;
        const size_t ro = (size_t)yr * p.res_ys;
#pragma unroll
        for (int ch = 0; ch < CH / 8; ++ch) {
          if (r1base) asm volatile("prefetch.global.L2 [%0];" ::"l"(r1base + ro + (size_t)ch * p.res_cs));
          if (r2base) asm volatile("prefetch.global.L2 [%0];" ::"l"(r2base + ro + (size_t)ch * p.res_cs));
        }
      };
      // residual inputs of one output row, straight into registers (issued a whole row stage before their use by the
      // CH = 16 epilogue below: the round trip hides behind the accumulator wait of the next row)
      auto load_res = [&](int y, uint4 (&s1v)[CH / 8], uint4 (&s2v)[CH / 8]) {
        if (!RES || !real || y < pc.ya) return;
        const int yr = DILV ? cres + y * dl : y;
        if (DILV && yr >= p.H) return;
        const size_t ro = (size_t)yr * p.res_ys;
#pragma unroll
        for (int ch = 0; ch < CH / 8; ++ch) {
          if (r1base && ch < nchunks) s1v[ch] = *reinterpret_cast<const uint4*>(r1base + ro + (size_t)ch * p.res_cs);
          if (r2base && ch < nchunks) s2v[ch] = *reinterpret_cast<const uint4*>(r2base + ro + (size_t)ch * p.res_cs);
        }
      };
      // use_pre: 0 = load the residuals here, 1 = preloaded by load_res
      auto store_row_impl = [&](int y, const float (&o)[CH], const int use_pre, const uint4 (&pre1)[CH / 8],
                                const uint4 (&pre2)[CH / 8]) {
#ifdef INNFER_EXPERIMENTS   // INNFER_ROWS_DX0=2 (wrong results): the epilogue only drains TMEM, no residual loads / stores
        if (p.dbg_dx0 == 2) return;
#endif
        if (!in_range || !(real || p.out_wide)) return;
        const int yr = DILV ? cres + y * dl : y;   // image row of virtual row y
        if (DILV && yr >= p.H) return;
        const size_t ro = (size_t)yr * p.res_ys;
        __half* op = obase + (size_t)yr * p.out_ys;
        // all residual loads of the row are issued before the first use: one exposed L2 round trip per row instead of
        // one per chunk (measured: 3 200 cycles per row for 4 chunks x 2 residuals when loaded chunk by chunk)
        uint4 s1v[CH / 8], s2v[CH / 8];
        if (RES && real) {
          if (use_pre == 1) {
#pragma unroll
            for (int ch = 0; ch < CH / 8; ++ch) {
              s1v[ch] = pre1[ch];
              s2v[ch] = pre2[ch];
            }
          } else {
#pragma unroll
            for (int ch = 0; ch < CH / 8; ++ch) {
              if (r1base && ch < nchunks) s1v[ch] = *reinterpret_cast<const uint4*>(r1base + ro + (size_t)ch * p.res_cs);
              if (r2base && ch < nchunks) s2v[ch] = *reinterpret_cast<const uint4*>(r2base + ro + (size_t)ch * p.res_cs);
            }
          }
        }
#pragma unroll
        for (int ch = 0; ch < CH / 8; ++ch) {
          if (ch >= nchunks) break;
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (real) {
            const uint4 s1 = s1v[ch], s2 = s2v[ch];
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = o[ch * 8 + e] + bias[ch * 8 + e];
            if (p.lrelu && !(DILV && p.act_after_res)) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * p.slope;
            }
            if (PAIR && p.idt) {   // the residual is already in the accumulator (as x / alpha1)
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] *= p.alpha1;
            }
            if (r1base) {
              const __half2* hp = reinterpret_cast<const __half2*>(&s1);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 v = __half22float2(hp[e]);
                if (DILV && p.res1_unact) {   // the side input is the activated copy of the value
                  const float inv = 1.f / p.slope;
                  v.x = v.x > 0.f ? v.x : v.x * inv;
                  v.y = v.y > 0.f ? v.y : v.y * inv;
                }
                f[2 * e] = f[2 * e] * p.alpha1 + v.x;
                f[2 * e + 1] = f[2 * e + 1] * p.alpha1 + v.y;
              }
            }
            if (r2base) {
              const __half2* hp = reinterpret_cast<const __half2*>(&s2);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = __half22float2(hp[e]);
                f[2 * e] = f[2 * e] * p.alpha2 + v.x;
                f[2 * e + 1] = f[2 * e + 1] * p.alpha2 + v.y;
              }
            }
            if constexpr (DILV) {
              if (p.raw) {   // pre-activation copy (PPON: the running sum the next dilated conv adds to)
                uint4 rk;
                const __half2 r0 = __floats2half2_rn(f[0], f[1]), r1 = __floats2half2_rn(f[2], f[3]);
                const __half2 r2 = __floats2half2_rn(f[4], f[5]), r3 = __floats2half2_rn(f[6], f[7]);
                rk.x = *reinterpret_cast<const uint32_t*>(&r0);
                rk.y = *reinterpret_cast<const uint32_t*>(&r1);
                rk.z = *reinterpret_cast<const uint32_t*>(&r2);
                rk.w = *reinterpret_cast<const uint32_t*>(&r3);
                *reinterpret_cast<uint4*>(p.raw + (size_t)(p.raw_chunk0 + grp * (CH / 8) + ch) * p.res_cs + ro + col) = rk;
              }
              if (p.lrelu && p.act_after_res) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * p.slope;
              }
            }
            const __half2 h0 = __floats2half2_rn(f[0], f[1]);
            const __half2 h1 = __floats2half2_rn(f[2], f[3]);
            const __half2 h2 = __floats2half2_rn(f[4], f[5]);
            const __half2 h3 = __floats2half2_rn(f[6], f[7]);
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            pk.z = *reinterpret_cast<const uint32_t*>(&h2);
            pk.w = *reinterpret_cast<const uint32_t*>(&h3);
          } else if (DILV && p.raw) {   // separator column: keep the zero padding of the second destination too
            *reinterpret_cast<uint4*>(p.raw + (size_t)(p.raw_chunk0 + grp * (CH / 8) + ch) * p.res_cs + ro + col) =
                make_uint4(0u, 0u, 0u, 0u);
          }
          if (p.out_compact4) *reinterpret_cast<uint2*>(op) = make_uint2(pk.x, pk.y);
          else *reinterpret_cast<uint4*>(op + (size_t)ch * p.out_cs) = pk;
        }
      };
      auto store_row = [&](int y, const float (&o)[CH]) {
        const uint4 none[CH / 8] = {};
        store_row_impl(y, o, 0, none, none);
      };
      // PPON's dilated convs (12 MMAs per row, their epilogue is the critical path) and conv5 of RDB3 on the CTA pair
      // (one residual left, known at compile time): residuals in registers one row ahead.  Elsewhere the extra
      // registers spill, so the L2 prefetch stays.
      constexpr bool kRegPrefetch = RES && (DILV || (PAIR && NR1));
      uint4 cur1[CH / 8] = {}, cur2[CH / 8] = {};   // kRegPrefetch: residuals of the row stored in this iteration
      for (int r = pc.r0; r <= pc.r1; ++r) {
        ROWS_TRACE(if (r > pc.r0 || ecount > 0) ++ecount);
        uint4 nxt1[CH / 8] = {}, nxt2[CH / 8] = {};
        if constexpr (kRegPrefetch) load_res(r, nxt1, nxt2);   // row r is stored one iteration (one row stage) later
        else load_side(r);   // row r is finished one iteration (one whole row stage) later: enough to cover an HBM miss
        mbar_wait(smem_u32(&tfull_bar[slot]), use & 1u);
        tc_fence_after();
        ROWS_TRACE(const bool tr = p.trace && blockIdx.x == 0 && threadIdx.x == 64 && ecount < 256;
                   if (tr) p.trace[2048 + ecount] = clock64());
        const uint32_t tacc = tmem_base + lane_base + (uint32_t)(slot * N + grp * CH);
        auto release_slot = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(map_to_cta(smem_u32(&slot_bar[slot]), 0));
            else mbar_arrive(smem_u32(&slot_bar[slot]));
          }
          if (++slot == NSLOT) {
            slot = 0;
            ++use;
          }
        };
        if constexpr (PAIR || (WE && COUT == 64)) {
          // 16 channels per thread through one 16-register TMEM buffer.
          uint32_t v[16];
          if constexpr (!RES) {
            // no residual arrays to hold: the finished row goes to `o`, the running sums move up and the slot returns
            // to the MMA warp BEFORE the row is converted and stored (with two slots per CTA the store used to sit
            // between two rows' MMAs)
            float o[16];
            tmem_ld16(tacc + 2 * COUT, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = accA[c] + __uint_as_float(v[c]);
            ROWS_TRACE(if (tr) p.trace[2304 + ecount] = clock64());
            tmem_ld16(tacc + COUT, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) accA[c] = accB[c] + __uint_as_float(v[c]);
            tmem_ld16(tacc, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) accB[c] = __uint_as_float(v[c]);
            release_slot();
            ROWS_TRACE(if (tr) p.trace[2560 + ecount] = clock64());
            if (r - 1 >= pc.ya) store_row(r - 1, o);
            ROWS_TRACE(if (tr) p.trace[2816 + ecount] = clock64());
          } else {
          // residual epilogues: the finished row is completed in accA and stored before the other two blocks are read
          // (96 registers per thread with 608 threads: a separate copy next to the residuals would spill)
          tmem_ld16(tacc + 2 * COUT, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) accA[c] += __uint_as_float(v[c]);
          ROWS_TRACE(if (tr) p.trace[2304 + ecount] = clock64());
          // (measured, round 2e: issuing the residual loads before the accumulator wait does not help -- the epilogue
          // is itself the critical path here, so there is no wait to hide them behind, and the extra live registers
          // spill at 96 per thread: 455 -> 462 us per 63-tile launch)
          if constexpr (kRegPrefetch) {   // conv5 of RDB3: the RRDB input row was loaded one row stage ago
            if (r - 1 >= pc.ya) store_row_impl(r - 1, accA, 1, cur1, cur2);
#pragma unroll
            for (int ch = 0; ch < CH / 8; ++ch) {
              cur1[ch] = nxt1[ch];
              cur2[ch] = nxt2[ch];
            }
          } else {
            if (r - 1 >= pc.ya) store_row(r - 1, accA);
          }
          ROWS_TRACE(if (tr) p.trace[2560 + ecount] = clock64());
          tmem_ld16(tacc + COUT, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) accA[c] = accB[c] + __uint_as_float(v[c]);
          tmem_ld16(tacc, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) accB[c] = __uint_as_float(v[c]);
          release_slot();
          ROWS_TRACE(if (tr) p.trace[2816 + ecount] = clock64());
          }
        } else if constexpr (CH == 8) {
          // dilated kernels: 8 channels per thread, all three blocks at once, slot released before the row is stored
          float o[CH];
          uint32_t v0[8], v1[8], v2[8];
          tmem_ld8(tacc, v0);
          tmem_ld8(tacc + COUT, v1);
          tmem_ld8(tacc + 2 * COUT, v2);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            o[c] = accA[c] + __uint_as_float(v2[c]);
            accA[c] = accB[c] + __uint_as_float(v1[c]);
            accB[c] = __uint_as_float(v0[c]);
          }
          release_slot();
          if constexpr (kRegPrefetch) {
            if (r - 1 >= pc.ya) store_row_impl(r - 1, o, 1, cur1, cur2);
#pragma unroll
            for (int ch = 0; ch < CH / 8; ++ch) {
              cur1[ch] = nxt1[ch];
              cur2[ch] = nxt2[ch];
            }
          } else {
            if (r - 1 >= pc.ya) store_row(r - 1, o);
          }
        } else if constexpr (CH == 16) {
          // all three blocks of this thread's channels are read at once and the slot goes back to the MMA warp
          // before the finished row is stored
          float o[CH];
          uint32_t v0[16], v1[16], v2[16];
          tmem_ld16(tacc, v0);
          tmem_ld16(tacc + COUT, v1);
          tmem_ld16(tacc + 2 * COUT, v2);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            o[c] = accA[c] + __uint_as_float(v2[c]);
            accA[c] = accB[c] + __uint_as_float(v1[c]);
            accB[c] = __uint_as_float(v0[c]);
          }
          release_slot();
          if constexpr (kRegPrefetch) {
            if (r - 1 >= pc.ya) store_row_impl(r - 1, o, 1, cur1, cur2);
#pragma unroll
            for (int ch = 0; ch < CH / 8; ++ch) {
              cur1[ch] = nxt1[ch];
              cur2[ch] = nxt2[ch];
            }
          } else {
            if (r - 1 >= pc.ya) store_row(r - 1, o);
          }
        } else {
          // COUT = 64: 3 x 32 running values per thread, read through one 16-register buffer.  The finished row is
          // completed IN accA and stored before the other two blocks are read, so that no separate copy of it is
          // live while the residuals of the row are in registers (no spills: they would go to L2 here); the slot is
          // released after the store, which the second TMEM slot and the ~3 500-cycle row stages absorb.
          uint32_t v[16];
#pragma unroll
          for (int g = 0; g < CH / 16; ++g) {
            tmem_ld16(tacc + 2 * COUT + g * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) accA[g * 16 + c] += __uint_as_float(v[c]);
          }
          if (r - 1 >= pc.ya) store_row(r - 1, accA);
#pragma unroll
          for (int g = 0; g < CH / 16; ++g) {
            tmem_ld16(tacc + COUT + g * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) accA[g * 16 + c] = accB[g * 16 + c] + __uint_as_float(v[c]);
            tmem_ld16(tacc + g * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) accB[g * 16 + c] = __uint_as_float(v[c]);
          }
          release_slot();
        }
      }
      if (pc.yb == Hv) {   // bottom row: the row below is zero padding
        if constexpr (kRegPrefetch) store_row_impl(Hv - 1, accA, 1, cur1, cur2);
        else store_row(Hv - 1, accA);
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA may leave while the other still signals / reads it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512u);
    else tmem_dealloc(tmem_base, 512u);
  }
  ROWS_TRACE(if (p.trace && threadIdx.x == 0) p.trace[3072 + blockIdx.x * 8 + 4] = clock64());
}

template <int COUT, int KSLABS, bool RES, bool PAIR, bool DILV = false, bool WE = false, bool NR1 = false>
int launch_rows_res(const CUtensorMap* tmap_in, const ConvRowsParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem_bytes = conv_rows_weight_bytes(p.nch, COUT) / (PAIR ? 2 : 1) + (PAIR ? kRowsIdtBytes : 0) +
                            (size_t)p.stages * conv_rows_stage_bytes(p.kc) + 1024;
  auto kern = conv_rows_kernel<COUT, KSLABS, RES, PAIR, DILV, WE, NR1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int dl = DILV ? p.dil : 1;
  const long long T = (long long)(PAIR ? (p.nstrips + 1) / 2 : p.nstrips * dl) * ((p.H + dl - 1) / dl);
  long long workers = PAIR ? num_sms / 2 : num_sms;   // PAIR: one cluster of two CTAs per worker
  if (T < workers) workers = T;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(PAIR ? 2 * workers : workers));
  cfg.blockDim = dim3(rows_threads(COUT, PAIR || DILV || WE));
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int na = 0;
  // INNFER_L2_PERSIST=1 (experiment): mark this conv's wide output chunks as persisting in L2, so that the next conv of
  // the dense block -- which reads them together with everything before them -- finds them there.  Pays only if a
  // batch's 32-channel slice fits the persisting carve-out (innfer_rrdb_set_max_batch / INNFER_MB <= ~28 tiles).
  static const int l2_persist = getenv("INNFER_L2_PERSIST") ? atoi(getenv("INNFER_L2_PERSIST")) : 0;
  if (l2_persist && p.out_wide) {
    static size_t max_window = 0, carve = 0;
    if (!max_window) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceProp prop;
      cudaGetDeviceProperties(&prop, dev);
      max_window = (size_t)prop.accessPolicyMaxWindowSize;
      carve = (size_t)prop.persistingL2CacheMaxSize;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      fprintf(stderr, "innfer: L2 persisting carve-out %zu MB, max window %zu MB\n", carve >> 20, max_window >> 20);
    }
    const size_t bytes = (size_t)p.out_nchunks * (size_t)p.out_cs * sizeof(__half);
    if (bytes <= max_window) {
      attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
      attr[na].val.accessPolicyWindow.base_ptr = p.out + (size_t)p.out_chunk0 * p.out_cs;
      attr[na].val.accessPolicyWindow.num_bytes = bytes;
      attr[na].val.accessPolicyWindow.hitRatio = bytes <= carve ? 1.f : (float)carve / (float)bytes;
      attr[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      ++na;
    }
  }
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (p.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return (int)cudaLaunchKernelEx(&cfg, kern, *tmap_in, p);
}

template <int COUT, int KSLABS>
int launch_rows_impl(const CUtensorMap* tmap_in, const ConvRowsParams& p, int num_sms, cudaStream_t stream) {
  // the residual-free variant (conv1..conv4 of the plain net) keeps its epilogue free of the side-input code
  const bool res = p.res1 || p.res2;
  if (p.act_after_res || p.raw || p.dil != 1) {   // PPON's dilated 64 -> 32 convs
    if constexpr (COUT == 32 && KSLABS == 4) {
      if (p.pair || p.res2 || p.dil < 1 || p.dil > 8) return (int)cudaErrorInvalidValue;
      return res ? launch_rows_res<COUT, KSLABS, true, false, true>(tmap_in, p, num_sms, stream)
                 : launch_rows_res<COUT, KSLABS, false, false, true>(tmap_in, p, num_sms, stream);
    }
    return (int)cudaErrorInvalidValue;
  }
  if (p.pair) {
    if constexpr (COUT == 64 && KSLABS == 6) {   // conv5 of the nf = 64 net is the one conv that needs (and gains from) the pair
      if (res && !p.res1) return launch_rows_res<COUT, KSLABS, true, true, false, false, true>(tmap_in, p, num_sms, stream);
      return res ? launch_rows_res<COUT, KSLABS, true, true>(tmap_in, p, num_sms, stream)
                 : launch_rows_res<COUT, KSLABS, false, true>(tmap_in, p, num_sms, stream);
    }
    return (int)cudaErrorInvalidValue;
  }
  if constexpr (COUT == 32 || COUT == 64) {
    // 16 epilogue warps of 8 channels for the Cout = 32 convs (conv1..conv4 of a dense block): INNFER_ROWS_WEPI bit 0
    // for the residual-free ones, bit 1 for those with residual inputs (ESRGAN+, nf = 32 nets); bits 2 / 3: the same
    // for the single-CTA Cout = 64 convs (16 channels per thread: SRResNet's block convs, PPON's c1, HR_conv0)
    static const int wepi = getenv("INNFER_ROWS_WEPI") ? atoi(getenv("INNFER_ROWS_WEPI")) : 15;
    constexpr int sh = COUT == 64 ? 2 : 0;
    if (!res && (wepi & (1 << sh))) return launch_rows_res<COUT, KSLABS, false, false, false, true>(tmap_in, p, num_sms, stream);
    if (res && (wepi & (2 << sh))) return launch_rows_res<COUT, KSLABS, true, false, false, true>(tmap_in, p, num_sms, stream);
  }
  return res ? launch_rows_res<COUT, KSLABS, true, false>(tmap_in, p, num_sms, stream)
             : launch_rows_res<COUT, KSLABS, false, false>(tmap_in, p, num_sms, stream);
}

template <int COUT>
int launch_rows_k(const CUtensorMap* tmap_in, const ConvRowsParams& p, int num_sms, cudaStream_t stream) {
  switch (p.kc / 2) {
    case 1: return launch_rows_impl<COUT, 1>(tmap_in, p, num_sms, stream);
    case 2: return launch_rows_impl<COUT, 2>(tmap_in, p, num_sms, stream);
    case 3: return launch_rows_impl<COUT, 3>(tmap_in, p, num_sms, stream);
    case 4: return launch_rows_impl<COUT, 4>(tmap_in, p, num_sms, stream);
    case 5: return launch_rows_impl<COUT, 5>(tmap_in, p, num_sms, stream);
    case 6: return launch_rows_impl<COUT, 6>(tmap_in, p, num_sms, stream);
    case 7: return launch_rows_impl<COUT, 7>(tmap_in, p, num_sms, stream);
    case 8: return launch_rows_impl<COUT, 8>(tmap_in, p, num_sms, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace

int conv_rows_stage_bytes(int kc) { return kc * kRowPx * 16; }
int conv_rows_weight_bytes(int nch, int cout) { return (nch / 2) * 3 * 2 * (3 * cout) * 16; }

int launch_conv_rows(const CUtensorMap* tmap_in, const ConvRowsParams& p, int cout, int num_sms, cudaStream_t stream) {
  if (cout == 16) return p.kc == 8 ? launch_rows_impl<16, 4>(tmap_in, p, num_sms, stream) : (int)cudaErrorInvalidValue;
  if (cout == 32) return launch_rows_k<32>(tmap_in, p, num_sms, stream);
  if (cout == 64) return launch_rows_k<64>(tmap_in, p, num_sms, stream);
  return (int)cudaErrorInvalidValue;
}

}  // namespace innfer
