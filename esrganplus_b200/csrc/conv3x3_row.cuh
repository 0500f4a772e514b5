// conv3x3_row.cuh — row-streaming, output-stationary fused 3x3 convolution for sm_100a (images wider
// than ~64 px: the benchmark path).
//
// Same contract as conv3x3_tc.cuh (the fused dense-block conv of block.py:260-268 / 287-291 and the
// plain conv_blocks of architecture.py:55-71), different decomposition:
//
//   M-tile  : ONE image row segment of 128 pixels (lane == column).  For input row r the MMAs compute
//                 Q_r[x, ky, co] = sum_{kx, ci} X[r, x+kx-1, ci] * W[co, ci, ky, kx]
//             i.e. the three kernel ROWS ky are stacked along N (N = 3*Cout = 96: 83 % of the tensor
//             pipe instead of 40 % at N = 32, see conv3x3_tc.cuh), and the three kernel COLUMNS are
//             three MMAs whose A operand is the same shared-memory row shifted by kx pixels (shifted
//             UMMA descriptors into a 130-pixel TMA box; the hardware swizzle is address based, so a
//             128-byte shift stays consistent with what TMA wrote).
//   Output  : out[y] = Q_{y-1}[ky=0] + Q_y[ky=1] + Q_{y+1}[ky=2].  Tensor memory is a ring of BN-column
//             blocks, one per OUTPUT row, laid out so that the blocks of rows r+1, r, r-1 are adjacent:
//             the N = 3*BN accumulator of input row r IS those three blocks, every MMA accumulates, and
//             the sum over ky happens in the tensor pipe.  The epilogue reads one block per output
//             row, zeroes it for its next occupant and hands it back.  (Two of every NBLK rows straddle
//             the end of the ring and are issued as two narrower MMAs.)
//   Schedule: CTA b owns the contiguous output rows [b*U/G, (b+1)*U/G) (U = n * column blocks * H) and
//             streams input rows ya-1 .. yb; the two halo rows per CTA are the only recomputation.
//   Warps   : 0-11 epilogue = 3 warpgroups taking output rows round-robin (thread == pixel, TMEM lane
//             quarter == warp % 4); 12 TMA producer (ring of row buffers, weights resident); 13-14 MMA
//             issuers.  tcgen05.mma issue is nearly synchronous (the pipe queues ~1-2 instructions), so
//             a single issuer leaves a bubble at every barrier wait / commit (85-99 vs 57-61 cycles per
//             N = 96 MMA, tools/ubench_row.cu).  With row_alt (ESRP_VARIANT_ROW_ALT, set by
//             the engine's inference plans) the two issuers ALTERNATE whole input rows: warp (I & 1) issues every tap of row I and passes a turn token; tcgen05.commit
//             only tracks the MMAs of the committing thread, so a block barrier collects the commit of
//             the issuer of its last row r+1 (which also issued r-1), the commit of the issuer of its
//             middle row r, and a plain arrival of the warp that idles during r+1 (count 3).  row_alt = 0
//             (default) is the earlier protocol: both warps split the taps of every row and both commit
//             (count 2).
//   Barriers: every barrier has in-order waiters that cannot be lapped, however long a waiter is delayed:
//             a row buffer / block is only recycled after BOTH issuers have observed every barrier of the
//             row that completes it (their commits resp. the idle warp's arrival; the producer waits on
//             that block barrier) and after the owning warpgroup released it (issuers).
//   Epilogue memory traffic: thread == pixel makes a plain NHWC access one LSU wavefront per lane.  fp32
//             operands may therefore be [n][h][c/4][w][4] (f32_planar: consecutive lanes touch consecutive
//             16 bytes), and bf16 outputs go through a 4x4 transpose inside each group of four lanes so
//             that one store instruction writes 8 x 64 contiguous bytes.
//   Slices  : a launch may co-schedule nsl slices of BN output channels (nsl = 2: conv5): CTAs
//             nsl*i .. nsl*i+nsl-1 walk the same rows with the weights / bias / channel offsets of their
//             slice, so the rows are fetched from DRAM once and every CTA owns nsl times more rows.
#pragma once
#ifndef ESRP_SYNCCHECK_PAD
// Per-row timeline events of the issuer / epilogue / producer of CTA 0 (tools/trace_conv.py): compiled in with
// -DESRP_TRACE_FINE only, they sit in the issuing thread's instruction stream.
#ifdef ESRP_TRACE_FINE
#define ESRP_FINE_TRACE(stmt) stmt
#else
#define ESRP_FINE_TRACE(stmt)
#endif
#define ESRP_SYNCCHECK_PAD 0  // experiment (tools/gpu_r2_p.sh): move tok[] off shared-memory offset 0x148
#endif
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"
#include "conv3x3_tc.cuh"  // load_residual, trace_ev, kMaxStages, kSmemFixed
#include "conv_params.h"
#include "esrp_philox.cuh"
#include "esrp_ptx.cuh"

namespace esrp {

constexpr int kRowWGs = 3;                     // epilogue warpgroups (output rows in flight)
constexpr int kRowEpiWarps = 4 * kRowWGs;
constexpr int kRowMmaWarps = 2;                // MMA issuer warps (they split the taps of every row)
constexpr int kRowThreads = 32 * (kRowEpiWarps + 1 + kRowMmaWarps);
constexpr int kRowTile = 128;                  // output columns per M-tile
constexpr int kMaxBlocks = 16;                 // TMEM output-row blocks in the ring

// Walks the row segments of this CTA: identical in the three warp roles.
struct SegWalk {
  int u, u_end;
  int img, x0, ya, yb;  // current segment: output rows [ya, yb) of column block x0 of image img
  // cta / ncta: position among the CTAs that split the rows (== blockIdx.x / gridDim.x unless the launch
  // co-schedules output slices, see ConvKParams::nsl)
  int img_off;           // CTA pairs: the second CTA walks the same rows of image img + n / 2
  __device__ __forceinline__ SegWalk(const ConvKParams& p, int cta, int ncta, int img_off_ = 0) {
    const long long U = p.units_total;
    u = static_cast<int>(U * cta / ncta);
    u_end = static_cast<int>(U * (cta + 1) / ncta);
    img = x0 = ya = yb = 0;
    img_off = img_off_;
  }
  __device__ __forceinline__ bool next(const ConvKParams& p) {
    if (u >= u_end) return false;
    const int col = u / p.h;
    ya = u - col * p.h;
    const int cnt = min(u_end - u, p.h - ya);
    yb = ya + cnt;
    img = col / p.x_tiles;
    x0 = (col - img * p.x_tiles) * kRowTile;
    img += img_off;
    u += cnt;
    return true;
  }
};

template <int GC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[GC]) {
  if constexpr (GC == 16) {
    tmem_ld_x16(taddr, v);
  } else {
    tmem_ld_x8(taddr, v);
  }
}

// Taps of one K-chunk issued by issuer warp MW: warp 0 takes kx = 1 and the first half of kx = 0, warp 1
// the rest (+ the conv1x1).  All MMAs accumulate: the blocks were zeroed by their previous reader.
// WRAP = false is the straight-line common case (one N = 3*BN MMA per tap, compile-time descriptors: the
// issue sequence must stay at ~2 integer ops per MMA because the tensor pipe queues almost nothing);
// WRAP = true splits every tap in two narrower MMAs where the three blocks straddle the end of the ring.
template <int KC, int BN, int MW, bool WRAP, int KSN = KC / 16, bool PAIR = false>  // KSN: K-slices of 16 channels to issue (the rest: zero weights)
__device__ __forceinline__ void issue_taps(uint32_t dA, uint32_t dB, uint32_t idA, uint32_t idB, uint32_t bB,
                                           uint32_t al, uint32_t bl, uint32_t desc_hi, uint32_t w_block_desc) {
  constexpr int RB = KC * 2, KS = KC / 16;
  constexpr uint32_t ID_FULL = PAIR ? umma_idesc_bf16_m256(3 * BN) : umma_idesc_bf16_m128(3 * BN);
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
    for (int ks = 0; ks < KSN; ++ks) {
      const bool warp0 = kx == 1 || (kx == 0 && ks < KS / 2);
      if (warp0 != (MW == 0)) continue;
      const uint32_t a_d = al + ((kx * RB + ks * 32) >> 4);  // the 130-pixel row shifted by kx pixels
      const uint32_t b_d = bl + kx * w_block_desc + ((ks * 32) >> 4);
      if (PAIR) {
        umma_f16_ss2_2sm(dA, a_d, desc_hi, b_d, desc_hi, ID_FULL, 1u);
      } else if (!WRAP) {
        umma_f16_ss2(dA, a_d, desc_hi, b_d, desc_hi, ID_FULL, 1u);
      } else {
        umma_f16_ss2(dA, a_d, desc_hi, b_d, desc_hi, idA, 1u);
        umma_f16_ss2(dB, a_d, desc_hi, b_d + bB, desc_hi, idB, 1u);
      }
    }
  }
}

// PAIR: launched as clusters of two CTAs (cta_group::2).  The two CTAs stream the SAME rows of two different images
// (img, img + n/2: identical segment structure, so ring positions, barrier phases and TMEM addresses coincide); every MMA is
// M = 256 x N = 3*BN issued by the leader's issuer warps, each CTA holding HALF of the weight rows (B is split along N), so
// per MMA a CTA reads 4 KB of A + 1.5 KB of B from shared memory instead of 4 + 3 (the shared-memory port is what bounds
// the single-CTA steady state) and the resident weights take half the space.  Barriers the issuers wait on live in the
// leader: the row tiles of both CTAs complete on the leader's full_bar (cta_group::2 TMA), both CTAs' epilogues arrive on the
// leader's blk_empty; commits are multicast to the blk_full of both CTAs (own producer / epilogue wait locally).
template <int KC, int BN, bool AUX, bool EXT, bool PAIR = false>
__global__ void __launch_bounds__(kRowThreads, 1)
conv3x3_row_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                   const __grid_constant__ ConvKParams p) {
  constexpr int RB = KC * 2;
  constexpr int KS = KC / 16;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * RB;
  constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
  // Ring of output-row blocks (positions descend with the output sequence number, so that the blocks of the rows r+1, r,
  // r-1 an input row accumulates into are adjacent).  Default: a power-of-two ring; the accumulator window of 2 of every
  // NBLK input rows straddles its end and is issued as two narrower MMAs per tap.
  // SHADOW (the CTA-pair variant, whose B halves must be the same rows for every MMA): NBLK ring positions plus two SHADOW
  // blocks behind them that stand in for positions 0 / 1 when the window would straddle the end, so the window is ALWAYS
  // one contiguous N = 3*BN accumulator.  The rows at positions 0 / 1 then have their sum spread over two blocks (main: the
  // contributions issued while the window was at the bottom of the ring, shadow: the ones issued from its top) and the
  // epilogue adds them.  Measured on the single-CTA kernel as well (round 2): 1.3 % faster on config 2 (13.04 -> 12.88 ms),
  // but WHICH rows are summed in two parts depends on how the rows are cut over the CTAs, i.e. a tile's result would
  // depend on its position in the batch at fp32-rounding level (amplified to bf16 noise by 345 requantising convs): not
  // worth the invariant, so the default path keeps the split MMAs.
  constexpr bool SHADOW = PAIR;
  constexpr int NBLK = SHADOW ? (AUX ? 7 : 14) : ((AUX || BN == 64) ? 8 : 16);  // ring positions
  constexpr int NMAIN = SHADOW ? NBLK + 2 : NBLK;          // main blocks incl. the two shadows
  constexpr int AUX_COL0 = NMAIN * BN;                     // conv1x1 blocks live behind the main blocks
  static_assert(!PAIR || (BN == 32 && KC == 64 && !EXT), "CTA pairs: dense-block convs of the inference plan only");
  constexpr int nb_rows = AUX ? 4 * BN : 3 * BN;           // B rows per tap block in global memory
  constexpr int w_block_bytes_g = nb_rows * RB;
  constexpr int w_chunk_bytes_g = 3 * w_block_bytes_g;
  constexpr int w_block_bytes = (PAIR ? nb_rows / 2 : nb_rows) * RB;  // resident in shared memory (PAIR: this CTA's half:
  constexpr int w_chunk_bytes = 3 * w_block_bytes;                    //   rows [48 r, 48 r + 48) (+ [96 + 16 r, + 16) conv1x1))
  constexpr int aux_row0 = PAIR ? 3 * BN / 2 : 3 * BN;     // first conv1x1 row of a resident tap block
  constexpr int GC = BN < 16 ? BN : 16;                    // output channels per epilogue round
  constexpr int ROUNDS = BN / GC;
  static_assert(NBLK <= kMaxBlocks && (NMAIN + (AUX ? NBLK : 0)) * BN <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  // slots 1019..1023 of the epilogue's trace row: clock64 / %globaltimer at kernel entry, at the start of the role loops, at exit
  ESRP_FINE_TRACE(if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) { p.trace[2 * 1024 + 1023] = clock64(); long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[2 * 1024 + 1022] = gt; })
  // per-CTA %globaltimer at entry (row 0, slots 512 + CTA) and exit (row 1): launch-to-launch gaps, tools/trace_gap.py
  ESRP_FINE_TRACE(if (p.trace && threadIdx.x == 0 && blockIdx.x < 512) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[512 + blockIdx.x] = gt; })

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages] one per row buffer
  uint64_t* blk_full = full_bar + kMaxStages;               // [kMaxBlocks] block complete (2 issuer commits)
  uint64_t* blk_empty = blk_full + kMaxBlocks;              // [kMaxBlocks] block read + zeroed (4 warps)
  uint64_t* wfull = blk_empty + kMaxBlocks;                 // [1]
  uint64_t* tok = wfull + (ESRP_SYNCCHECK_PAD ? 2 : 1);     // [2] issue turn: tok[w] = "warp w may issue"
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + 512);  // (its own 128-byte line: tcgen05.alloc writes it asynchronously)
  uint32_t* buf_bar = tmem_holder + 1;                      // [kMaxStages] producer: barrier/parity that
  uint32_t* buf_par = buf_bar + kMaxStages;                 // [kMaxStages]   frees each row buffer
  float* bias_s = reinterpret_cast<float*>(smem + 1024);    // [BN]

  const int w_res_bytes = p.num_chunks * w_chunk_bytes;     // weights are always resident
  uint8_t* w_res = smem + kSmemFixed;
  uint8_t* stage0 = w_res + w_res_bytes;
  const int a_bytes = p.a_stage_bytes;                      // one chunk tile (TMA box rounded up to 1 KB)
  const int row_bytes = a_bytes * p.num_chunks;             // one row buffer
  const int D = p.stages;                                   // row buffers

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Co-scheduled output slices (a 64-channel conv over a K too large for one resident weight set): CTAs
  // nsl*i .. nsl*i+nsl-1 walk the SAME rows, each with the weights / bias / channel offsets of its own slice.
  // They run in step on neighbouring SMs, so the input rows come from DRAM once and from L2 afterwards, and
  // every CTA owns nsl times more rows (half the halo recomputation of nsl separate launches).
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0: leader of the pair
  const int bid = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nbid = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int img_off = PAIR ? static_cast<int>(rank) * (p.n >> 1) : 0;
  const int nsl = p.nsl > 1 ? p.nsl : 1;
  const int sl = nsl > 1 ? bid % nsl : 0;
  const int cta = bid / nsl, ncta = nbid / nsl;
  const int csh = sl * BN;                                  // channel shift of every global channel offset
  const uint8_t* const w_src = p.w_packed + static_cast<size_t>(sl) * p.sl_stride;

  if (warp == kRowEpiWarps && lane == 0) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < (p.chunk_bars ? D * p.num_chunks : D); ++i) mbar_init(&full_bar[i], 1);
    for (int i = 0; i < kMaxBlocks; ++i) {
      // row-alternating issue (row_alt == 2): + one plain arrival of the warp that does NOT issue the completing row
      mbar_init(&blk_full[i], (PAIR && p.pair_single) ? 1 : p.row_alt == 2 ? kRowMmaWarps + 1 : kRowMmaWarps);
      mbar_init(&blk_empty[i], PAIR ? 8 : 4);  // (PAIR: the four warps of a warpgroup of BOTH CTAs; only the leader's is used)
      mbar_arrive_cnt(&blk_empty[i], PAIR ? 8 : 4);  // phase 0 = "the block is free": complete from the start (no wait relies on the
                                          // parity of a phase that never existed; compute-sanitizer synccheck flags those)
    }
    mbar_init(wfull, (PAIR && rank == 0) ? 2 : 1);  // leader of a pair: + the peer's "my half of B has landed"
    mbar_init(&tok[0], 1);
    mbar_init(&tok[1], 1);
    mbar_arrive(&tok[0]);                 // warp 0 holds the first turn
    fence_barrier_init();
  }
  if (warp == kRowEpiWarps + 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_holder, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_holder, 512);
      tmem_relinquish();
    }
  }
  grid_dep_launch_dependents();
  if (threadIdx.x < BN)  // (weights / bias: written long ago)
    bias_s[threadIdx.x] =
        p.bias ? reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.bias) + static_cast<size_t>(sl) * p.sl_stride)[threadIdx.x]
               : 0.f;
  tcgen05_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // (PAIR: the peer's barriers are initialised too)
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // every block starts zeroed: all MMAs accumulate
  if (warp < kRowEpiWarps) {  // lane quarter warp % 4, a third of the columns each
    const uint32_t la = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = (warp >> 2) * 16; c < NMAIN * BN; c += 16 * kRowWGs) tmem_st_zero_x16(la + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // (PAIR: the peer's blocks are zeroed before the leader issues)
  tcgen05_fence_after();

  ESRP_FINE_TRACE(if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[2 * 1024 + 1019] = clock64();)
  // ring position (block index) of output sequence number O: descending, so that the blocks of the output
  // rows r+1, r, r-1 an input row r accumulates into are adjacent with ascending columns (ky = 0, 1, 2)
  auto pos = [](uint32_t O) -> uint32_t { return (NBLK - 1) - (O % NBLK); };
  auto use = [](uint32_t O) -> uint32_t { return (O / NBLK) & 1; };

  if (p.dbg & ESRP_DBG_EMPTY) {
    // timing experiment: prologue + teardown only
  } else if (warp == kRowEpiWarps) {
    // ===================================== TMA producer =====================================
    // Row-buffer ring: buffer b holds the num_chunks K-chunk tiles of one input row.  It is free again
    // when both issuers have committed that row, which they signal on the blk_full barrier of the output
    // row that input completes (one commit per row and warp serves the epilogue and the producer).
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(w_res_bytes));
      if constexpr (PAIR) {
        for (int c = 0; c < p.num_chunks; ++c)
          for (int kx = 0; kx < 3; ++kx) {
            const uint8_t* src = w_src + static_cast<size_t>(c) * w_chunk_bytes_g + kx * w_block_bytes_g;
            uint8_t* dst = w_res + c * w_chunk_bytes + kx * w_block_bytes;
            bulk_load_1d(dst, src + rank * (3 * BN / 2) * RB, (3 * BN / 2) * RB, wfull);
            if (AUX) bulk_load_1d(dst + aux_row0 * RB, src + (3 * BN + rank * (BN / 2)) * RB, (BN / 2) * RB, wfull);
          }
      } else {
        for (int c = 0; c < p.num_chunks; ++c)
          bulk_load_1d(w_res + c * w_chunk_bytes, w_src + static_cast<size_t>(c) * w_chunk_bytes,
                       w_chunk_bytes, wfull);
      }
      uint32_t tn = 0;
      trace_ev(p, 0, tn);
      grid_dep_wait();  // activations of the previous kernel must be complete before the first TMA load
      const int nch = p.num_chunks;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.a_box_bytes) * nch;
      int b = 0;
      uint32_t I = 0, O0 = 0;
      uint8_t* st = stage0;
      SegWalk sw(p, cta, ncta, img_off);
      const uint32_t full0 = PAIR ? mapa_u32(smem_u32(full_bar), 0) : smem_u32(full_bar);  // (PAIR: the LEADER's barriers)
      while (sw.next(p)) {
        const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
        for (int r = r0; r <= r1; ++r, ++I) {
          if (I >= static_cast<uint32_t>(D)) mbar_wait(&blk_full[buf_bar[b]], buf_par[b]);
          ESRP_FINE_TRACE(trace_ev(p, 0, tn);)  // buffer free, TMA issued next
          const uint32_t Oc = O0 + (r - r0);  // the output row this input row completes
          buf_bar[b] = pos(Oc);
          buf_par[b] = use(Oc);
          if (p.dbg & ESRP_DBG_NO_TMA) {
            mbar_arrive(&full_bar[b]);
            st += row_bytes;
            if (++b == D) { b = 0; st = stage0; }
            continue;
          }
          if constexpr (PAIR) {
            // both CTAs' tiles are counted on the leader's barrier (which expects twice the bytes); the peer only loads
            if (p.chunk_bars) {
              for (int c = 0; c < nch; ++c) {
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[b * nch + c], 2u * static_cast<uint32_t>(p.a_box_bytes));
                tma_load_4d_hint_2sm(st + c * a_bytes, p.chunk_src[c] ? &tm1 : &tm0, full0 + 8u * (b * nch + c), p.chunk_c0[c],
                                     sw.x0 - 1, r, sw.img, kL2EvictLast);
              }
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[b], 2u * tx_bytes);
              for (int c = 0; c < nch; ++c)
                tma_load_4d_hint_2sm(st + c * a_bytes, p.chunk_src[c] ? &tm1 : &tm0, full0 + 8u * b, p.chunk_c0[c], sw.x0 - 1, r,
                                     sw.img, kL2EvictLast);
            }
          } else if (p.chunk_bars) {  // one barrier per chunk tile: full_bar[b * nch + c]
            for (int c = 0; c < nch; ++c) {
              mbar_arrive_expect_tx(&full_bar[b * nch + c], static_cast<uint32_t>(p.a_box_bytes));
              tma_load_4d_hint(st + c * a_bytes, p.chunk_src[c] ? &tm1 : &tm0, &full_bar[b * nch + c], p.chunk_c0[c], sw.x0 - 1, r,
                               sw.img, kL2EvictLast);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[b], tx_bytes);
            for (int c = 0; c < nch; ++c)
              tma_load_4d_hint(st + c * a_bytes, p.chunk_src[c] ? &tm1 : &tm0, &full_bar[b], p.chunk_c0[c], sw.x0 - 1, r,
                               sw.img, kL2EvictLast);  // bf16 activations are re-read by the next convs: keep in L2
          }
          st += row_bytes;
          if (++b == D) { b = 0; st = stage0; }
        }
        O0 += static_cast<uint32_t>(r1 - r0 + 3);
      }
      trace_ev(p, 0, tn);
    }
  } else if (PAIR && rank != 0 && warp > kRowEpiWarps) {
    // peer CTA of a pair: nothing to issue; one warp tells the leader when this CTA's half of B has landed
    if (warp == kRowEpiWarps + 1) {
      mbar_wait(wfull, 0);
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(wfull), 0));
    }
  } else if (warp > kRowEpiWarps) {
    // ====================================== MMA issuers ======================================
    const int mw = warp - (kRowEpiWarps + 1);
    mbar_wait(wfull, 0);
    const int nch = p.num_chunks, naux = p.aux_chunks;
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(stage0));
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(w_res));
    const uint32_t row_step = static_cast<uint32_t>(row_bytes) >> 4, chunk_step = static_cast<uint32_t>(a_bytes) >> 4;
    constexpr uint32_t w_step = static_cast<uint32_t>(w_chunk_bytes) >> 4;
    constexpr uint32_t w_block_desc = static_cast<uint32_t>(w_block_bytes) >> 4;
    const uint32_t idesc_aux = PAIR ? umma_idesc_bf16_m256(BN) : umma_idesc_bf16_m128(BN);
    // commit / plain arrival on a block barrier (PAIR: of both CTAs)
    auto commit = [](uint64_t* bar) { if constexpr (PAIR) umma_commit_2sm(bar); else umma_commit(bar); };
    auto arrive_blk = [](uint64_t* bar) {
      mbar_arrive(bar);
      if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 1));
    };
    uint32_t tn = 0;
    if (mw == 0 && lane == 0) trace_ev(p, 1, tn);
    int b = 0;
    uint32_t fph = 0, tph = 0, O0 = 0, I = 0;
    const bool alt = p.row_alt != 0;
    const int nfull = p.last_half ? nch - 1 : nch;  // chunks issued over all of their K-slices
    uint32_t a_lo = a_lo0;
    if (PAIR && p.pair_single) {
      // Experiment (ESRP_PAIR_SINGLE=1): ONE issuer thread for every row and ONE multicast commit per row (a commit tracks
      // all earlier MMAs of its thread, so the commit of a row's "final" block covers the rows before it): block barriers
      // count 1, the second issuer warp does nothing.  With cta_group::2 the issue is asynchronous (the pipe queues whole
      // rows), so a single thread does not leave the bubbles it leaves on single CTAs.
      if (mw == 0) {
        SegWalk sw1(p, cta, ncta);
        while (sw1.next(p)) {
          const int ni = min(sw1.yb, p.h - 1) - max(sw1.ya - 1, 0) + 1;
          for (int k = 0; k < ni; ++k) {
            mbar_wait(&full_bar[b], fph);
            const uint32_t On = O0 + k + 2;
            if (k == 0) {
              mbar_wait(&blk_empty[pos(O0)], use(O0));
              mbar_wait(&blk_empty[pos(O0 + 1)], use(O0 + 1));
            }
            mbar_wait(&blk_empty[pos(On)], use(On));
            tcgen05_fence_after();
            const uint32_t dA = tmem_base + pos(On) * BN;
            const uint32_t d_aux = tmem_base + AUX_COL0 + pos(O0 + k + 1) * BN;
            if (elect_one()) {
              uint32_t al = a_lo, bl = w_lo0;
              for (int c = 0; c < nfull; ++c, al += chunk_step, bl += w_step) {
                issue_taps<KC, BN, 0, false, KS, PAIR>(dA, dA, 0u, 0u, 0u, al, bl, DESC_HI, w_block_desc);
                issue_taps<KC, BN, 1, false, KS, PAIR>(dA, dA, 0u, 0u, 0u, al, bl, DESC_HI, w_block_desc);
              }
              if (nfull < nch) {
                issue_taps<KC, BN, 0, false, KS / 2, PAIR>(dA, dA, 0u, 0u, 0u, al, bl, DESC_HI, w_block_desc);
                issue_taps<KC, BN, 1, false, KS / 2, PAIR>(dA, dA, 0u, 0u, 0u, al, bl, DESC_HI, w_block_desc);
              }
              if (AUX) {
                al = a_lo;
                bl = w_lo0;
                for (int c = 0; c < naux; ++c, al += chunk_step, bl += w_step) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks)
                    umma_f16_ss2_2sm(d_aux, al + ((1 * RB + ks * 32) >> 4), DESC_HI,
                                     bl + w_block_desc + ((aux_row0 * RB + ks * 32) >> 4), DESC_HI, idesc_aux,
                                     (c | ks) != 0 ? 1u : 0u);
                }
              }
              commit(&blk_full[pos(O0 + k)]);
              if (k == ni - 1) {
                commit(&blk_full[pos(O0 + k + 1)]);
                commit(&blk_full[pos(O0 + k + 2)]);
              }
            }
            __syncwarp();
            a_lo += row_step;
            if (++b == D) { b = 0; fph ^= 1; a_lo = a_lo0; }
          }
          O0 += static_cast<uint32_t>(ni + 2);
        }
      }
    } else {
    SegWalk sw(p, cta, ncta);
    while (sw.next(p)) {
      const int ni = min(sw.yb, p.h - 1) - max(sw.ya - 1, 0) + 1;
      for (int k = 0; k < ni; ++k) {
        // chunk_bars: both warps meet the row's FIRST chunk barrier here (like the row barrier), the issuing thread the others
        // inside its issue loop (tests/test_row_protocol_model.py::test_chunk_barriers_are_live_and_safe).
        mbar_wait(&full_bar[p.chunk_bars ? b * nch : b], fph);
        ESRP_FINE_TRACE(if (mw == 0 && lane == 0 && (I & 1u) == 0u) trace_ev(p, 1, tn);)  // full_bar ok (own rows)
        // blocks first touched by this input row must have been read + zeroed by their previous occupant
        const uint32_t On = O0 + k + 2;  // output row r+1 (ky = 0): always new
        if (k == 0) {
          mbar_wait(&blk_empty[pos(O0)], use(O0));
          mbar_wait(&blk_empty[pos(O0 + 1)], use(O0 + 1));
        }
        mbar_wait(&blk_empty[pos(On)], use(On));
        tcgen05_fence_after();
        // accumulator = blocks pos(On), +1, +2 (SHADOW: +1 / +2 may be the shadow blocks NBLK / NBLK + 1 of positions 0 / 1);
        // without shadows it is split in two MMAs where it straddles the end of the ring
        const uint32_t P = pos(On);
        const uint32_t nA = (SHADOW || P + 3 <= NBLK) ? 3u * BN : (NBLK - P) * BN;  // columns before the wrap
        const uint32_t nB = 3u * BN - nA;
        const uint32_t dA = tmem_base + P * BN, dB = tmem_base;
        const uint32_t idA = umma_idesc_bf16_m128(nA), idB = umma_idesc_bf16_m128(nB ? nB : 16u);
        const uint32_t bB = (nA * RB) >> 4;                                  // B rows of the second part
        const uint32_t d_aux = tmem_base + AUX_COL0 + pos(O0 + k + 1) * BN;  // conv1x1 of output row r
        // The issuers take strict turns (warp 0, warp 1, warp 0, ...): left alone they fall into lock-step
        // (both wait, both issue interleaved, both commit) and their per-row overhead is exposed; in turns,
        // the waits / commits of one warp overlap the MMAs of the other.
        if (alt) {
          // Row-alternating issue: warp (I & 1) issues ALL taps of input row I, then passes the turn: one hand-over
          // per row instead of two, and a whole row of MMAs hides the other warp's waits / commits.  A block still
          // collects two commits: one from the issuer of the row that completes it ("final", after row r+1) and one
          // from the issuer of its middle row r (commits only track the committing thread's MMAs; rows r-1 and r+1
          // belong to the same warp).  Segment ends: the missing contributor's commit is issued by the same thread.
          if ((I & 1u) == static_cast<uint32_t>(mw)) {
            ESRP_FINE_TRACE(if (mw == 0 && lane == 0) trace_ev(p, 1, tn);)  // waits done
            mbar_wait(&tok[mw], tph);
            ESRP_FINE_TRACE(if (mw == 0 && lane == 0) trace_ev(p, 1, tn);)  // turn taken
            if (elect_one()) {
              uint32_t al = a_lo, bl = w_lo0;
              for (int c = 0; c < nfull; ++c, al += chunk_step, bl += w_step) {
                if (p.chunk_bars && c > 0) { mbar_wait(&full_bar[b * nch + c], fph); tcgen05_fence_after(); }
                if (nB == 0) {
                  issue_taps<KC, BN, 0, false, KS, PAIR>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                  issue_taps<KC, BN, 1, false, KS, PAIR>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                } else {
                  issue_taps<KC, BN, 0, true>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                  issue_taps<KC, BN, 1, true>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                }
              }
              if (p.chunk_bars && nfull < nch && nfull > 0) { mbar_wait(&full_bar[b * nch + nfull], fph); tcgen05_fence_after(); }
              if (nfull < nch) {  // last chunk: only its first half carries weights (K = 96 / 160 in 64-channel chunks)
                if (nB == 0) {
                  issue_taps<KC, BN, 0, false, KS / 2, PAIR>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                  issue_taps<KC, BN, 1, false, KS / 2, PAIR>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                } else {
                  issue_taps<KC, BN, 0, true, KS / 2>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                  issue_taps<KC, BN, 1, true, KS / 2>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                }
              }
              if (AUX) {
                al = a_lo;
                bl = w_lo0;
                for (int c = 0; c < naux; ++c, al += chunk_step, bl += w_step) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks)
                    if constexpr (PAIR)
                      umma_f16_ss2_2sm(d_aux, al + ((1 * RB + ks * 32) >> 4), DESC_HI,
                                       bl + w_block_desc + ((aux_row0 * RB + ks * 32) >> 4), DESC_HI, idesc_aux,
                                       (c | ks) != 0 ? 1u : 0u);
                    else
                    umma_f16_ss2(d_aux, al + ((1 * RB + ks * 32) >> 4), DESC_HI,
                                 bl + w_block_desc + ((aux_row0 * RB + ks * 32) >> 4), DESC_HI, idesc_aux,
                                 (c | ks) != 0 ? 1u : 0u);
                }
              }
              ESRP_FINE_TRACE(if (mw == 0) trace_ev(p, 1, tn);)  // MMAs issued
              mbar_arrive(&tok[mw ^ 1]);                // the other warp's turn
              commit(&blk_full[pos(O0 + k)]);      // final contributor of output row r-1
              commit(&blk_full[pos(O0 + k + 1)]);  // middle contributor of output row r
              if (k == 0) commit(&blk_full[pos(O0)]);  // (dummy) first block: no earlier row
              if (k == ni - 1) {                        // last input row of the segment: no later row
                commit(&blk_full[pos(O0 + k + 1)]);
                commit(&blk_full[pos(O0 + k + 2)]);
                commit(&blk_full[pos(O0 + k + 2)]);
              }
            }
            __syncwarp();
            ESRP_FINE_TRACE(if (mw == 0 && lane == 0) trace_ev(p, 1, tn);)  // committed
            tph ^= 1;
          } else if (p.row_alt == 2) {
            // The idle warp has now observed every barrier of row I as well.  Its arrival keeps the block (and, through
            // the producer's wait on it, the row buffer) from being recycled before that: without it a warp that fell
            // a full ring behind could miss a phase of full_bar / blk_empty and wait for ever.
            // The last output row of a segment (a real row at the bottom of an image) has only two contributors: row
            // k-1, issued by THIS warp, and row k.  No later row of this warp commits on its block, so the arrival is a
            // commit here: it tracks this thread's MMAs of row k-1 (found by tests/test_row_protocol_model.py).
            if (elect_one()) {
              arrive_blk(&blk_full[pos(O0 + k)]);
              if (k == ni - 1) {
                commit(&blk_full[pos(O0 + k + 1)]);
                arrive_blk(&blk_full[pos(O0 + k + 2)]);
              }
            }
            __syncwarp();
          }
          ++I;
          a_lo += row_step;
          if (++b == D) { b = 0; fph ^= 1; a_lo = a_lo0; }
          continue;
        }
        mbar_wait(&tok[mw], tph);
        if (elect_one()) {
          uint32_t al = a_lo, bl = w_lo0;
          if (nB == 0) {
            if (mw == 0) {
              for (int c = 0; c < nch; ++c, al += chunk_step, bl += w_step)
                issue_taps<KC, BN, 0, false>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
            } else {
              for (int c = 0; c < nch; ++c, al += chunk_step, bl += w_step)
                issue_taps<KC, BN, 1, false>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
            }
          } else {
            if (mw == 0) {
              for (int c = 0; c < nch; ++c, al += chunk_step, bl += w_step)
                issue_taps<KC, BN, 0, true>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
            } else {
              for (int c = 0; c < nch; ++c, al += chunk_step, bl += w_step)
                issue_taps<KC, BN, 1, true>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
            }
          }
          if (AUX && mw == 1) {
            al = a_lo;
            bl = w_lo0;
            for (int c = 0; c < naux; ++c, al += chunk_step, bl += w_step) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks)
                umma_f16_ss2(d_aux, al + ((1 * RB + ks * 32) >> 4), DESC_HI,
                             bl + w_block_desc + ((aux_row0 * RB + ks * 32) >> 4), DESC_HI, idesc_aux,
                             (c | ks) != 0 ? 1u : 0u);
            }
          }
          mbar_arrive(&tok[mw ^ 1]);             // the other warp's turn
          commit(&blk_full[pos(O0 + k)]);  // output row r-1 has all its contributions (from this warp)
          if (k == ni - 1) {                    // last input row of the segment completes the other two as well
            commit(&blk_full[pos(O0 + k + 1)]);
            commit(&blk_full[pos(O0 + k + 2)]);
          }
        }
        __syncwarp();
        tph ^= 1;
        a_lo += row_step;
        if (++b == D) { b = 0; fph ^= 1; a_lo = a_lo0; }
      }
      O0 += static_cast<uint32_t>(ni + 2);
    }
    }
    if (mw == 0 && lane == 0) trace_ev(p, 1, tn);
  } else {
    // ======================================= epilogue =======================================
    // One warp per scheduler cannot hide the latency of this instruction stream on its own, so three
    // warpgroups take the output rows round-robin.  Thread == pixel: it reads the finished block of its
    // output row, zeroes it, hands it back, then applies the fused tail and stores all BN channels.
    const int wg = warp >> 2;
    const int q = warp & 3;                                     // TMEM lane quarter
    const int xl = q * 32 + lane;                               // column within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t O0 = 0, tn = 0;
    int turn = 0;
    if (threadIdx.x == 0) trace_ev(p, 2, tn);
    grid_dep_wait();  // residual reads / output writes must not race with the previous kernel
    SegWalk sw(p, cta, ncta, img_off);
    // "block read + zeroed": the issuers (of the leader) wait for it
    const uint32_t empty0 = (PAIR && rank != 0) ? mapa_u32(smem_u32(blk_empty), 0) : 0u;
    auto release_blk = [&](uint32_t P_) {
      if (PAIR && rank != 0) mbar_arrive_cluster(empty0 + 8u * P_); else mbar_arrive(&blk_empty[P_]);
    };
    while (sw.next(p)) {
      const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
      const int xs = sw.x0 + xl;
      const bool col_ok = xs < p.w;
      const int no = r1 - r0 + 3;  // output sequence of this segment: rows r0-1 .. r1+1 (the ends are dummies)
#pragma unroll 1
      for (int j = 0; j < no; ++j) {
        const bool mine = turn == wg;
        if (++turn == kRowWGs) turn = 0;
        if (!mine) continue;
        const uint32_t O = O0 + j;
        const int y = r0 - 1 + j;
        const bool real = y >= sw.ya && y < sw.yb && !(p.dbg & ESRP_DBG_NO_EPI);
        const bool store = real && col_ok;
        const uint32_t P = pos(O);
        const uint32_t blk = lane_addr + P * BN;
        const bool sh = SHADOW && P < 2;                   // part of this row's sum lives in the shadow block of its position
        const uint32_t blk2 = lane_addr + (NBLK + P) * BN;
        const size_t rowid = static_cast<size_t>(sw.img) * p.h + (real ? y : sw.ya);
        const size_t pix = rowid * p.w + (col_ok ? xs : 0);
        // Element offset of channel c (multiple of 4) of this pixel in an fp32 operand of `ct` channels.  NHWC, or
        // with f32_planar the engine-private [n][h][c/4][w][4] layout: consecutive lanes (pixels) then touch
        // consecutive 16 bytes, one instruction covers 4 cache lines instead of 32 (the thread == pixel mapping
        // makes NHWC fp32 accesses cost one LSU wavefront per lane, which bounded conv5).
        const bool planar = p.f32_planar != 0;
        const size_t f4_step = planar ? static_cast<size_t>(p.w) : 1;
        auto off32 = [&](int ct, int c) -> size_t {
          return planar ? ((rowid * (ct >> 2) + (c >> 2)) * p.w + (col_ok ? xs : 0)) * 4 : pix * ct + c;
        };
        auto off_res = [&](int is_f32, int ct, int c) -> size_t { return is_f32 ? off32(ct, c) : pix * ct + c; };
        // residuals of the first round: in flight while we wait for the accumulator
        float r1v[GC], r2v[GC];
        if (real) {  // every lane (the shuffles of the bf16 store need the whole warp): pix is clamped for columns >= w
          if (p.r1) load_residual<GC>(p.r1, p.r1_is_f32, off_res(p.r1_is_f32, p.r1_ctotal, p.r1_c0 + csh), r1v, f4_step);
          if (p.r2) load_residual<GC>(p.r2, p.r2_is_f32, off_res(p.r2_is_f32, p.r2_ctotal, p.r2_c0 + csh), r2v, f4_step);
        }
        ESRP_FINE_TRACE(if (threadIdx.x == 0) trace_ev(p, 2, tn);)  // row start
        mbar_wait(&blk_full[pos(O)], use(O));
        ESRP_FINE_TRACE(if (threadIdx.x == 0) trace_ev(p, 2, tn);)  // blk_full ok
        tcgen05_fence_after();
        if (!real) {  // dummy row at a segment end: just recycle the block
#pragma unroll
          for (int c = 0; c < BN; c += GC) {
            if constexpr (GC == 16) tmem_st_zero_x16(blk + c); else tmem_st_zero_x8(blk + c);
            if (sh) { if constexpr (GC == 16) tmem_st_zero_x16(blk2 + c); else tmem_st_zero_x8(blk2 + c); }
          }
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) release_blk(P);
          continue;
        }
        uint32_t pend[8];  // bf16 output of an even round, stored together with the following odd round
        const bool quad = (ROUNDS % 2 == 0) && (p.cout % (2 * GC) == 0) && !p.no_quad;
        // GC channels per round: load, zero, (after the last round: hand the block back), fused tail, store.
        // The ring is 8-16 blocks deep, so releasing after the last load costs nothing.
#pragma unroll
        for (int g = 0; g < ROUNDS; ++g) {
          const int ch0 = g * GC;
          uint32_t acc[GC], ax[GC];
          tmem_ld_cols<GC>(blk + ch0, acc);
          if (AUX) tmem_ld_cols<GC>(blk + AUX_COL0 + ch0, ax);
          if (sh) {  // (warp-uniform) 2 of every NBLK rows: add the shadow block's part of the sum
            uint32_t a2[GC];
            tmem_ld_cols<GC>(blk2 + ch0, a2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < GC; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(a2[i]));
            if constexpr (GC == 16) tmem_st_zero_x16(blk2 + ch0); else tmem_st_zero_x8(blk2 + ch0);
          } else {
            tmem_ld_wait();
          }
          if constexpr (GC == 16) tmem_st_zero_x16(blk + ch0); else tmem_st_zero_x8(blk + ch0);
          if (g == ROUNDS - 1) {
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) release_blk(P);
            ESRP_FINE_TRACE(if (threadIdx.x == 0) trace_ev(p, 2, tn);)  // block released
          }
          if (ch0 >= p.cout) continue;  // (warp-uniform; lanes of columns >= w compute along and store nothing)
          const int gch = ch0 + csh;  // channel relative to the *_c0 offsets of the descriptor
          float v[GC];
          const float4* bias4 = reinterpret_cast<const float4*>(bias_s + ch0);
#pragma unroll
          for (int i = 0; i < GC / 4; ++i) {
            const float4 b4 = bias4[i];
            v[4 * i] = __uint_as_float(acc[4 * i]) + b4.x;
            v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + b4.y;
            v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + b4.z;
            v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + b4.w;
          }
          if constexpr (EXT) {
            if (store) ext_mask_store<GC>(p, pix, gch, v);
          }
          if (p.act) {
            const float slope = p.act == 2 ? 0.f : 0.2f;  // LeakyReLU(0.2) / ReLU
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaxf(v[i], slope * v[i]);
          }
          if (p.s0 != 1.0f) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] *= p.s0;
          }
          if (AUX) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] += __uint_as_float(ax[i]);
          }
          if (p.r1) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s1, r1v[i], v[i]);
            if (g + 1 < ROUNDS) load_residual<GC>(p.r1, p.r1_is_f32, off_res(p.r1_is_f32, p.r1_ctotal, p.r1_c0 + gch + GC), r1v, f4_step);
          }
          if constexpr (EXT) {
            if (p.r2 && p.r2_pre) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] += r2v[i];
              if (g + 1 < ROUNDS) load_residual<GC>(p.r2, p.r2_is_f32, off_res(p.r2_is_f32, p.r2_ctotal, p.r2_c0 + gch + GC), r2v, f4_step);
            }
            if (store) ext_pre_and_mask<GC>(p, pix, gch, v);
          }
          if (p.noise) {
            const unsigned long long nseed = p.seed_ptr ? __ldg(p.seed_ptr) : p.seed;  // graph replays read the key from memory
#pragma unroll 1
            for (int i = 0; i < GC; i += 4) {
              float z[4];
              philox_normal4(nseed,
                             p.offset + (pix * static_cast<unsigned long long>(p.noise_ctotal) + p.noise_c0 + gch + i) / 4, z);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                for (int kk = 0; kk < GC; kk += 4)  // static register indexing
                  if (kk == i) v[kk + jj] = fmaf(z[jj] * p.sigma, v[kk + jj], v[kk + jj]);
              }
            }
          }
          if (EXT && p.r2_pre) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] *= p.s2;
          } else if (p.r2) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s2, v[i], r2v[i]);
            if (g + 1 < ROUNDS) load_residual<GC>(p.r2, p.r2_is_f32, off_res(p.r2_is_f32, p.r2_ctotal, p.r2_c0 + gch + GC), r2v, f4_step);
          }
          if (p.out_bf16) {
            uint32_t pk[GC / 2];
#pragma unroll
            for (int i = 0; i < GC / 2; ++i) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            if (!quad) {
              if (store) {
                uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.ob_ctotal + p.ob_c0 + gch);
#pragma unroll
                for (int i = 0; i < GC / 8; ++i)  // next conv's operand: keep in L2
                  st_global_u4_hint(op + i, make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]), kL2EvictLast);
              }
            } else if ((g & 1) == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pend[i] = pk[i];
            } else {
              // 32 channels = four 16-byte chunks per pixel.  Stored by their owner, one instruction touches 32 cache
              // lines (pixel stride = ob_ctotal * 2 bytes); after a 4x4 transpose inside each group of four lanes,
              // lane j holds chunk j of the group's four pixels and an instruction writes 8 x 64 contiguous bytes.
              const bool hi = (lane & 2) != 0, lo = (lane & 1) != 0;
              uint32_t d0[8], d1[8];  // [pixel hi bit][chunk lo bit][4 words] after the exchange with lane ^ 2
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t snd = hi ? pend[i] : pk[i], kp = hi ? pk[i] : pend[i];
                const uint32_t rc = __shfl_xor_sync(0xffffffffu, snd, 2);
                d0[i] = hi ? rc : kp;
                d1[i] = hi ? kp : rc;
              }
              uint32_t e[4][4];       // [pixel of the group][4 words] after the exchange with lane ^ 1
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t s0_ = lo ? d0[i] : d0[4 + i], k0_ = lo ? d0[4 + i] : d0[i];
                const uint32_t s1_ = lo ? d1[i] : d1[4 + i], k1_ = lo ? d1[4 + i] : d1[i];
                const uint32_t r0_ = __shfl_xor_sync(0xffffffffu, s0_, 1), r1_ = __shfl_xor_sync(0xffffffffu, s1_, 1);
                e[0][i] = lo ? r0_ : k0_;
                e[1][i] = lo ? k0_ : r0_;
                e[2][i] = lo ? r1_ : k1_;
                e[3][i] = lo ? k1_ : r1_;
              }
              const int xg = sw.x0 + q * 32 + (lane & ~3);  // first pixel of this lane's group
              __nv_bfloat16* ob = p.out_bf16 + (rowid * p.w + xg) * p.ob_ctotal + p.ob_c0 + (gch - GC) + (lane & 3) * 8;
#pragma unroll
              for (int m = 0; m < 4; ++m)
                if (xg + m < p.w)
                  st_global_u4_hint(reinterpret_cast<uint4*>(ob + static_cast<size_t>(m) * p.ob_ctotal),
                                    make_uint4(e[m][0], e[m][1], e[m][2], e[m][3]), kL2EvictLast);
            }
          }
          if (p.out_f32 && store) {
            float4* op = reinterpret_cast<float4*>(p.out_f32 + off32(p.of_ctotal, p.of_c0 + gch));
#pragma unroll
            for (int i = 0; i < GC / 4; ++i)  // fp32 trunk: read once, a whole dense block later -> stream
              st_global_f4_hint(op + i * f4_step, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), kL2EvictFirst);
          }
          if (p.out_nchw && store) {
            const size_t plane = static_cast<size_t>(p.h) * p.w;
            float* op = p.out_nchw + (static_cast<size_t>(sw.img) * p.cout) * plane + static_cast<size_t>(y) * p.w + xs;
#pragma unroll
            for (int i = 0; i < GC; ++i)
              if (ch0 + i < p.cout) op[static_cast<size_t>(ch0 + i) * plane] = v[i];
          }
        }
      }
      O0 += static_cast<uint32_t>(no);
    }
    if (threadIdx.x == 0) trace_ev(p, 2, tn);
  }

  tcgen05_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();  // (PAIR: no CTA leaves while its peer may still reach into it)
  if (warp == kRowEpiWarps + 1) {
    tcgen05_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
  ESRP_FINE_TRACE(if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) { p.trace[2 * 1024 + 1021] = clock64(); long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[2 * 1024 + 1020] = gt; })
  ESRP_FINE_TRACE(if (p.trace && threadIdx.x == 0 && blockIdx.x < 512) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); p.trace[1024 + 512 + blockIdx.x] = gt; })
}

}  // namespace esrp
