// conv3x3_row.cuh — row-streaming fused 3x3 convolution for sm_100a (images at least ~64 px wide).
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
//   Output  : out[y] = Q_{y-1}[ky=0] + Q_y[ky=1] + Q_{y+1}[ky=2] — three accumulators that live in
//             DIFFERENT tensor-memory slots at the SAME lane, so the epilogue is three tcgen05.ld and
//             two adds per value: no shuffles, no shared-memory exchange.
//   Schedule: CTA b owns the contiguous output rows [b*U/G, (b+1)*U/G) (U = n * column blocks * H) and
//             streams input rows ya-1 .. yb through a ring of TMEM slots (5 x 96 or 4 x 128 columns);
//             the two halo rows per CTA are the only recomputation (~14 % at 14 rows per CTA).
//   Stages  : ring of row buffers (one image row x all K-chunks, 16.6 KB per 64-channel chunk); every
//             dense-block conv keeps its weights resident in shared memory.  One tcgen05.commit per row.
//   Warps   : 0-11 epilogue = 3 warpgroups taking output rows round-robin (thread == pixel; TMEM lane
//             quarter == warp % 4), 12 TMA producer, 13-14 MMA issuers (input rows round-robin).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"
#include "conv3x3_tc.cuh"  // lrelu02, load_residual, trace_ev, kMaxStages, kSmemFixed
#include "conv_params.h"
#include "esrp_philox.cuh"
#include "esrp_ptx.cuh"

namespace esrp {

constexpr int kRowWGs = 3;                     // epilogue warpgroups (rows in flight)
constexpr int kRowEpiWarps = 4 * kRowWGs;
constexpr int kRowMmaWarps = 2;                // MMA issuer warps (input rows round-robin)
constexpr int kRowThreads = 32 * (kRowEpiWarps + 1 + kRowMmaWarps);
constexpr int kRowTile = 128;  // output columns per M-tile
constexpr int kMaxSlots = 8;

// Walks the row segments of this CTA: identical in the three warp roles.
struct SegWalk {
  int u, u_end;
  int img, x0, ya, yb;  // current segment: output rows [ya, yb) of column block x0 of image img
  __device__ __forceinline__ explicit SegWalk(const ConvKParams& p) {
    const long long U = p.units_total;
    u = static_cast<int>(U * blockIdx.x / gridDim.x);
    u_end = static_cast<int>(U * (blockIdx.x + 1) / gridDim.x);
    img = x0 = ya = yb = 0;
  }
  __device__ __forceinline__ bool next(const ConvKParams& p) {
    if (u >= u_end) return false;
    const int col = u / p.h;
    ya = u - col * p.h;
    const int cnt = min(u_end - u, p.h - ya);
    yb = ya + cnt;
    img = col / p.x_tiles;
    x0 = (col - img * p.x_tiles) * kRowTile;
    u += cnt;
    return true;
  }
};

template <int GCH>
__device__ __forceinline__ void tmem_ld_half(uint32_t taddr, uint32_t (&v)[GCH]) {
  if constexpr (GCH == 32) {
    tmem_ld_x32(taddr, v);
  } else if constexpr (GCH == 16) {
    tmem_ld_x16(taddr, v);
  } else {
    tmem_ld_x8(taddr, v);
  }
}

template <int KC, int BN, bool AUX>
__global__ void __launch_bounds__(kRowThreads, 1)
conv3x3_row_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                   const __grid_constant__ ConvKParams p) {
  constexpr int RB = KC * 2;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * RB;
  constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
  constexpr int GC = BN < 16 ? BN : 16;  // output channels per epilogue round
  constexpr int ROUNDS = BN / GC;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages] one per row buffer
  uint64_t* q_full = full_bar + kMaxStages;                 // [kMaxSlots]
  uint64_t* q_empty = q_full + kMaxSlots;                   // [kMaxSlots]
  uint64_t* wfull = q_empty + kMaxSlots;                    // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wfull + 1);
  float* bias_s = reinterpret_cast<float*>(smem + 1024);    // [BN]

  constexpr int nb_rows = AUX ? 4 * BN : 3 * BN;
  constexpr int w_block_bytes = nb_rows * RB;
  constexpr int w_chunk_bytes = 3 * w_block_bytes;
  const int w_res_bytes = p.num_chunks * w_chunk_bytes;     // weights are always resident
  uint8_t* w_res = smem + kSmemFixed;
  uint8_t* stage0 = w_res + w_res_bytes;
  const int a_bytes = p.a_stage_bytes;                      // one chunk tile (TMA box rounded up to 1 KB)
  const int row_bytes = a_bytes * p.num_chunks;             // one row buffer
  const int D = p.stages;                                   // row buffers (< NS)
  const int NS = p.mt;  // TMEM slots
  const int NT = p.nt;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kRowEpiWarps && lane == 0) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < D; ++i) mbar_init(&full_bar[i], 1);
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&q_full[i], 1);
      // a slot is recycled after 3 consuming output rows x 4 warps have read it (weighted at segment ends)
      // AND all 12 epilogue warps have observed its q_full phase (so none of them can be lapped)
      mbar_init(&q_empty[i], 12 + kRowEpiWarps);
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == kRowEpiWarps + 1) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x < BN) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (p.dbg & ESRP_DBG_EMPTY) {
    // timing experiment: prologue + teardown only
  } else if (warp == kRowEpiWarps) {
    // ===================================== TMA producer =====================================
    // Row-buffer ring: buffer b = row % D holds the num_chunks K-chunk tiles of one input row.  A
    // buffer is reused once the row that last occupied it has been fully multiplied, which the MMA
    // warps signal on q_full (the same commit that wakes the epilogue): one commit per row in total.
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(w_res_bytes));
      for (int c = 0; c < p.num_chunks; ++c)
        bulk_load_1d(w_res + c * w_chunk_bytes, p.w_packed + static_cast<size_t>(c) * w_chunk_bytes,
                     w_chunk_bytes, wfull);
      uint32_t tn = 0;
      trace_ev(p, 0, tn);
      const int nch = p.num_chunks;
      const uint32_t tx_bytes = static_cast<uint32_t>(p.a_box_bytes) * nch;
      int b = 0;                 // row buffer of the row being loaded
      uint32_t ri = 0;           // index of the row being loaded
      int ds = 0;                // q_full slot of row ri - D (the row whose completion frees buffer b)
      uint32_t dph = 0;
      uint8_t* st = stage0;
      SegWalk sw(p);
      while (sw.next(p)) {
        const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
        for (int r = r0; r <= r1; ++r, ++ri) {
          if (ri >= static_cast<uint32_t>(D)) {
            mbar_wait(&q_full[ds], dph);
            if (++ds == NS) { ds = 0; dph ^= 1; }
          }
          if (p.dbg & ESRP_DBG_NO_TMA) {
            mbar_arrive(&full_bar[b]);
          } else {
            mbar_arrive_expect_tx(&full_bar[b], tx_bytes);
            for (int c = 0; c < nch; ++c)
              tma_load_4d(st + c * a_bytes, p.chunk_src[c] ? &tm1 : &tm0, &full_bar[b], p.chunk_c0[c], sw.x0 - 1, r,
                          sw.img);
          }
          st += row_bytes;
          if (++b == D) { b = 0; st = stage0; }
        }
      }
      trace_ev(p, 0, tn);
    }
  } else if (warp > kRowEpiWarps) {
    // ====================================== MMA issuers ======================================
    // kRowMmaWarps warps take the input rows round-robin.  tcgen05.mma issue is nearly synchronous
    // (the tensor pipe queues only ~1-2 instructions), so with a single issuer every barrier wait /
    // commit between two rows is a bubble in the pipe (measured: 85-99 cycles per N=96 MMA with one
    // issuer, 57-61 with two; tools/ubench_row.cu).  Each warp stays converged and one elected lane
    // issues, which keeps the MMA sequence on the uniform datapath.
    const int mw = warp - (kRowEpiWarps + 1);
    mbar_wait(wfull, 0);
    const uint32_t idesc_main = umma_idesc_bf16_m128(3 * BN);
    const uint32_t idesc_aux = umma_idesc_bf16_m128(4 * BN);
    const int nch = p.num_chunks, naux = p.aux_chunks;
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(stage0));
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(w_res));
    const uint32_t row_step = static_cast<uint32_t>(row_bytes) >> 4, chunk_step = static_cast<uint32_t>(a_bytes) >> 4;
    constexpr uint32_t w_step = static_cast<uint32_t>(w_chunk_bytes) >> 4;
    uint32_t tn = 0;
    if (mw == 0 && lane == 0) trace_ev(p, 1, tn);
    int b = 0, slot = 0, turn = 0;
    uint32_t fph = 0, qph = 1;            // parities to wait on full[b] / q_empty[slot]
    uint32_t a_lo = a_lo0, d_tmem = tmem_base;
    SegWalk sw(p);
    while (sw.next(p)) {
      const int nrows = min(sw.yb, p.h - 1) - max(sw.ya - 1, 0) + 1;
      for (int r = 0; r < nrows; ++r) {
        // The ring lengths D (row buffers) and NS (TMEM slots) are multiples of kRowMmaWarps, so a buffer /
        // slot only ever serves ONE issuer warp: each warp waits on its own barriers only and sees every
        // phase of them in order.
        if (turn == mw) {
          mbar_wait(&full_bar[b], fph);
          mbar_wait(&q_empty[slot], qph);
        }
        if (turn == mw) {
          tcgen05_fence_after();
          if (elect_one()) {
            if (p.dbg & ESRP_DBG_NO_MMA) {
              mbar_arrive(&q_full[slot]);
            } else {
              uint32_t al = a_lo, bl = w_lo0;
              for (int c = 0; c < nch; ++c) {
                const bool aux_c = c < naux;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                  const int kx = kk == 0 ? 1 : (kk == 1 ? 0 : 2);  // centre column first: it owns the conv1x1 columns
                  const uint32_t a_off = kx * RB;                  // A = the 130-pixel row shifted by kx pixels
                  const uint32_t b_off = kx * w_block_bytes;
                  const uint32_t idesc = (aux_c && kx == 1) ? idesc_aux : idesc_main;
#pragma unroll
                  for (int ks = 0; ks < KC / 16; ++ks) {
                    umma_f16_ss2(d_tmem, al + ((a_off + ks * 32) >> 4), DESC_HI, bl + ((b_off + ks * 32) >> 4),
                                 DESC_HI, idesc, (c | kk | ks) != 0 ? 1u : 0u);
                  }
                }
                al += chunk_step;
                bl += w_step;
              }
              umma_commit(&q_full[slot]);  // row complete: wakes the epilogue and frees the row buffer
            }
          }
          __syncwarp();
        }
        if (++turn == kRowMmaWarps) turn = 0;
        a_lo += row_step;
        if (++b == D) { b = 0; fph ^= 1; a_lo = a_lo0; }
        d_tmem += NT;
        if (++slot == NS) { slot = 0; qph ^= 1; d_tmem = tmem_base; }
      }
    }
    if (mw == 0 && lane == 0) trace_ev(p, 1, tn);
  } else {
    // ======================================= epilogue =======================================
    // kRowWGs warpgroups take the output rows of a segment round-robin, so several rows are in flight
    // (one warp per scheduler cannot hide the latency of this instruction stream on its own).
    // Thread == pixel: it folds the three partial rows, applies the fused tail and stores all BN
    // channels of its pixel, GC channels per round.
    const int wg = warp >> 2;                                   // which rows
    const int q = warp & 3;                                     // TMEM lane quarter
    const int xl = q * 32 + lane;                               // column within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t ri_base = 0, seen = 0, tn = 0;
    if (threadIdx.x == 0) trace_ev(p, 2, tn);
    SegWalk sw(p);
    while (sw.next(p)) {
      const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
      const int xs = sw.x0 + xl;
      const bool col_ok = xs < p.w;
      int turn = 0;
#pragma unroll 1
      for (int y = sw.ya; y < sw.yb; ++y) {
        const bool has_up = y - 1 >= r0, has_dn = y + 1 <= r1;
        const uint32_t i_mid = ri_base + (y - r0);
        const uint32_t i_last = has_dn ? i_mid + 1 : i_mid;
        // Every warp walks EVERY output row and observes every q_full phase in order, also for the rows other
        // warpgroups consume: a parity wait is only correct if the waiter checks phase k after phase k-1
        // completed and before phase k+1 completes.  Each observation is acknowledged on q_empty, so a slot
        // cannot be recycled (and its q_full phase advance) before all epilogue warps have seen it.
        for (; seen <= i_last; ++seen) {
          mbar_wait(&q_full[seen % NS], (seen / NS) & 1);
          if (lane == 0) mbar_arrive(&q_empty[seen % NS]);  // "observed": see the q_empty count
        }
        const bool mine = turn == wg;
        if (++turn == kRowWGs) turn = 0;
        if (!mine) continue;
        const size_t pix = (static_cast<size_t>(sw.img) * p.h + y) * p.w + xs;
        const uint32_t a_mid = lane_addr + (i_mid % NS) * NT;
        const uint32_t a_up = lane_addr + ((i_mid - 1) % NS) * NT;
        const uint32_t a_dn = lane_addr + ((i_mid + 1) % NS) * NT;
        tcgen05_fence_after();
        if (threadIdx.x == 0) trace_ev(p, 2, tn);
        if (p.dbg & ESRP_DBG_NO_EPI) {  // timing experiment: release the slots without reading them
          __syncwarp();
          if (lane == 0) {
            const int lo = (y == sw.ya) ? 1 : 0, hi = (y == sw.yb - 1) ? 1 : 0;
            if (has_up) mbar_arrive_cnt(&q_empty[(i_mid - 1) % NS], 1 + 2 * lo);
            mbar_arrive_cnt(&q_empty[i_mid % NS], 1 + lo + hi);
            if (has_dn) mbar_arrive_cnt(&q_empty[(i_mid + 1) % NS], 1 + 2 * hi);
          }
          continue;
        }
#pragma unroll
        for (int g = 0; g < ROUNDS; ++g) {
          const int ch0 = g * GC;
          float r1v[GC], r2v[GC];
          if (p.r1 && col_ok) load_residual<GC>(p.r1, p.r1_is_f32, pix * p.r1_ctotal + p.r1_c0 + ch0, r1v);
          if (p.r2 && col_ok) load_residual<GC>(p.r2, p.r2_is_f32, pix * p.r2_ctotal + p.r2_c0 + ch0, r2v);
          uint32_t pu[GC], pm[GC], pd[GC], ax[GC];
          tmem_ld_half<GC>(a_mid + 1 * BN + ch0, pm);
          if (has_up) tmem_ld_half<GC>(a_up + 0 * BN + ch0, pu);
          if (has_dn) tmem_ld_half<GC>(a_dn + 2 * BN + ch0, pd);
          if (AUX) tmem_ld_half<GC>(a_mid + 3 * BN + ch0, ax);
          tmem_ld_wait();
          if (g == ROUNDS - 1) {
            // This output row has consumed its three accumulator rows.  Each slot is released by 3
            // outputs x 4 warps = 12 arrivals; the first / last output of a segment also arrives for
            // the neighbours that do not exist in it.  (tcgen05.wait::ld is warp-wide: lane 0 arrives.)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              const int lo = (y == sw.ya) ? 1 : 0, hi = (y == sw.yb - 1) ? 1 : 0;
              if (has_up) mbar_arrive_cnt(&q_empty[(i_mid - 1) % NS], 1 + 2 * lo);
              mbar_arrive_cnt(&q_empty[i_mid % NS], 1 + lo + hi);
              if (has_dn) mbar_arrive_cnt(&q_empty[(i_mid + 1) % NS], 1 + 2 * hi);
            }
          }
          float v[GC];
          const float4* bias4 = reinterpret_cast<const float4*>(bias_s + ch0);
          if (has_up && has_dn) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = (__uint_as_float(pm[i]) + __uint_as_float(pu[i])) + __uint_as_float(pd[i]);
          } else {
#pragma unroll
            for (int i = 0; i < GC; ++i) {
              float a = __uint_as_float(pm[i]);
              if (has_up) a += __uint_as_float(pu[i]);
              if (has_dn) a += __uint_as_float(pd[i]);
              v[i] = a;
            }
          }
#pragma unroll
          for (int i = 0; i < GC / 4; ++i) {
            const float4 b4 = bias4[i];
            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          }
          if (p.act) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaxf(v[i], 0.2f * v[i]);  // LeakyReLU(0.2)
          }
          if (p.s0 != 1.0f) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] *= p.s0;
          }
          if (AUX) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] += __uint_as_float(ax[i]);
          }
          if (p.r1 && col_ok) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s1, r1v[i], v[i]);
          }
          if (p.noise) {
#pragma unroll 1
            for (int i = 0; i < GC; i += 4) {
              float z[4];
              philox_normal4(p.seed,
                             p.offset + (pix * static_cast<unsigned long long>(p.noise_ctotal) + p.noise_c0 + ch0 + i) / 4, z);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // static register indexing: select the 4 values of this iteration
#pragma unroll
                for (int k = 0; k < GC; k += 4)
                  if (k == i) v[k + j] = fmaf(z[j] * p.sigma, v[k + j], v[k + j]);
              }
            }
          }
          if (p.r2 && col_ok) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s2, v[i], r2v[i]);
          }
          if (col_ok && ch0 < p.cout) {
            if (p.out_bf16) {
              uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.ob_ctotal + p.ob_c0 + ch0);
#pragma unroll
              for (int i = 0; i < GC / 8; ++i) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
                  pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                op[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
            if (p.out_f32) {
              float4* op = reinterpret_cast<float4*>(p.out_f32 + pix * p.of_ctotal + p.of_c0 + ch0);
#pragma unroll
              for (int i = 0; i < GC / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.out_nchw) {
              const size_t plane = static_cast<size_t>(p.h) * p.w;
              float* op = p.out_nchw + (static_cast<size_t>(sw.img) * p.cout) * plane + static_cast<size_t>(y) * p.w + xs;
#pragma unroll
              for (int i = 0; i < GC; ++i)
                if (ch0 + i < p.cout) op[static_cast<size_t>(ch0 + i) * plane] = v[i];
            }
          }
        }
        if (threadIdx.x == 0) trace_ev(p, 2, tn);
      }
      ri_base += static_cast<uint32_t>(r1 - r0 + 1);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kRowEpiWarps + 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace esrp
