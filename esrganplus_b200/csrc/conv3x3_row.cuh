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
//   Stages  : one image row x one 64-channel chunk = 16.6 KB, so the ring is deep (up to 8 stages)
//             and every dense-block conv keeps its weights resident in shared memory.
//   Warps   : 0-11 epilogue = 3 warpgroups taking output rows round-robin (thread == pixel; TMEM lane
//             quarter == warp % 4), 12 TMA producer, 13 MMA issuer / TMEM owner.
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
constexpr int kRowThreads = 32 * (kRowEpiWarps + 2);
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

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;              // [kMaxStages]
  uint64_t* q_full = empty_bar + kMaxStages;                // [kMaxSlots]
  uint64_t* q_empty = q_full + kMaxSlots;                   // [kMaxSlots]
  uint64_t* wfull = q_empty + kMaxSlots;                    // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wfull + 1);
  float* bias_s = reinterpret_cast<float*>(smem + 1024);    // [BN]

  constexpr int nb_rows = AUX ? 4 * BN : 3 * BN;
  constexpr int w_block_bytes = nb_rows * RB;
  constexpr int w_chunk_bytes = 3 * w_block_bytes;
  const int w_res_bytes = p.w_resident ? p.num_chunks * w_chunk_bytes : 0;
  uint8_t* w_res = smem + kSmemFixed;
  uint8_t* stage0 = w_res + w_res_bytes;
  const int a_bytes = p.a_stage_bytes;
  const int stage_bytes = a_bytes + (p.w_resident ? 0 : w_chunk_bytes);
  const int S = p.stages;
  const int NS = p.mt;  // TMEM slots
  const int NT = p.nt;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kRowEpiWarps && lane == 0) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 12);  // 3 consuming output rows x 4 warps (weighted at segment ends)
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

  if (warp == kRowEpiWarps) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      if (p.w_resident) {
        mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(w_res_bytes));
        for (int c = 0; c < p.num_chunks; ++c)
          bulk_load_1d(w_res + c * w_chunk_bytes, p.w_packed + static_cast<size_t>(c) * w_chunk_bytes,
                       w_chunk_bytes, wfull);
      }
      uint32_t it = 0, tn = 0;
      trace_ev(p, 0, tn);
      SegWalk sw(p);
      while (sw.next(p)) {
        const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
        for (int r = r0; r <= r1; ++r) {
          for (int c = 0; c < p.num_chunks; ++c, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            trace_ev(p, 0, tn);
            uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
            if (p.dbg & ESRP_DBG_NO_TMA) {
              mbar_arrive(&full_bar[s]);
              continue;
            }
            mbar_arrive_expect_tx(&full_bar[s],
                                  static_cast<uint32_t>(p.a_box_bytes + (p.w_resident ? 0 : w_chunk_bytes)));
            tma_load_4d(st, p.chunk_src[c] ? &tm1 : &tm0, &full_bar[s], p.chunk_c0[c],
                        sw.x0 - ((p.dbg & ESRP_DBG_NO_XHALO) ? 0 : 1), r, sw.img);
            if (!p.w_resident)
              bulk_load_1d(st + a_bytes, p.w_packed + static_cast<size_t>(c) * w_chunk_bytes, w_chunk_bytes,
                           &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == kRowEpiWarps + 1) {
    // ====================================== MMA issuer ======================================
    if (p.w_resident) mbar_wait(wfull, 0);
    const uint32_t idesc_main = umma_idesc_bf16_m128(3 * BN);
    const uint32_t idesc_aux = umma_idesc_bf16_m128(4 * BN);
    uint32_t it = 0, ri = 0, tn = 0;
    if (lane == 0) trace_ev(p, 1, tn);
    SegWalk sw(p);
    while (sw.next(p)) {
      const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
      for (int r = r0; r <= r1; ++r, ++ri) {
        const int slot = ri % NS;
        const uint32_t qph = (ri / NS) & 1;
        const uint32_t d_tmem = tmem_base + slot * NT;
        for (int c = 0; c < p.num_chunks; ++c, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full_bar[s], ph);
          if (c == 0) mbar_wait(&q_empty[slot], qph ^ 1);
          tcgen05_fence_after();
          if (lane == 0) trace_ev(p, 1, tn);
          uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
          const uint32_t a_lo = umma_desc_lo(smem_u32(st));
          const uint32_t b_lo =
              umma_desc_lo(p.w_resident ? smem_u32(w_res + c * w_chunk_bytes) : smem_u32(st + a_bytes));
          const bool aux_c = c < p.aux_chunks;
          if (p.dbg & ESRP_DBG_NO_MMA) {
            if (elect_one()) {
              mbar_arrive(&empty_bar[s]);
              if (c == p.num_chunks - 1) mbar_arrive(&q_full[slot]);
            }
            __syncwarp();
            continue;
          }
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              const int kx = kk == 0 ? 1 : (kk == 1 ? 0 : 2);  // centre column first: it owns the conv1x1 columns
              const uint32_t a_off = kx * RB;                  // A = the 130-pixel row shifted by kx pixels
              const uint32_t b_off = kx * w_block_bytes;
              const uint32_t idesc = (aux_c && kx == 1) ? idesc_aux : idesc_main;
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                umma_f16_ss2(d_tmem, a_lo + ((a_off + ks * 32) >> 4), DESC_HI, b_lo + ((b_off + ks * 32) >> 4),
                             DESC_HI, idesc, (c | kk | ks) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[s]);
            if (c == p.num_chunks - 1) umma_commit(&q_full[slot]);
          }
          __syncwarp();
          if (lane == 0) trace_ev(p, 1, tn);
        }
      }
    }
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
#pragma unroll 1
      for (int y = sw.ya + wg; y < sw.yb; y += kRowWGs) {
        const size_t pix = (static_cast<size_t>(sw.img) * p.h + y) * p.w + xs;
        const bool has_up = y - 1 >= r0, has_dn = y + 1 <= r1;
        const uint32_t i_mid = ri_base + (y - r0);
        const uint32_t i_last = has_dn ? i_mid + 1 : i_mid;
        const uint32_t a_mid = lane_addr + (i_mid % NS) * NT;
        const uint32_t a_up = lane_addr + ((i_mid - 1) % NS) * NT;
        const uint32_t a_dn = lane_addr + ((i_mid + 1) % NS) * NT;
        // Observe EVERY row's completion in order, also the rows other warpgroups consume: a parity wait
        // is only meaningful if the waiter never skips a phase of that barrier.
        for (; seen <= i_last; ++seen) mbar_wait(&q_full[seen % NS], (seen / NS) & 1);
        tcgen05_fence_after();
        if (threadIdx.x == 0) trace_ev(p, 2, tn);
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
