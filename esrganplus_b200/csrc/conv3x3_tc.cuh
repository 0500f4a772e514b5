// conv3x3_tc.cuh — fused 3x3 convolution for sm_100a: TMA-staged NHWC row tiles in shared memory
// feeding tcgen05.mma implicit-GEMM tiles with fp32 accumulators in tensor memory, and the
// dense-block elementwise tail (bias, LeakyReLU, conv1x1 add, residual scale/add, noise, RRDB
// residual) fused into the TMEM->register epilogue.
//
// Replaces (reference): block.py:260-268 (ResidualDenseBlock_5C.forward: conv_k(torch.cat(...)),
// lrelu, + conv1x1, + x2, *0.2 + x, noise), block.py:287-291 (RRDB residual) and the plain
// conv_blocks of architecture.py:55-71.  torch.cat never materialises: a conv reads its K
// dimension as a list of channel chunks taken from up to two NHWC tensors.
//
// Why the taps are stacked along N
//   Every dense-block conv but the last has Cout = 32.  A tcgen05.mma M128 x N x K16 costs
//   max(N/2, ~32 + N/4) cycles (measured, profiles/r01_ubench_mma_ss_ts.jsonl): the A tile
//   (128 px x 16 ch) takes ~32 cycles to stream however small N is, so N = 32 runs at 40 % of the
//   tensor pipe.  Instead of 9 shifted-A MMAs with N = Cout per K-slice, this kernel issues 3 MMAs
//   (one per kernel row ky) with N = 3*Cout: the three kernel columns kx are three blocks of
//   B rows, all multiplied with the SAME un-shifted A tile,
//       P[px, kx, co] = sum_{ky, ci} X[y+ky-1, x', ci] * W[co, ci, ky, kx]        (x' = source column)
//       out[y, x, co] = P[x-1, 0, co] + P[x, 1, co] + P[x+1, 2, co]
//   and the +-1 column shift is resolved in the epilogue with two warp shuffles per value (lane ==
//   source column).  N = 96 runs at 83 % of the pipe, N = 192 at 98 %.
//
// Decomposition
//   M-tile   : 128 pixels = RM rows x CW columns (CW = 128/64/32/16, RM = 128/CW), lane = ry*CW + rx.
//   CTA tile : up to `mt` vertically stacked M-tiles of one column block; shared memory holds the
//              (mt*RM + 2) x CW x KC halo-in-y tile of one K-chunk per stage (one TMA box, zero fill
//              outside the image = the conv's zero padding).  The A operand of (m, ky) is that tile
//              at pixel offset (m*RM + ky)*CW: dense, swizzle-atom aligned, no shifted descriptors.
//   Columns  : W <= CW -> one column block; else blocks advance by CW-2 source columns and lanes whose
//              neighbour lies in another block produce no output (x halo by recompute).
//   Weights  : pre-swizzled bf16 [chunk][ky][NB rows = kx*BN+co (+ BN conv1x1 rows)][KC]; resident in
//              shared memory when they fit, else streamed with the chunk.
//   Schedule : CTA b owns the contiguous range of M-tile units [b*U/G, (b+1)*U/G) and cuts it into
//              tiles of <= mt units, so all 148 SMs get the same work to within one M-tile.
//   Warps    : 0-3 epilogue (TMEM lane quarter == warp id), 4 TMA producer, 5 MMA issuer/TMEM owner.
//   TMEM     : mt accumulators of NT columns, reused in a rolling fashion: the epilogue of M-tile m
//              overlaps the MMAs of M-tiles m+1.. and of the next tile (per-slot full/empty barriers).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"
#include "conv_params.h"
#include "esrp_philox.cuh"
#include "esrp_ptx.cuh"

namespace esrp {

constexpr int kNumEpiWarps = 4;
constexpr int kConvThreads = 32 * (kNumEpiWarps + 2);
constexpr int kMaxStages = 8;
constexpr int kSmemFixed = 4096;  // barriers + bias (2 KB) and the epilogue exchange buffer (2 KB)

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

// Optional in-kernel timeline (debug/profiling): CTA 0 records clock64() per pipeline event into
// p.trace[role * 1024 + n]; role 0 = producer, 1 = MMA issuer, 2 = epilogue thread 0.
__device__ __forceinline__ void trace_ev(const ConvKParams& p, int role, uint32_t& n) {
  if (p.trace != nullptr && blockIdx.x == 0 && n < 1024u) p.trace[role * 1024 + n++] = clock64();
}

// Walks the tiles of this CTA: identical in the three warp roles.
struct TileWalk {
  int u, u_end;
  int img, x_start, y0, cnt;  // current tile
  __device__ __forceinline__ explicit TileWalk(const ConvKParams& p) {
    const long long U = p.units_total;
    u = static_cast<int>(U * blockIdx.x / gridDim.x);
    u_end = static_cast<int>(U * (blockIdx.x + 1) / gridDim.x);
    img = x_start = y0 = cnt = 0;
  }
  __device__ __forceinline__ bool next(const ConvKParams& p) {
    if (u >= u_end) return false;
    const int col = u / p.units_per_col;
    const int yu = u - col * p.units_per_col;
    cnt = min(min(p.mt, u_end - u), p.units_per_col - yu);
    img = col / p.x_tiles;
    x_start = (col - img * p.x_tiles) * p.x_step;
    y0 = yu * p.rm;
    u += cnt;
    return true;
  }
};

// GC consecutive values of a residual tensor (fp32 or bf16 storage) into fp32 registers.
template <int GC>
__device__ __forceinline__ void load_residual(const void* base, int is_f32, size_t elem_off, float (&r)[GC],
                                              size_t f4_step = 1) {  // fp32: distance between consecutive float4s
  if (is_f32) {
    const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
#pragma unroll
    for (int i = 0; i < GC / 4; ++i) {
      const float4 t = ld_global_f4_hint(rp + i * f4_step, kL2EvictFirst);  // fp32 trunk: streamed, do not displace bf16 operands
      r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
    }
  } else {
    const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem_off);
#pragma unroll
    for (int i = 0; i < GC / 8; ++i) {
      const uint4 t = __ldg(rp + i);
      const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[8 * i + 2 * j] = __uint_as_float(u[j] << 16);
        r[8 * i + 2 * j + 1] = __uint_as_float(u[j] & 0xFFFF0000u);
      }
    }
  }
}

template <int GC>
__device__ __forceinline__ void tmem_ld_group(uint32_t taddr, uint32_t (&v)[GC]) {
  if constexpr (GC == 32) {
    tmem_ld_x32(taddr, v);
  } else {
    tmem_ld_x16(taddr, v);
  }
}


// ---- training extensions of the fused tail (EXT kernels; see include/esrp.h) ----
// Forward: save the LeakyReLU derivative selector, one bit per activation (v = acc + bias, pre-activation).
template <int GC>
__device__ __forceinline__ void ext_mask_store(const ConvKParams& p, size_t pix, int ch0, const float (&v)[GC]) {
  if (p.mask_out == nullptr) return;
  unsigned short* mp = p.mask_out + ((pix * static_cast<size_t>(p.mo_ctotal) + p.mo_c0 + ch0) >> 4);
#pragma unroll
  for (int j = 0; j < GC / 16; ++j) {
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) bits |= (v[16 * j + i] > 0.f ? 1u : 0u) << i;
    mp[j] = static_cast<unsigned short>(bits);
  }
}
// Backward: copies of the full gradient, then the LeakyReLU derivative from the saved bits.
template <int GC>
__device__ __forceinline__ void ext_pre_and_mask(const ConvKParams& p, size_t pix, int ch0, float (&v)[GC]) {
  if (p.pre_bf16) {
    uint4* op = reinterpret_cast<uint4*>(p.pre_bf16 + pix * p.pb_ctotal + p.pb_c0 + ch0);
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
  if (p.pre_f32) {
    float4* op = reinterpret_cast<float4*>(p.pre_f32 + pix * p.pf_ctotal + p.pf_c0 + ch0);
#pragma unroll
    for (int i = 0; i < GC / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (p.mask_in) {
    const unsigned short* mp = p.mask_in + ((pix * static_cast<size_t>(p.mi_ctotal) + p.mi_c0 + ch0) >> 4);
#pragma unroll
    for (int j = 0; j < GC / 16; ++j) {
      const uint32_t bits = mp[j];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[16 * j + i] *= ((bits >> i) & 1u) ? 1.0f : 0.2f;
    }
  }
}

template <int KC, int BN, bool EXT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                  const __grid_constant__ ConvKParams p) {
  constexpr int RB = KC * 2;                              // bytes per pixel row of a chunk
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;       // 128 B / 64 B swizzle
  constexpr uint32_t SBO = 8 * RB;                        // dense 8-row core-matrix groups
  constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
  constexpr int GC = BN < 32 ? BN : 32;                   // output channels per epilogue round
  constexpr int ROUNDS = BN / GC;

  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned carve-up: [barriers 1 KB][bias 1 KB][exchange 2 KB][resident weights][stages...]
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);   // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;              // [kMaxStages]
  uint64_t* acc_full = empty_bar + kMaxStages;              // [ESRP_MAX_MT]
  uint64_t* acc_empty = acc_full + ESRP_MAX_MT;             // [ESRP_MAX_MT]
  uint64_t* wfull = acc_empty + ESRP_MAX_MT;                // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wfull + 1);
  float* bias_s = reinterpret_cast<float*>(smem + 1024);    // [BN]
  float* xchg = reinterpret_cast<float*>(smem + 2048);      // [2 bufs][4 warps][2 sides][GC]

  const bool has_aux = p.aux_chunks > 0;
  const int nb_rows = has_aux ? 4 * BN : 3 * BN;            // B rows per (chunk, ky) block
  const int w_block_bytes = nb_rows * RB;
  const int w_chunk_bytes = 3 * w_block_bytes;
  const int w_res_bytes = p.w_resident ? p.num_chunks * w_chunk_bytes : 0;
  uint8_t* w_res = smem + kSmemFixed;
  uint8_t* stage0 = w_res + w_res_bytes;
  const int a_bytes = p.a_stage_bytes;                      // box bytes rounded up to 1024
  const int stage_bytes = a_bytes + (p.w_resident ? 0 : w_chunk_bytes);
  const int S = p.stages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_arrive(&empty_bar[i]);  // phase 0 = "the stage is free": complete from the start (no wait relies on the parity
                                   // of a phase that never existed; compute-sanitizer synccheck flags those as missing init)
    }
    for (int i = 0; i < ESRP_MAX_MT; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 32 * kNumEpiWarps);
      mbar_arrive_cnt(&acc_empty[i], 32 * kNumEpiWarps);  // phase 0 = "the accumulator slot is free"
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  grid_dep_launch_dependents();
  if (threadIdx.x < BN) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int NT = p.nt;  // TMEM columns per M-tile accumulator
  if (warp != 5) grid_dep_wait();  // producer (activation loads) and epilogue (residuals / stores)

  if (warp == 4) {
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
      TileWalk tw(p);
      while (tw.next(p)) {
        for (int c = 0; c < p.num_chunks; ++c, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty_bar[s], ph);
          trace_ev(p, 0, tn);
          uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s],
                                static_cast<uint32_t>(p.a_box_bytes + (p.w_resident ? 0 : w_chunk_bytes)));
          tma_load_4d(st, p.chunk_src[c] ? &tm1 : &tm0, &full_bar[s], p.chunk_c0[c], tw.x_start, tw.y0 - 1,
                      tw.img);
          if (!p.w_resident)
            bulk_load_1d(st + a_bytes, p.w_packed + static_cast<size_t>(c) * w_chunk_bytes, w_chunk_bytes,
                         &full_bar[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ====================================== MMA issuer ======================================
    // The whole warp walks the pipeline (converged), one elected lane issues the MMAs.
    if (p.w_resident) mbar_wait(wfull, 0);
    const uint32_t idesc_main = umma_idesc_bf16_m128(3 * BN);
    const uint32_t idesc_aux = umma_idesc_bf16_m128(4 * BN);
    const uint32_t mtile_bytes = 128u * RB;
    const uint32_t row_bytes = static_cast<uint32_t>(p.cw) * RB;
    uint32_t it = 0, tl = 0, tn = 0;
    if (lane == 0) trace_ev(p, 1, tn);
    TileWalk tw(p);
    while (tw.next(p)) {
      const uint32_t accph = tl & 1;
      for (int c = 0; c < p.num_chunks; ++c, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        if (lane == 0) trace_ev(p, 1, tn);
        uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
        const uint32_t a_lo = umma_desc_lo(smem_u32(st));
        const uint32_t b_lo =
            umma_desc_lo(p.w_resident ? smem_u32(w_res + c * w_chunk_bytes) : smem_u32(st + a_bytes));
        const bool last = (c == p.num_chunks - 1);
        const bool aux_c = c < p.aux_chunks;
        for (int m = 0; m < p.mt; ++m) {
          if (c == 0) {
            mbar_wait(&acc_empty[m], accph);
            tcgen05_fence_after();
          }
          const bool leader = elect_one();
          if (leader && m < tw.cnt) {
            const uint32_t d_tmem = tmem_base + m * NT;
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              const int ky = kk == 0 ? 1 : (kk == 1 ? 0 : 2);  // centre row first: it owns the conv1x1 columns
              const uint32_t a_off = m * mtile_bytes + ky * row_bytes;
              const uint32_t b_off = ky * w_block_bytes;
              const uint32_t idesc = (aux_c && ky == 1) ? idesc_aux : idesc_main;
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                umma_f16_ss2(d_tmem, a_lo + ((a_off + ks * 32) >> 4), DESC_HI, b_lo + ((b_off + ks * 32) >> 4),
                             DESC_HI, idesc, (c | kk | ks) != 0 ? 1u : 0u);
              }
            }
          }
          if (leader && last) umma_commit(&acc_full[m]);  // accumulator m complete (or unused in this tile)
          __syncwarp();
        }
        if (elect_one()) umma_commit(&empty_bar[s]);  // smem slot free once these MMAs have read it
        __syncwarp();
        if (lane == 0) trace_ev(p, 1, tn);
      }
      ++tl;
    }
  } else {
    // ======================================= epilogue =======================================
    uint32_t tl = 0, tn = 0;
    const int row = threadIdx.x;               // 0..127 == TMEM lane == M row
    const int ry = row >> p.cw_log2, rx = row & (p.cw - 1);
    const bool need_xchg = p.cw > 32;
    const bool has_left = rx > 0, has_right = rx < p.cw - 1;
    if (threadIdx.x == 0) trace_ev(p, 2, tn);
    uint32_t xbuf = 0;
    TileWalk tw(p);
    while (tw.next(p)) {
      const uint32_t accph = tl & 1;
      const int xs = tw.x_start + rx;           // source == output column of this lane
      const bool col_ok = xs < p.w && (has_left || xs == 0) && (has_right || xs == p.w - 1);
#pragma unroll 1
      for (int m = 0; m < p.mt; ++m) {
        const int py = tw.y0 + m * p.rm + ry;
        const bool valid = (m < tw.cnt) && col_ok && (py < p.h);
        const size_t pix = (static_cast<size_t>(tw.img) * p.h + py) * p.w + xs;
        // residual prefetch (addresses do not depend on the accumulator): hides the HBM/L2 latency
        float r1v[GC], r2v[GC];
        if (ROUNDS == 1) {
          if (p.r1 && valid) load_residual<GC>(p.r1, p.r1_is_f32, pix * p.r1_ctotal + p.r1_c0, r1v);
          if (p.r2 && valid) load_residual<GC>(p.r2, p.r2_is_f32, pix * p.r2_ctotal + p.r2_c0, r2v);
        }
        mbar_wait(&acc_full[m], accph);
        tcgen05_fence_after();
        if (threadIdx.x == 0) trace_ev(p, 2, tn);
        if (m >= tw.cnt) {  // slot unused in this (partial) tile: just hand it back
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[m]);
          continue;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + m * NT;
#pragma unroll 1
        for (int g = 0; g < ROUNDS; ++g) {
          uint32_t p0[GC], p1[GC], p2[GC], ax[GC];
          tmem_ld_group<GC>(taddr + 0 * BN + g * GC, p0);
          tmem_ld_group<GC>(taddr + 1 * BN + g * GC, p1);
          tmem_ld_group<GC>(taddr + 2 * BN + g * GC, p2);
          if (has_aux) tmem_ld_group<GC>(taddr + 3 * BN + g * GC, ax);
          if (ROUNDS > 1) {
            if (p.r1 && valid) load_residual<GC>(p.r1, p.r1_is_f32, pix * p.r1_ctotal + p.r1_c0 + g * GC, r1v);
            if (p.r2 && valid) load_residual<GC>(p.r2, p.r2_is_f32, pix * p.r2_ctotal + p.r2_c0 + g * GC, r2v);
          }
          tmem_ld_wait();
          if (g == ROUNDS - 1) {  // accumulator is in registers: hand the TMEM slot back to the MMA warp
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[m]);
          }
          // ---- column shift: out[x] = P0[x-1] + P1[x] + P2[x+1] ----
          float* xb = xchg + xbuf * (kNumEpiWarps * 2 * GC);
          if (need_xchg) {
            if (lane == 31) {
#pragma unroll
              for (int i = 0; i < GC; ++i) xb[(warp * 2 + 0) * GC + i] = __uint_as_float(p0[i]);
            }
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < GC; ++i) xb[(warp * 2 + 1) * GC + i] = __uint_as_float(p2[i]);
            }
            named_bar_sync(1, 32 * kNumEpiWarps);
          }
          float v[GC];
#pragma unroll
          for (int i = 0; i < GC; ++i) {
            float l = __shfl_up_sync(0xffffffffu, __uint_as_float(p0[i]), 1);
            float r = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[i]), 1);
            if (need_xchg) {
              if (lane == 0 && has_left) l = xb[((warp - 1) * 2 + 0) * GC + i];
              if (lane == 31 && has_right) r = xb[((warp + 1) * 2 + 1) * GC + i];
            }
            if (!has_left) l = 0.f;
            if (!has_right) r = 0.f;
            v[i] = __uint_as_float(p1[i]) + l + r;
          }
          xbuf ^= 1;
          if (threadIdx.x == 0) trace_ev(p, 2, tn);
          if (!valid) continue;
          const int ch0 = g * GC;
#pragma unroll
          for (int i = 0; i < GC; ++i) v[i] += bias_s[ch0 + i];
          if constexpr (EXT) ext_mask_store<GC>(p, pix, ch0, v);
          if (p.act) {
            const float slope = p.act == 2 ? 0.f : 0.2f;  // LeakyReLU(0.2) / ReLU
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaxf(v[i], slope * v[i]);
          }
          if (p.s0 != 1.0f) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] *= p.s0;
          }
          if (has_aux) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] += __uint_as_float(ax[i]);
          }
          if (p.r1) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s1, r1v[i], v[i]);
          }
          if constexpr (EXT) {
            if (p.r2 && p.r2_pre) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] += r2v[i];
            }
            ext_pre_and_mask<GC>(p, pix, ch0, v);
          }
          if (p.noise) {
            const unsigned long long nseed = p.seed_ptr ? __ldg(p.seed_ptr) : p.seed;  // graph replays read the key from memory
            // y = t + N(0,1) * sigma * t  (block.py:117-121); one Philox counter per 4 channels of
            // element index e = pixel * noise_ctotal + noise_c0 + channel.
#pragma unroll
            for (int i = 0; i < GC; i += 4) {
              float z[4];
              philox_normal4(nseed,
                             p.offset + (pix * static_cast<unsigned long long>(p.noise_ctotal) + p.noise_c0 + ch0 + i) / 4, z);
#pragma unroll
              for (int j = 0; j < 4; ++j) v[i + j] = fmaf(z[j] * p.sigma, v[i + j], v[i + j]);
            }
          }
          if (EXT && p.r2_pre) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] *= p.s2;
          } else if (p.r2) {
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s2, v[i], r2v[i]);
          }
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
            if (p.out_lo) {  // split precision: what the bf16 rounding above dropped, as a second bf16 tensor slice
              uint4* ol = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.ob_ctotal + p.ob_lo_c0 + ch0);
#pragma unroll
              for (int i = 0; i < GC / 8; ++i) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float a = v[8 * i + 2 * j], b = v[8 * i + 2 * j + 1];
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                  const float2 hf = __bfloat1622float2(h2);
                  const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
                  pk[j] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                ol[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(p.out_f32 + pix * p.of_ctotal + p.of_c0 + ch0);
#pragma unroll
            for (int i = 0; i < GC / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (p.out_nchw) {
            const size_t plane = static_cast<size_t>(p.h) * p.w;
            float* op = p.out_nchw + (static_cast<size_t>(tw.img) * p.cout) * plane + static_cast<size_t>(py) * p.w + xs;
#pragma unroll
            for (int i = 0; i < GC; ++i)
              if (ch0 + i < p.cout) op[static_cast<size_t>(ch0 + i) * plane] = v[i];
          }
        }
        if (threadIdx.x == 0) trace_ev(p, 2, tn);
      }
      ++tl;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace esrp
