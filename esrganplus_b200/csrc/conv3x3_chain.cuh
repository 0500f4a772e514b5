// conv3x3_chain.cuh — persistent chain of row-streaming 3x3 convolutions for sm_100a: ONE launch runs a whole
// sequence of fused convs (the 5 convs of every ResidualDenseBlock_5C of the RRDB trunk, block.py:260-268 /
// 287-291: 345 phases for nb = 23), each phase being exactly the computation of conv3x3_row_kernel
// (conv3x3_row.cuh: ky stacked along N, shifted-descriptor kx taps, output-stationary TMEM ring, fused tail).
//
// Why: at 16 x 128 x 128 a dense-block conv is 6-36 us of tensor-pipe work; as separate launches each pays ~7 us of
// prologue (barrier init, TMEM allocation, weight fetch, pipeline fill / drain, launch gap: 220 KB of shared memory
// per CTA forbids overlap with the previous launch) — 19 % of the step.  Here a CTA keeps its barriers,
// its tensor memory and its row range over all phases; between phases it only swaps the resident weights.
//
// Inter-CTA dependencies (no grid barrier): CTA c publishes flags[c] = ph + 1 (release, gpu scope) when all its
// outputs of phase ph are written.  Phase ph of a CTA reads input rows [u_lo - 1, u_hi] and overwrites buffers whose
// rows [u_lo - 1, u_hi] its neighbours may still be reading in phase ph - 1, so before its first activation load it
// waits (acquire) for flags[c'] >= ph of every CTA c' whose phase-(ph-1) row range intersects [u_lo - 1, u_hi]
// (chain_dep_range, conv_chain_dep.h: normally c - 1, c, c + 1) — this covers both the read-after-write and the
// write-after-read hazards, and a CTA can run ahead of a distant one by as many phases as they are apart.
// Generic-proxy stores -> async-proxy (TMA) loads cross a proxy: fence.proxy.async on both sides of the flag.
// All CTAs must be co-resident (grid <= number of SMs, one CTA per SM by shared-memory size).
//
// Intra-CTA protocol (simplified with respect to conv3x3_row.cuh; every barrier has in-order waiters that wait for
// EVERY phase they could otherwise be lapped on):
//   full_bar[b]  producer -> the issuer of that row (TMA bytes of row buffer b)
//   buf_free[b]  issuer -> producer: tcgen05.commit after the last MMA that reads row buffer b (dedicated barrier:
//                the producer is never a waiter of a block barrier, so no ring-size condition ties it to the blocks)
//   blk_full[X]  2 arrivals = one tcgen05.commit of EACH issuer thread after its last contribution to the block
//   blk_empty[X] 4 epilogue warps read + zeroed block X -> the first issuer that touches its next occupant
//   tok[w]       issue turn: the two issuer threads alternate whole input rows (row I by warp I & 1); the hand-over
//                is bracketed by tcgen05.fence::before/after_thread_sync, so MMAs reach the pipe in row order and the
//                accumulation order (hence the result) is deterministic
//   wfull/wfree  producer -> issuers: weights of the phase resident; issuers -> producer: everything of the phase issued
#pragma once
#include "conv3x3_row.cuh"
#include "conv_chain_dep.h"

namespace esrp {

constexpr int kChainBlocks = 8;  // TMEM ring: 8 output-row blocks of BN columns + 8 conv1x1 blocks behind them

struct alignas(64) ChainPhase {
  CUtensorMap tm0, tm1;  // read by TMA straight from global memory (host-written before the launch, never modified)
  ConvKParams p;
};

struct ChainArgs {
  const ChainPhase* phases;
  unsigned int* flags;  // [gridDim.x], zero at launch
  int num_phases;
  int dep_all;          // wait for every CTA instead of the row neighbours (images wider than one 128-pixel tile)
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// generic proxy <-> async proxy (TMA) ordering for global memory
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// Residual operands inside the chain were written earlier in the SAME launch: no non-coherent (ld.global.nc) loads.
template <int GC>
__device__ __forceinline__ void load_residual_coherent(const void* base, int is_f32, size_t elem_off, float (&r)[GC],
                                                       size_t f4_step) {
  if (is_f32) {
    const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
#pragma unroll
    for (int i = 0; i < GC / 4; ++i) {
      float4 t;
      asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                   : "l"(rp + i * f4_step), "l"(kL2EvictFirst)
                   : "memory");
      r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
    }
  } else {
    const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem_off);
#pragma unroll
    for (int i = 0; i < GC / 8; ++i) {
      uint4 t;
      asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                   : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                   : "l"(rp + i)
                   : "memory");
      const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[8 * i + 2 * j] = __uint_as_float(u[j] << 16);
        r[8 * i + 2 * j + 1] = __uint_as_float(u[j] & 0xFFFF0000u);
      }
    }
  }
}

// All taps of one K-chunk of one input row (KSN of the KC/16 K-slices carry weights).  WRAP: the three blocks
// straddle the end of the ring and every tap is two narrower MMAs.
template <int KC, bool WRAP, int KSN>
__device__ __forceinline__ void chain_issue_chunk(uint32_t dA, uint32_t dB, uint32_t idA, uint32_t idB, uint32_t bB,
                                                  uint32_t al, uint32_t bl, uint32_t desc_hi, uint32_t w_block_desc) {
  constexpr int RB = KC * 2;
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
    for (int ks = 0; ks < KSN; ++ks) {
      const uint32_t a_d = al + ((kx * RB + ks * 32) >> 4);  // the 130-pixel row shifted by kx pixels
      const uint32_t b_d = bl + kx * w_block_desc + ((ks * 32) >> 4);
      umma_f16_ss2(dA, a_d, desc_hi, b_d, desc_hi, idA, 1u);
      if (WRAP) umma_f16_ss2(dB, a_d, desc_hi, b_d + bB, desc_hi, idB, 1u);
    }
  }
}

// Spin until flags[c] >= want for c in [c_lo, c_hi] (one flag per lane and round).
__device__ __forceinline__ void chain_wait_flags(const unsigned int* flags, int c_lo, int c_hi, unsigned int want, int lane) {
  for (int c = c_lo + lane; c <= c_hi; c += 32) {
    if (ld_acquire_gpu(flags + c) >= want) continue;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (ld_acquire_gpu(flags + c) < want) {
      __nanosleep(64);
      if ((++spins & 0x3FF) == 0 && (clock64() - t0) > 40000000000LL) asm volatile("trap;\n");
    }
  }
  __syncwarp();
}

template <int KC, int BN, bool EXT>
__global__ void __launch_bounds__(kRowThreads, 1) conv3x3_chain_kernel(const __grid_constant__ ChainArgs a) {
  static_assert(BN == 32, "the chain kernel is the N = 96 dense-block shape");
  constexpr int RB = KC * 2;
  constexpr int KS = KC / 16;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * RB;
  constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
  constexpr int NBLK = kChainBlocks;
  constexpr int AUX_COL0 = NBLK * BN;
  constexpr int GC = 16;
  constexpr int ROUNDS = BN / GC;
  static_assert(2 * NBLK * BN <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kMaxStages]
  uint64_t* buf_free = full_bar + kMaxStages;              // [kMaxStages]
  uint64_t* blk_full = buf_free + kMaxStages;              // [NBLK]
  uint64_t* blk_empty = blk_full + NBLK;                   // [NBLK]
  uint64_t* wfull = blk_empty + NBLK;                      // [1]
  uint64_t* wfree = wfull + 1;                             // [1]
  uint64_t* tok = wfree + 1;                               // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tok + 2);
  float* bias_s = reinterpret_cast<float*>(smem + 1024);   // [kRowEpiWarps][BN]: a private copy per epilogue warp
  uint8_t* const w_res = smem + kSmemFixed;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nph = a.num_phases;

  if (warp == kRowEpiWarps && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&buf_free[i], 1);
    }
    for (int i = 0; i < NBLK; ++i) {
      mbar_init(&blk_full[i], kRowMmaWarps);
      mbar_init(&blk_empty[i], 4);
    }
    mbar_init(wfull, 1);
    mbar_init(wfree, kRowMmaWarps);
    mbar_init(&tok[0], 1);
    mbar_init(&tok[1], 1);
    fence_barrier_init();
  }
  if (warp == kRowEpiWarps + 1) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  grid_dep_launch_dependents();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (warp < kRowEpiWarps) {  // every block starts zeroed: all MMAs accumulate
    const uint32_t la = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = (warp >> 2) * 16; c < NBLK * BN; c += 16 * kRowWGs) tmem_st_zero_x16(la + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  // ring position of output sequence number O: descending, so that the blocks of output rows r+1, r, r-1 (ky = 0, 1, 2)
  // have ascending columns
  auto pos = [](uint32_t O) -> uint32_t { return (NBLK - 1) - (O & (NBLK - 1)); };
  auto use = [](uint32_t O) -> uint32_t { return (O / NBLK) & 1; };

  if (warp == kRowEpiWarps) {
    // ===================================== TMA producer =====================================
    uint32_t bmask = 0;  // per row buffer: parity of the next buf_free phase to wait for
    int prev_d = 0, prev_rows = 0;
    grid_dep_wait();  // activations (and repacked weights) of the previous kernels
    for (int ph = 0; ph < nph; ++ph) {
      const ChainPhase& P = a.phases[ph];
      const ConvKParams& p = P.p;
      const int nsl = p.nsl > 1 ? p.nsl : 1;
      const int sl = nsl > 1 ? static_cast<int>(blockIdx.x) % nsl : 0;
      const int cta = static_cast<int>(blockIdx.x) / nsl, ncta = static_cast<int>(gridDim.x) / nsl;
      const int nch = p.num_chunks;
      const int w_chunk_bytes = 3 * (p.aux_chunks > 0 ? 4 : 3) * BN * RB;
      const int w_res_bytes = nch * w_chunk_bytes;
      uint8_t* const stage0 = w_res + w_res_bytes;
      const int a_bytes = p.a_stage_bytes;
      const int row_bytes = a_bytes * nch;
      const int D = p.stages;
      if (lane == 0) {
        if (ph > 0) {
          // the weights and row buffers of the previous phase are reusable when both issuers have issued all of it ...
          mbar_wait(wfree, static_cast<uint32_t>(ph - 1) & 1u);
          // ... and the MMAs of its outstanding rows have completed
          const int outst = prev_rows < prev_d ? prev_rows : prev_d;
          for (int b = 0; b < outst; ++b) {
            mbar_wait(&buf_free[b], (bmask >> b) & 1u);
            bmask ^= 1u << b;
          }
        }
        const uint8_t* const w_src = p.w_packed + static_cast<size_t>(sl) * p.sl_stride;
        mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(w_res_bytes));
        for (int c = 0; c < nch; ++c)
          bulk_load_1d(w_res + c * w_chunk_bytes, w_src + static_cast<size_t>(c) * w_chunk_bytes, w_chunk_bytes, wfull);
        tma_prefetch_desc(&P.tm0);
        tma_prefetch_desc(&P.tm1);
      }
      if (ph > 0) {
        // producers of my input rows / readers of the rows I am about to overwrite (see the header)
        const ConvKParams& q = a.phases[ph - 1].p;
        const int nslq = q.nsl > 1 ? q.nsl : 1;
        const int ngq = static_cast<int>(gridDim.x) / nslq;
        int g_lo = 0, g_hi = ngq - 1;
        if (!a.dep_all) chain_dep_range(p.units_total, cta, ncta, ngq, &g_lo, &g_hi);
        chain_wait_flags(a.flags, g_lo * nslq, g_hi * nslq + nslq - 1, static_cast<unsigned int>(ph), lane);
      }
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async_all();
        const uint32_t tx_bytes = static_cast<uint32_t>(p.a_box_bytes) * nch;
        int b = 0, ip = 0;
        uint8_t* st = stage0;
        SegWalk sw(p, cta, ncta);
        while (sw.next(p)) {
          const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, p.h - 1);
          for (int r = r0; r <= r1; ++r, ++ip) {
            if (ip >= D) {
              mbar_wait(&buf_free[b], (bmask >> b) & 1u);
              bmask ^= 1u << b;
            }
            mbar_arrive_expect_tx(&full_bar[b], tx_bytes);
            for (int c = 0; c < nch; ++c)
              tma_load_4d_hint(st + c * a_bytes, p.chunk_src[c] ? &P.tm1 : &P.tm0, &full_bar[b], p.chunk_c0[c], sw.x0 - 1, r,
                               sw.img, kL2EvictLast);  // bf16 activations are re-read by the next convs: keep in L2
            st += row_bytes;
            if (++b == D) { b = 0; st = stage0; }
          }
        }
        prev_d = D;
        prev_rows = ip;
      }
    }
    if (lane == 0) {  // leave no asynchronous arrival pending at exit
      mbar_wait(wfree, static_cast<uint32_t>(nph - 1) & 1u);
      const int outst = prev_rows < prev_d ? prev_rows : prev_d;
      for (int b = 0; b < outst; ++b) mbar_wait(&buf_free[b], (bmask >> b) & 1u);
    }
  } else if (warp > kRowEpiWarps) {
    // ====================================== MMA issuers ======================================
    const uint32_t mw = static_cast<uint32_t>(warp - (kRowEpiWarps + 1));
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(w_res));
    constexpr uint32_t ID_FULL = umma_idesc_bf16_m128(3 * BN);
    const uint32_t idesc_aux = umma_idesc_bf16_m128(BN);
    uint32_t fmask = 0;  // per row buffer: parity of its next full_bar phase (both warps count every row)
    uint32_t tph = 0, O0 = 0, I = 0;
    for (int ph = 0; ph < nph; ++ph) {
      const ConvKParams& p = a.phases[ph].p;
      const int nsl = p.nsl > 1 ? p.nsl : 1;
      const int cta = static_cast<int>(blockIdx.x) / nsl, ncta = static_cast<int>(gridDim.x) / nsl;
      const int nch = p.num_chunks, naux = p.aux_chunks;
      const int nfull = p.last_half ? nch - 1 : nch;  // chunks issued over all of their K-slices
      const uint32_t w_block_bytes = static_cast<uint32_t>((naux > 0 ? 4 : 3) * BN * RB);
      const uint32_t w_chunk_bytes = 3u * w_block_bytes;
      const uint32_t w_step = w_chunk_bytes >> 4, w_block_desc = w_block_bytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(w_res + static_cast<uint32_t>(nch) * w_chunk_bytes));
      const uint32_t chunk_step = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
      const uint32_t row_step = chunk_step * static_cast<uint32_t>(nch);
      const int D = p.stages;
      mbar_wait(wfull, static_cast<uint32_t>(ph) & 1u);
      int b = 0;
      uint32_t a_lo = a_lo0;
      SegWalk sw(p, cta, ncta);
      while (sw.next(p)) {
        const int ni = min(sw.yb, p.h - 1) - max(sw.ya - 1, 0) + 1;
        for (int k = 0; k < ni; ++k) {
          if ((I & 1u) == mw) {
            mbar_wait(&full_bar[b], (fmask >> b) & 1u);
            // blocks first touched by this input row must have been read + zeroed by their previous occupant
            const uint32_t On = O0 + k + 2;  // output row r+1 (ky = 0): always new
            if (k == 0) {
              mbar_wait(&blk_empty[pos(O0)], use(O0) ^ 1u);
              mbar_wait(&blk_empty[pos(O0 + 1)], use(O0 + 1) ^ 1u);
            }
            mbar_wait(&blk_empty[pos(On)], use(On) ^ 1u);
            // my turn: the other warp has issued every MMA of the previous row
            mbar_wait(&tok[mw], tph ^ (mw == 0 ? 1u : 0u));
            tcgen05_fence_after();
            if (elect_one()) {
              // accumulator = blocks pos(On), +1, +2; split in two MMAs where it straddles the end of the ring
              const uint32_t Pb = pos(On);
              const uint32_t nA = (Pb + 3 <= NBLK) ? 3u * BN : (NBLK - Pb) * BN;  // columns before the wrap
              const uint32_t nB = 3u * BN - nA;
              const uint32_t dA = tmem_base + Pb * BN, dB = tmem_base;
              const uint32_t bB = (nA * RB) >> 4;  // B rows of the second part
              uint32_t al = a_lo, bl = w_lo0;
              if (nB == 0) {
                for (int c = 0; c < nfull; ++c, al += chunk_step, bl += w_step)
                  chain_issue_chunk<KC, false, KS>(dA, dB, ID_FULL, 0u, 0u, al, bl, DESC_HI, w_block_desc);
                if (nfull < nch)  // last chunk: only its first half carries weights (K = 96 / 160 in 64-channel chunks)
                  chain_issue_chunk<KC, false, KS / 2>(dA, dB, ID_FULL, 0u, 0u, al, bl, DESC_HI, w_block_desc);
              } else {
                const uint32_t idA = umma_idesc_bf16_m128(nA), idB = umma_idesc_bf16_m128(nB);
                for (int c = 0; c < nfull; ++c, al += chunk_step, bl += w_step)
                  chain_issue_chunk<KC, true, KS>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
                if (nfull < nch) chain_issue_chunk<KC, true, KS / 2>(dA, dB, idA, idB, bB, al, bl, DESC_HI, w_block_desc);
              }
              if (naux > 0) {  // conv1x1 of output row r: centre pixel x B rows [3*BN, 4*BN) of the kx = 1 block
                const uint32_t d_aux = tmem_base + AUX_COL0 + pos(O0 + k + 1) * BN;
                al = a_lo;
                bl = w_lo0;
                for (int c = 0; c < naux; ++c, al += chunk_step, bl += w_step) {
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks)
                    umma_f16_ss2(d_aux, al + ((1 * RB + ks * 32) >> 4), DESC_HI,
                                 bl + w_block_desc + ((3 * BN * RB + ks * 32) >> 4), DESC_HI, idesc_aux,
                                 (c | ks) != 0 ? 1u : 0u);
                }
              }
              tcgen05_fence_before();
              mbar_arrive(&tok[mw ^ 1u]);  // the other warp's turn
              // Completion tracking.  A commit covers the MMAs of THIS thread only, and block y collects the rows
              // y-1, y+1 (one warp) and y (the other): each warp commits on a block after its last contribution.
              umma_commit(&blk_full[pos(O0 + k + 1)]);                  // output row r: my only contribution
              umma_commit(&blk_full[pos(O0 + k)]);                      // output row r-1: my last contribution
              if (k == 0) umma_commit(&blk_full[pos(O0)]);              // (dummy block above the segment: nobody else)
              if (k == ni - 2) umma_commit(&blk_full[pos(O0 + k + 2)]);  // last output row of the segment: no row r+2 follows
              if (k == ni - 1) {
                umma_commit(&blk_full[pos(O0 + k + 2)]);                // (dummy block below the segment: nobody else)
                umma_commit(&blk_full[pos(O0 + k + 2)]);
                if (ni == 1) umma_commit(&blk_full[pos(O0 + k + 1)]);   // one-row segment: nobody else either
              }
              umma_commit(&buf_free[b]);                                // row buffer b may be refilled
            }
            __syncwarp();
            tph ^= 1u;
          }
          fmask ^= 1u << b;
          ++I;
          a_lo += row_step;
          if (++b == D) { b = 0; a_lo = a_lo0; }
        }
        O0 += static_cast<uint32_t>(ni + 2);
      }
      if (lane == 0) mbar_arrive(wfree);
    }
  } else {
    // ======================================= epilogue =======================================
    // Three warpgroups take the output rows round-robin.  Thread == pixel: it reads the finished block of its output
    // row, zeroes it, hands it back, then applies the fused tail and stores all BN channels (conv3x3_row.cuh).
    const int wg = warp >> 2;
    const int q = warp & 3;        // TMEM lane quarter
    const int xl = q * 32 + lane;  // column within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* const bias_w = bias_s + warp * BN;
    uint32_t O0 = 0;
    int turn = 0;
    grid_dep_wait();  // residual reads / output writes must not race with the previous kernel
    for (int ph = 0; ph < nph; ++ph) {
      const ConvKParams& p = a.phases[ph].p;
      const int nsl = p.nsl > 1 ? p.nsl : 1;
      const int sl = nsl > 1 ? static_cast<int>(blockIdx.x) % nsl : 0;
      const int cta = static_cast<int>(blockIdx.x) / nsl, ncta = static_cast<int>(gridDim.x) / nsl;
      const int csh = sl * BN;  // channel shift of every global channel offset
      const bool aux = p.aux_chunks > 0;
      __syncwarp();
      bias_w[lane] = p.bias ? reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.bias) +
                                                               static_cast<size_t>(sl) * p.sl_stride)[lane]
                            : 0.f;
      __syncwarp();
      SegWalk sw(p, cta, ncta);
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
          const bool real = y >= sw.ya && y < sw.yb;
          const bool store = real && col_ok;
          const uint32_t blk = lane_addr + pos(O) * BN;
          const size_t rowid = static_cast<size_t>(sw.img) * p.h + (real ? y : sw.ya);
          const size_t pix = rowid * p.w + (col_ok ? xs : 0);
          const bool planar = p.f32_planar != 0;
          const size_t f4_step = planar ? static_cast<size_t>(p.w) : 1;
          auto off32 = [&](int ct, int c) -> size_t {
            return planar ? ((rowid * (ct >> 2) + (c >> 2)) * p.w + (col_ok ? xs : 0)) * 4 : pix * ct + c;
          };
          auto off_res = [&](int is_f32, int ct, int c) -> size_t { return is_f32 ? off32(ct, c) : pix * ct + c; };
          // residuals of the first round: in flight while we wait for the accumulator
          float r1v[GC], r2v[GC];
          if (real) {  // every lane (the shuffles of the bf16 store need the whole warp): pix is clamped for columns >= w
            if (p.r1) load_residual_coherent<GC>(p.r1, p.r1_is_f32, off_res(p.r1_is_f32, p.r1_ctotal, p.r1_c0 + csh), r1v, f4_step);
            if (p.r2) load_residual_coherent<GC>(p.r2, p.r2_is_f32, off_res(p.r2_is_f32, p.r2_ctotal, p.r2_c0 + csh), r2v, f4_step);
          }
          mbar_wait(&blk_full[pos(O)], use(O));
          tcgen05_fence_after();
          if (!real) {  // dummy row at a segment end: just recycle the block
#pragma unroll
            for (int c = 0; c < BN; c += GC) tmem_st_zero_x16(blk + c);
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&blk_empty[pos(O)]);
            continue;
          }
          uint32_t pend[8];  // bf16 output of an even round, stored together with the following odd round
          const bool quad = (ROUNDS % 2 == 0) && (p.cout % (2 * GC) == 0) && !p.no_quad;
#pragma unroll
          for (int g = 0; g < ROUNDS; ++g) {
            const int ch0 = g * GC;
            uint32_t acc[GC], ax[GC];
            tmem_ld_x16(blk + ch0, acc);
            if (aux) tmem_ld_x16(blk + AUX_COL0 + ch0, ax);
            tmem_ld_wait();
            tmem_st_zero_x16(blk + ch0);
            if (g == ROUNDS - 1) {
              tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&blk_empty[pos(O)]);
            }
            if (ch0 >= p.cout) continue;  // (warp-uniform; lanes of columns >= w compute along and store nothing)
            const int gch = ch0 + csh;    // channel relative to the *_c0 offsets of the descriptor
            float v[GC];
            const float4* bias4 = reinterpret_cast<const float4*>(bias_w + ch0);
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
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] = fmaxf(v[i], 0.2f * v[i]);  // LeakyReLU(0.2)
            }
            if (p.s0 != 1.0f) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] *= p.s0;
            }
            if (aux) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] += __uint_as_float(ax[i]);
            }
            if (p.r1) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] = fmaf(p.s1, r1v[i], v[i]);
              if (g + 1 < ROUNDS)
                load_residual_coherent<GC>(p.r1, p.r1_is_f32, off_res(p.r1_is_f32, p.r1_ctotal, p.r1_c0 + gch + GC), r1v, f4_step);
            }
            if constexpr (EXT) {
              if (p.r2 && p.r2_pre) {
#pragma unroll
                for (int i = 0; i < GC; ++i) v[i] += r2v[i];
                if (g + 1 < ROUNDS)
                  load_residual_coherent<GC>(p.r2, p.r2_is_f32, off_res(p.r2_is_f32, p.r2_ctotal, p.r2_c0 + gch + GC), r2v, f4_step);
              }
              if (store) ext_pre_and_mask<GC>(p, pix, gch, v);
            }
            if (p.noise) {
              const unsigned long long nseed = p.seed_ptr ? __ldg(p.seed_ptr) : p.seed;
#pragma unroll 1
              for (int i = 0; i < GC; i += 4) {
                float z[4];
                philox_normal4(nseed, p.offset + (pix * static_cast<unsigned long long>(p.noise_ctotal) + p.noise_c0 + gch + i) / 4, z);
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
              if (g + 1 < ROUNDS)
                load_residual_coherent<GC>(p.r2, p.r2_is_f32, off_res(p.r2_is_f32, p.r2_ctotal, p.r2_c0 + gch + GC), r2v, f4_step);
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
                // 4x4 transpose inside each group of four lanes: one store instruction then writes 8 x 64 contiguous
                // bytes instead of 32 x 16 (conv3x3_row.cuh)
                const bool hi = (lane & 2) != 0, lo = (lane & 1) != 0;
                uint32_t d0[8], d1[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const uint32_t snd = hi ? pend[i] : pk[i], kp = hi ? pk[i] : pend[i];
                  const uint32_t rc = __shfl_xor_sync(0xffffffffu, snd, 2);
                  d0[i] = hi ? rc : kp;
                  d1[i] = hi ? kp : rc;
                }
                uint32_t e[4][4];
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
          }
        }
        O0 += static_cast<uint32_t>(no);
      }
      // all outputs of this phase written: publish (the neighbours' producers acquire the flag before their TMA loads)
      __threadfence();
      named_bar_sync(1, kRowEpiWarps * 32);
      if (threadIdx.x == 0) {
        fence_proxy_async_all();
        st_release_gpu(a.flags + blockIdx.x, static_cast<unsigned int>(ph + 1));
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kRowEpiWarps + 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace esrp
