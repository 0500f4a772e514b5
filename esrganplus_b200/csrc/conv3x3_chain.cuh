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
// The phase table (64 bytes per conv + the distinct tensor maps) is the kernel's PARAMETER block (constant bank, up to
// 32 KB with CUDA 12.1+): as a table in global memory every field access of the epilogue was an L2 round trip, because
// the gpu-scope fences of the flag protocol keep invalidating L1 (ncu: 35 % of all stall samples).
//
// Intra-CTA protocol (warps 0-11 epilogue = 3 warpgroups, 12 TMA producer, 13-15 MMA issuers):
//   THREE issuer threads work concurrently on every input row, thread kx on the taps of kernel column kx.  Measured in
//   this kernel with loads and epilogue switched off (tools/chain_trace.py): one thread alone issues an N = 96 MMA every
//   85-105 cycles while the pipe needs ~48, two or three threads together reach the pipe's rate; every row costs a thread
//   ~900 cycles of barrier waits, commit and loop control during which it issues nothing.
//   full_bar[t]  producer -> issuers: TMA bytes of activation tile t (ring of `stages` rows x num_chunks tiles)
//   blk_full[X]  3 arrivals = one tcgen05.commit of EACH issuer thread after its share of the input row that completes
//                output block X.  Waited for by the epilogue warpgroup of that row AND by the producer: the tiles of
//                input row r are free when the block that row completes is (planner: rows in the ring + 2 per segment
//                end < blocks in the ring, so the producer cannot be lapped on a block barrier)
//   blk_empty[X] 4 epilogue warps read + zeroed block X -> every issuer, before it touches the block's next occupant
//   wfull/wfree  producer -> issuers: weights of the phase resident; issuers -> producer: everything of the phase issued
//   Every waiter derives the parity of a barrier phase from a per-barrier use count that all roles advance in the same
//   order (bit masks), so the block ring restarts at every phase: 16 blocks of 32 columns, or 8 + 8 conv1x1 blocks in the
//   conv2 phase (2 of every `blocks` rows straddle the ring end and cost two narrower MMAs per tap: 88 instead of 56 cycles).
//   No barrier can be lapped: a tile / block is only recycled after ALL issuer threads have committed on it.  The order in
//   which the threads' MMAs reach the pipe is not fixed, so sums are reproducible up to fp32 addition order only
//   (esrp_rrdbnet_set_chain(h, 0) selects the one-launch-per-conv path, which reproduces bit for bit).
#pragma once
#include "conv3x3_row.cuh"
#include "conv_chain_dep.h"

namespace esrp {

constexpr int kChainBlocks = 16;      // TMEM ring: 16 output-row blocks of 32 columns (8 + 8 conv1x1 blocks in the conv2 phase)
constexpr int kChainMaxPhases = 440;  // phases per launch: the whole table travels in the kernel's parameter space (< 32 KB)
constexpr int kChainMaxMaps = 16;
constexpr int kChainMaxTiles = 12;    // activation tiles in the shared-memory ring
constexpr int kChainIssuers = 3;      // one per kernel column kx
constexpr int kChainThreads = 32 * (kRowEpiWarps + 1 + kChainIssuers);

// One conv of the chain, 64 bytes.  Pointers are 16-byte units relative to ChainArgs::act_base / w_base.
struct ChainPhaseC {
  uint32_t w_off16, bias_off16;                  // packed weights / bias of slice 0 (w_base)
  uint32_t sl_stride16;                          // distance to the next co-scheduled slice
  uint32_t r1_off16, r2_off16, ob_off16, of_off16;  // residuals / outputs (act_base); kChainNull = absent
  float s0, s2;                                  // (s1 == 1 for every dense-block conv)
  uint16_t r1_ctotal, r1_c0, r2_ctotal, r2_c0, ob_ctotal, ob_c0, of_ctotal, of_c0;
  uint8_t chunk_c0_8[4];                         // first channel of each K-chunk / 8
  uint8_t chunk_src;                             // bit c: chunk c is read through tensor map tm1
  uint8_t tm0, tm1;                              // indices into ChainArgs::tmaps
  uint8_t num_chunks, aux_chunks, nsl, stages;   // stages: activation ROWS in the ring (stages * num_chunks <= kChainMaxTiles)
  uint8_t flags;                                 // kChainF*
};
static_assert(sizeof(ChainPhaseC) == 64, "ChainPhaseC layout");
constexpr uint32_t kChainNull = 0xFFFFFFFFu;
constexpr int kChainFAct = 1, kChainFR1F32 = 2, kChainFR2F32 = 4, kChainFPlanar = 8, kChainFLastHalf = 16, kChainFNoQuad = 32;

struct ChainArgs {
  CUtensorMap tmaps[kChainMaxMaps];
  ChainPhaseC ph[kChainMaxPhases];
  const uint8_t* w_base;
  uint8_t* act_base;
  unsigned int* flags;  // [gridDim.x], zero at launch
  long long units_total;
  int n, h, w, x_tiles;
  int a_box_bytes, a_stage_bytes;
  int num_phases;
  int dep_all;          // wait for every CTA instead of the row neighbours (images wider than one 128-pixel tile)
  long long* trace;     // optional [num_phases][16] clock64 timeline of CTA trace_cta (ESRP_CHAIN_TRACE, tools/chain_trace.py)
  int trace_cta;
  int dbg;              // timing / bisecting experiments (ESRP_CHAIN_DBG; results are WRONG with any of them): 2 = do not wait
                        // for the neighbours' flags, 4 = no __threadfence before the flag, 8 = no activation loads,
                        // 256 = no epilogue, 8192 = issuers start every phase together
};
static_assert(sizeof(ChainArgs) <= 32764, "kernel parameter space");

// Row segments of one CTA (SegWalk of conv3x3_row.cuh over the chain's common geometry).
struct ChainWalk {
  int u, u_end, h, x_tiles;
  int img, x0, ya, yb;
  __device__ __forceinline__ ChainWalk(long long U, int h_, int x_tiles_, int cta, int ncta) : h(h_), x_tiles(x_tiles_) {
    u = static_cast<int>(U * cta / ncta);
    u_end = static_cast<int>(U * (cta + 1) / ncta);
    img = x0 = ya = yb = 0;
  }
  __device__ __forceinline__ bool next() {
    if (u >= u_end) return false;
    const int col = u / h;
    ya = u - col * h;
    const int cnt = min(u_end - u, h - ya);
    yb = ya + cnt;
    img = col / x_tiles;
    x0 = (col - img * x_tiles) * kRowTile;
    u += cnt;
    return true;
  }
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

// The taps of ONE kernel column kx of one K-chunk of one input row (KSN of the KC/16 K-slices carry weights).  WRAP: the
// three blocks straddle the end of the ring and every tap is two narrower MMAs.  All MMAs accumulate (the blocks were
// zeroed by their previous reader).
template <int KC, bool WRAP, int KSN>
__device__ __forceinline__ void chain_issue_chunk(uint32_t dA, uint32_t dB, uint32_t idA, uint32_t idB, uint32_t bB,
                                                  uint32_t a_kx, uint32_t b_kx, uint32_t desc_hi) {
#pragma unroll
  for (int ks = 0; ks < KSN; ++ks) {
    const uint32_t a_d = a_kx + ((ks * 32) >> 4);
    const uint32_t b_d = b_kx + ((ks * 32) >> 4);
    umma_f16_ss2(dA, a_d, desc_hi, b_d, desc_hi, idA, 1u);
    if (WRAP) umma_f16_ss2(dB, a_d, desc_hi, b_d + bB, desc_hi, idB, 1u);
  }
}
template <int KC>
__device__ __forceinline__ void chain_issue_chunk_any(bool wrap, bool half, uint32_t dA, uint32_t dB, uint32_t idA,
                                                      uint32_t idB, uint32_t bB, uint32_t a_kx, uint32_t b_kx,
                                                      uint32_t desc_hi) {
  constexpr int KS = KC / 16;
  if (!wrap) {
    if (!half) chain_issue_chunk<KC, false, KS>(dA, dB, idA, idB, bB, a_kx, b_kx, desc_hi);
    else chain_issue_chunk<KC, false, KS / 2>(dA, dB, idA, idB, bB, a_kx, b_kx, desc_hi);
  } else {
    if (!half) chain_issue_chunk<KC, true, KS>(dA, dB, idA, idB, bB, a_kx, b_kx, desc_hi);
    else chain_issue_chunk<KC, true, KS / 2>(dA, dB, idA, idB, bB, a_kx, b_kx, desc_hi);
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

#define CHAIN_TRACE(ph, slot) \
  do { if (a.trace != nullptr && static_cast<int>(blockIdx.x) == a.trace_cta) a.trace[(ph) * 16 + (slot)] = clock64(); } while (0)

template <int KC, int BN>
__global__ void __launch_bounds__(kChainThreads, 1) conv3x3_chain_kernel(const __grid_constant__ ChainArgs a) {
  static_assert(BN == 32 && KC == 64, "the chain kernel is the N = 96 dense-block shape over 64-channel chunks");
  constexpr int RB = KC * 2;
  constexpr int KS = KC / 16;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * RB;
  constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (LAYOUT << 29);
  constexpr int NBLK_MAX = kChainBlocks;
  constexpr int AUX_COL0 = (NBLK_MAX / 2) * BN;  // conv1x1 blocks of the 8-block ring live in the columns of blocks 8..15
  constexpr int GC = 16;
  constexpr int ROUNDS = BN / GC;
  static_assert(NBLK_MAX * BN <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [kChainMaxTiles]
  uint64_t* blk_full = full_bar + kChainMaxTiles;          // [NBLK_MAX]
  uint64_t* blk_empty = blk_full + NBLK_MAX;               // [NBLK_MAX]
  uint64_t* wfull = blk_empty + NBLK_MAX;                  // [1]
  uint64_t* wfree = wfull + 1;                             // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wfree + 1);
  uint32_t* row_rec = tmem_holder + 1;                     // [4] producer: block barrier (index | parity << 8) that frees each row of tiles
  float* bias_s = reinterpret_cast<float*>(smem + 1024);   // [kRowEpiWarps][BN]: a private copy per epilogue warp
  uint8_t* const w_res = smem + kSmemFixed;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nph = a.num_phases;
  const int img_h = a.h, img_w = a.w;

  if (warp == kRowEpiWarps && lane == 0) {
    for (int i = 0; i < kChainMaxTiles; ++i) mbar_init(&full_bar[i], 1);
    for (int i = 0; i < NBLK_MAX; ++i) {
      mbar_init(&blk_full[i], kChainIssuers);
      mbar_init(&blk_empty[i], 4);
      mbar_arrive_cnt(&blk_empty[i], 4);  // phase 0 = "the block is free": complete from the start
    }
    mbar_init(wfull, 1);
    mbar_init(wfree, kChainIssuers);
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
  if (warp < kRowEpiWarps) {  // every block starts zeroed: all MMAs accumulate, readers zero what they read
    const uint32_t la = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (int c = (warp >> 2) * 16; c < NBLK_MAX * BN; c += 16 * kRowWGs) tmem_st_zero_x16(la + c);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();

  // Ring position of output sequence number O (restarting at 0 in every phase) in a ring of nblk blocks: descending, so
  // that the blocks of output rows r+1, r, r-1 (ky = 0, 1, 2) have ascending columns.
  auto pos = [](uint32_t O, uint32_t nblk) -> uint32_t { return (nblk - 1) - (O & (nblk - 1)); };

  if (warp == kRowEpiWarps) {
    // ===================================== TMA producer =====================================
    uint32_t pcnt = 0;   // per block: use count & 1 == parity of the blk_full phase of its current use
    int prev_out = 0;    // rows of the previous phase whose completion has not been waited for yet (slots 0 .. prev_out-1)
    grid_dep_wait();     // activations (and repacked weights) of the previous kernels
    for (int ph = 0; ph < nph; ++ph) {
      const ChainPhaseC& P = a.ph[ph];
      const int nsl = P.nsl;
      const int sl = nsl > 1 ? static_cast<int>(blockIdx.x) % nsl : 0;
      const int cta = static_cast<int>(blockIdx.x) / nsl, ncta = static_cast<int>(gridDim.x) / nsl;
      const int nch = P.num_chunks;
      const uint32_t nblk = P.aux_chunks > 0 ? NBLK_MAX / 2 : NBLK_MAX;
      const int w_chunk_bytes = 3 * (P.aux_chunks > 0 ? 4 : 3) * BN * RB;
      const int w_res_bytes = nch * w_chunk_bytes;
      uint8_t* const stage0 = w_res + w_res_bytes;
      const int a_bytes = a.a_stage_bytes;
      const int DR = P.stages;
      if (lane == 0) {
        CHAIN_TRACE(ph, 0);
        if (ph > 0) {
          // the weights and tiles of the previous phase are reusable when every issuer has issued all of it ...
          mbar_wait(wfree, static_cast<uint32_t>(ph - 1) & 1u);
          CHAIN_TRACE(ph, 1);
          // ... and the MMAs of its outstanding rows have completed
          for (int s = 0; s < prev_out; ++s) mbar_wait(&blk_full[row_rec[s] & 0xFFu], row_rec[s] >> 8);
        }
        const uint8_t* const w_src = a.w_base + (static_cast<size_t>(P.w_off16) + static_cast<size_t>(sl) * P.sl_stride16) * 16;
        CHAIN_TRACE(ph, 2);
        mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(w_res_bytes));
        for (int c = 0; c < nch; ++c)
          bulk_load_1d(w_res + c * w_chunk_bytes, w_src + static_cast<size_t>(c) * w_chunk_bytes, w_chunk_bytes, wfull);
      }
      if (ph > 0 && !(a.dbg & 2)) {
        // producers of my input rows / readers of the rows I am about to overwrite (see the header)
        const int nslq = a.ph[ph - 1].nsl;
        const int ngq = static_cast<int>(gridDim.x) / nslq;
        int g_lo = 0, g_hi = ngq - 1;
        if (!a.dep_all) chain_dep_range(a.units_total, cta, ncta, ngq, &g_lo, &g_hi);
        chain_wait_flags(a.flags, g_lo * nslq, g_hi * nslq + nslq - 1, static_cast<unsigned int>(ph), lane);
      }
      __syncwarp();
      if (lane == 0) {
        CHAIN_TRACE(ph, 3);
        fence_proxy_async_all();
        const uint32_t tx_bytes = static_cast<uint32_t>(a.a_box_bytes);
        const CUtensorMap* const t0 = &a.tmaps[P.tm0];
        const CUtensorMap* const t1 = &a.tmaps[P.tm1];
        const uint32_t csrc = P.chunk_src;
        int cc0[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) cc0[c] = static_cast<int>(P.chunk_c0_8[c]) * 8;
        int slot = 0, ip = 0, t = 0;
        uint8_t* st = stage0;
        uint32_t O0 = 0;
        ChainWalk sw(a.units_total, img_h, a.x_tiles, cta, ncta);
        while (sw.next()) {
          const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, img_h - 1);
          for (int r = r0; r <= r1; ++r, ++ip) {
            if (ip >= DR) mbar_wait(&blk_full[row_rec[slot] & 0xFFu], row_rec[slot] >> 8);
            // this row's tiles are free again when the output block it completes (row r-1: sequence O0 + k) is
            const uint32_t X = pos(O0 + static_cast<uint32_t>(r - r0), nblk);
            row_rec[slot] = X | (((pcnt >> X) & 1u) << 8);
            pcnt ^= 1u << X;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < nch) {
                if (a.dbg & 8) {  // timing experiment: no activation loads (results are garbage)
                  mbar_arrive(&full_bar[t]);
                } else {
                  mbar_arrive_expect_tx(&full_bar[t], tx_bytes);
                  tma_load_4d_hint(st, ((csrc >> c) & 1u) ? t1 : t0, &full_bar[t], cc0[c], sw.x0 - 1, r, sw.img,
                                   kL2EvictLast);  // bf16 activations are re-read by the next convs: keep in L2
                }
                st += a_bytes;
                ++t;
              }
            }
            if (++slot == DR) { slot = 0; t = 0; st = stage0; }
          }
          // the two blocks below the last row of the segment are uses of the ring as well
          const uint32_t ni = static_cast<uint32_t>(r1 - r0 + 1);
          pcnt ^= 1u << pos(O0 + ni, nblk);
          pcnt ^= 1u << pos(O0 + ni + 1, nblk);
          O0 += ni + 2;
        }
        prev_out = ip < DR ? ip : DR;
        CHAIN_TRACE(ph, 4);
      }
    }
    if (lane == 0) {  // leave no asynchronous arrival pending at exit
      mbar_wait(wfree, static_cast<uint32_t>(nph - 1) & 1u);
      for (int s = 0; s < prev_out; ++s) mbar_wait(&blk_full[row_rec[s] & 0xFFu], row_rec[s] >> 8);
    }
  } else if (warp > kRowEpiWarps) {
    // ============================ MMA issuers (warps 13 .. 13 + kChainIssuers - 1) ============================
    // (The loop is warp-converged with one elected lane issuing: descriptors then stay on the uniform datapath.  A
    // `lane == 0` thread loop was measured 35 % slower per MMA.)
    const uint32_t mw = static_cast<uint32_t>(warp - (kRowEpiWarps + 1));  // == kx: the tap column this thread issues
    const uint32_t w_lo0 = umma_desc_lo(smem_u32(w_res));
    constexpr uint32_t ID_FULL = umma_idesc_bf16_m128(3 * BN);
    const uint32_t idesc_aux = umma_idesc_bf16_m128(BN);
    uint32_t fmask = 0;  // per tile: parity of its next full_bar phase
    uint32_t ucnt = 0;   // per block: use count & 1 (phase 0 of blk_empty is complete from the start, the release of use c completes phase c + 1)
    for (int ph = 0; ph < nph; ++ph) {
      const ChainPhaseC& P = a.ph[ph];
      const int nsl = P.nsl;
      const int cta = static_cast<int>(blockIdx.x) / nsl, ncta = static_cast<int>(gridDim.x) / nsl;
      const int nch = P.num_chunks, naux = P.aux_chunks;
      const uint32_t nblk = naux > 0 ? NBLK_MAX / 2 : NBLK_MAX;
      const bool last_half = (P.flags & kChainFLastHalf) != 0;  // the last chunk carries weights in its first half only
      const uint32_t w_block_bytes = static_cast<uint32_t>((naux > 0 ? 4 : 3) * BN * RB);
      const uint32_t w_chunk_bytes = 3u * w_block_bytes;
      const uint32_t w_step = w_chunk_bytes >> 4, w_block_desc = w_block_bytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(w_res + static_cast<uint32_t>(nch) * w_chunk_bytes));
      const uint32_t chunk_step = static_cast<uint32_t>(a.a_stage_bytes) >> 4;
      const int DR = P.stages;
      mbar_wait(wfull, static_cast<uint32_t>(ph) & 1u);
      if (mw > 0 && !(a.dbg & 8192)) {
        // The pipe serves the three threads evenly, so threads released together reach every row boundary together and
        // their per-row bookkeeping leaves the pipe idle.  Start thread kx a third / two thirds of a row late: a thread
        // that is ahead stays ahead, and its bookkeeping hides behind the others' MMAs (worth 3-4 %).
        const int mmas = nch * 3 * KS - (last_half ? 3 * KS / 2 : 0);
        const long long lag = static_cast<long long>(mw) * (mmas * 52 + 700) / 3;
        const long long t0 = clock64();
        while (clock64() - t0 < lag) {}
      }
      bool first_row = true;
      int slot = 0, t = 0;
      uint32_t O0 = 0;
      // Before the first MMA into block X: its previous occupant has been read + zeroed (the blk_empty phase its previous
      // release completed; phase 0 is complete from the start), and so has whatever last lived in the same columns under the other
      // ring layout (block X ^ 8: the conv1x1 blocks of the 8-block ring are the columns of blocks 8..15).
      int fresh = NBLK_MAX;  // touches of this phase that may still meet a block of the previous phase's layout
      auto touch = [&](uint32_t X) {
        mbar_wait(&blk_empty[X], (ucnt >> X) & 1u);
        if (fresh > 0) {
          mbar_wait(&blk_empty[X ^ 8u], (ucnt >> (X ^ 8u)) & 1u);
          --fresh;
        }
        ucnt ^= 1u << X;
      };
      ChainWalk sw(a.units_total, img_h, a.x_tiles, cta, ncta);
      while (sw.next()) {
        const int ni = min(sw.yb, img_h - 1) - max(sw.ya - 1, 0) + 1;
        for (int k = 0; k < ni; ++k) {
          const uint32_t On = O0 + k + 2;  // output row r+1 (ky = 0): always new
          if (k == 0) {
            touch(pos(O0, nblk));
            touch(pos(O0 + 1, nblk));
          }
          touch(pos(On, nblk));
          mbar_wait(&full_bar[t], (fmask >> t) & 1u);  // first tile of the row (the others: in the issue loop)
          tcgen05_fence_after();
          if (first_row && lane == 0) { CHAIN_TRACE(ph, 7 + static_cast<int>(mw)); first_row = false; }
          if (elect_one()) {
            // accumulator = blocks pos(On), +1, +2; split in two MMAs where it straddles the end of the ring
            const uint32_t Pb = pos(On, nblk);
            const uint32_t nA = (Pb + 3 <= nblk) ? 3u * BN : (nblk - Pb) * BN;  // columns before the wrap
            const uint32_t nB = 3u * BN - nA;
            const uint32_t dA = tmem_base + Pb * BN, dB = tmem_base;
            const uint32_t bB = (nA * RB) >> 4;  // B rows of the second part
            const uint32_t idA = nB ? umma_idesc_bf16_m128(nA) : ID_FULL, idB = umma_idesc_bf16_m128(nB ? nB : 16u);
            const uint32_t d_aux = tmem_base + AUX_COL0 + pos(O0 + k + 1, nblk) * BN;  // conv1x1 of output row r
            uint32_t bl = w_lo0;
            for (int c = 0; c < nch; ++c, bl += w_step) {
              if (c > 0) mbar_wait(&full_bar[t + c], (fmask >> (t + c)) & 1u);
              const uint32_t al = a_lo0 + static_cast<uint32_t>(t + c) * chunk_step;
              const bool half = last_half && c == nch - 1;
              const uint32_t a_kx = al + ((mw * RB) >> 4);   // the 130-pixel row shifted by kx pixels
              const uint32_t b_kx = bl + mw * w_block_desc;  // weights of tap column kx
              chain_issue_chunk_any<KC>(nB != 0, half, dA, dB, idA, idB, bB, a_kx, b_kx, DESC_HI);
              if (c < naux && mw == 1) {  // conv1x1: centre pixel (kx = 1) x B rows [3*BN, 4*BN) of the kx = 1 block
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                  umma_f16_ss2(d_aux, a_kx + ((ks * 32) >> 4), DESC_HI, b_kx + ((3 * BN * RB + ks * 32) >> 4), DESC_HI,
                               idesc_aux, 1u);
              }
            }
            // ONE commit per thread and row (a commit covers the MMAs of the committing thread only): block y collects
            // rows y-1, y, y+1, each issued by all threads, so after its share of row r a thread's contributions to
            // output row r-1 are complete.  The same barrier frees the tiles of row r.
            umma_commit(&blk_full[pos(O0 + k, nblk)]);
            if (k == ni - 1) {  // last input row of the segment: nothing follows for the last two blocks
              umma_commit(&blk_full[pos(O0 + k + 1, nblk)]);
              umma_commit(&blk_full[pos(O0 + k + 2, nblk)]);
            }
          }
          __syncwarp();
          for (int c = 0; c < nch; ++c) fmask ^= 1u << (t + c);
          t += nch;
          if (++slot == DR) { slot = 0; t = 0; }
        }
        O0 += static_cast<uint32_t>(ni + 2);
      }
      if (lane == 0) { CHAIN_TRACE(ph, 9 + static_cast<int>(mw)); mbar_arrive(wfree); }
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
    uint32_t fcnt = 0;  // per block: use count & 1 == parity of the blk_full phase of its current use
    int turn = 0;
    grid_dep_wait();  // residual reads / output writes must not race with the previous kernel
    for (int ph = 0; ph < nph; ++ph) {
      // ---- per-phase fields: constant bank -> registers, once ----
      const ChainPhaseC& P = a.ph[ph];
      const int nsl = P.nsl;
      const int sl = nsl > 1 ? static_cast<int>(blockIdx.x) % nsl : 0;
      const int cta = static_cast<int>(blockIdx.x) / nsl, ncta = static_cast<int>(gridDim.x) / nsl;
      const int csh = sl * BN;  // channel shift of every global channel offset
      const uint32_t fl = P.flags;
      const bool aux = P.aux_chunks > 0;
      const uint32_t nblk = aux ? NBLK_MAX / 2 : NBLK_MAX;
      const bool act = (fl & kChainFAct) != 0, planar = (fl & kChainFPlanar) != 0;
      const bool r1_f32 = (fl & kChainFR1F32) != 0, r2_f32 = (fl & kChainFR2F32) != 0;
      const bool quad = !(fl & kChainFNoQuad);
      const float s0 = P.s0, s2 = P.s2;
      const uint8_t* const r1p = P.r1_off16 != kChainNull ? a.act_base + static_cast<size_t>(P.r1_off16) * 16 : nullptr;
      const uint8_t* const r2p = P.r2_off16 != kChainNull ? a.act_base + static_cast<size_t>(P.r2_off16) * 16 : nullptr;
      __nv_bfloat16* const obp = P.ob_off16 != kChainNull ? reinterpret_cast<__nv_bfloat16*>(a.act_base + static_cast<size_t>(P.ob_off16) * 16) : nullptr;
      float* const ofp = P.of_off16 != kChainNull ? reinterpret_cast<float*>(a.act_base + static_cast<size_t>(P.of_off16) * 16) : nullptr;
      const int r1_ct = P.r1_ctotal, r1_c0 = P.r1_c0 + csh, r2_ct = P.r2_ctotal, r2_c0 = P.r2_c0 + csh;
      const int ob_ct = P.ob_ctotal, ob_c0 = P.ob_c0 + csh, of_ct = P.of_ctotal, of_c0 = P.of_c0 + csh;
      __syncwarp();
      bias_w[lane] = reinterpret_cast<const float*>(a.w_base + (static_cast<size_t>(P.bias_off16) + static_cast<size_t>(sl) * P.sl_stride16) * 16)[lane];
      __syncwarp();
      uint32_t O0 = 0;
      ChainWalk sw(a.units_total, img_h, a.x_tiles, cta, ncta);
      while (sw.next()) {
        const int r0 = max(sw.ya - 1, 0), r1 = min(sw.yb, img_h - 1);
        const int xs = sw.x0 + xl;
        const bool col_ok = xs < img_w;
        const int xc = col_ok ? xs : 0;
        const int no = r1 - r0 + 3;  // output sequence of this segment: rows r0-1 .. r1+1 (the ends are dummies)
#pragma unroll 1
        for (int j = 0; j < no; ++j) {
          const uint32_t X = pos(O0 + static_cast<uint32_t>(j), nblk);
          const uint32_t par = (fcnt >> X) & 1u;
          fcnt ^= 1u << X;  // every warpgroup counts every use
          const bool mine = turn == wg;
          if (++turn == kRowWGs) turn = 0;
          if (!mine) continue;
          const int y = r0 - 1 + j;
          const bool real = y >= sw.ya && y < sw.yb;
          const bool store = real && col_ok;
          const uint32_t blk = lane_addr + X * BN;
          const size_t rowid = static_cast<size_t>(sw.img) * img_h + (real ? y : sw.ya);
          const size_t pix = rowid * img_w + xc;
          const size_t f4_step = planar ? static_cast<size_t>(img_w) : 1;
          // element offset of channel c (multiple of 4) in an fp32 operand of ct channels: NHWC or [n][h][c/4][w][4]
          auto off32 = [&](int ct, int c) -> size_t {
            return planar ? ((rowid * (ct >> 2) + (c >> 2)) * img_w + xc) * 4 : pix * ct + c;
          };
          // residuals of the first round: in flight while we wait for the accumulator
          float r1v[GC], r2v[GC];
          if (real) {  // every lane (the shuffles of the bf16 store need the whole warp): pix is clamped for columns >= w
            if (r1p) load_residual_coherent<GC>(r1p, r1_f32, r1_f32 ? off32(r1_ct, r1_c0) : pix * r1_ct + r1_c0, r1v, f4_step);
            if (r2p) load_residual_coherent<GC>(r2p, r2_f32, r2_f32 ? off32(r2_ct, r2_c0) : pix * r2_ct + r2_c0, r2v, f4_step);
          }
          mbar_wait(&blk_full[X], par);
          tcgen05_fence_after();
          if (!real || (a.dbg & 256)) {  // dummy row at a segment end (or timing experiment 256: no epilogue): just recycle the block
#pragma unroll
            for (int c = 0; c < BN; c += GC) tmem_st_zero_x16(blk + c);
            if (aux) {  // halo rows collect a conv1x1 product as well
#pragma unroll
              for (int c = 0; c < BN; c += GC) tmem_st_zero_x16(blk + AUX_COL0 + c);
            }
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&blk_empty[X]);
            continue;
          }
          uint32_t pend[8];  // bf16 output of an even round, stored together with the following odd round
#pragma unroll
          for (int g = 0; g < ROUNDS; ++g) {
            const int ch0 = g * GC;
            uint32_t acc[GC], ax[GC];
            tmem_ld_x16(blk + ch0, acc);
            if (aux) tmem_ld_x16(blk + AUX_COL0 + ch0, ax);
            tmem_ld_wait();
            tmem_st_zero_x16(blk + ch0);
            if (aux) tmem_st_zero_x16(blk + AUX_COL0 + ch0);
            if (g == ROUNDS - 1) {
              tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&blk_empty[X]);
            }
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
            if (act) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] = fmaxf(v[i], 0.2f * v[i]);  // LeakyReLU(0.2)
            }
#pragma unroll
            for (int i = 0; i < GC; ++i) v[i] *= s0;
            if (aux) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] += __uint_as_float(ax[i]);
            }
            if (r1p) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] += r1v[i];
              if (g + 1 < ROUNDS)
                load_residual_coherent<GC>(r1p, r1_f32, r1_f32 ? off32(r1_ct, r1_c0 + ch0 + GC) : pix * r1_ct + r1_c0 + ch0 + GC, r1v, f4_step);
            }
            if (r2p) {
#pragma unroll
              for (int i = 0; i < GC; ++i) v[i] = fmaf(s2, v[i], r2v[i]);
              if (g + 1 < ROUNDS)
                load_residual_coherent<GC>(r2p, r2_f32, r2_f32 ? off32(r2_ct, r2_c0 + ch0 + GC) : pix * r2_ct + r2_c0 + ch0 + GC, r2v, f4_step);
            }
            if (obp) {
              uint32_t pk[GC / 2];
#pragma unroll
              for (int i = 0; i < GC / 2; ++i) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                pk[i] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              if (!quad) {
                if (store) {
                  uint4* op = reinterpret_cast<uint4*>(obp + pix * ob_ct + ob_c0 + ch0);
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
                __nv_bfloat16* ob = obp + (rowid * img_w + xg) * ob_ct + ob_c0 + (ch0 - GC) + (lane & 3) * 8;
#pragma unroll
                for (int m = 0; m < 4; ++m)
                  if (xg + m < img_w)
                    st_global_u4_hint(reinterpret_cast<uint4*>(ob + static_cast<size_t>(m) * ob_ct),
                                      make_uint4(e[m][0], e[m][1], e[m][2], e[m][3]), kL2EvictLast);
              }
            }
            if (ofp && store) {
              float4* op = reinterpret_cast<float4*>(ofp + off32(of_ct, of_c0 + ch0));
#pragma unroll
              for (int i = 0; i < GC / 4; ++i)  // fp32 trunk: read once, a whole dense block later -> stream
                st_global_f4_hint(op + i * f4_step, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), kL2EvictFirst);
            }
          }
        }
        O0 += static_cast<uint32_t>(no);
      }
      // all outputs of this phase written: publish (the neighbours' producers acquire the flag before their TMA loads)
      if (!(a.dbg & 4)) __threadfence();
      named_bar_sync(1, kRowEpiWarps * 32);
      if (threadIdx.x == 0) {
        fence_proxy_async_all();
        st_release_gpu(a.flags + blockIdx.x, static_cast<unsigned int>(ph + 1));
        CHAIN_TRACE(ph, 12);
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
