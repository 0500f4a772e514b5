// conv3x3_tc.cuh — fused 3x3 convolution for sm_100a: TMA-staged NHWC halo tiles in shared memory
// feeding tcgen05.mma implicit-GEMM tiles with fp32 accumulators in tensor memory, and the
// dense-block elementwise tail (bias, LeakyReLU, conv1x1 add, residual scale/add, noise, RRDB
// residual) fused into the TMEM->register epilogue.
//
// Replaces (reference): block.py:260-268 (ResidualDenseBlock_5C.forward: conv_k(torch.cat(...)),
// lrelu, + conv1x1, + x2, *0.2 + x, noise), block.py:287-291 (RRDB residual) and the plain
// conv_blocks of architecture.py:55-71.  torch.cat never materialises: a conv reads its K
// dimension as a list of channel chunks taken from up to two NHWC tensors.
//
// Decomposition
//   GEMM view: D[pixel, cout] = sum over (chunk, tap, k) A[pixel+tap, chunk*KC+k] * W[cout, ...]
//   CTA tile : 16 rows x (8*MT) cols of output pixels = MT UMMA M-tiles of 128 pixels
//              (M-tile = 16 rows x 8 cols, so every 8-row core-matrix group is 8 consecutive
//              pixels of one image row: contiguous KC*2-byte rows in the halo tile).
//   Stage    : one K-chunk (KC channels) of the (16+2) x (8*MT+2) halo tile, loaded ONCE by TMA
//              (zero fill outside the image = the conv's zero padding); all 9 taps are shifted
//              UMMA descriptors into that same tile, so L2->SMEM traffic is 1.27x the ideal
//              instead of 9x.  (ESRP_VARIANT_ALIGNED loads one box per kx so every operand is
//              swizzle-atom aligned; kept as a cross-check of the shifted-descriptor scheme.)
//   Weights  : pre-swizzled bf16 [chunk][tap][BN][KC]; resident in SMEM for the whole persistent
//              CTA when they fit, else streamed with the chunk.
//   Warps    : 0-3 epilogue (TMEM lane quarter == warp id), 4 TMA producer, 5 MMA issuer/TMEM owner
//   TMEM     : 2 accumulator buffers x MT x (BN [+BN aux]) fp32 columns -> epilogue of tile i
//              overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"
#include "conv_params.h"
#include "esrp_philox.cuh"
#include "esrp_ptx.cuh"

namespace esrp {

constexpr int kTileH = 16;
constexpr int kNumEpiWarps = 4;
constexpr int kConvThreads = 32 * (kNumEpiWarps + 2);

template <int KC, int MT, bool HALO>
struct ConvGeom {
  static constexpr int RB = KC * 2;                     // bytes per pixel row of a chunk
  static constexpr int TW = 8 * MT;                     // tile width in pixels
  static constexpr int HW = HALO ? (TW + 2) : TW;       // smem tile width in pixels
  static constexpr int HH = kTileH + 2;                 // smem tile height in pixels
  static constexpr int SUB_BYTES = HH * HW * RB;        // one TMA box
  static constexpr int A_BYTES_RAW = HALO ? SUB_BYTES : 3 * SUB_BYTES;
  static constexpr int A_BYTES = (A_BYTES_RAW + 1023) / 1024 * 1024;
  static constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;  // 128 B / 64 B swizzle
  static constexpr uint32_t SBO_A = HW * RB;            // next 8-pixel group = next image row
  static constexpr uint32_t SBO_B = 8 * RB;             // weights: dense rows
  static_assert(HALO || (SUB_BYTES % 1024 == 0), "aligned variant needs atom-aligned boxes");
};

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

template <int KC, int BN, int MT, bool HALO>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                  const __grid_constant__ ConvKParams p) {
  using G = ConvGeom<KC, MT, HALO>;
  constexpr int RB = G::RB;
  constexpr int W_CHUNK_BYTES = 9 * BN * RB;  // all taps of one chunk
  constexpr int W_AUX_BYTES = BN * RB;        // 1x1 weights of one chunk
  constexpr uint32_t IDESC = umma_idesc_bf16_m128(BN);

  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned carve-up: [barriers 1 KB][resident weights][stage 0][stage 1]...
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);  // [stages]   (<= 16)
  uint64_t* empty_bar = full_bar + 16;                     // [stages]
  uint64_t* tmem_full = empty_bar + 16;                    // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint64_t* wfull = tmem_empty + 2;                        // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wfull + 1);

  const bool has_aux = p.aux_chunks > 0;
  const int bnt = has_aux ? 2 * BN : BN;  // TMEM columns per M-tile accumulator
  const int w_res_bytes =
      p.w_resident ? (p.num_chunks * W_CHUNK_BYTES + p.aux_chunks * W_AUX_BYTES) : 0;
  uint8_t* w_res = smem + 1024;
  uint8_t* stage0 = w_res + w_res_bytes;
  const int stage_bytes = G::A_BYTES + (p.w_resident ? 0 : (W_CHUNK_BYTES + (has_aux ? W_AUX_BYTES : 0)));
  const int S = p.stages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm0);
    tma_prefetch_desc(&tm1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 32 * kNumEpiWarps);
    mbar_init(&tmem_empty[1], 32 * kNumEpiWarps);
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_holder, p.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 4) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      if (p.w_resident) {
        mbar_arrive_expect_tx(wfull, static_cast<uint32_t>(w_res_bytes));
        for (int c = 0; c < p.num_chunks; ++c)
          bulk_load_1d(w_res + c * W_CHUNK_BYTES, p.w_packed + static_cast<size_t>(c) * W_CHUNK_BYTES,
                       W_CHUNK_BYTES, wfull);
        for (int c = 0; c < p.aux_chunks; ++c)
          bulk_load_1d(w_res + p.num_chunks * W_CHUNK_BYTES + c * W_AUX_BYTES,
                       p.w_aux + static_cast<size_t>(c) * W_AUX_BYTES, W_AUX_BYTES, wfull);
      }
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int img = tile / tiles_per_img;
        const int rem = tile - img * tiles_per_img;
        const int ty = rem / p.tiles_x;
        const int tx = rem - ty * p.tiles_x;
        const int y0 = ty * kTileH - 1;
        const int x0 = tx * G::TW - 1;
        for (int c = 0; c < p.num_chunks; ++c, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
          uint32_t tx_bytes = G::A_BYTES_RAW;
          if (!p.w_resident) tx_bytes += W_CHUNK_BYTES + ((c < p.aux_chunks) ? W_AUX_BYTES : 0);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          const CUtensorMap* tm = p.chunk_src[c] ? &tm1 : &tm0;
          if (HALO) {
            tma_load_4d(st, tm, &full_bar[s], p.chunk_c0[c], x0, y0, img);
          } else {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              tma_load_4d(st + kx * G::SUB_BYTES, tm, &full_bar[s], p.chunk_c0[c], x0 + kx, y0, img);
          }
          if (!p.w_resident) {
            bulk_load_1d(st + G::A_BYTES, p.w_packed + static_cast<size_t>(c) * W_CHUNK_BYTES,
                         W_CHUNK_BYTES, &full_bar[s]);
            if (c < p.aux_chunks)
              bulk_load_1d(st + G::A_BYTES + W_CHUNK_BYTES,
                           p.w_aux + static_cast<size_t>(c) * W_AUX_BYTES, W_AUX_BYTES, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ====================================== MMA issuer ======================================
    // The whole warp walks the pipeline (converged), one elected lane issues: this keeps the
    // tcgen05.mma sequence free of per-instruction divergence handling (2 UIADD3 + UTCHMMA each).
    {
      constexpr uint32_t A_HI = (G::SBO_A >> 4) | (1u << 14) | (G::LAYOUT << 29);
      constexpr uint32_t B_HI = (G::SBO_B >> 4) | (1u << 14) | (G::LAYOUT << 29);
      if (p.w_resident) mbar_wait(wfull, 0);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const uint32_t ab = tl & 1, abph = (tl >> 1) & 1;
        mbar_wait(&tmem_empty[ab], abph ^ 1);
        tcgen05_fence_after();
        for (int c = 0; c < p.num_chunks; ++c, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
          const uint32_t a_lo = umma_desc_lo(smem_u32(st));
          const uint32_t b_lo = umma_desc_lo(
              p.w_resident ? smem_u32(w_res + c * W_CHUNK_BYTES) : smem_u32(st + G::A_BYTES));
          const uint32_t baux_lo = umma_desc_lo(
              p.w_resident ? smem_u32(w_res + p.num_chunks * W_CHUNK_BYTES + c * W_AUX_BYTES)
                           : smem_u32(st + G::A_BYTES + W_CHUNK_BYTES));
          const bool last = (c == p.num_chunks - 1);
          if (elect_one()) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              const uint32_t d_tmem = tmem_base + (ab * MT + m) * bnt;
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                const int ky = t / 3, kx = t % 3;
                // byte offset of the first pixel of this M-tile for this tap inside the smem tile
                const uint32_t a_off = HALO ? (ky * G::HW + m * 8 + kx) * RB
                                            : kx * G::SUB_BYTES + (ky * G::HW + m * 8) * RB;
                const uint32_t b_off = t * BN * RB;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  umma_f16_ss2(d_tmem, a_lo + ((a_off + ks * 32) >> 4), A_HI, b_lo + ((b_off + ks * 32) >> 4),
                               B_HI, IDESC, (c | t | ks) != 0 ? 1u : 0u);
                }
              }
              if (c < p.aux_chunks) {
                const uint32_t a_off = HALO ? (1 * G::HW + m * 8 + 1) * RB
                                            : 1 * G::SUB_BYTES + (1 * G::HW + m * 8) * RB;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  umma_f16_ss2(d_tmem + BN, a_lo + ((a_off + ks * 32) >> 4), A_HI, baux_lo + ((ks * 32) >> 4),
                               B_HI, IDESC, (c | ks) != 0 ? 1u : 0u);
                }
              }
            }
            umma_commit(&empty_bar[s]);              // smem slot free once these MMAs have read it
            if (last) umma_commit(&tmem_full[ab]);   // accumulators complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================================= epilogue =======================================
    uint32_t tl = 0;
    const int row = threadIdx.x;  // 0..127 == TMEM lane == M row
    const int ry = row >> 3, rx = row & 7;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const uint32_t ab = tl & 1, abph = (tl >> 1) & 1;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      mbar_wait(&tmem_full[ab], abph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int py = ty * kTileH + ry;
        const int px = tx * G::TW + m * 8 + rx;
        const bool valid = (py < p.h) && (px < p.w);
        const size_t pix = (static_cast<size_t>(img) * p.h + py) * p.w + px;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + (ab * MT + m) * bnt;
#pragma unroll 1
        for (int g = 0; g < BN / 16; ++g) {
          uint32_t acc[16];
          uint32_t aux[16];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated tail
          tmem_ld_x16(taddr + g * 16, acc);
          if (has_aux) tmem_ld_x16(taddr + BN + g * 16, aux);
          tmem_ld_wait();
          if (!valid) continue;
          const int ch0 = g * 16;
          if (ch0 >= p.cout) continue;
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + i));
              v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
          }
          if (p.act) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = lrelu02(v[i]);
          }
          if (p.s0 != 1.0f) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= p.s0;
          }
          if (has_aux) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(aux[i]);
          }
          if (p.r1) {
            float r[16];
            if (p.r1_is_f32) {
              const float4* rp = reinterpret_cast<const float4*>(
                  static_cast<const float*>(p.r1) + pix * p.r1_ctotal + p.r1_c0 + ch0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 t = rp[i];
                r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
              }
            } else {
              const uint4* rp = reinterpret_cast<const uint4*>(
                  static_cast<const __nv_bfloat16*>(p.r1) + pix * p.r1_ctotal + p.r1_c0 + ch0);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint4 t = rp[i];
                const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  r[8 * i + 2 * j] = __uint_as_float(u[j] << 16);
                  r[8 * i + 2 * j + 1] = __uint_as_float(u[j] & 0xFFFF0000u);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(p.s1, r[i], v[i]);
          }
          if (p.noise) {
            // y = t + N(0,1) * sigma * t  (block.py:117-121); one Philox counter per 4 channels.
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float z[4];
              philox_normal4(p.seed, p.offset + (pix * static_cast<unsigned long long>(p.cout) + ch0 + i) / 4, z);
#pragma unroll
              for (int j = 0; j < 4; ++j) v[i + j] = fmaf(z[j] * p.sigma, v[i + j], v[i + j]);
            }
          }
          if (p.r2) {
            float r[16];
            if (p.r2_is_f32) {
              const float4* rp = reinterpret_cast<const float4*>(
                  static_cast<const float*>(p.r2) + pix * p.r2_ctotal + p.r2_c0 + ch0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 t = rp[i];
                r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
              }
            } else {
              const uint4* rp = reinterpret_cast<const uint4*>(
                  static_cast<const __nv_bfloat16*>(p.r2) + pix * p.r2_ctotal + p.r2_c0 + ch0);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint4 t = rp[i];
                const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  r[8 * i + 2 * j] = __uint_as_float(u[j] << 16);
                  r[8 * i + 2 * j + 1] = __uint_as_float(u[j] & 0xFFFF0000u);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(p.s2, v[i], r[i]);
          }
          if (p.out_bf16) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              pk[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.ob_ctotal + p.ob_c0 + ch0);
            op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(p.out_f32 + pix * p.of_ctotal + p.of_c0 + ch0);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (p.out_nchw) {
            const size_t plane = static_cast<size_t>(p.h) * p.w;
            float* op = p.out_nchw + (static_cast<size_t>(img) * p.cout) * plane + static_cast<size_t>(py) * p.w + px;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ch0 + i < p.cout) op[static_cast<size_t>(ch0 + i) * plane] = v[i];
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[ab]);
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
