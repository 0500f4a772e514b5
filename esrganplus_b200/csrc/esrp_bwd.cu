// esrp_bwd.cu — backward-pass kernels of the ESRGAN+ hot path (the autograd backward the reference gets
// from PyTorch for block.py:260-268 / architecture.py:47-78, driven by SRRaGAN_model.py:140,167):
//
//   * esrp_pack_dgrad_weights : data-gradient operator of one or several convs as ONE packed conv
//     (the gradient of a dense block w.r.t. one of its feature slices is a conv whose K dimension is the
//     concatenation of the later convs' output gradients — a dense block again — so it runs on the same
//     tcgen05 kernels as the forward pass);
//   * conv3x3_wgrad_kernel    : weight gradient dW[co][ci][ky][kx] = sum_px dY[px][co] * X[px + tap][ci]
//     as warp-level bf16 MMAs (mma.sync m16n8k16, fp32 accumulate) straight from the NHWC tensors:
//     the contraction index is the PIXEL, which is the slow index of both operands, so both fragments are
//     fetched with ldmatrix.trans from padded (bank-conflict-free) shared-memory tiles filled by cp.async;
//     split-K over pixel tiles across CTAs, vectorised fp32 reductions into per-unit accumulator blocks;
//     the bias gradient (column sums of dY) rides along;
//   * wgrad_scatter_kernel    : accumulator blocks -> OIHW fp32 gradient tensors (the reference's layout);
//   * conv1x1_bwd_kernel      : both gradients of the bias-free 1x1 conv of block.py:244,263;
//   * upsample2x_bwd_kernel   : nn.Upsample(nearest, x2) backward (2x2 sum) fused with the LeakyReLU mask.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/esrp.h"
#include "esrp_bwd.h"
#include "esrp_host.h"
#include "esrp_ptx.cuh"

namespace esrp {

// ------------------------------------------------------------------------------------------------
// dgrad weight pack
// ------------------------------------------------------------------------------------------------
struct DgradTable {
  esrp_dgrad_group_t g[2 * ESRP_MAX_CHUNKS];
};

__device__ __forceinline__ int swz_elem(int row, int k, int kc) {
  const int chunk16 = k >> 3;
  const int x = (kc == 64) ? (row & 7) : ((row >> 1) & 3);
  return row * kc + (((chunk16 ^ x) << 3) | (k & 7));
}

// out[chunk][outer][row = blk*bn + r][kc]; K channel (chunk, k) belongs to group (chunk*kc + k) / 32 whose
// source conv weight W[w_o, w_i, 3, 3] contributes  scale * W[co0 + k%32][row0 + r][2-ky][2-kx].
__global__ void pack_dgrad_kernel(DgradTable tab, int num_groups, int layout, int row0, int rows, int kc, int bn,
                                  int num_chunks, __nv_bfloat16* __restrict__ out) {
  const int nb_rows = 3 * bn;
  const int total = num_chunks * 3 * nb_rows * kc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % kc;
    const int row = (i / kc) % nb_rows;
    const int outer = (i / (kc * nb_rows)) % 3;
    const int chunk = i / (kc * nb_rows * 3);
    const int blk = row / bn;
    const int r = row - blk * bn;
    const int ky = layout == ESRP_LAYOUT_ROW ? blk : outer;
    const int kx = layout == ESRP_LAYOUT_ROW ? outer : blk;
    const int kk = chunk * kc + k;
    const int gi = kk >> 5;
    float v = 0.f;
    if (r < rows && gi < num_groups) {
      const esrp_dgrad_group_t& g = tab.g[gi];
      const int co = g.co0 + (kk & 31);
      const int ci = row0 + r;
      if (g.w != nullptr && co < g.w_o && ci < g.w_i)
        v = g.scale * g.w[((static_cast<size_t>(co) * g.w_i + ci) * 3 + (2 - ky)) * 3 + (2 - kx)];
    }
    out[static_cast<size_t>(chunk * 3 + outer) * nb_rows * kc + swz_elem(row, k, kc)] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 256;
constexpr int kWgTilePx = 128;            // pixels (MMA K) per tile
constexpr int kWgXPitch = 80;             // bytes per pixel row of the X tile (32 ch bf16 + 16 pad): 5 x 16 B, odd -> no bank conflicts
constexpr int kWgYPitch = 144;            // bytes per pixel row of the dY tile (64 ch bf16 + 16 pad)

struct WgradParams {
  esrp_wgrad_unit_t u[kMmaMaxUnits];
  int num_units, splits;
  int n, h, w;
  int tw, tw_log2, tr;       // tile = tr rows x tw columns, tr * tw == kWgTilePx
  int tiles_x, tiles_y;
  int tiles_total;
  int x_tile_px;             // (tr + 2) * (tw + 2)
  int stage_bytes;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One CTA = (unit, split): accumulates dW[tap][co 0..63][ci 0..31] of its unit over its share of the pixel
// tiles in registers (warp w: co rows 16*(w&3).., ci columns 16*(w>>2)..; 9 taps x 2 n-tiles x 4 = 72 fp32),
// then adds them to the unit's accumulator block with 8-byte vector reductions.
__global__ void __launch_bounds__(kWgThreads, 2) conv3x3_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(128) uint8_t wsm[];
  const int unit = blockIdx.x / p.splits;
  const int split = blockIdx.x - unit * p.splits;
  const esrp_wgrad_unit_t& U = p.u[unit];
  const int t_begin = static_cast<int>(static_cast<long long>(p.tiles_total) * split / p.splits);
  const int t_end = static_cast<int>(static_cast<long long>(p.tiles_total) * (split + 1) / p.splits);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mi = warp & 3, nh = warp >> 2;
  const int xw = p.tw + 2;  // halo tile width in pixels

  const __nv_bfloat16* xg = static_cast<const __nv_bfloat16*>(U.x);
  const __nv_bfloat16* yg = static_cast<const __nv_bfloat16*>(U.dy);
  // 16-byte segments of the dY row that exist in the tensor (a block may hang over the channel count)
  const int y_valid_segs = min(8, max(0, (U.dy_ctotal - U.dy_c0) / 8));

  auto load_tile = [&](int t, int stage) {
    uint8_t* xs = wsm + stage * p.stage_bytes;
    uint8_t* ys = xs + p.x_tile_px * kWgXPitch;
    const int tx = t % p.tiles_x;
    const int ty = (t / p.tiles_x) % p.tiles_y;
    const int img = t / (p.tiles_x * p.tiles_y);
    const int x0 = tx * p.tw, y0 = ty * p.tr;
    // X halo tile: (tr+2) x (tw+2) pixels x 4 segments
    const int xsegs = p.x_tile_px * 4;
    for (int i = tid; i < xsegs; i += kWgThreads) {
      const int seg = i & 3, px = i >> 2;
      const int ry = px / xw, rx = px - ry * xw;
      const int gy = y0 + ry - 1, gx = x0 + rx - 1;
      const bool ok = gy >= 0 && gy < p.h && gx >= 0 && gx < p.w;
      const size_t gpix = (static_cast<size_t>(img) * p.h + (ok ? gy : 0)) * p.w + (ok ? gx : 0);
      cp_async16_zfill(smem_u32(xs + px * kWgXPitch + seg * 16), xg + gpix * U.x_ctotal + U.x_c0 + seg * 8, ok);
    }
    // dY tile: tr x tw pixels x 8 segments
    for (int i = tid; i < kWgTilePx * 8; i += kWgThreads) {
      const int seg = i & 7, px = i >> 3;
      const int ry = px >> p.tw_log2, rx = px & (p.tw - 1);
      const int gy = y0 + ry, gx = x0 + rx;
      const bool ok = gy < p.h && gx < p.w && seg < y_valid_segs;
      const size_t gpix = (static_cast<size_t>(img) * p.h + (gy < p.h ? gy : 0)) * p.w + (gx < p.w ? gx : 0);
      cp_async16_zfill(smem_u32(ys + px * kWgYPitch + seg * 16), yg + gpix * U.dy_ctotal + U.dy_c0 + (ok ? seg : 0) * 8, ok);
    }
  };

  float acc[9][2][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][j][i] = 0.f;
  float bsum = 0.f;  // bias gradient: thread -> column tid & 63, pixels (tid >> 6) * 32 ..

  // per-lane ldmatrix row offsets
  // A (dY^T): matrix j = lane / 8: pixel (lane % 8) + (j / 2) * 8, channel 16 * mi + (j % 2) * 8
  const int a_px = (lane & 7) + ((lane >> 4) << 3);
  const int a_off = a_px * kWgYPitch + (mi * 16 + ((lane >> 3) & 1) * 8) * 2;
  // B (X): matrix j: pixel (lane % 8) + (j % 2) * 8, channel 16 * nh + (j / 2) * 8
  const int b_px = (lane & 7) + (((lane >> 3) & 1) << 3);
  const int b_ch_off = (nh * 16 + (lane >> 4) * 8) * 2;

  if (t_begin < t_end) {
    load_tile(t_begin, 0);
    cp_async_commit();
  }
  for (int t = t_begin; t < t_end; ++t) {
    const int stage = (t - t_begin) & 1;
    if (t + 1 < t_end) load_tile(t + 1, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint8_t* xs = wsm + stage * p.stage_bytes;
    const uint8_t* ys = xs + p.x_tile_px * kWgXPitch;
    const uint32_t xs_a = smem_u32(xs), ys_a = smem_u32(ys);
    if (U.bias_acc != nullptr) {
      const int col = tid & 63, part = tid >> 6;
      const __nv_bfloat16* yc = reinterpret_cast<const __nv_bfloat16*>(ys) + col;
#pragma unroll 8
      for (int px = part * 32; px < part * 32 + 32; ++px) bsum += __bfloat162float(yc[px * (kWgYPitch / 2)]);
    }
#pragma unroll 1
    for (int ks = 0; ks < kWgTilePx / 16; ++ks) {
      const int k0 = ks * 16;
      const int ry = k0 >> p.tw_log2, rx = k0 & (p.tw - 1);
      uint32_t a[4];
      ldmatrix_x4_trans(ys_a + k0 * kWgYPitch + a_off, a);
      const uint32_t xb = xs_a + ((ry * xw + rx + b_px) * kWgXPitch) + b_ch_off;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          uint32_t b[4];
          ldmatrix_x4_trans(xb + (ky * xw + kx) * kWgXPitch, b);
          mma_bf16_16816(acc[ky * 3 + kx][0], a, b[0], b[1]);
          mma_bf16_16816(acc[ky * 3 + kx][1], a, b[2], b[3]);
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();

  // accumulate: block layout [tap][co 64][ci 32]; thread holds (co = 16 mi + lane/4 (+8), ci = 16 nh + 8 j + 2 (lane%4) (+1))
  float* ab = U.acc;
  const int co = mi * 16 + (lane >> 2);
  const int ci = nh * 16 + (lane & 3) * 2;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float2* d0 = reinterpret_cast<float2*>(ab + (static_cast<size_t>(t) * 64 + co) * 32 + ci + j * 8);
      float2* d1 = reinterpret_cast<float2*>(ab + (static_cast<size_t>(t) * 64 + co + 8) * 32 + ci + j * 8);
      atomicAdd(d0, make_float2(acc[t][j][0], acc[t][j][1]));
      atomicAdd(d1, make_float2(acc[t][j][2], acc[t][j][3]));
    }
  }
  if (U.bias_acc != nullptr) {
    // 4 partial sums per column (one per 32-pixel quarter of the tile, in 4 different warps)
    atomicAdd(U.bias_acc + (tid & 63), bsum);
  }
}

// ------------------------------------------------------------------------------------------------
// scatter: accumulator blocks -> reference-layout gradient tensors
// ------------------------------------------------------------------------------------------------
// kind 0: conv weight: dst[(co0 + c) * w_i + ci0 + i][tap] = scale * acc[tap][col0 + c][i]   c < ncols, i < nci
// kind 1: flat copy:   dst[i] = scale * acc[i]                                              i < ncols
__global__ void __launch_bounds__(256) wgrad_scatter_kernel(const esrp_scatter_entry_t* __restrict__ tab, int num,
                                                            float* const* __restrict__ dst_ptrs) {
  // block = (entry, group of 8 accumulator columns): the [9][8][32] slice is staged through shared memory so that both
  // the reads (32 consecutive input channels) and the writes (runs of nci * 9 or nci * 16 floats per output channel)
  // are coalesced
  __shared__ float tile[9][8][33];
  const int e = blockIdx.x >> 3, cg = blockIdx.x & 7;
  if (e >= num) return;
  const esrp_scatter_entry_t s = tab[e];
  float* dst = dst_ptrs ? dst_ptrs[s.dst_index] : s.dst;
  if (dst == nullptr) return;
  if (s.kind == 1) {
    for (int i = cg * 256 + threadIdx.x; i < s.ncols; i += 8 * 256) dst[s.dst_off + i] = s.scale * s.acc[i];
    return;
  }
  const int c0 = cg * 8;
  if (c0 >= s.ncols) return;
  const int nc = min(8, s.ncols - c0);
  for (int i = threadIdx.x; i < 9 * 8 * 32; i += 256) {
    const int ci = i & 31, c = (i >> 5) & 7, tap = i >> 8;
    tile[tap][c][ci] = (c < nc && ci < s.nci) ? s.acc[(static_cast<size_t>(tap) * 64 + s.col0 + c0 + c) * 32 + ci] : 0.f;
  }
  __syncthreads();
  if (s.kind == 2) {
    // 4x4 / stride-2 conv evaluated as a 3x3 conv over the space-to-depth tensor (esrp_s2d_pad_nhwc_bf16):
    // unit channel ci0 + i = (a*2+b) * w_i + ci, tap (A+1, B+1) -> dW4[co][ci][2A+a][2B+b]; taps with A or B < 0 are zero
    const int ab = s.ci0 / s.w_i, cib = s.ci0 - ab * s.w_i;
    const int a = ab >> 1, b = ab & 1;
    for (int i = threadIdx.x; i < nc * s.nci * 4; i += 256) {
      const int q = i & 3, ci = (i >> 2) % s.nci, c = (i >> 2) / s.nci;
      const int A = q >> 1, B = q & 1;
      dst[s.dst_off + (static_cast<size_t>(s.co0 + c0 + c) * s.w_i + cib + ci) * 16 + (2 * A + a) * 4 + (2 * B + b)] =
          s.scale * tile[(A + 1) * 3 + (B + 1)][c][ci];
    }
    return;
  }
  const int run = s.nci * 9;
  for (int i = threadIdx.x; i < nc * run; i += 256) {
    const int c = i / run, r = i - c * run;
    const int ci = r / 9, tap = r - ci * 9;
    dst[s.dst_off + (static_cast<size_t>(s.co0 + c0 + c) * s.w_i + s.ci0) * 9 + r] = s.scale * tile[tap][c][ci];
  }
}

// ------------------------------------------------------------------------------------------------
// conv1x1 backward (block.py:244,263: x2 = lrelu(conv2(..)) + U x, U = [gc, nf], no bias)
//   g[px][c]  += sum_k dx2[px][k] * U[k][c]  (+ extra[px][c])          data gradient, in place on the fp32 trunk gradient
//   dU[k][c]  += sum_px dx2[px][k] * x[px][c]                           weight gradient (atomics per CTA)
// ------------------------------------------------------------------------------------------------
// Both products run on warp-level bf16 MMAs (mma.sync m16n8k16, fp32 accumulate) from padded shared-memory tiles of 64
// pixels: O = D . U (A = D row-major via ldmatrix, B = U^T staged once per CTA as bf16 — the forward's packed 1x1 is bf16
// too) and dU = D^T . X (contraction over pixels, both fragments via ldmatrix.trans like conv3x3_wgrad_kernel).  The
// kernel is then bound by its HBM traffic (x, dx2 in; the fp32 trunk gradient in and out).
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int NF>
__global__ void __launch_bounds__(256, 3) conv1x1_bwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ctotal,
                                                             const __nv_bfloat16* __restrict__ dx2, int d_ctotal, int d_c0,
                                                             const float* __restrict__ u, float* __restrict__ g,
                                                             const float* __restrict__ extra, float* __restrict__ du_acc,
                                                             long long npx) {
  constexpr int TP = 64;                 // pixels per tile
  constexpr int DP = 80;                 // bytes per row of the D tile / of U^T (32 bf16 + 16 pad: conflict-free ldmatrix)
  constexpr int XP = NF * 2 + 16;        // bytes per row of the X tile
  constexpr int NT1 = NF / 16;           // n8 tiles per warp in O = D . U   (warp = 16 pixels x NF/2 channels)
  __shared__ __align__(16) uint8_t ut[NF * DP];
  __shared__ __align__(16) uint8_t dsm[TP * DP];
  __shared__ __align__(16) uint8_t xsm[TP * XP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32 * NF; i += 256) {
    const int k = i / NF, c = i % NF;
    *reinterpret_cast<__nv_bfloat16*>(ut + c * DP + k * 2) = __float2bfloat16_rn(u[i]);
  }
  // part 1 (O = D . U): warp -> pixel rows m0 .., channels n0 ..
  const int m0 = 16 * (warp & 3), n0 = (NF / 2) * (warp >> 2);
  // part 2 (dU = D^T . X): warp -> dx2 channels 16 * mi .., x channels 16 * nh ..
  const int mi = warp & 1, nh = warp >> 1;
  const bool du_warp = du_acc != nullptr && nh * 16 < NF;
  float du[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) du[j][i] = 0.f;
  const uint32_t ut_a = smem_u32(ut), ds_a = smem_u32(dsm), xs_a = smem_u32(xsm);
  const long long ntiles = (npx + TP - 1) / TP;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long p0 = t * TP;
    __syncthreads();
    for (int i = tid; i < TP * 4; i += 256) {
      const int seg = i & 3, px = i >> 2;
      const bool ok = p0 + px < npx;
      cp_async16_zfill(ds_a + px * DP + seg * 16, dx2 + (ok ? p0 + px : 0) * d_ctotal + d_c0 + seg * 8, ok);
    }
    for (int i = tid; i < TP * (NF / 8); i += 256) {
      const int seg = i % (NF / 8), px = i / (NF / 8);
      const bool ok = p0 + px < npx;
      cp_async16_zfill(xs_a + px * XP + seg * 16, x + (ok ? p0 + px : 0) * x_ctotal + seg * 8, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // ---- data gradient: g[px][c] += sum_k D[px][k] U[k][c] (+ extra)
    {
      float o[NT1][4];
#pragma unroll
      for (int j = 0; j < NT1; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a[4];
        ldmatrix_x4(ds_a + (m0 + (lane & 15)) * DP + (kk * 16 + (lane >> 4) * 8) * 2, a);
#pragma unroll
        for (int np = 0; np < NT1 / 2; ++np) {
          uint32_t b[4];
          ldmatrix_x4(ut_a + (n0 + np * 16 + (lane & 7) + (lane >> 4) * 8) * DP + (kk * 16 + ((lane >> 3) & 1) * 8) * 2, b);
          mma_bf16_16816(o[2 * np], a, b[0], b[1]);
          mma_bf16_16816(o[2 * np + 1], a, b[2], b[3]);
        }
      }
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const long long px = p0 + m0 + (lane >> 2) + hrow * 8;
        if (px < npx) {
#pragma unroll
          for (int j = 0; j < NT1; ++j) {
            const size_t off = static_cast<size_t>(px) * NF + n0 + j * 8 + (lane & 3) * 2;
            float2 v = *reinterpret_cast<float2*>(g + off);
            v.x += o[j][hrow * 2];
            v.y += o[j][hrow * 2 + 1];
            if (extra != nullptr) {
              const float2 e2 = *reinterpret_cast<const float2*>(extra + off);
              v.x += e2.x;
              v.y += e2.y;
            }
            *reinterpret_cast<float2*>(g + off) = v;
          }
        }
      }
    }
    // ---- weight gradient: dU[k][c] += sum_px D[px][k] X[px][c]
    if (du_warp) {
      const int a_px = (lane & 7) + ((lane >> 4) << 3);
      const int a_off = a_px * DP + (mi * 16 + ((lane >> 3) & 1) * 8) * 2;
      const int b_px = (lane & 7) + (((lane >> 3) & 1) << 3);
      const int b_off = b_px * XP + (nh * 16 + (lane >> 4) * 8) * 2;
#pragma unroll
      for (int ks = 0; ks < TP / 16; ++ks) {
        uint32_t a[4], b[4];
        ldmatrix_x4_trans(ds_a + ks * 16 * DP + a_off, a);
        ldmatrix_x4_trans(xs_a + ks * 16 * XP + b_off, b);
        mma_bf16_16816(du[0], a, b[0], b[1]);
        mma_bf16_16816(du[1], a, b[2], b[3]);
      }
    }
  }
  if (du_warp) {
    const int k = mi * 16 + (lane >> 2);
    const int c = nh * 16 + (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      atomicAdd(du_acc + k * NF + c + j * 8, du[j][0]);
      atomicAdd(du_acc + k * NF + c + j * 8 + 1, du[j][1]);
      atomicAdd(du_acc + (k + 8) * NF + c + j * 8, du[j][2]);
      atomicAdd(du_acc + (k + 8) * NF + c + j * 8 + 1, du[j][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample backward: d[n,y,x,c] = sum_{a,b} dup[n,2y+a,2x+b,c], then optional LeakyReLU mask
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_bwd_kernel(const uint4* __restrict__ dup, int n, int h, int w, int cv,
                                      const unsigned short* __restrict__ mask, int m_ctotal, int m_c0,
                                      uint4* __restrict__ out_bf16, float* __restrict__ out_f32) {
  const size_t total = static_cast<size_t>(n) * h * w * cv;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t t = i / cv;
    const int x = static_cast<int>(t % w); t /= w;
    const int y = static_cast<int>(t % h);
    const int img = static_cast<int>(t / h);
    const size_t ow = static_cast<size_t>(2) * w;
    const size_t base = ((static_cast<size_t>(img) * 2 * h + 2 * y) * ow + 2 * x) * cv + v;
    const uint4 q[4] = {__ldg(dup + base), __ldg(dup + base + cv), __ldg(dup + base + ow * cv), __ldg(dup + base + ow * cv + cv)};
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t u[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[2 * j] += __uint_as_float(u[j] << 16);
        s[2 * j + 1] += __uint_as_float(u[j] & 0xFFFF0000u);
      }
    }
    const size_t pix = (static_cast<size_t>(img) * h + y) * w + x;
    if (out_f32) {
      float4* op = reinterpret_cast<float4*>(out_f32 + (pix * cv + v) * 8);
      op[0] = make_float4(s[0], s[1], s[2], s[3]);
      op[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    if (mask) {
      const int bit0 = m_c0 + v * 8;
      const uint32_t bits = mask[(pix * m_ctotal + bit0) >> 4] >> (bit0 & 15);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] *= ((bits >> j) & 1u) ? 1.0f : 0.2f;
    }
    if (out_bf16) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(s[2 * j], s[2 * j + 1]);
        pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      out_bf16[pix * cv + v] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int plan_wgrad(const esrp_wgrad_unit_t* units, int num_units, int n, int h, int w, int splits, WgradLaunch* out) {
  if (!units || num_units < 1 || num_units > ESRP_WGRAD_MAX_UNITS) return set_error("wgrad: num_units=%d out of range (1..%d)", num_units, ESRP_WGRAD_MAX_UNITS);
  if (n < 1 || h < 1 || w < 1) return set_error("wgrad: bad shape");
  // default: the tcgen05 kernel (esrp_wgrad_tc.cu); ESRP_WGRAD_MMA=1 or an explicit split count selects the mma.sync
  // kernel below, which also takes over where a tensor map cannot be encoded for the tcgen05 tiling
  static const bool force_mma = getenv("ESRP_WGRAD_MMA") != nullptr;
  if (!force_mma && splits <= 0) {
    bool ok = true;
    for (int i = 0; i < num_units; ++i)
      if (!units[i].x || !units[i].dy || !units[i].acc || (units[i].x_c0 % 32) || (units[i].dy_c0 % 64)) ok = false;
    static const bool strict = getenv("ESRP_WGRAD_TC_STRICT") != nullptr;  // tests: a refused tcgen05 plan is an error
    if (ok) {
      if (plan_wgrad_tc(units, num_units, n, h, w, out) == 0) return 0;
      if (strict) return 1;
    }
  }
  if (num_units > kMmaMaxUnits) return set_error("wgrad: the mma.sync kernel takes at most %d units per launch (got %d)", kMmaMaxUnits, num_units);
  out->tc = 0;
  out->num_bias = 0;
  WgradParams& p = *reinterpret_cast<WgradParams*>(out->params);
  static_assert(sizeof(WgradParams) <= sizeof(out->params), "WgradLaunch::params too small");
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < num_units; ++i) {
    const esrp_wgrad_unit_t& u = units[i];
    if (!u.x || !u.dy || !u.acc) return set_error("wgrad: unit %d has a null pointer", i);
    if ((u.x_ctotal % 8) || (u.x_c0 % 8) || u.x_c0 + 32 > u.x_ctotal) return set_error("wgrad: unit %d x channel range", i);
    if ((u.dy_ctotal % 8) || (u.dy_c0 % 8) || u.dy_c0 >= u.dy_ctotal) return set_error("wgrad: unit %d dy channel range", i);
    p.u[i] = u;
  }
  p.num_units = num_units;
  p.n = n; p.h = h; p.w = w;
  int twl = 4;
  while (twl < 6 && (1 << twl) < w) ++twl;   // 16, 32 or 64 columns
  p.tw_log2 = twl;
  p.tw = 1 << twl;
  p.tr = kWgTilePx / p.tw;
  p.tiles_x = (w + p.tw - 1) / p.tw;
  p.tiles_y = (h + p.tr - 1) / p.tr;
  const long long tt = static_cast<long long>(n) * p.tiles_x * p.tiles_y;
  if (tt > 0x7fffffffLL) return set_error("wgrad: problem too large");
  p.tiles_total = static_cast<int>(tt);
  p.x_tile_px = (p.tr + 2) * (p.tw + 2);
  p.stage_bytes = (p.x_tile_px * kWgXPitch + kWgTilePx * kWgYPitch + 127) / 128 * 128;
  const int sms = sm_count();
  if (sms <= 0) return set_error("wgrad: no CUDA device");
  if (splits <= 0) splits = (2 * sms + num_units - 1) / num_units;
  if (splits > p.tiles_total) splits = p.tiles_total;
  p.splits = splits;
  out->grid = num_units * splits;
  out->smem = 2 * p.stage_bytes;
  if (ensure_max_smem(reinterpret_cast<const void*>(conv3x3_wgrad_kernel), 2 * 48 * 1024)) return 1;
  if (out->smem > 2 * 48 * 1024) return set_error("wgrad: internal: stage too large");
  return 0;
}

int run_wgrad(const WgradLaunch& L, cudaStream_t stream) {
  if (L.tc) return run_wgrad_tc(L, stream);
  const WgradParams& p = *reinterpret_cast<const WgradParams*>(L.params);
  conv3x3_wgrad_kernel<<<L.grid, kWgThreads, L.smem, stream>>>(p);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int run_conv1x1_bwd(int nf, const void* x, int x_ctotal, const void* dx2, int d_ctotal, int d_c0, const float* u,
                    float* g, const float* extra, float* du_acc, long long npx, cudaStream_t stream) {
  const int sms = sm_count();
  long long tiles = (npx + 63) / 64;
  // every CTA ends with 32 x nf atomics into dU: >= 4 tiles per CTA before adding CTAs, at most 3 CTAs per SM
  long long want = (tiles + 3) / 4;
  if (want > 3LL * sms) want = 3LL * sms;
  if (want < 1) want = 1;
  int grid = static_cast<int>(want);
  if (grid < 1) return 0;
  if (nf == 64)
    conv1x1_bwd_kernel<64><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ctotal,
                                                       static_cast<const __nv_bfloat16*>(dx2), d_ctotal, d_c0, u, g, extra, du_acc, npx);
  else if (nf == 32)
    conv1x1_bwd_kernel<32><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), x_ctotal,
                                                       static_cast<const __nv_bfloat16*>(dx2), d_ctotal, d_c0, u, g, extra, du_acc, npx);
  else
    return set_error("conv1x1_bwd: nf=%d unsupported (32 or 64)", nf);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int run_scatter(const esrp_scatter_entry_t* tab_dev, int num, float* const* dst_ptrs_dev, cudaStream_t stream) {
  if (num < 1) return 0;
  wgrad_scatter_kernel<<<num * 8, 256, 0, stream>>>(tab_dev, num, dst_ptrs_dev);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace esrp

using namespace esrp;

extern "C" {

int esrp_pack_dgrad_weights(const esrp_dgrad_group_t* groups_host, int32_t num_groups, int32_t layout, int32_t row0,
                            int32_t rows, int32_t kc, int32_t bn, void* out, void* stream) {
  if (!groups_host || !out) return set_error("pack_dgrad: null pointer");
  if (kc != 32 && kc != 64) return set_error("pack_dgrad: kc must be 32 or 64");
  if (layout != ESRP_LAYOUT_TILE && layout != ESRP_LAYOUT_ROW) return set_error("pack_dgrad: unknown layout %d", layout);
  if (bn != 16 && bn != 32 && bn != 64) return set_error("pack_dgrad: bn must be 16, 32 or 64");
  if (rows < 1 || rows > bn || row0 < 0) return set_error("pack_dgrad: bad row slice");
  if (num_groups < 1 || num_groups > 2 * ESRP_MAX_CHUNKS) return set_error("pack_dgrad: num_groups=%d", num_groups);
  const int num_chunks = (num_groups * 32 + kc - 1) / kc;
  if (num_chunks > ESRP_MAX_CHUNKS) return set_error("pack_dgrad: too many chunks");
  DgradTable tab;
  memset(&tab, 0, sizeof(tab));
  for (int i = 0; i < num_groups; ++i) tab.g[i] = groups_host[i];
  const int total = num_chunks * 9 * bn * kc;
  int blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  pack_dgrad_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(tab, num_groups, layout, row0, rows, kc, bn,
                                                                          num_chunks, static_cast<__nv_bfloat16*>(out));
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_conv3x3_wgrad(const esrp_wgrad_unit_t* units_host, int32_t num_units, int32_t n, int32_t h, int32_t w,
                       int32_t splits, void* stream) {
  // same plan cache as esrp_conv3x3_nhwc: planning encodes two tensor maps per job
  if (!units_host || num_units < 1 || num_units > ESRP_WGRAD_MAX_UNITS) return set_error("wgrad: num_units=%d out of range (1..%d)", num_units, ESRP_WGRAD_MAX_UNITS);
  static std::mutex mu;
  static std::unordered_map<std::string, WgradLaunch>* cache = new std::unordered_map<std::string, WgradLaunch>();
  std::string key(reinterpret_cast<const char*>(units_host), sizeof(esrp_wgrad_unit_t) * num_units);
  const int32_t shape[4] = {n, h, w, splits};
  key.append(reinterpret_cast<const char*>(shape), sizeof(shape));
  WgradLaunch L;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache->find(key);
    if (it == cache->end()) {
      if (plan_wgrad(units_host, num_units, n, h, w, splits, &L)) return 1;
      if (cache->size() >= 2048) cache->clear();
      cache->emplace(std::move(key), L);
    } else {
      L = it->second;
    }
  }
  return run_wgrad(L, static_cast<cudaStream_t>(stream));
}

int esrp_wgrad_scatter(const esrp_scatter_entry_t* entries_host, int32_t num, void* stream) {
  if (!entries_host || num < 1) return set_error("wgrad_scatter: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  esrp_scatter_entry_t* dev = nullptr;
  ESRP_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&dev), sizeof(esrp_scatter_entry_t) * num, s));
  ESRP_CUDA_OK(cudaMemcpyAsync(dev, entries_host, sizeof(esrp_scatter_entry_t) * num, cudaMemcpyHostToDevice, s));
  int rc = run_scatter(dev, num, nullptr, s);
  ESRP_CUDA_OK(cudaFreeAsync(dev, s));
  return rc;
}

int esrp_conv1x1_bwd(int32_t nf, const void* x, int32_t x_ctotal, const void* dx2, int32_t d_ctotal, int32_t d_c0,
                     const float* u, float* g, const float* extra, float* du_acc, int64_t npx, void* stream) {
  if (!x || !dx2 || !u || !g) return set_error("conv1x1_bwd: null pointer");
  return run_conv1x1_bwd(nf, x, x_ctotal, dx2, d_ctotal, d_c0, u, g, extra, du_acc, npx, static_cast<cudaStream_t>(stream));
}

int esrp_upsample2x_bwd_nhwc_bf16(const void* dup, int32_t n, int32_t h, int32_t w, int32_t c, const void* mask,
                                  int32_t mask_ctotal, int32_t mask_c0, void* out_bf16, float* out_f32, void* stream) {
  if (!dup || (c % 8) || (!out_bf16 && !out_f32)) return set_error("upsample2x_bwd: bad arguments");
  if (mask && ((mask_ctotal % 16) || (mask_c0 % 8))) return set_error("upsample2x_bwd: mask alignment");
  const int cv = c / 8;
  const size_t total = static_cast<size_t>(n) * h * w * cv;
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 0;
  upsample2x_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dup), n, h, w, cv, static_cast<const unsigned short*>(mask), mask_ctotal, mask_c0,
      static_cast<uint4*>(out_bf16), out_f32);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
