// esrp_wgrad_tc.cu — weight gradient of the 3x3 convs on the 5th-gen tensor cores (tcgen05), sm_100a.
//
//   dW[tap][co][ci] = sum_px dY[px][co] * X[px + tap][ci]          (autograd convolution_backward of block.py:253-258)
//
// is a GEMM whose contraction index is the PIXEL, the slow index of both NHWC operands.  tcgen05.mma takes such
// operands directly: with the "MN-major" bit of the instruction descriptor set, an operand tile is [K rows of 128 B]
// = [pixel][64 channels] exactly as TMA writes an NHWC box with the 128-byte swizzle, so no transposition pass exists:
//
//   A (M = 128 output-gradient channels) : dY tile, 128 pixels x 2 panels of 64 channels       (TMA box 64 x TW x TR)
//   B (N = 64 input channels)            : X halo tile, (TR+1) x (TW+2) pixels x 64 channels    (TMA box 64 x TW+2 x TR+1:
//                                          the two kernel rows the CTA's tap group touches)
//   tap (ky, kx)                         : the SAME X tile with the descriptor start address advanced by
//                                          (ky * (TW+2) + kx) pixel rows — the swizzle is a function of the address bits,
//                                          so a 128-byte shift stays consistent with what TMA wrote (the trick of
//                                          conv3x3_row.cuh, here along K instead of M)
//   D                                    : one 128-lane x 64-column fp32 accumulator per tap in tensor memory; it stays
//                                          there over ALL pixel tiles of the CTA (output-stationary over K), so the main
//                                          loop is TMA + MMA only and the epilogue runs once.
// Nine taps x 64 columns do not fit the 512 TMEM columns, so a CTA owns 5 or 4 taps ("tap group"); grid = jobs x 2 tap
// groups x pixel splits.  A job = (64-channel chunk of X, 128-channel slab of dY) = up to four of the 32 x 64 "units" of
// include/esrp.h; the epilogue adds the accumulators into the units' fp32 blocks [9][64][32] with 16-byte vector
// reductions (the same blocks the mma.sync kernel of esrp_bwd.cu fills, so scatter and callers are unchanged).
// Bias gradients (column sums of dY) come from a small HBM-bound reduction kernel.
//
// Warps: 0 = TMA producer, 1 and 6 = MMA issuers (they split the taps; warp 1 also allocates TMEM), 2-5 = epilogue (TMEM
// lane quarter = warp % 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "../../include/esrp.h"
#include "esrp_bwd.h"
#include "esrp_host.h"
#include "esrp_ptx.cuh"

namespace esrp {

namespace {

constexpr int kTcThreads = 224;      // warps: 0 TMA, 1 + 6 MMA issuers, 2-5 epilogue
constexpr int kTcMaxStages = 4;
constexpr int kTcYPanel = 128 * 128;  // bytes of one 64-channel dY panel (128 pixels x 128 B)

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
  // MN-major operand, 128-byte swizzle: ((8,n),(8,k)) : ((16 B, LBO), (128 B, SBO = 1024 B))
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16, BF16 x BF16 -> F32, A and B MN-major, M = 128, N = 64
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// MMA issue loop of one issuer thread for the local taps [TL0, TL1) of its CTA's tap group.  The issue stream must stay
// at one or two integer ops per MMA (the tensor pipe queues almost nothing), so everything but the stage base is
// hoisted: descriptor hi words are constants, the lo word is (stage base + k-step offset) + tap offset in 16-byte units.
template <int TL0, int TL1>
__device__ __forceinline__ void issue_loop(const WgTcParams& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                           uint64_t* done_bar, uint32_t tmem_base, uint32_t y_off, int t_begin, int t_end,
                                           int tap0, int tg) {
  const uint32_t d_hi = static_cast<uint32_t>(mn_desc(0, kTcYPanel) >> 32);
  const uint32_t lo_lbo = static_cast<uint32_t>(mn_desc(0, kTcYPanel) & 0xFFFF0000u);
  uint32_t ks_off[8], tap_off[TL1 - TL0];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int k0 = ks * 16;
    ks_off[ks] = static_cast<uint32_t>(((k0 >> p.tw_log2) * p.xw + (k0 & (p.tw - 1))) * 8);
  }
#pragma unroll
  for (int tl = TL0; tl < TL1; ++tl) {
    const int tap = tap0 + tl;
    const int ky = tap / 3, kx = tap - ky * 3;
    tap_off[tl - TL0] = static_cast<uint32_t>(((ky - tg) * p.xw + kx) * 8);
  }
  for (int t = t_begin; t < t_end; ++t) {
    const int i = t - t_begin, s = i % p.stages;
    mbar_wait(&full_bar[s], (i / p.stages) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t xs = smem_u32(smem + s * p.stage_bytes);
    const uint32_t b_lo = ((xs >> 4) & 0x3FFFu) | lo_lbo;
    const uint32_t a_lo = (((xs + y_off) >> 4) & 0x3FFFu) | lo_lbo;
    const uint32_t first = i > 0 ? 1u : 0u;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t a = a_lo + ks * 128;  // 16 pixel rows of 128 B
      const uint32_t bk = b_lo + ks_off[ks];
      const uint32_t accf = ks > 0 ? 1u : first;
#pragma unroll
      for (int tl = TL0; tl < TL1; ++tl)
        umma_f16_ss2(tmem_base + tl * 64, a, d_hi, bk + tap_off[tl - TL0], d_hi, kIdesc, accf);
    }
    umma_commit(&empty_bar[s]);
  }
  umma_commit(done_bar);
}

__global__ void __launch_bounds__(kTcThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgTcParams p) {
  extern __shared__ uint8_t tc_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kTcMaxStages], empty_bar[kTcMaxStages], done_bar;
  __shared__ uint32_t tmem_holder;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_job = 2 * p.splits;
  const int job = blockIdx.x / per_job;
  const int rem = blockIdx.x - job * per_job;
  const int tg = rem / p.splits;
  const int split = rem - tg * p.splits;
  const int t_begin = static_cast<int>(static_cast<long long>(p.tiles_total) * split / p.splits);
  const int t_end = static_cast<int>(static_cast<long long>(p.tiles_total) * (split + 1) / p.splits);
  const int tap0 = tg == 0 ? 0 : 5, ntaps = tg == 0 ? 5 : 4;
  const WgTcJobInfo& J = p.job[job];

  if (threadIdx.x == 0) {
    // a tcgen05.commit only tracks the MMAs of the committing thread: "stage consumed" == both issuers committed
    for (int s = 0; s < kTcMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
    mbar_init(&done_bar, 2);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tma_prefetch_desc(&p.tmx[J.mx]);
    tma_prefetch_desc(&p.tmy[J.my]);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_holder, 512);
    tmem_relinquish();
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_holder;
  const uint32_t y_off = static_cast<uint32_t>((p.x_bytes + 1023) / 1024 * 1024);

  if (warp == 0) {
    if (elect_one()) {
      for (int t = t_begin; t < t_end; ++t) {
        const int i = t - t_begin, s = i % p.stages;
        if (i >= p.stages) mbar_wait(&empty_bar[s], ((i / p.stages) - 1) & 1);
        const int tx = t % p.tiles_x;
        const int ty = (t / p.tiles_x) % p.tiles_y;
        const int img = t / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.tw, y0 = ty * p.tr;
        uint8_t* xs = smem + s * p.stage_bytes;
        uint8_t* ys = xs + y_off;
        // the tap group needs two of the three kernel rows: rows y0-1 .. y0+tr-1 (taps 0-4) or y0 .. y0+tr (taps 5-8);
        // a slab whose upper 64 channels lie beyond the tensor skips that panel (its accumulator rows are never read)
        mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(p.x_bytes + J.y_panels * kTcYPanel));
        tma_load_4d(xs, &p.tmx[J.mx], &full_bar[s], J.xc0, x0 - 1, y0 - 1 + tg, img);
        tma_load_4d(ys, &p.tmy[J.my], &full_bar[s], J.dyc0, x0, y0, img);
        if (J.y_panels > 1) tma_load_4d(ys + kTcYPanel, &p.tmy[J.my], &full_bar[s], J.dyc0 + 64, x0, y0, img);
      }
    }
  } else if (warp == 1 || warp == 6) {
    if (elect_one()) {
      // two issuer warps split the taps of the group (tcgen05.mma issue is nearly synchronous: one issuer leaves a
      // bubble at every barrier wait, tools/ubench_row.cu)
      const int half = (ntaps + 1) / 2;
      const bool first_issuer = warp == 1;
      if (ntaps == 5) {
        if (first_issuer) issue_loop<0, 3>(p, smem, full_bar, empty_bar, &done_bar, tmem_base, y_off, t_begin, t_end, tap0, tg);
        else issue_loop<3, 5>(p, smem, full_bar, empty_bar, &done_bar, tmem_base, y_off, t_begin, t_end, tap0, tg);
      } else {
        if (first_issuer) issue_loop<0, 2>(p, smem, full_bar, empty_bar, &done_bar, tmem_base, y_off, t_begin, t_end, tap0, tg);
        else issue_loop<2, 4>(p, smem, full_bar, empty_bar, &done_bar, tmem_base, y_off, t_begin, t_end, tap0, tg);
      }
      (void)half;
    }
  } else if (t_begin < t_end) {
    // epilogue: TMEM lane = output-gradient channel within the 128-channel slab, column = (tap, input channel)
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int blk = row >> 6, co = row & 63;
    float* a0 = J.acc[0][blk];
    float* a1 = J.acc[1][blk];
    for (int tl = 0; tl < ntaps; ++tl) {
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tl * 64;
      tmem_ld_x32(taddr, v0);
      tmem_ld_x32(taddr + 32, v1);
      tmem_ld_wait();
      const size_t off = (static_cast<size_t>(tap0 + tl) * 64 + co) * 32;
      if (a0 != nullptr) {
        float4* d = reinterpret_cast<float4*>(a0 + off);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          atomicAdd(d + j, make_float4(__uint_as_float(v0[4 * j]), __uint_as_float(v0[4 * j + 1]), __uint_as_float(v0[4 * j + 2]),
                                       __uint_as_float(v0[4 * j + 3])));
      }
      if (a1 != nullptr) {
        float4* d = reinterpret_cast<float4*>(a1 + off);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          atomicAdd(d + j, make_float4(__uint_as_float(v1[4 * j]), __uint_as_float(v1[4 * j + 1]), __uint_as_float(v1[4 * j + 2]),
                                       __uint_as_float(v1[4 * j + 3])));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    tmem_dealloc(tmem_base, 512);
  }
}

// bias gradients: column sums of dY over all pixels for up to 8 requested 64-channel blocks of ONE tensor, in a single
// HBM-bound pass (16-byte loads; thread = 8 channels x a pixel lane).
struct ColsumArgs {
  int c0[8];
  float* out[8];
  int num;
};
__global__ void __launch_bounds__(256) colsum_kernel(const uint4* __restrict__ dy, int ctotal, long long npx, ColsumArgs a) {
  __shared__ float red[256][9];
  const int vecs = ctotal >> 3;                 // 16-byte vectors per pixel
  const int lanes = 256 / vecs;                 // pixel lanes per block
  const int v = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (pl < lanes) {
    for (long long px = static_cast<long long>(blockIdx.x) * lanes + pl; px < npx; px += static_cast<long long>(gridDim.x) * lanes) {
      const uint4 q = __ldg(dy + px * vecs + v);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[2 * j] += __uint_as_float(u[j] << 16);
        s[2 * j + 1] += __uint_as_float(u[j] & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = pl < lanes ? s[j] : 0.f;
  __syncthreads();
  // thread t < ctotal sums channel t over the pixel lanes
  for (int c = threadIdx.x; c < ctotal; c += 256) {
    const int vv = c >> 3, j = c & 7;
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += red[l * vecs + vv][j];
    for (int k = 0; k < a.num; ++k)
      if (c >= a.c0[k] && c < a.c0[k] + 64) atomicAdd(a.out[k] + (c - a.c0[k]), t);
  }
}

}  // namespace

int plan_wgrad_tc(const esrp_wgrad_unit_t* units, int num_units, int n, int h, int w, WgradLaunch* out) {
  WgTcParams& p = *reinterpret_cast<WgTcParams*>(out->params);
  static_assert(sizeof(WgTcParams) <= sizeof(out->params), "WgradLaunch::params too small for WgTcParams");
  memset(&p, 0, sizeof(p));
  out->num_bias = 0;
  struct Key { const void* x; int xct, xc0; const void* dy; int dct, dyc0; };
  struct Ten { const void* p; int ct; };
  Key keys[kTcMaxJobs];
  Ten xt[kTcMaxMaps], yt[kTcMaxMaps];
  int nj = 0, nxt = 0, nyt = 0;
  auto ten_index = [](Ten* tab, int* cnt, const void* ptr, int ct) -> int {
    for (int i = 0; i < *cnt; ++i)
      if (tab[i].p == ptr && tab[i].ct == ct) return i;
    if (*cnt == kTcMaxMaps) return -1;
    tab[*cnt] = Ten{ptr, ct};
    return (*cnt)++;
  };
  for (int i = 0; i < num_units; ++i) {
    const esrp_wgrad_unit_t& u = units[i];
    if ((u.x_c0 % 32) || (u.dy_c0 % 64)) return set_error("wgrad(tc): unit %d channel offsets must be multiples of 32 / 64", i);
    Key k{u.x, u.x_ctotal, u.x_c0 - u.x_c0 % 64, u.dy, u.dy_ctotal, 0};
    // 128-aligned slab of dY holding the unit's 64-column block (channels beyond the tensor are zero-filled by TMA)
    k.dyc0 = u.dy_c0 - u.dy_c0 % 128;
    int j = 0;
    for (; j < nj; ++j)
      if (keys[j].x == k.x && keys[j].xc0 == k.xc0 && keys[j].dy == k.dy && keys[j].dyc0 == k.dyc0) break;
    if (j == nj) {
      if (nj == kTcMaxJobs) return set_error("wgrad(tc): more than %d (chunk, slab) jobs in one launch", kTcMaxJobs);
      const int mx = ten_index(xt, &nxt, k.x, k.xct), my = ten_index(yt, &nyt, k.dy, k.dct);
      if (mx < 0 || my < 0) return set_error("wgrad(tc): more than %d distinct tensors in one launch", kTcMaxMaps);
      keys[nj] = k;
      p.job[nj].xc0 = k.xc0;
      p.job[nj].dyc0 = k.dyc0;
      p.job[nj].y_panels = (k.dyc0 + 64 < k.dct) ? 2 : 1;
      p.job[nj].mx = static_cast<short>(mx);
      p.job[nj].my = static_cast<short>(my);
      ++nj;
    }
    const int g = (u.x_c0 % 64) / 32, b = (u.dy_c0 - k.dyc0) / 64;
    if (p.job[j].acc[g][b] != nullptr) return set_error("wgrad(tc): duplicate unit %d", i);
    p.job[j].acc[g][b] = u.acc;
    if (u.bias_acc != nullptr) {
      if (out->num_bias == 8) return set_error("wgrad(tc): too many bias accumulators");
      if ((u.dy_ctotal % 8) || u.dy_ctotal > 2048) return set_error("wgrad(tc): dy channel count %d unsupported for the bias reduction", u.dy_ctotal);
      out->bias[out->num_bias++] = WgradLaunch::Bias{u.dy, u.dy_ctotal, u.dy_c0, u.bias_acc};
    }
  }
  p.num_jobs = nj;
  p.n = n; p.h = h; p.w = w;
  int twl = 4;
  while (twl < 7 && (1 << twl) < w) ++twl;  // 16, 32, 64 or 128 columns
  p.tw_log2 = twl;
  p.tw = 1 << twl;
  p.tr = 128 / p.tw;
  p.xw = p.tw + 2;
  p.tiles_x = (w + p.tw - 1) / p.tw;
  p.tiles_y = (h + p.tr - 1) / p.tr;
  const long long tt = static_cast<long long>(n) * p.tiles_x * p.tiles_y;
  if (tt > 0x7fffffffLL) return set_error("wgrad(tc): problem too large");
  p.tiles_total = static_cast<int>(tt);
  p.x_bytes = (p.tr + 1) * p.xw * 128;
  p.stage_bytes = (p.x_bytes + 1023) / 1024 * 1024 + 2 * kTcYPanel;
  p.stages = (kMaxSmem - 2048) / p.stage_bytes;
  if (p.stages > kTcMaxStages) p.stages = kTcMaxStages;
  if (p.stages < 2) return set_error("wgrad(tc): internal: fewer than two stages fit in shared memory");
  const int sms = sm_count();
  if (sms <= 0) return set_error("wgrad(tc): no CUDA device");
  int splits = sms / (2 * nj);
  if (splits < 1) splits = 1;
  if (splits > p.tiles_total) splits = p.tiles_total;
  p.splits = splits;
  for (int i = 0; i < nxt; ++i)
    if (make_nhwc_tmap(&p.tmx[i], xt[i].p, n, h, w, xt[i].ct, 64, p.xw, p.tr + 1)) return 1;
  for (int i = 0; i < nyt; ++i)
    if (make_nhwc_tmap(&p.tmy[i], yt[i].p, n, h, w, yt[i].ct, 64, p.tw, p.tr)) return 1;
  out->grid = nj * 2 * splits;
  out->smem = p.stages * p.stage_bytes + 1024;
  out->tc = 1;
  out->npx = static_cast<long long>(n) * h * w;
  if (ensure_max_smem(reinterpret_cast<const void*>(wgrad_tc_kernel), kMaxSmem - 1024)) return 1;
  if (out->smem > kMaxSmem - 1024) return set_error("wgrad(tc): internal: stages do not fit in shared memory");
  return 0;
}

int run_wgrad_tc(const WgradLaunch& L, cudaStream_t stream) {
  const WgTcParams& p = *reinterpret_cast<const WgTcParams*>(L.params);
  wgrad_tc_kernel<<<L.grid, kTcThreads, L.smem, stream>>>(p);
  ESRP_CUDA_OK(cudaGetLastError());
  const int sms = sm_count();
  bool done[8] = {false, false, false, false, false, false, false, false};
  for (int i = 0; i < L.num_bias; ++i) {
    if (done[i]) continue;
    ColsumArgs a;
    a.num = 0;
    for (int j = i; j < L.num_bias; ++j)
      if (!done[j] && L.bias[j].dy == L.bias[i].dy) {
        a.c0[a.num] = L.bias[j].c0;
        a.out[a.num] = L.bias[j].out;
        ++a.num;
        done[j] = true;
      }
    const int ct = L.bias[i].ctotal;
    const int lanes = 256 / (ct / 8);
    // every block ends with one atomic per channel: give each pixel lane >= 16 pixels before adding blocks
    long long blocks = (L.npx + lanes * 16 - 1) / (lanes * 16);
    if (blocks > 4LL * sms) blocks = 4LL * sms;
    if (blocks < 1) blocks = 1;
    colsum_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const uint4*>(L.bias[i].dy), ct, L.npx, a);
  }
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace esrp
