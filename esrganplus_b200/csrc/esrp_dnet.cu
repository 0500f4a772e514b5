// esrp_dnet.cu — the non-GEMM pieces of Discriminator_VGG_128 (reference: architecture.py:87-129):
// space-to-depth for the 4x4 / stride-2 convs, BatchNorm2d batch statistics + normalise + LeakyReLU,
// and the two small Linear layers.  The convolutions themselves run on the tcgen05 conv kernels
// (esrp_conv3x3_nhwc); all kernels here are HBM-bound elementwise / reduction passes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"
#include "esrp_host.h"

namespace esrp {

// 4x4 stride-2 pad-1 conv == 2x2 conv over S[n, Y, X, (a*2+b)*c + ch] = in[n, 2Y+a-1, 2X+b-1, ch]
// (Y in [0, h/2], X in [0, w/2], zero outside the image).  16-byte vectors along channels.
__global__ void s2d_pad_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n, int h, int w, int cv) {
  const int ho = h / 2 + 1, wo = w / 2 + 1;
  const size_t total = static_cast<size_t>(n) * ho * wo * 4 * cv;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t t = i / cv;
    const int ab = static_cast<int>(t % 4); t /= 4;
    const int X = static_cast<int>(t % wo); t /= wo;
    const int Y = static_cast<int>(t % ho);
    const int img = static_cast<int>(t / ho);
    const int y = 2 * Y + (ab >> 1) - 1, x = 2 * X + (ab & 1) - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < h && x >= 0 && x < w) val = __ldg(src + ((static_cast<size_t>(img) * h + y) * w + x) * cv + v);
    dst[i] = val;
  }
}

// Per-channel sum and sum of squares over the valid [h, w] region of x [n, hp, wp, c] fp32.
// One block per (channel group of 32, pixel slab); partial sums via atomics on double accumulators.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int n, int h, int w, int hp, int wp, int c,
                                                       double* __restrict__ sums) {
  // thread = 4 channels (one 16-byte load) x one of 32 pixel lanes; block = 32 channels
  const int c4 = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int ch = blockIdx.x * 32 + c4 * 4;
  const long long npx = static_cast<long long>(n) * h * w;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  if (ch < c) {
    for (long long pi = static_cast<long long>(blockIdx.y) * 32 + pl; pi < npx; pi += static_cast<long long>(gridDim.y) * 32) {
      const int xw = static_cast<int>(pi % w);
      const long long t = pi / w;
      const int yh = static_cast<int>(t % h);
      const int img = static_cast<int>(t / h);
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((static_cast<size_t>(img) * hp + yh) * wp + xw) * c + ch));
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      ss[0] = fmaf(v.x, v.x, ss[0]); ss[1] = fmaf(v.y, v.y, ss[1]); ss[2] = fmaf(v.z, v.z, ss[2]); ss[3] = fmaf(v.w, v.w, ss[3]);
    }
  }
  __shared__ float sh[2][32][33];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[0][pl][c4 * 4 + j] = s[j]; sh[1][pl][c4 * 4 + j] = ss[j]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, cc = threadIdx.x & 31;
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sh[which][i][cc];
    if (blockIdx.x * 32 + cc < c) atomicAdd(&sums[which * c + blockIdx.x * 32 + cc], t);
  }
}

// y = lrelu(x * scale[ch] + shift[ch]) over the valid region; writes NHWC bf16 [n,h,w,c] and/or the
// NCHW-flattened fp32 [n, c*h*w] the classifier consumes (architecture.py:127 x.view(B, -1)).
__global__ void bn_apply_kernel(const float* __restrict__ x, int n, int h, int w, int hp, int wp, int c,
                                const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_nchw) {
  const size_t total = static_cast<size_t>(n) * h * w * c;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    size_t t = i / c;
    const int xw = static_cast<int>(t % w); t /= w;
    const int yh = static_cast<int>(t % h);
    const int img = static_cast<int>(t / h);
    float v = x[((static_cast<size_t>(img) * hp + yh) * wp + xw) * c + ch] * scale[ch] + shift[ch];
    if (act) v = v > 0.f ? v : 0.2f * v;
    if (out_bf16) out_bf16[i] = __float2bfloat16_rn(v);
    if (out_nchw) out_nchw[((static_cast<size_t>(img) * c + ch) * h + yh) * w + xw] = v;
  }
}

// y[b, o] = act(sum_k x[b, k] * W[o, k] + bias[o]); one warp per output.
__global__ void linear_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                              float* __restrict__ y, int b, int k, int o, int act) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= b * o) return;
  const int bi = warp / o, oi = warp % o;
  const float* xr = x + static_cast<size_t>(bi) * k;
  const float* wr = wgt + static_cast<size_t>(oi) * k;
  float acc = 0.f;
  for (int i = lane; i < k; i += 32) acc = fmaf(xr[i], wr[i], acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    float v = acc + (bias ? bias[oi] : 0.f);
    if (act) v = v > 0.f ? v : 0.2f * v;
    y[static_cast<size_t>(bi) * o + oi] = v;
  }
}

}  // namespace esrp

using namespace esrp;

extern "C" {

int esrp_s2d_pad_nhwc_bf16(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c, void* stream) {
  if (!src || !dst || (c % 8) || (h % 2) || (w % 2)) return set_error("s2d_pad: c %% 8, h %% 2, w %% 2 must be 0");
  const int cv = c / 8;
  const size_t total = static_cast<size_t>(n) * (h / 2 + 1) * (w / 2 + 1) * 4 * cv;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  s2d_pad_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src),
                                                                         static_cast<uint4*>(dst), n, h, w, cv);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_bn_stats_nhwc_f32(const float* x, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, int32_t c,
                           double* sums2c, void* stream) {
  if (!x || !sums2c || h > hp || w > wp) return set_error("bn_stats: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ESRP_CUDA_OK(cudaMemsetAsync(sums2c, 0, sizeof(double) * 2 * c, s));
  const long long npx = static_cast<long long>(n) * h * w;
  // latency-bound streaming reduction: many slabs (32 pixels per block iteration), capped at 4 blocks per SM and group
  int slabs = static_cast<int>((npx + 127) / 128);
  if (slabs > 592) slabs = 592;
  if (slabs < 1) slabs = 1;
  dim3 grid((c + 31) / 32, slabs);
  bn_stats_kernel<<<grid, 256, 0, s>>>(x, n, h, w, hp, wp, c, sums2c);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_bn_apply_nhwc(const float* x, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, int32_t c,
                       const float* scale, const float* shift, int32_t act, void* out_bf16, float* out_nchw_f32,
                       void* stream) {
  if (!x || !scale || !shift || (!out_bf16 && !out_nchw_f32)) return set_error("bn_apply: bad arguments");
  const size_t total = static_cast<size_t>(n) * h * w * c;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_apply_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, h, w, hp, wp, c, scale, shift, act, static_cast<__nv_bfloat16*>(out_bf16), out_nchw_f32);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_linear_f32(const float* x, const float* w, const float* bias, float* y, int32_t b, int32_t k, int32_t o,
                    int32_t act, void* stream) {
  if (!x || !w || !y) return set_error("linear: null pointer");
  const int warps = b * o;
  const int blocks = (warps * 32 + 255) / 256;
  linear_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, b, k, o, act);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// =================================================================================================
// Backward pieces of Discriminator_VGG_128 (autograd of architecture.py:87-129 as driven by
// SRRaGAN_model.py:140,167): BatchNorm2d(train) + LeakyReLU backward as a reduction pass and an apply
// pass, the inverse of the space-to-depth rearrangement, and the two Linear layers.
// =================================================================================================
namespace esrp {

// coef layout: [7][c] fp32 = mean, rstd, scale (gamma*rstd), shift, g_rs (gamma*rstd), a (sum dzb / N), b (sum dzb*xhat / N)
__device__ __forceinline__ float bn_dzb(float z, float d, float scale, float shift) {
  const float zb = fmaf(z, scale, shift);
  return zb > 0.f ? d : 0.2f * d;
}

// sums[ch] += sum dzb, sums[c+ch] += sum dzb * xhat over the valid [h, w] region.
// dout: bf16 NHWC [n,h,w,c] (dout_nchw == NULL) or fp32 NCHW-flat [n, c*h*w].
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ z, const __nv_bfloat16* __restrict__ dout,
                                                            const float* __restrict__ dout_nchw, int n, int h, int w, int hp, int wp,
                                                            int c, const float* __restrict__ coef, double* __restrict__ sums) {
  const int c4 = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int ch = blockIdx.x * 32 + c4 * 4;
  const long long npx = static_cast<long long>(n) * h * w;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (ch < c) {
    float mean[4], rstd[4], scale[4], shift[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mean[j] = coef[ch + j]; rstd[j] = coef[c + ch + j]; scale[j] = coef[2 * c + ch + j]; shift[j] = coef[3 * c + ch + j];
    }
    for (long long pi = static_cast<long long>(blockIdx.y) * 32 + pl; pi < npx; pi += static_cast<long long>(gridDim.y) * 32) {
      const int xw = static_cast<int>(pi % w);
      const long long t = pi / w;
      const int yh = static_cast<int>(t % h);
      const int img = static_cast<int>(t / h);
      const float4 zq = __ldg(reinterpret_cast<const float4*>(z + ((static_cast<size_t>(img) * hp + yh) * wp + xw) * c + ch));
      const float zv[4] = {zq.x, zq.y, zq.z, zq.w};
      float d[4];
      if (dout_nchw) {
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = dout_nchw[((static_cast<size_t>(img) * c + ch + j) * h + yh) * w + xw];
      } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(dout + static_cast<size_t>(pi) * c + ch));
        d[0] = __uint_as_float(q.x << 16); d[1] = __uint_as_float(q.x & 0xFFFF0000u);
        d[2] = __uint_as_float(q.y << 16); d[3] = __uint_as_float(q.y & 0xFFFF0000u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dzb = bn_dzb(zv[j], d[j], scale[j], shift[j]);
        s1[j] += dzb;
        s2[j] = fmaf(dzb, (zv[j] - mean[j]) * rstd[j], s2[j]);
      }
    }
  }
  __shared__ float sh[2][32][33];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[0][pl][c4 * 4 + j] = s1[j]; sh[1][pl][c4 * 4 + j] = s2[j]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int which = threadIdx.x >> 5, cc = threadIdx.x & 31;
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sh[which][i][cc];
    if (blockIdx.x * 32 + cc < c) atomicAdd(&sums[which * c + blockIdx.x * 32 + cc], t);
  }
}

// dz[n,hp,wp,c] (bf16, the conv's output grid; zero outside the valid region) = g_rs * (dzb - a - xhat * b)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ z, const __nv_bfloat16* __restrict__ dout,
                                    const float* __restrict__ dout_nchw, int n, int h, int w, int hp, int wp, int c,
                                    const float* __restrict__ coef, __nv_bfloat16* __restrict__ dz) {
  const size_t total = static_cast<size_t>(n) * hp * wp * c;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    size_t t = i / c;
    const int xw = static_cast<int>(t % wp); t /= wp;
    const int yh = static_cast<int>(t % hp);
    const int img = static_cast<int>(t / hp);
    float v = 0.f;
    if (yh < h && xw < w) {
      const float zv = z[i];
      const float d = dout_nchw ? dout_nchw[((static_cast<size_t>(img) * c + ch) * h + yh) * w + xw]
                                : __bfloat162float(dout[((static_cast<size_t>(img) * h + yh) * w + xw) * c + ch]);
      const float dzb = bn_dzb(zv, d, coef[2 * c + ch], coef[3 * c + ch]);
      const float xhat = (zv - coef[ch]) * coef[c + ch];
      v = coef[4 * c + ch] * (dzb - coef[5 * c + ch] - xhat * coef[6 * c + ch]);
    }
    dz[i] = __float2bfloat16_rn(v);
  }
}

// Inverse of s2d_pad_kernel: din[n, y, x, ch] = ds[n, (y+1)>>1, (x+1)>>1, (((y+1)&1)*2 + ((x+1)&1))*c + ch],
// optionally times the LeakyReLU derivative selected by the sign of `ref` (the bf16 activation that was fed forward).
__global__ void s2d_pad_bwd_kernel(const uint4* __restrict__ ds, uint4* __restrict__ din, const uint4* __restrict__ ref,
                                   int n, int h, int w, int cv) {
  const int ho = h / 2 + 1, wo = w / 2 + 1;
  const size_t total = static_cast<size_t>(n) * h * w * cv;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t t = i / cv;
    const int x = static_cast<int>(t % w); t /= w;
    const int y = static_cast<int>(t % h);
    const int img = static_cast<int>(t / h);
    const int Y = (y + 1) >> 1, a = (y + 1) & 1, X = (x + 1) >> 1, b = (x + 1) & 1;
    uint4 q = __ldg(ds + ((static_cast<size_t>(img) * ho + Y) * wo + X) * 4 * cv + (a * 2 + b) * cv + v);
    if (ref) {
      const uint4 r = __ldg(ref + i);
      uint32_t qs[4] = {q.x, q.y, q.z, q.w};
      const uint32_t rs[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo = __uint_as_float(qs[j] << 16), hi = __uint_as_float(qs[j] & 0xFFFF0000u);
        const float rlo = __uint_as_float(rs[j] << 16), rhi = __uint_as_float(rs[j] & 0xFFFF0000u);
        if (!(rlo > 0.f)) lo *= 0.2f;
        if (!(rhi > 0.f)) hi *= 0.2f;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
        qs[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      q = make_uint4(qs[0], qs[1], qs[2], qs[3]);
    }
    din[i] = q;
  }
}

// Linear backward.  dym = dy * lrelu'(yout) when yout != NULL (yout = the layer's activated output).
// mode 0: dx[b,k] = sum_o dym[b,o] w[o,k];  mode 1: dw[o,k] = sum_b dym[b,o] x[b,k];  mode 2: db[o] = sum_b dym[b,o]
__global__ void linear_bwd_kernel(int mode, const float* __restrict__ dy, const float* __restrict__ yout,
                                  const float* __restrict__ x, const float* __restrict__ wgt, float* __restrict__ out,
                                  int b, int k, int o) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto dym = [&](int bi, int oi) {
    const float d = dy[static_cast<size_t>(bi) * o + oi];
    return (yout && !(yout[static_cast<size_t>(bi) * o + oi] > 0.f)) ? 0.2f * d : d;
  };
  if (mode == 0) {
    if (idx >= static_cast<size_t>(b) * k) return;
    const int bi = static_cast<int>(idx / k), ki = static_cast<int>(idx % k);
    float acc = 0.f;
    for (int oi = 0; oi < o; ++oi) acc = fmaf(dym(bi, oi), wgt[static_cast<size_t>(oi) * k + ki], acc);
    out[idx] = acc;
  } else if (mode == 1) {
    if (idx >= static_cast<size_t>(o) * k) return;
    const int oi = static_cast<int>(idx / k), ki = static_cast<int>(idx % k);
    float acc = 0.f;
    for (int bi = 0; bi < b; ++bi) acc = fmaf(dym(bi, oi), x[static_cast<size_t>(bi) * k + ki], acc);
    out[idx] = acc;
  } else {
    if (idx >= static_cast<size_t>(o)) return;
    float acc = 0.f;
    for (int bi = 0; bi < b; ++bi) acc += dym(bi, static_cast<int>(idx));
    out[idx] = acc;
  }
}

}  // namespace esrp

extern "C" {

int esrp_bn_bwd_reduce(const float* z, const void* dout_bf16, const float* dout_nchw_f32, int32_t n, int32_t h, int32_t w,
                       int32_t hp, int32_t wp, int32_t c, const float* coef7c, double* sums2c, void* stream) {
  if (!z || (!dout_bf16 && !dout_nchw_f32) || !coef7c || !sums2c || h > hp || w > wp) return esrp::set_error("bn_bwd_reduce: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ESRP_CUDA_OK(cudaMemsetAsync(sums2c, 0, sizeof(double) * 2 * c, s));
  const long long npx = static_cast<long long>(n) * h * w;
  // latency-bound streaming reduction: many slabs (32 pixels per block iteration), capped at 4 blocks per SM and group
  int slabs = static_cast<int>((npx + 127) / 128);
  if (slabs > 592) slabs = 592;
  if (slabs < 1) slabs = 1;
  dim3 grid((c + 31) / 32, slabs);
  esrp::bn_bwd_reduce_kernel<<<grid, 256, 0, s>>>(z, static_cast<const __nv_bfloat16*>(dout_bf16), dout_nchw_f32, n, h, w, hp,
                                                  wp, c, coef7c, sums2c);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_bn_bwd_apply(const float* z, const void* dout_bf16, const float* dout_nchw_f32, int32_t n, int32_t h, int32_t w,
                      int32_t hp, int32_t wp, int32_t c, const float* coef7c, void* dz_bf16, void* stream) {
  if (!z || (!dout_bf16 && !dout_nchw_f32) || !coef7c || !dz_bf16 || h > hp || w > wp) return esrp::set_error("bn_bwd_apply: bad arguments");
  const size_t total = static_cast<size_t>(n) * hp * wp * c;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  esrp::bn_bwd_apply_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, static_cast<const __nv_bfloat16*>(dout_bf16), dout_nchw_f32, n, h, w, hp, wp, c, coef7c,
      static_cast<__nv_bfloat16*>(dz_bf16));
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_s2d_pad_bwd_nhwc_bf16(const void* ds, void* din, const void* lrelu_ref, int32_t n, int32_t h, int32_t w, int32_t c,
                               void* stream) {
  if (!ds || !din || (c % 8) || (h % 2) || (w % 2)) return esrp::set_error("s2d_pad_bwd: c %% 8, h %% 2, w %% 2 must be 0");
  const int cv = c / 8;
  const size_t total = static_cast<size_t>(n) * h * w * cv;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  esrp::s2d_pad_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(ds), static_cast<uint4*>(din), static_cast<const uint4*>(lrelu_ref), n, h, w, cv);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_linear_bwd_f32(const float* dy, const float* yout_act, const float* x, const float* w, float* dx, float* dw, float* db,
                        int32_t b, int32_t k, int32_t o, void* stream) {
  if (!dy || !x || !w) return esrp::set_error("linear_bwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dx) esrp::linear_bwd_kernel<<<(static_cast<size_t>(b) * k + 255) / 256, 256, 0, s>>>(0, dy, yout_act, x, w, dx, b, k, o);
  if (dw) esrp::linear_bwd_kernel<<<(static_cast<size_t>(o) * k + 255) / 256, 256, 0, s>>>(1, dy, yout_act, x, w, dw, b, k, o);
  if (db) esrp::linear_bwd_kernel<<<(o + 255) / 256, 256, 0, s>>>(2, dy, yout_act, x, w, db, b, k, o);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// BatchNorm2d bookkeeping on [C]-sized vectors, one launch each (instead of a dozen framework ops per layer):
// -------------------------------------------------------------------------------------------------
namespace esrp {

// sums2c = (sum, sum of squares) over `count` samples per channel (esrp_bn_stats_nhwc_f32).
// training != 0: batch statistics (biased variance for the normalisation, unbiased for running_var,
//                torch.nn.BatchNorm2d semantics, block.py:32) + running-stat update with `momentum`;
// training == 0: running statistics.  coef7c rows 0..4 <- mean, rstd, scale, shift, g_rs (rows 5, 6 zeroed).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int training,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int c,
                                   float* __restrict__ coef) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double mean, var;
  if (training) {
    mean = sums[ch] / count;
    var = sums[c + ch] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean != nullptr) {
      running_mean[ch] = static_cast<float>((1.0 - momentum) * running_mean[ch] + momentum * mean);
      const double unbiased = var * (count / (count > 1.0 ? count - 1.0 : 1.0));
      running_var[ch] = static_cast<float>((1.0 - momentum) * running_var[ch] + momentum * unbiased);
    }
  } else {
    mean = running_mean[ch];
    var = running_var[ch];
  }
  const double rstd = rsqrt(var + static_cast<double>(eps));
  const double g = gamma ? gamma[ch] : 1.0, b = beta ? beta[ch] : 0.0;
  coef[ch] = static_cast<float>(mean);
  coef[c + ch] = static_cast<float>(rstd);
  coef[2 * c + ch] = static_cast<float>(g * rstd);
  coef[3 * c + ch] = static_cast<float>(b - mean * g * rstd);
  coef[4 * c + ch] = static_cast<float>(g * rstd);
  coef[5 * c + ch] = 0.f;
  coef[6 * c + ch] = 0.f;
}

// After esrp_bn_bwd_reduce: d_gamma = sum dzb*xhat, d_beta = sum dzb; batch-statistics layers also get the
// a = sum dzb / N, b = sum dzb*xhat / N rows of coef7c that the apply pass subtracts.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double count, int batch_stats, int c,
                                       float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  if (batch_stats) {
    coef[5 * c + ch] = static_cast<float>(sums[ch] / count);
    coef[6 * c + ch] = static_cast<float>(sums[c + ch] / count);
  }
  if (dgamma) dgamma[ch] = static_cast<float>(sums[c + ch]);
  if (dbeta) dbeta[ch] = static_cast<float>(sums[ch]);
}

}  // namespace esrp

extern "C" {

int esrp_bn_finalize(const double* sums2c, double count, const float* gamma, const float* beta, float eps, float momentum,
                     int32_t training, float* running_mean, float* running_var, int32_t c, float* coef7c, void* stream) {
  if (!coef7c || (training && !sums2c) || (!training && (!running_mean || !running_var))) return esrp::set_error("bn_finalize: bad arguments");
  esrp::bn_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sums2c, count, gamma, beta, eps, momentum, training, running_mean, running_var, c, coef7c);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_bn_bwd_finalize(const double* sums2c, double count, int32_t batch_stats, int32_t c, float* coef7c, float* dgamma,
                         float* dbeta, void* stream) {
  if (!sums2c || !coef7c) return esrp::set_error("bn_bwd_finalize: bad arguments");
  esrp::bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sums2c, count, batch_stats, c,
                                                                                                coef7c, dgamma, dbeta);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// Fork / join over a small pool of side streams: the output-channel slices of a deep discriminator layer are
// independent launches that each fill only 7-70 of the 148 SMs; issued on different streams they run concurrently.
// -------------------------------------------------------------------------------------------------
namespace esrp {
static cudaEvent_t next_event() {
  static thread_local cudaEvent_t ring[64];
  static thread_local int pos = 0, made = 0;
  if (made < 64) {
    if (cudaEventCreateWithFlags(&ring[made], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return ring[made++];
  }
  // a recorded event may be re-recorded as soon as the waits on it have been ENQUEUED (they captured its state)
  cudaEvent_t e = ring[pos];
  pos = (pos + 1) & 63;
  return e;
}
}  // namespace esrp

extern "C" {

int esrp_streams_create(int32_t n, void** out_streams) {
  if (!out_streams || n < 1 || n > 16) return esrp::set_error("streams_create: n must be in 1..16");
  for (int i = 0; i < n; ++i) {
    cudaStream_t s;
    ESRP_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    out_streams[i] = s;
  }
  return 0;
}

int esrp_streams_fork(void* main_stream, void* const* sides, int32_t n) {
  cudaEvent_t e = esrp::next_event();
  if (!e) return esrp::set_error("streams_fork: cudaEventCreate failed");
  ESRP_CUDA_OK(cudaEventRecord(e, static_cast<cudaStream_t>(main_stream)));
  for (int i = 0; i < n; ++i) ESRP_CUDA_OK(cudaStreamWaitEvent(static_cast<cudaStream_t>(sides[i]), e, 0));
  return 0;
}

int esrp_streams_join(void* main_stream, void* const* sides, int32_t n) {
  for (int i = 0; i < n; ++i) {
    cudaEvent_t e = esrp::next_event();
    if (!e) return esrp::set_error("streams_join: cudaEventCreate failed");
    ESRP_CUDA_OK(cudaEventRecord(e, static_cast<cudaStream_t>(sides[i])));
    ESRP_CUDA_OK(cudaStreamWaitEvent(static_cast<cudaStream_t>(main_stream), e, 0));
  }
  return 0;
}

}  // extern "C"
