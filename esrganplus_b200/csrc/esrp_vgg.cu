// esrp_vgg.cu — the element-wise pieces of the perceptual branch: `VGGFeatureExtractor`
// codes/models/modules/architecture.py:279-307 = input normalisation (:304-305) + torchvision vgg19().features[:35]
// (16 x [3x3 conv, ReLU] with four 2x2 max-pools, cut before the last ReLU).  The convs run on the tcgen05 conv kernels
// (act = 2); what is left is HBM-bound bookkeeping on NHWC bf16 tensors, 16 bytes (8 channels) per thread access:
//   * esrp_nchw_f32_to_nhwc_bf16_affine: (x - mean) / std while the image is re-laid out and padded to 32 channels,
//   * esrp_maxpool2x2_nhwc_bf16: nn.MaxPool2d(2, 2),
//   * esrp_relu_bwd_nhwc_bf16: gradient through a ReLU from its OUTPUT (y > 0 <=> pre-activation > 0),
//   * esrp_maxpool2x2_relu_bwd_nhwc_bf16: gradient through [ReLU, MaxPool2d(2,2)] in one pass: the window's first maximum
//     (torch's scan order: (0,0), (0,1), (1,0), (1,1)) receives the pooled gradient if it is positive.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/esrp.h"
#include "esrp_host.h"

namespace esrp {

namespace {

__global__ void nchw_to_nhwc_affine_kernel(const float* __restrict__ src, const float* __restrict__ scale,
                                           const float* __restrict__ shift, __nv_bfloat16* __restrict__ dst, int c, long long hw,
                                           int c_pad, long long total_px) {
  // one thread per pixel and group of 8 output channels (a 16-byte store); reads are coalesced along the pixel index
  const int groups = c_pad / 8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total_px * groups; t += stride) {
    const long long px = t % total_px;   // pixel-major inside a group: consecutive threads read consecutive floats
    const int g = static_cast<int>(t / total_px);
    const long long img = px / hw, q = px - img * hw;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = g * 8 + i;
      float x = 0.f;
      if (ch < c) {
        x = src[(img * c + ch) * hw + q];
        if (scale) x = x * scale[ch] + shift[ch];
      }
      v[i] = __float2bfloat16(x);
    }
    *reinterpret_cast<uint4*>(dst + px * c_pad + g * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__global__ void maxpool2x2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int c8) {
  const int ho = h / 2, wo = w / 2;
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int g = static_cast<int>(t % c8);
    long long r = t / c8;
    const int xo = static_cast<int>(r % wo);
    r /= wo;
    const int yo = static_cast<int>(r % ho);
    const int img = static_cast<int>(r / ho);
    const long long base = ((static_cast<long long>(img) * h + 2 * yo) * w + 2 * xo) * c8 + g;
    float a[8], b[8], m[8];
    unpack8(x[base], m);
    unpack8(x[base + c8], a);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], a[i]);
    unpack8(x[base + static_cast<long long>(w) * c8], a);
    unpack8(x[base + static_cast<long long>(w) * c8 + c8], b);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], fmaxf(a[i], b[i]));
    y[t] = pack8(m);
  }
}

__global__ void relu_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy, uint4* __restrict__ dz, long long n8) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < n8; t += stride) {
    float a[8], d[8];
    unpack8(y[t], a);
    unpack8(dy[t], d);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = a[i] > 0.f ? d[i] : 0.f;
    dz[t] = pack8(d);
  }
}

__global__ void maxpool2x2_relu_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dpool, uint4* __restrict__ dz, int n,
                                           int h, int w, int c8) {
  const int ho = h / 2, wo = w / 2;
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int g = static_cast<int>(t % c8);
    long long r = t / c8;
    const int xo = static_cast<int>(r % wo);
    r /= wo;
    const int yo = static_cast<int>(r % ho);
    const int img = static_cast<int>(r / ho);
    const long long p00 = ((static_cast<long long>(img) * h + 2 * yo) * w + 2 * xo) * c8 + g;
    const long long p01 = p00 + c8, p10 = p00 + static_cast<long long>(w) * c8, p11 = p10 + c8;
    float v00[8], v01[8], v10[8], v11[8], d[8];
    unpack8(y[p00], v00);
    unpack8(y[p01], v01);
    unpack8(y[p10], v10);
    unpack8(y[p11], v11);
    unpack8(dpool[t], d);
    float o00[8], o01[8], o10[8], o11[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // first maximum in scan order (a later element wins only if strictly greater), then the ReLU mask of the winner
      int k = 0;
      float m = v00[i];
      if (v01[i] > m) { m = v01[i]; k = 1; }
      if (v10[i] > m) { m = v10[i]; k = 2; }
      if (v11[i] > m) { m = v11[i]; k = 3; }
      const float gsel = m > 0.f ? d[i] : 0.f;
      o00[i] = k == 0 ? gsel : 0.f;
      o01[i] = k == 1 ? gsel : 0.f;
      o10[i] = k == 2 ? gsel : 0.f;
      o11[i] = k == 3 ? gsel : 0.f;
    }
    dz[p00] = pack8(o00);
    dz[p01] = pack8(o01);
    dz[p10] = pack8(o10);
    dz[p11] = pack8(o11);
  }
}

int grid_for(long long work_items) {
  const int sms = sm_count();
  long long blocks = (work_items + 255) / 256;
  const long long cap = static_cast<long long>(sms > 0 ? sms : 148) * 8;   // a multiple of the SM count, grid-stride loops
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace esrp

using namespace esrp;

extern "C" {

int esrp_nchw_f32_to_nhwc_bf16_affine(const float* src, const float* scale, const float* shift, void* dst, int32_t n, int32_t c,
                                      int32_t h, int32_t w, int32_t c_pad, void* stream) {
  if (!src || !dst || n < 1 || c < 1 || h < 1 || w < 1 || c_pad < c || (c_pad % 8)) return set_error("nchw_f32_to_nhwc_bf16_affine: bad arguments");
  if ((scale == nullptr) != (shift == nullptr)) return set_error("nchw_f32_to_nhwc_bf16_affine: scale and shift go together");
  const long long hw = static_cast<long long>(h) * w, px = hw * n;
  nchw_to_nhwc_affine_kernel<<<grid_for(px * (c_pad / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, scale, shift, static_cast<__nv_bfloat16*>(dst), c, hw, c_pad, px);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_maxpool2x2_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream) {
  if (!x || !y || n < 1 || h < 2 || w < 2 || (h % 2) || (w % 2) || c < 8 || (c % 8)) return set_error("maxpool2x2_nhwc_bf16: bad arguments (even h, w; c %% 8 == 0)");
  maxpool2x2_kernel<<<grid_for(static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), n, h, w, c / 8);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_relu_bwd_nhwc_bf16(const void* y, const void* dy, void* dz, int64_t count, void* stream) {
  if (!y || !dy || !dz || count < 8 || (count % 8)) return set_error("relu_bwd_nhwc_bf16: bad arguments (count %% 8 == 0)");
  relu_bwd_kernel<<<grid_for(count / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(y), static_cast<const uint4*>(dy),
                                                                                     static_cast<uint4*>(dz), count / 8);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_maxpool2x2_relu_bwd_nhwc_bf16(const void* y, const void* dpool, void* dz, int32_t n, int32_t h, int32_t w, int32_t c,
                                       void* stream) {
  if (!y || !dpool || !dz || n < 1 || h < 2 || w < 2 || (h % 2) || (w % 2) || c < 8 || (c % 8))
    return set_error("maxpool2x2_relu_bwd_nhwc_bf16: bad arguments (even h, w; c %% 8 == 0)");
  maxpool2x2_relu_bwd_kernel<<<grid_for(static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(y), static_cast<const uint4*>(dpool), static_cast<uint4*>(dz), n, h, w, c / 8);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
