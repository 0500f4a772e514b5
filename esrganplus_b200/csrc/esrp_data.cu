// esrp_data.cu — the training data path of one minibatch on the device (reference:
// codes/data/LRHR_dataset.py:83-121 with on-the-fly LR, codes/data/util.py:94-106 `augment`, :211-274 / :345-412
// MATLAB-style antialiased bicubic `imresize_np`).  Per sample the reference resizes the WHOLE HR image on one CPU core
// (two passes of out_len mat-vec products), crops 32x32 / 128x128, flips / rotates and transposes.  Here one CTA per
// sample computes exactly the LR pixels of the crop (the whole-image result restricted to the crop: same weights, same
// symmetric boundary) and gathers the HR crop, with the augmentation, the BGR->RGB swap and the HWC->CHW transpose folded
// into the addressing.  HBM-bound: ~150 KB read per sample.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/esrp.h"
#include "esrp_host.h"

namespace esrp {

// padded coordinate a of the reference's symmetric copy (util.py:371-383) -> coordinate in the image
__device__ __forceinline__ int reflect_index(int a, int sym_s, int n) {
  if (a < sym_s) return sym_s - 1 - a;
  a -= sym_s;
  return a < n ? a : n - 1 - (a - n);
}

// destination (y, x) of the augmented crop -> source (y, x) inside the crop (util.py:100-104: hflip, vflip, then transpose)
__device__ __forceinline__ void augment_source(int y, int x, int size, int hflip, int vflip, int rot90, int* sy, int* sx) {
  int a = rot90 ? x : y, b = rot90 ? y : x;
  if (vflip) a = size - 1 - a;
  if (hflip) b = size - 1 - b;
  *sy = a;
  *sx = b;
}

constexpr int kDataRows = 8;  // LR rows of the crop per pass through shared memory

__global__ void lrhr_batch_kernel(const esrp_lrhr_job_t* __restrict__ jobs, int scale, int hr_size, int xw_max,
                                  float* __restrict__ lr_out, float* __restrict__ hr_out) {
  extern __shared__ float tmp[];  // [kDataRows][xw_max][3]: the vertical pass over the columns the crop's LR pixels read
  const esrp_lrhr_job_t J = jobs[blockIdx.x];
  const int lr_size = hr_size / scale;
  float* lr = lr_out + static_cast<size_t>(blockIdx.x) * 3 * lr_size * lr_size;
  float* hr = hr_out + static_cast<size_t>(blockIdx.x) * 3 * hr_size * hr_size;
  // ---- HR crop: gather with augmentation, BGR -> RGB, HWC -> CHW, / 255 (LRHR_dataset.py:103-105,116-121) ----
  for (int i = threadIdx.x; i < 3 * hr_size * hr_size; i += blockDim.x) {
    const int x = i % hr_size, y = (i / hr_size) % hr_size, c = i / (hr_size * hr_size);
    int sy, sx;
    augment_source(y, x, hr_size, J.hflip, J.vflip, J.rot90, &sy, &sx);
    const size_t src = (static_cast<size_t>(J.rnd_h * scale + sy) * J.w + (J.rnd_w * scale + sx)) * 3 + (2 - c);
    hr[i] = static_cast<float>(J.img[src]) / 255.0f;
  }
  // ---- LR crop = rows [rnd_h, +lr_size) x columns [rnd_w, +lr_size) of imresize_np(img / 255, 1 / scale) ----
  const int xa0 = J.iw[J.rnd_w];                                   // first padded column any LR pixel of the crop reads
  const int xw = J.iw[J.rnd_w + lr_size - 1] + J.pw - xa0;         // padded columns spanned (<= xw_max)
  for (int r0 = 0; r0 < lr_size; r0 += kDataRows) {
    const int rows = min(kDataRows, lr_size - r0);
    __syncthreads();
    // H pass (util.py:385-391): tmp[r][xa][c] = sum_k wh[i][k] * img_aug[ih[i] + k][x][c]
    for (int t = threadIdx.x; t < rows * xw * 3; t += blockDim.x) {
      const int c = t % 3, xa = (t / 3) % xw, r = t / (3 * xw);
      const int i = J.rnd_h + r0 + r;
      const int x = reflect_index(xa0 + xa, J.sym_ws, J.w);
      const float* wrow = J.wh + static_cast<size_t>(i) * J.ph;
      const int ya = J.ih[i];
      float acc = 0.f;
      for (int k = 0; k < J.ph; ++k) {
        const int y = reflect_index(ya + k, J.sym_hs, J.h);
        acc += wrow[k] * (static_cast<float>(J.img[(static_cast<size_t>(y) * J.w + x) * 3 + c]) / 255.0f);
      }
      tmp[(r * xw_max + xa) * 3 + c] = acc;
    }
    __syncthreads();
    // W pass (util.py:405-410) + augmentation + channel swap + transpose
    for (int t = threadIdx.x; t < rows * lr_size * 3; t += blockDim.x) {
      const int c = t % 3, jj = (t / 3) % lr_size, r = t / (3 * lr_size);
      const int j = J.rnd_w + jj;
      const float* wrow = J.ww + static_cast<size_t>(j) * J.pw;
      const int xb = J.iw[j] - xa0;
      float acc = 0.f;
      for (int k = 0; k < J.pw; ++k) acc += tmp[(r * xw_max + xb + k) * 3 + c] * wrow[k];
      // (r0 + r, jj) is the SOURCE position inside the crop; find where the augmentation puts it
      int a = r0 + r, b = jj;
      if (J.hflip) b = lr_size - 1 - b;
      if (J.vflip) a = lr_size - 1 - a;
      const int dy = J.rot90 ? b : a, dx = J.rot90 ? a : b;
      lr[(static_cast<size_t>(2 - c) * lr_size + dy) * lr_size + dx] = acc;
    }
  }
}

}  // namespace esrp

using namespace esrp;

extern "C" {

int esrp_lrhr_batch(const esrp_lrhr_job_t* jobs_dev, int32_t count, int32_t scale, int32_t hr_size, int32_t xw_max, float* lr_out,
                    float* hr_out, void* stream) {
  if (!jobs_dev || !lr_out || !hr_out || count < 0 || scale < 1 || hr_size < scale || (hr_size % scale) || xw_max < 1)
    return set_error("lrhr_batch: bad arguments");
  if (count == 0) return 0;
  const size_t smem = static_cast<size_t>(kDataRows) * xw_max * 3 * sizeof(float);
  if (smem > 48 * 1024) return set_error("lrhr_batch: xw_max=%d needs %zu bytes of shared memory (> 48 KB)", xw_max, smem);
  lrhr_batch_kernel<<<count, 256, smem, static_cast<cudaStream_t>(stream)>>>(jobs_dev, scale, hr_size, xw_max, lr_out, hr_out);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int32_t esrp_sizeof_lrhr_job(void) { return static_cast<int32_t>(sizeof(esrp_lrhr_job_t)); }

}  // extern "C"
