// esrp_pack.cu — batched weight repack kernel (see esrp_pack.h).  Same element maps as pack_conv_weights_kernel
// (esrp_api.cu) and pack_dgrad_kernel (esrp_bwd.cu): out[chunk][outer][row = blk*bn + r][kc], rows pre-swizzled.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "esrp_host.h"
#include "esrp_pack.h"

namespace esrp {

namespace {
constexpr int kBlocksPerJob = 8;

__device__ __forceinline__ int swz(int row, int k, int kc) {
  const int chunk16 = k >> 3;
  const int x = (kc == 64) ? (row & 7) : ((row >> 1) & 3);
  return row * kc + (((chunk16 ^ x) << 3) | (k & 7));
}

__global__ void pack_batch_kernel(const PackJob* __restrict__ jobs) {
  const PackJob& J = jobs[blockIdx.y];
  const int kc = J.kc, bn = J.bn;
  const bool has_aux = J.type == 0 && J.aux_chunks > 0;
  const int nb_rows = (has_aux ? 4 : 3) * bn;
  const int total = J.num_chunks * 3 * nb_rows * kc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % kc;
    const int row = (i / kc) % nb_rows;
    const int outer = (i / (kc * nb_rows)) % 3;
    const int chunk = i / (kc * nb_rows * 3);
    const int blk = row / bn;
    const int r = row - blk * bn;
    const int ky = J.layout == ESRP_LAYOUT_ROW ? blk : outer;
    const int kx = J.layout == ESRP_LAYOUT_ROW ? outer : blk;
    float v = 0.f;
    if (r < J.rows) {
      if (J.type == 0) {
        const int ci = J.lc0[chunk] + k;
        if (blk < 3) {
          if (J.row0 + r < J.w_o && ci < J.w_i) v = J.w[((static_cast<size_t>(J.row0 + r) * J.w_i + ci) * 3 + ky) * 3 + kx];
        } else if (outer == 1 && chunk < J.aux_chunks && J.aux != nullptr) {
          if (J.row0 + r < J.w_o && ci < J.aux_cin) v = J.aux[static_cast<size_t>(J.row0 + r) * J.aux_cin + ci];
        }
      } else {
        const int kk = chunk * kc + k;
        const int gi = kk >> 5;
        if (gi < J.num_groups) {
          const esrp_dgrad_group_t& g = J.g[gi];
          const int co = g.co0 + (kk & 31);
          const int ci = J.row0 + r;
          if (g.w != nullptr && co < g.w_o && ci < g.w_i)
            v = g.scale * g.w[((static_cast<size_t>(co) * g.w_i + ci) * 3 + (2 - ky)) * 3 + (2 - kx)];
        }
      }
    }
    J.out[static_cast<size_t>(chunk * 3 + outer) * nb_rows * kc + swz(row, k, kc)] = __float2bfloat16_rn(v);
  }
  if (J.type == 0 && J.bias_dst != nullptr && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < bn; i += blockDim.x)
      J.bias_dst[i] = (J.bias_src != nullptr && i < J.rows && J.row0 + i < J.w_o) ? J.bias_src[J.row0 + i] : 0.f;
  }
}
}  // namespace

int run_pack_batch(const PackJob* jobs_dev, int num_jobs, cudaStream_t stream) {
  if (num_jobs < 1) return 0;
  pack_batch_kernel<<<dim3(kBlocksPerJob, num_jobs), 256, 0, stream>>>(jobs_dev);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace esrp
