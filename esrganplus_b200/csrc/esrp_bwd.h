// esrp_bwd.h — host-side launch records of the backward kernels (esrp_bwd.cu), shared with the engines.
#pragma once
#include <cuda_runtime.h>

#include "../../include/esrp.h"

namespace esrp {

struct WgradLaunch {
  alignas(16) unsigned char params[ESRP_WGRAD_MAX_UNITS * sizeof(esrp_wgrad_unit_t) + 128];
  int grid = 0;
  int smem = 0;
};
int plan_wgrad(const esrp_wgrad_unit_t* units, int num_units, int n, int h, int w, int splits, WgradLaunch* out);
int run_wgrad(const WgradLaunch& L, cudaStream_t stream);
int run_conv1x1_bwd(int nf, const void* x, int x_ctotal, const void* dx2, int d_ctotal, int d_c0, const float* u,
                    float* g, const float* extra, float* du_acc, long long npx, cudaStream_t stream);
int run_scatter(const esrp_scatter_entry_t* tab_dev, int num, float* const* dst_ptrs_dev, cudaStream_t stream);

}  // namespace esrp
