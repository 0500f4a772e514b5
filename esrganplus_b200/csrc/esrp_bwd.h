// esrp_bwd.h — host-side launch records of the backward kernels (esrp_bwd.cu), shared with the engines.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"

namespace esrp {

// tcgen05 weight-gradient kernel (esrp_wgrad_tc.cu): a job = one 64-channel chunk of X against one 128-channel slab of dY
constexpr int kTcMaxJobs = 64;
constexpr int kTcMaxMaps = 4;   // distinct X / dY tensors per launch (jobs share tensor maps)
constexpr int kMmaMaxUnits = 16;  // units per launch of the mma.sync kernel
struct WgTcJobInfo {
  int xc0, dyc0;       // first channel of the X chunk / of the dY slab
  short y_panels;      // 64-channel panels of the slab that exist in the tensor (1 or 2)
  short mx, my;        // tensor-map indices
  short pad_;
  float* acc[2][2];    // [32-channel group of the chunk][64-column block of the slab] unit accumulators, or nullptr
};
struct alignas(64) WgTcParams {
  CUtensorMap tmx[kTcMaxMaps];
  CUtensorMap tmy[kTcMaxMaps];
  WgTcJobInfo job[kTcMaxJobs];
  int num_jobs, splits;
  int n, h, w;
  int tw, tw_log2, tr, xw;   // tile = tr rows x tw columns (tr * tw == 128); xw = tw + 2
  int tiles_x, tiles_y, tiles_total;
  int x_bytes, stage_bytes, stages;
};

struct WgradLaunch {
  alignas(64) unsigned char params[sizeof(WgTcParams) > 1024 ? sizeof(WgTcParams) : 1024];
  int grid = 0;
  int smem = 0;
  int tc = 0;                // 1: tcgen05 kernel (params holds WgTcParams), 0: mma.sync kernel (WgradParams)
  struct Bias { const void* dy; int ctotal, c0; float* out; };
  Bias bias[8];
  int num_bias = 0;
  long long npx = 0;
};
int plan_wgrad_tc(const esrp_wgrad_unit_t* units, int num_units, int n, int h, int w, WgradLaunch* out);
int run_wgrad_tc(const WgradLaunch& L, cudaStream_t stream);
int plan_wgrad(const esrp_wgrad_unit_t* units, int num_units, int n, int h, int w, int splits, WgradLaunch* out);
int run_wgrad(const WgradLaunch& L, cudaStream_t stream);
int run_conv1x1_bwd(int nf, const void* x, int x_ctotal, const void* dx2, int d_ctotal, int d_c0, const float* u,
                    float* g, const float* extra, float* du_acc, long long npx, cudaStream_t stream);
int run_scatter(const esrp_scatter_entry_t* tab_dev, int num, float* const* dst_ptrs_dev, cudaStream_t stream);

}  // namespace esrp
