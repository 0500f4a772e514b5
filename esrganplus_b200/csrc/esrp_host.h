// esrp_host.h — host-side helpers shared by the translation units of libesrp.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"
#include "conv_params.h"

namespace esrp {
// printf-style; stores a thread-local message for esrp_last_error() and returns 1.
int set_error(const char* fmt, ...);
// Encode a 4-D TMA map over an NHWC bf16 tensor; returns 0 on success.
int make_nhwc_tmap(CUtensorMap* tm, const void* ptr, int n, int h, int w, int c_total, int kc,
                   int box_w, int box_h);
int sm_count();  // of the CURRENT device (cached per device ordinal)
// Opt the kernel in to `bytes` of dynamic shared memory on the CURRENT device (the attribute is per device: cached per
// (device, kernel), so a second GPU in the same process gets its own call).  Returns 0 on success.
int ensure_max_smem(const void* kernel, int bytes = 232448);

// A fully planned conv launch: kernel instantiation, TMA maps, parameters.  Planning (argument
// validation, tensor-map encoding, shared-memory/TMEM budgeting) happens once per shape; replay is
// a single cudaLaunchKernel, so a whole-network forward is a flat list of these.
struct ConvLaunch {
  const void* kernel = nullptr;
  CUtensorMap tm0, tm1;
  ConvKParams params;
  int grid = 0;
  int threads = 0;
  int smem = 0;
  // which kernel family / instantiation `kernel` is (the chain builder merges compatible row-kernel launches)
  int fam = 0;  // 0: tile kernel (conv3x3_tc.cuh), 1: row kernel (conv3x3_row.cuh)
  int kc = 0, bn = 0, ext = 0;
  int cluster = 1;  // CTAs per thread-block cluster (2: the row kernel's cta_group::2 pairs)
};
int plan_conv(const esrp_conv3x3_t& d, ConvLaunch* out);
// per-family planners (esrp_conv_{row,tile}{,_ext}.cu); *_ext carry the training extensions of the fused tail
int plan_row_base(const esrp_conv3x3_t& d, ConvLaunch* out);
int plan_row_ext(const esrp_conv3x3_t& d, ConvLaunch* out);
int plan_tile_base(const esrp_conv3x3_t& d, ConvLaunch* out);
int plan_tile_ext(const esrp_conv3x3_t& d, ConvLaunch* out);
void copy_common(const esrp_conv3x3_t& d, ConvKParams* pp);
// A persistent chain of row-kernel convs executed by ONE launch (conv3x3_chain.cuh): the phase table (tensor maps +
// a compact record per conv) is the kernel's parameter block, kept on the host; the per-CTA completion flags live in
// device memory owned by this object.
struct ChainLaunch {
  const void* kernel = nullptr;
  void* args_host = nullptr;   // ChainArgs (malloc)
  unsigned int* dev_flags = nullptr;
  int num_phases = 0;
  int dep_all = 0;
  int grid = 0, threads = 0, smem = 0;
};
constexpr int kChainMaxPhasesHost = 440;  // == kChainMaxPhases (conv3x3_chain.cuh): phases per chain launch
// true when the launch can be a phase of a chain (row kernel, kc = 64, bn = 32, NHWC outputs only)
bool chain_compatible(const ConvLaunch& L);
int plan_chain(const ConvLaunch* const* convs, int count, ChainLaunch* out);
int run_chain(const ChainLaunch& L, cudaStream_t stream);
void free_chain(ChainLaunch* L);
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100
int run_conv(const ConvLaunch& L, cudaStream_t stream);

#define ESRP_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return ::esrp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
}  // namespace esrp
