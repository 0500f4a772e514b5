// esrp_conv_chain.cu — planning + launch of the persistent conv chain (conv3x3_chain.cuh): a run of row-kernel conv
// launches of one engine plan (the 5 convs of every ResidualDenseBlock_5C of the RRDB trunk, block.py:260-291) becomes
// ONE kernel launch whose phases are the original launches.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/esrp.h"
#include "conv3x3_chain.cuh"
#include "esrp_host.h"

namespace esrp {

bool chain_compatible(const ConvLaunch& L) {
  const ConvKParams& p = L.params;
  return L.fam == 1 && L.kc == 64 && L.bn == 32 && L.ext == 0 && p.out_nchw == nullptr && p.trace == nullptr && p.dbg == 0 &&
         p.w_resident == 1 && L.grid >= 1;
}

int plan_chain(const ConvLaunch* const* convs, int count, ChainLaunch* out) {
  if (count < 1) return set_error("conv chain: empty");
  std::vector<ChainPhase> host(static_cast<size_t>(count));
  const ConvKParams& p0 = convs[0]->params;
  int smem = 0;
  for (int i = 0; i < count; ++i) {
    const ConvLaunch& L = *convs[i];
    const ConvKParams& p = L.params;
    if (!chain_compatible(L)) return set_error("conv chain: phase %d is not a kc=64 bn=32 row-kernel conv", i);
    if (p.n != p0.n || p.h != p0.h || p.w != p0.w || p.units_total != p0.units_total || L.grid != convs[0]->grid)
      return set_error("conv chain: phase %d has another shape / grid", i);
    const int nsl = p.nsl > 1 ? p.nsl : 1;
    if (L.grid % nsl) return set_error("conv chain: grid %d is not a multiple of the %d co-scheduled slices of phase %d", L.grid, nsl, i);
    ChainPhase& ph = host[static_cast<size_t>(i)];
    memset(&ph, 0, sizeof(ph));
    ph.tm0 = L.tm0;
    ph.tm1 = L.tm1;
    ph.p = p;
    // row buffers: as many as fit beside the resident weights (the chain's producer has dedicated buffer barriers, so
    // the ring-size condition of conv3x3_row.cuh does not apply)
    const int w_all = p.num_chunks * 3 * (p.aux_chunks > 0 ? 4 : 3) * 32 * 128;
    const int row_bytes = p.a_stage_bytes * p.num_chunks;
    int nbuf = (kMaxSmem - kSmemFixed - 1024 - w_all) / row_bytes;
    if (nbuf > kMaxStages) nbuf = kMaxStages;
    if (nbuf < 2) return set_error("conv chain: phase %d: weights + 2 row buffers do not fit in shared memory", i);
    ph.p.stages = nbuf;
    const int need = kSmemFixed + 1024 + w_all + nbuf * row_bytes;
    if (need > smem) smem = need;
  }
  auto kern = conv3x3_chain_kernel<64, 32, false>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern))) return 1;
  out->kernel = reinterpret_cast<const void*>(kern);
  out->num_phases = count;
  out->grid = convs[0]->grid;
  out->threads = kRowThreads;
  out->smem = smem;
  out->dep_all = p0.x_tiles > 1 ? 1 : 0;  // a 130-pixel box reads one pixel of the neighbouring column block
  if (out->grid > sm_count()) return set_error("conv chain: grid %d exceeds the %d SMs (all CTAs must be co-resident)", out->grid, sm_count());
  ESRP_CUDA_OK(cudaMalloc(&out->dev_phases, sizeof(ChainPhase) * static_cast<size_t>(count)));
  ESRP_CUDA_OK(cudaMemcpy(out->dev_phases, host.data(), sizeof(ChainPhase) * static_cast<size_t>(count), cudaMemcpyHostToDevice));
  ESRP_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&out->dev_flags), sizeof(unsigned int) * static_cast<size_t>(out->grid)));
  return 0;
}

void free_chain(ChainLaunch* L) {
  if (L->dev_phases) cudaFree(L->dev_phases);
  if (L->dev_flags) cudaFree(L->dev_flags);
  L->dev_phases = nullptr;
  L->dev_flags = nullptr;
}

int run_chain(const ChainLaunch& L, cudaStream_t stream) {
  if (L.grid < 1 || L.num_phases < 1) return 0;
  ESRP_CUDA_OK(cudaMemsetAsync(L.dev_flags, 0, sizeof(unsigned int) * static_cast<size_t>(L.grid), stream));
  ChainArgs a;
  a.phases = static_cast<const ChainPhase*>(L.dev_phases);
  a.flags = L.dev_flags;
  a.num_phases = L.num_phases;
  a.dep_all = L.dep_all;
  void* args[1] = {&a};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(L.threads);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  // no programmatic early start: the flags must be zero before any CTA of this launch publishes or polls them
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  ESRP_CUDA_OK(cudaLaunchKernelExC(&cfg, L.kernel, args));
  return 0;
}

}  // namespace esrp
