// esrp_conv_chain.cu — planning + launch of the persistent conv chain (conv3x3_chain.cuh): a run of row-kernel conv
// launches of one engine plan (the 5 convs of every ResidualDenseBlock_5C of the RRDB trunk, block.py:260-291) becomes
// ONE kernel launch whose phases are the original launches.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/esrp.h"
#include "conv3x3_chain.cuh"
#include "esrp_host.h"

namespace esrp {

static_assert(kChainMaxPhasesHost == kChainMaxPhases, "keep esrp_host.h in step with conv3x3_chain.cuh");

bool chain_compatible(const ConvLaunch& L) {
  const ConvKParams& p = L.params;
  return L.fam == 1 && L.cluster == 1 && L.kc == 64 && L.bn == 32 && L.ext == 0 && p.out_nchw == nullptr && p.trace == nullptr && p.dbg == 0 &&
         p.w_resident == 1 && L.grid >= 1 && p.cout == 32 && p.act != 2 && p.num_chunks <= 4 && p.noise == 0 && p.seed_ptr == nullptr &&
         (p.r1 == nullptr || p.s1 == 1.0f) && p.bias != nullptr && (p.nsl <= 1 || p.sl_stride % 16 == 0);
}

namespace {
bool same_map(const CUtensorMap& x, const CUtensorMap& y) { return memcmp(&x, &y, sizeof(CUtensorMap)) == 0; }
}  // namespace

int plan_chain(const ConvLaunch* const* convs, int count, ChainLaunch* out) {
  if (count < 1 || count > kChainMaxPhases) return set_error("conv chain: %d phases (1..%d)", count, kChainMaxPhases);
  ChainArgs* A = static_cast<ChainArgs*>(calloc(1, sizeof(ChainArgs)));
  if (!A) return set_error("conv chain: out of host memory");
  out->args_host = A;
  const ConvKParams& p0 = convs[0]->params;
  // base addresses: every pointer of a phase is stored as a 32-bit count of 16-byte units above them
  uintptr_t act_lo = ~uintptr_t(0), w_lo = ~uintptr_t(0);
  auto lower = [](uintptr_t* lo, const void* ptr) { if (ptr && reinterpret_cast<uintptr_t>(ptr) < *lo) *lo = reinterpret_cast<uintptr_t>(ptr); };
  for (int i = 0; i < count; ++i) {
    const ConvKParams& p = convs[i]->params;
    lower(&act_lo, p.r1); lower(&act_lo, p.r2); lower(&act_lo, p.out_bf16); lower(&act_lo, p.out_f32);
    lower(&w_lo, p.w_packed); lower(&w_lo, p.bias);
  }
  act_lo &= ~uintptr_t(15);
  w_lo &= ~uintptr_t(15);
  auto off16 = [&](uintptr_t base, const void* ptr, uint32_t* dst) -> int {
    if (!ptr) { *dst = kChainNull; return 0; }
    const uintptr_t d = reinterpret_cast<uintptr_t>(ptr) - base;
    if ((d & 15) || (d >> 4) >= kChainNull) return set_error("conv chain: pointer %p is not 16-byte aligned / out of range of its base", ptr);
    *dst = static_cast<uint32_t>(d >> 4);
    return 0;
  };
  int nmaps = 0;
  auto map_index = [&](const CUtensorMap& tm) -> int {
    for (int i = 0; i < nmaps; ++i)
      if (same_map(A->tmaps[i], tm)) return i;
    if (nmaps == kChainMaxMaps) return -1;
    A->tmaps[nmaps] = tm;
    return nmaps++;
  };
  int smem = 0;
  for (int i = 0; i < count; ++i) {
    const ConvLaunch& L = *convs[i];
    const ConvKParams& p = L.params;
    if (!chain_compatible(L)) return set_error("conv chain: phase %d is not a kc=64 bn=32 row-kernel conv", i);
    if (p.n != p0.n || p.h != p0.h || p.w != p0.w || p.units_total != p0.units_total || L.grid != convs[0]->grid ||
        p.a_box_bytes != p0.a_box_bytes || p.a_stage_bytes != p0.a_stage_bytes)
      return set_error("conv chain: phase %d has another shape / grid", i);
    const int nsl = p.nsl > 1 ? p.nsl : 1;
    if (L.grid % nsl) return set_error("conv chain: grid %d is not a multiple of the %d co-scheduled slices of phase %d", L.grid, nsl, i);
    ChainPhaseC& c = A->ph[i];
    if (off16(w_lo, p.w_packed, &c.w_off16) || off16(w_lo, p.bias, &c.bias_off16) || off16(act_lo, p.r1, &c.r1_off16) ||
        off16(act_lo, p.r2, &c.r2_off16) || off16(act_lo, p.out_bf16, &c.ob_off16) || off16(act_lo, p.out_f32, &c.of_off16))
      return 1;
    c.sl_stride16 = static_cast<uint32_t>(p.sl_stride / 16);
    c.s0 = p.s0; c.s2 = p.s2;
    auto u16 = [](int v) { return static_cast<uint16_t>(v); };
    if ((p.r1_ctotal | p.r1_c0 | p.r2_ctotal | p.r2_c0 | p.ob_ctotal | p.ob_c0 | p.of_ctotal | p.of_c0) >> 16)
      return set_error("conv chain: channel counts above 65535");
    c.r1_ctotal = u16(p.r1_ctotal); c.r1_c0 = u16(p.r1_c0); c.r2_ctotal = u16(p.r2_ctotal); c.r2_c0 = u16(p.r2_c0);
    c.ob_ctotal = u16(p.ob_ctotal); c.ob_c0 = u16(p.ob_c0); c.of_ctotal = u16(p.of_ctotal); c.of_c0 = u16(p.of_c0);
    c.chunk_src = 0;
    for (int k = 0; k < p.num_chunks; ++k) {
      if ((p.chunk_c0[k] & 7) || p.chunk_c0[k] / 8 > 255) return set_error("conv chain: chunk channel offset %d", p.chunk_c0[k]);
      c.chunk_c0_8[k] = static_cast<uint8_t>(p.chunk_c0[k] / 8);
      if (p.chunk_src[k]) c.chunk_src |= static_cast<uint8_t>(1u << k);
    }
    const int i0 = map_index(L.tm0), i1 = map_index(L.tm1);
    if (i0 < 0 || i1 < 0) return set_error("conv chain: more than %d distinct tensor maps", kChainMaxMaps);
    c.tm0 = static_cast<uint8_t>(i0); c.tm1 = static_cast<uint8_t>(i1);
    c.num_chunks = static_cast<uint8_t>(p.num_chunks);
    c.aux_chunks = static_cast<uint8_t>(p.aux_chunks);
    c.nsl = static_cast<uint8_t>(nsl);
    c.flags = static_cast<uint8_t>((p.act ? kChainFAct : 0) | (p.r1 && p.r1_is_f32 ? kChainFR1F32 : 0) |
                                   (p.r2 && p.r2_is_f32 ? kChainFR2F32 : 0) | (p.f32_planar ? kChainFPlanar : 0) |
                                   (p.last_half ? kChainFLastHalf : 0) | (p.no_quad ? kChainFNoQuad : 0));
    // activation ring: as many ROWS of K-chunk tiles as fit beside the resident weights.  A row's tiles are handed back
    // through the barrier of the output block that row completes, so the producer must not be lapped on a block
    // barrier: rows in the ring + 2 blocks per segment end inside that window < blocks in the ring (16, or 8 with conv1x1)
    const int w_all = p.num_chunks * 3 * (p.aux_chunks > 0 ? 4 : 3) * 32 * 128;
    int rows = (kMaxSmem - kSmemFixed - 1024 - w_all) / (p.a_stage_bytes * p.num_chunks);
    if (rows > kChainMaxTiles / p.num_chunks) rows = kChainMaxTiles / p.num_chunks;
    const int row_cap = p.aux_chunks > 0 ? (p.h < 3 ? 2 : 3) : 4;
    if (rows > row_cap) rows = row_cap;
    if (rows < 2) return set_error("conv chain: phase %d: weights + 2 activation rows do not fit in shared memory", i);
    c.stages = static_cast<uint8_t>(rows);
    const int nbuf = rows * p.num_chunks;
    const int need = kSmemFixed + 1024 + w_all + nbuf * p.a_stage_bytes;
    if (need > smem) smem = need;
  }
  A->w_base = reinterpret_cast<const uint8_t*>(w_lo);
  A->act_base = reinterpret_cast<uint8_t*>(act_lo);
  A->units_total = p0.units_total;
  A->n = p0.n; A->h = p0.h; A->w = p0.w; A->x_tiles = p0.x_tiles;
  A->a_box_bytes = p0.a_box_bytes; A->a_stage_bytes = p0.a_stage_bytes;
  A->num_phases = count;
  auto kern = conv3x3_chain_kernel<64, 32>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern))) return 1;
  out->kernel = reinterpret_cast<const void*>(kern);
  out->num_phases = count;
  out->grid = convs[0]->grid;
  out->threads = kChainThreads;
  out->smem = smem;
  out->dep_all = p0.x_tiles > 1 ? 1 : 0;  // a 130-pixel box reads one pixel of the neighbouring column block
  if (out->grid > sm_count()) return set_error("conv chain: grid %d exceeds the %d SMs (all CTAs must be co-resident)", out->grid, sm_count());
  ESRP_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&out->dev_flags), sizeof(unsigned int) * static_cast<size_t>(out->grid)));
  A->flags = out->dev_flags;
  return 0;
}

void free_chain(ChainLaunch* L) {
  if (L->args_host) free(L->args_host);
  if (L->dev_flags) cudaFree(L->dev_flags);
  L->args_host = nullptr;
  L->dev_flags = nullptr;
}

int run_chain(const ChainLaunch& L, cudaStream_t stream) {
  if (L.grid < 1 || L.num_phases < 1 || !L.args_host) return 0;
  ESRP_CUDA_OK(cudaMemsetAsync(L.dev_flags, 0, sizeof(unsigned int) * static_cast<size_t>(L.grid), stream));
  ChainArgs* A = static_cast<ChainArgs*>(L.args_host);
  static const int dbg = [] { const char* e = getenv("ESRP_CHAIN_DBG"); return e ? atoi(e) : 0; }();
  static const bool dep_all_env = getenv("ESRP_CHAIN_DEP_ALL") != nullptr;
  A->dep_all = (L.dep_all || dep_all_env) ? 1 : 0;
  A->dbg = dbg;
  // ESRP_CHAIN_TRACE=<file>: per-phase clock64 timeline of one CTA (ESRP_CHAIN_TRACE_CTA, default 70), written after
  // every launch (synchronises: diagnosis only)
  static const char* trace_path = getenv("ESRP_CHAIN_TRACE");
  static long long* trace_dev = nullptr;
  if (trace_path) {
    static const int cta = [] { const char* e = getenv("ESRP_CHAIN_TRACE_CTA"); return e ? atoi(e) : 70; }();
    if (!trace_dev) ESRP_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&trace_dev), sizeof(long long) * 16 * kChainMaxPhases));
    ESRP_CUDA_OK(cudaMemsetAsync(trace_dev, 0, sizeof(long long) * 16 * kChainMaxPhases, stream));
    A->trace = trace_dev;
    A->trace_cta = cta < L.grid ? cta : 0;
  } else {
    A->trace = nullptr;
  }
  void* args[1] = {A};  // the whole phase table is a kernel parameter (copied at launch / graph capture)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(L.threads);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  // no programmatic early start: the flags must be zero before any CTA of this launch publishes or polls them
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  ESRP_CUDA_OK(cudaLaunchKernelExC(&cfg, L.kernel, args));
  if (trace_path) {
    std::vector<long long> host(static_cast<size_t>(16) * L.num_phases);
    ESRP_CUDA_OK(cudaStreamSynchronize(stream));
    ESRP_CUDA_OK(cudaMemcpy(host.data(), trace_dev, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "w")) {
      for (int i = 0; i < L.num_phases; ++i) {
        for (int k = 0; k < 16; ++k) fprintf(f, "%lld%c", host[static_cast<size_t>(i) * 16 + k], k == 15 ? '\n' : ' ');
      }
      fclose(f);
    }
  }
  return 0;
}

}  // namespace esrp
