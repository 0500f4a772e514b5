// ubench_row.cu — tcgen05.mma throughput for the access pattern of conv3x3_row.cuh (sm_100a).
//   A: one 130-pixel row per stage (8-stage ring), operand start = stage + kx*128 + ks*32, SBO as given
//   B: weights [chunk][kx][N rows x 128 B]
//   D: ring of TMEM slots with a given column stride
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_row tools/ubench_row.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../esrganplus_b200/csrc/esrp_ptx.cuh"
using namespace esrp;

struct Cfg { int n, sbo, dstride, nslots, chunks, commit_every, stage_bytes, kx_shift, same_stage, issuers, interleave, loop_mode; };

__global__ void __launch_bounds__(128, 1) k(Cfg c, int rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* a0 = smem + 1024;                       // 8 stages
  uint8_t* b0 = a0 + 8 * c.stage_bytes;            // chunks*3*N*128
  for (int i = threadIdx.x; i < (8 * c.stage_bytes + c.chunks * 3 * c.n * 128) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(a0)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, c.issuers); mbar_init(bar + 1, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(holder, 512); tmem_relinquish(); }
  fence_proxy_async(); tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tmem = *holder;
  const uint32_t idesc = umma_idesc_bf16_m128(c.n);
  if (warp < c.issuers) {
    const uint32_t alo0 = umma_desc_lo(smem_u32(a0)), blo0 = umma_desc_lo(smem_u32(b0));
    const uint32_t ahi = (c.sbo >> 4) | (1u << 14) | (2u << 29), bhi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0;
    for (int rep = -1; rep < 4; ++rep) {
      if (rep == 0) t0 = clock64();
      int s = 0, slot = 0, cnt = 0;
      if (c.loop_mode == 1) {
        // ONE elected thread runs the whole row loop (no per-row elect / reconvergence)
        if (elect_one()) {
          for (int r = 0; r < rows; ++r) {
            if ((r % c.issuers) != warp) { if (++slot == c.nslots) slot = 0; s = (s + c.chunks) % 8; continue; }
            for (int ch = 0; ch < c.chunks; ++ch) {
              const uint32_t a_lo = alo0 + ((c.same_stage ? 0 : s) * c.stage_bytes >> 4);
              const uint32_t b_lo = blo0 + ((ch * 3 * c.n * 128) >> 4);
              const uint32_t bstep = (c.n * 128) >> 4, astep = c.kx_shift >> 4;
              const uint32_t d = tmem + slot * c.dstride;
              ++cnt;
#pragma unroll
              for (int kk = 0; kk < 3; ++kk) {
                const int kx = kk == 0 ? 1 : (kk == 1 ? 0 : 2);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16_ss2(d, a_lo + kx * astep + ks * 2, ahi, b_lo + kx * bstep + ks * 2, bhi, idesc, (ch | kk | ks) != 0);
              }
              if (c.commit_every && (cnt % c.commit_every) == 0) umma_commit(bar + 1);
              if (++s == 8) s = 0;
            }
            if (++slot == c.nslots) slot = 0;
          }
          umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, (rep + 1) & 1);
        tcgen05_fence_after();
        continue;
      }
      for (int r = 0; r < rows; ++r) {
        if ((r % c.issuers) != warp) { if (++slot == c.nslots) slot = 0; s = (s + c.chunks) % 8; continue; }
        for (int ch = 0; ch < c.chunks; ++ch) {
          const uint32_t a_lo = alo0 + ((c.same_stage ? 0 : s) * c.stage_bytes >> 4);
          const uint32_t b_lo = blo0 + ((ch * 3 * c.n * 128) >> 4);
          const uint32_t bstep = (c.n * 128) >> 4, astep = c.kx_shift >> 4;
          // interleave: rows alternate between two independent rings (columns 0.. and 256..) of sliding windows
          const uint32_t d = c.interleave ? tmem + (r & 1) * 256 + ((r >> 1) % c.nslots) * c.dstride : tmem + slot * c.dstride;
          ++cnt;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              const int kx = kk == 0 ? 1 : (kk == 1 ? 0 : 2);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_f16_ss2(d, a_lo + kx * astep + ks * 2, ahi, b_lo + kx * bstep + ks * 2, bhi, idesc, (ch | kk | ks) != 0);
            }
            if (c.commit_every && (cnt % c.commit_every) == 0) umma_commit(bar + 1);
          }
          if (c.loop_mode != 2) __syncwarp();
          if (++s == 8) s = 0;
        }
        if (++slot == c.nslots) slot = 0;
      }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, (rep + 1) & 1);
      tcgen05_fence_after();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

void run(const char* name, Cfg c) {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int rows = 64;
  const int smem = 2048 + 8 * c.stage_bytes + c.chunks * 3 * c.n * 128;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<148, 128, smem>>>(c, rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double mmas = 4.0 * rows * c.chunks * 12;
  printf("{\"name\": \"%s\", \"n\": %d, \"sbo\": %d, \"dstride\": %d, \"slots\": %d, \"chunks\": %d, \"commit_every\": %d, \"kx_shift\": %d, \"same_stage\": %d, \"issuers\": %d, \"cycles_per_mma\": %.2f, \"err\": \"%s\"}\n",
         name, c.n, c.sbo, c.dstride, c.nslots, c.chunks, c.commit_every, c.kx_shift, c.same_stage, c.issuers, h[0] / mmas, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  //              n  sbo  dstr slots ch commit stage  kxsh same
  run("1 issuer, no commit",      {96, 1024, 96, 5, 1, 0, 17408, 128, 0, 1});
  run("1 issuer, commit/row",     {96, 1024, 96, 5, 1, 1, 17408, 128, 0, 1});
  run("2 issuers, no commit",     {96, 1024, 96, 5, 1, 0, 17408, 128, 0, 2});
  run("2 issuers, commit/row",    {96, 1024, 96, 5, 1, 1, 17408, 128, 0, 2});
  run("3 issuers, commit/row",    {96, 1024, 96, 5, 1, 1, 17408, 128, 0, 3});
  run("2 issuers, 2 chunks commit/chunk", {96, 1024, 96, 5, 2, 1, 17408, 128, 0, 2});
  run("2 issuers, 3 chunks commit/row",   {96, 1024, 96, 5, 3, 3, 17408, 128, 0, 2});
  run("1 issuer, 3 chunks commit/row",    {96, 1024, 96, 5, 3, 3, 17408, 128, 0, 1});
  run("2 issuers n128 commit/row",{128, 1024, 128, 4, 1, 1, 17408, 128, 0, 2});
  run("2 issuers n32 commit/row", {32, 1024, 96, 5, 1, 1, 17408, 128, 0, 2});
  // output-stationary ring of conv3x3_row.cuh: consecutive rows accumulate into windows that overlap by 2 of 3 blocks
  run("1 issuer, sliding D (stride 32)",  {96, 1024, 32, 14, 1, 1, 17408, 128, 0, 1});
  run("2 issuers, sliding D (stride 32)", {96, 1024, 32, 14, 1, 1, 17408, 128, 0, 2});
  run("2 issuers, sliding D, 2 chunks",   {96, 1024, 32, 14, 2, 2, 17408, 128, 0, 2});
  run("2 issuers, sliding D, no commit",  {96, 1024, 32, 14, 1, 0, 17408, 128, 0, 2});
  run("2 issuers, sliding D, two interleaved rings", {96, 1024, 32, 6, 1, 1, 17408, 128, 0, 2, 1});
  run("2 issuers, sliding D, two interleaved rings, 2 chunks", {96, 1024, 32, 6, 2, 2, 17408, 128, 0, 2, 1});
  run("1 issuer, sliding D, two interleaved rings", {96, 1024, 32, 6, 1, 1, 17408, 128, 0, 1, 1});
  run("2 issuers, disjoint D stride 96, 2 chunks", {96, 1024, 96, 5, 2, 2, 17408, 128, 0, 2});
  // round 2: what does the per-row elect / reconvergence cost a single issuer?
  run("1 issuer, sliding D, one elected thread runs the row loop", {96, 1024, 32, 14, 1, 1, 17408, 128, 0, 1, 0, 1});
  run("1 issuer, sliding D, one elected thread, no commit",        {96, 1024, 32, 14, 1, 0, 17408, 128, 0, 1, 0, 1});
  run("1 issuer, sliding D, elect per row without syncwarp",       {96, 1024, 32, 14, 1, 1, 17408, 128, 0, 1, 0, 2});
  run("1 issuer, sliding D, one elected thread, 2 chunks",         {96, 1024, 32, 14, 2, 2, 17408, 128, 0, 1, 0, 1});
  run("1 issuer, sliding D, one elected thread, 3 chunks",         {96, 1024, 32, 14, 3, 3, 17408, 128, 0, 1, 0, 1});
  run("1 issuer, sliding D (stride 32), 3 chunks",                 {96, 1024, 32, 14, 3, 3, 17408, 128, 0, 1});
  return 0;
}
