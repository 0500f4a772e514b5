// ubench_tmem.cu — tcgen05.ld throughput/latency microbenchmark (sm_100a).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_tmem tools/ubench_tmem.cu
// Modes: loads per iteration x shape, number of warps, optional scattered global stores in the loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../esrganplus_b200/csrc/esrp_ptx.cuh"
using namespace esrp;

template <int WARPS, int SHAPE /*8,16,32*/, int NLOADS, bool STORES>
__global__ void __launch_bounds__(32 * WARPS, 1) k(int iters, long long* out, uint4* sink) {
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t base = holder + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int l = 0; l < NLOADS; ++l) {
      if constexpr (SHAPE == 32) { uint32_t v[32]; tmem_ld_x32(base + ((it + l) % 4) * 96 + l * 32, v); 
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i]; }
      else if constexpr (SHAPE == 16) { uint32_t v[16]; tmem_ld_x16(base + ((it + l) % 4) * 96 + l * 32 + (warp >> 2) * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i]; }
      else { uint32_t v[8]; tmem_ld_x8(base + ((it + l) % 4) * 96 + l * 32 + (warp >> 2) * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= v[i]; }
    }
    tmem_ld_wait();
    if (STORES) {
      uint4* op = sink + ((size_t)blockIdx.x * 4096 + (size_t)(it % 16) * 256 + threadIdx.x % 128) * 16 + (warp >> 2) * 2;
      op[0] = make_uint4(acc, 1, 2, 3); op[1] = make_uint4(acc, 4, 5, 6);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) sink[0] = make_uint4(acc, 0, 0, 0);
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(holder, 512); }
}

template <int WARPS, int SHAPE, int NLOADS, bool STORES>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8);
  uint4* sink; cudaMalloc(&sink, (size_t)148 * 4096 * 16 * 16);
  const int iters = 2000;
  k<WARPS, SHAPE, NLOADS, STORES><<<148, 32 * WARPS>>>(iters, d, sink);
  k<WARPS, SHAPE, NLOADS, STORES><<<148, 32 * WARPS>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = (double)h[0] / iters;
  double bytes = (double)WARPS * 32 * SHAPE * 4 * NLOADS;
  printf("{\"name\":\"%s\",\"warps\":%d,\"shape\":\"x%d\",\"loads\":%d,\"stores\":%d,\"cycles_per_iter\":%.1f,\"bytes_per_cycle\":%.1f,\"err\":\"%s\"}\n",
         name, WARPS, SHAPE, NLOADS, (int)STORES, cyc, bytes / cyc, cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  run<4, 16, 1, false>("4w x16 x1");
  run<4, 16, 3, false>("4w x16 x3");
  run<8, 16, 3, false>("8w x16 x3");
  run<4, 32, 3, false>("4w x32 x3");
  run<8, 32, 3, false>("8w x32 x3");
  run<8, 8, 3, false>("8w x8 x3");
  run<8, 16, 3, true>("8w x16 x3 +stores");
  run<4, 32, 3, true>("4w x32 x3 +stores");
  run<16, 16, 3, false>("16w x16 x3");
  return 0;
}
