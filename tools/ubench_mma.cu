// ubench_mma.cu — tcgen05.mma issue/throughput microbenchmark for B200 (sm_100a).
// Measures cycles per 128xNx16 bf16 MMA (operands in shared memory, or A in tensor memory) as a
// function of N, with the same descriptor walk the conv kernel uses (36 distinct A and B tiles per
// round).  Operand contents are irrelevant (zeros).  Prints one JSON line per configuration.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../esrganplus_b200/csrc/esrp_ptx.cuh"

using namespace esrp;

__device__ __forceinline__ void umma_ss(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d),
      "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d),
      "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: SS, A tiles walk a halo tile like the conv kernel.  mode 1: TS (A from TMEM).
// mode 2: SS with a constant A tile (smem re-read of the same rows).
template <int N>
__global__ void __launch_bounds__(128, 1) bench_kernel(int mode, int rounds, long long* out_cycles, long long* out_ns) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* a_tile = smem + 1024;            // 42 KB halo tile
  uint8_t* b_tile = smem + 1024 + 43008;    // 3 x N rows x 128 B
  // zero operands
  for (int i = threadIdx.x; i < (43008 + 3 * 256 * 128) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *holder;
  constexpr uint32_t IDESC = umma_idesc_bf16_m128(N);
  if (warp == 0) {
    const uint32_t alo0 = ((smem_u32(a_tile) >> 4) & 0x3FFF) | 0x10000u;
    const uint32_t blo0 = ((smem_u32(b_tile) >> 4) & 0x3FFF) | 0x10000u;
    const uint32_t ahi = (2304u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t bhi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0, t1 = 0, n0 = 0, n1 = 0;
    // warm-up round + timed rounds
    for (int r = -1; r < rounds; ++r) {
      if (r == 0) {
        t0 = clock64();
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n0));
      }
      if (elect_one()) {
#pragma unroll 1
       for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t aoff = ((t / 3) * 18 + m * 8 + (t % 3)) * 128 + ks * 32;
              const uint32_t boff = (t % 3) * N * 128 + ks * 32;
              if (mode == 0)
                umma_ss(tmem + m * 256, alo0 + (aoff >> 4), ahi, blo0 + (boff >> 4), bhi, IDESC, (t | ks) != 0);
              else if (mode == 2)
                umma_ss(tmem + m * 256, alo0 + ((ks * 32) >> 4), ahi, blo0 + (boff >> 4), bhi, IDESC, (t | ks) != 0);
              else
                umma_ts(tmem + m * 128, tmem + 448 + (t % 4) * 8, blo0 + (boff >> 4), bhi, IDESC, (t | ks) != 0);
            }
          }
        }
       }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, (r + 1) & 1);
      tcgen05_fence_after();
    }
    t1 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n1));
    if (threadIdx.x == 0) {
      out_cycles[blockIdx.x] = t1 - t0;
      out_ns[blockIdx.x] = n1 - n0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N>
void run(int mode, int grid, int rounds) {
  long long *dc, *dn;
  cudaMalloc(&dc, grid * sizeof(long long));
  cudaMalloc(&dn, grid * sizeof(long long));
  const int smem = 1024 + 1024 + 43008 + 3 * 256 * 128;
  cudaFuncSetAttribute(bench_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench_kernel<N><<<grid, 128, smem>>>(mode, rounds, dc, dn);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"n\": %d, \"mode\": %d, \"grid\": %d, \"error\": \"%s\"}\n", N, mode, grid, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<long long> hc(grid), hn(grid);
  cudaMemcpy(hc.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(hn.data(), dn, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0, ns = 0;
  for (int i = 0; i < grid; ++i) {
    cyc += hc[i];
    ns += hn[i];
  }
  cyc /= grid;
  ns /= grid;
  const double mmas = 288.0 * rounds;
  const double cpm = cyc / mmas;
  const double ghz = cyc / ns;
  const double tflops = 2.0 * 128 * N * 16 * mmas * grid / (ns * 1e-9) * 1e-12;
  printf("{\"n\": %d, \"mode\": %d, \"grid\": %d, \"cycles_per_mma\": %.2f, \"floor_cycles\": %.1f, \"sm_ghz\": %.3f, \"tflops\": %.1f}\n",
         N, mode, grid, cpm, N / 2.0, ghz, tflops);
  cudaFree(dc);
  cudaFree(dn);
}

int main(int argc, char** argv) {
  const int rounds = 500;
  for (int grid : {1, 148}) {
    for (int mode : {0, 2, 1}) {
      run<16>(mode, grid, rounds);
      run<32>(mode, grid, rounds);
      run<64>(mode, grid, rounds);
      run<96>(mode, grid, rounds);
      run<128>(mode, grid, rounds);
      if (mode != 1) {
        run<192>(mode, grid, rounds);
        run<256>(mode, grid, rounds);
      }
    }
  }
  return 0;
}
