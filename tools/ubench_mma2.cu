// ubench_mma2.cu — cycles per tcgen05.mma for cta_group::1 (M = 128) vs cta_group::2 (M = 256, cluster of two CTAs, the
// leader issues; each CTA supplies its 128 rows of A and N/2 rows of B) as a function of N, bf16 operands in shared memory,
// the descriptor walk of the row kernel (three kx-shifted A tiles x four K-slices, B tiles per tap).  Operand contents are
// irrelevant (zeros).  One JSON line per configuration.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma2 tools/ubench_mma2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../esrganplus_b200/csrc/esrp_ptx.cuh"

using namespace esrp;

template <int N, bool PAIR>
__global__ void __launch_bounds__(128, 1) bench_kernel(int rounds, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 64);
  uint8_t* a_tile = smem + 1024;          // 130-pixel row x 64 channels (17 KB, rounded) x 2 chunks
  uint8_t* b_tile = a_tile + 2 * 17408;   // 3 taps x N rows x 128 B (both chunks read the same B tiles: 227 KB limit)
  for (int i = threadIdx.x; i < (2 * 17408 + 3 * 256 * 128) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(a_tile)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2sm(holder, 512); tmem_relinquish_2sm(); } else { tmem_alloc(holder, 512); tmem_relinquish(); }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *holder;
  constexpr uint32_t IDESC = PAIR ? umma_idesc_bf16_m256(N) : umma_idesc_bf16_m128(N);
  constexpr int NB = PAIR ? N / 2 : N;  // resident B rows per tap
  if (warp == 0 && rank == 0) {
    const uint32_t alo0 = umma_desc_lo(smem_u32(a_tile));
    const uint32_t blo0 = umma_desc_lo(smem_u32(b_tile));
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = 0, t1 = 0;
    for (int r = -1; r < rounds; ++r) {
      if (r == 0) t0 = clock64();
      if (elect_one()) {
#pragma unroll 1
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t aoff = c * 17408 + kx * 128 + ks * 32;
                const uint32_t boff = kx * NB * 128 + ks * 32;
                if (PAIR)
                  umma_f16_ss2_2sm(tmem + (rep & 1) * 256, alo0 + (aoff >> 4), hi, blo0 + (boff >> 4), hi, IDESC, 1u);
                else
                  umma_f16_ss2(tmem + (rep & 1) * 256, alo0 + (aoff >> 4), hi, blo0 + (boff >> 4), hi, IDESC, 1u);
              }
            }
          }
        }
        if (PAIR) umma_commit_2sm(bar); else umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, (r + 1) & 1);
      tcgen05_fence_after();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

template <int N, bool PAIR>
void run(int grid, int rounds) {
  long long* dc;
  cudaMalloc(&dc, grid * sizeof(long long));
  cudaMemset(dc, 0, grid * sizeof(long long));
  const int smem = 1024 + 1024 + 2 * 17408 + 3 * 256 * 128;
  auto kern = bench_kernel<N, PAIR>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, rounds, dc);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"n\": %d, \"cta_group\": %d, \"grid\": %d, \"error\": \"%s\"}\n", N, PAIR ? 2 : 1, grid, cudaGetErrorString(e));
    exit(1);
  }
  std::vector<long long> hc(grid);
  cudaMemcpy(hc.data(), dc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0;
  int cnt = 0;
  for (int i = 0; i < grid; ++i)
    if (hc[i] > 0) { cyc += hc[i]; ++cnt; }
  const double per = cyc / cnt / (static_cast<double>(rounds) * 4 * 2 * 3 * 4);
  // useful MACs per SM and cycle: 128 rows x N x 16 per MMA (cta_group::2: per SM of the pair)
  printf("{\"n\": %d, \"cta_group\": %d, \"grid\": %d, \"cycles_per_mma\": %.1f, \"ideal_cycles\": %.1f, \"pipe_frac\": %.3f}\n", N,
         PAIR ? 2 : 1, grid, per, N / 2.0, (N / 2.0) / per);
  cudaFree(dc);
}

template <int N>
void both(int grid, int rounds) {
  run<N, false>(grid, rounds);
  run<N, true>(grid, rounds);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const int rounds = argc > 2 ? atoi(argv[2]) : 200;
  both<32>(grid, rounds);
  both<64>(grid, rounds);
  both<96>(grid, rounds);
  both<128>(grid, rounds);
  both<192>(grid, rounds);
  both<256>(grid, rounds);
  return 0;
}
