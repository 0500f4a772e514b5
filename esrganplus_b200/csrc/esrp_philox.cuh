// esrp_philox.cuh — counter-based Gaussian noise for the nESRGAN+ GaussianNoise layer
// (reference: block.py:110-122, `self.noise.repeat(*x.size()).normal_() * scale`).
// Philox4x32-10 keyed by (seed), counter = (offset + element_index/4); Box-Muller turns the four
// uniform words into four N(0,1) samples.  Stateless: the backward pass regenerates the same
// samples from (seed, offset), so no noise tensor is ever stored.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace esrp {

__host__ __device__ inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(M0) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(M1) * c[2];
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    const uint32_t n0 = hi1 ^ c[1] ^ k0;
    const uint32_t n1 = lo1;
    const uint32_t n2 = hi0 ^ c[3] ^ k1;
    const uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0;
    k1 += W1;
  }
}

// Four independent N(0,1) samples for 64-bit counter `ctr` under key `seed`.
__host__ __device__ inline void philox_normal4(unsigned long long seed, unsigned long long ctr,
                                               float z[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0x6e455352u, 0u};
  philox4x32_10(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), c);
  const float k2pi = 6.28318530717958647692f;
  const float inv32 = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float u1 = (static_cast<float>(c[2 * i]) + 1.0f) * inv32;  // (0, 1]
    const float u2 = static_cast<float>(c[2 * i + 1]) * inv32;        // [0, 1)
#ifdef __CUDA_ARCH__
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, co;
    __sincosf(k2pi * u2, &s, &co);
#else
    const float r = sqrtf(-2.0f * logf(u1));
    const float s = sinf(k2pi * u2), co = cosf(k2pi * u2);
#endif
    z[2 * i] = r * co;
    z[2 * i + 1] = r * s;
  }
}

}  // namespace esrp
