// conv_chain_dep.h — which CTAs a phase of the persistent conv chain (conv3x3_chain.cuh) depends on.  Plain integer
// arithmetic shared by the kernel and the host-side test (tests/test_chain_dep.py compiles this header with g++).
#pragma once

#if defined(__CUDACC__)
#define ESRP_HD __host__ __device__ __forceinline__
#else
#define ESRP_HD inline
#endif

namespace esrp {

// SegWalk (conv3x3_row.cuh) gives row group g of ng the units [U*g/ng, U*(g+1)/ng) (unit = one 128-pixel row segment).
ESRP_HD long long chain_group_start(long long U, int g, int ng) { return U * g / ng; }

// The group that owns unit u: the largest g with U*g/ng <= u, i.e. ceil((u+1)*ng/U) - 1.
ESRP_HD int chain_group_of(long long U, long long u, int ng) { return static_cast<int>(((u + 1) * ng + U - 1) / U) - 1; }

// Row group `g` of `ng` in this phase reads (and will overwrite buffers read at) the units one halo row above and
// below its own range: the groups [*g_lo, *g_hi] of the previous phase's split into ngq groups cover them.
// (Units of different images / column blocks are adjacent in this numbering; treating them as neighbours only adds a
// dependency.)
ESRP_HD void chain_dep_range(long long U, int g, int ng, int ngq, int* g_lo, int* g_hi) {
  long long lo = chain_group_start(U, g, ng) - 1, hi = chain_group_start(U, g + 1, ng);
  if (lo < 0) lo = 0;
  if (hi > U - 1) hi = U - 1;
  if (hi < lo) hi = lo;
  *g_lo = chain_group_of(U, lo, ngq);
  *g_hi = chain_group_of(U, hi, ngq);
}

}  // namespace esrp
