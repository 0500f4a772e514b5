// esrp_pack.h — batched weight repacking: every derived bf16 weight tile of a network (forward operators and
// data-gradient operators) is rebuilt by ONE kernel launch driven by a device-resident job table, so the cost of
// following an optimizer step is one small H2D copy + one launch instead of ~1300 tiny launches.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/esrp.h"

namespace esrp {

constexpr int kPackMaxChunks = 8;  // the generator's convs have <= 5 chunks / 6 data-gradient groups; keeps the job table small

struct PackJob {
  int type;  // 0: forward operator (esrp_pack_conv3x3_weights, transpose = 0), 1: data-gradient operator (esrp_pack_dgrad_weights)
  int layout, row0, rows, kc, bn, num_chunks;
  __nv_bfloat16* out;
  // type 0
  const float* w;
  int w_o, w_i;
  int lc0[kPackMaxChunks];
  const float* aux;
  int aux_cin, aux_chunks;
  const float* bias_src;  // [w_o] or NULL
  float* bias_dst;        // [bn], zero padded, or NULL
  // type 1
  int num_groups;
  esrp_dgrad_group_t g[2 * kPackMaxChunks];
};

int run_pack_batch(const PackJob* jobs_dev, int num_jobs, cudaStream_t stream);

}  // namespace esrp
