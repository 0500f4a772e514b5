// conv_params.h — kernel parameter block of conv3x3_tc_kernel (plain data; shared by host planning
// code and the device kernel).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

#include "../../include/esrp.h"

namespace esrp {

struct ConvKParams {
  int n, h, w;
  int tiles_x, tiles_y, num_tiles;
  int num_chunks;
  int chunk_src[ESRP_MAX_CHUNKS];
  int chunk_c0[ESRP_MAX_CHUNKS];
  int aux_chunks;
  int cout;
  const uint8_t* w_packed;
  const uint8_t* w_aux;
  const float* bias;
  int w_resident;
  int stages;
  uint32_t tmem_cols;
  int act;
  float s0;
  const void* r1;
  int r1_is_f32, r1_ctotal, r1_c0;
  float s1;
  const void* r2;
  int r2_is_f32, r2_ctotal, r2_c0;
  float s2;
  int noise;
  float sigma;
  unsigned long long seed, offset;
  __nv_bfloat16* out_bf16;
  int ob_ctotal, ob_c0;
  float* out_f32;
  int of_ctotal, of_c0;
  float* out_nchw;
};

}  // namespace esrp
