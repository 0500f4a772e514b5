// conv_params.h — kernel parameter block of conv3x3_tc_kernel (plain data; shared by host planning
// code and the device kernel).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

#include "../../include/esrp.h"

#define ESRP_MAX_MT 5  // M-tile accumulator slots per CTA (TMEM: mt * nt <= 512 columns)

namespace esrp {

struct ConvKParams {
  int n, h, w;
  // tiling (see conv3x3_tc.cuh): M-tile = rm rows x cw columns, CTA tile = up to mt M-tiles
  int cw, cw_log2, rm, mt;
  int x_tiles, x_step;          // column blocks per image and their source-column advance
  int units_per_col;            // ceil(h / rm)
  long long units_total;        // n * x_tiles * units_per_col
  int nt;                       // TMEM columns per accumulator: 3*BN (+BN with conv1x1)
  int a_box_bytes, a_stage_bytes;
  int num_chunks;
  int chunk_src[ESRP_MAX_CHUNKS];
  int chunk_c0[ESRP_MAX_CHUNKS];
  int aux_chunks;
  int cout;
  const uint8_t* w_packed;
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
  int noise, noise_ctotal, noise_c0;
  float sigma;
  unsigned long long seed, offset;
  const unsigned long long* seed_ptr;  // when non-null the Philox key is read from device memory (CUDA-graph replays)
  __nv_bfloat16* out_bf16;
  int ob_ctotal, ob_c0;
  float* out_f32;
  int of_ctotal, of_c0;
  float* out_nchw;
  // training extensions (kernels instantiated with EXT = true only)
  unsigned short* mask_out;
  int mo_ctotal, mo_c0;
  const unsigned short* mask_in;
  int mi_ctotal, mi_c0;
  int r2_pre;
  __nv_bfloat16* pre_bf16;
  int pb_ctotal, pb_c0;
  float* pre_f32;
  int pf_ctotal, pf_c0;
  int nsl;                 // > 1: the launch co-schedules nsl output slices of bn channels (row kernel only)
  long long sl_stride;     // bytes from one slice's packed weights (and bias) to the next
  int f32_planar;          // fp32 operands (r1/r2/out_f32) are [n][h][c/4][w][4] instead of NHWC (row kernel only)
  int last_half;           // row kernel (row_alt): only the first kc/2 channels of the last chunk carry weights
  int row_alt;             // row kernel: the two MMA issuers alternate whole rows instead of splitting the taps of every row
  int no_quad;             // row kernel: store bf16 outputs pixel by pixel (timing experiments: ESRP_NO_QUAD)
  int out_lo, ob_lo_c0;    // tile kernel: also store bf16(v - bf16(v)) at channel ob_lo_c0 of out_bf16 (esrp_conv3x3_t::out_lo)
  int chunk_bars;          // row kernel (row_alt, even ring depth, stages * num_chunks <= 8): one "data landed" barrier per K-chunk
                           // tile instead of one per row, so the MMAs of chunk c start when chunk c is there
  int pair_single;         // row kernel, CTA pairs: one issuer thread, one multicast commit per row (experiment, ESRP_PAIR_SINGLE=1)
  int dbg;           // timing experiments only (ESRP_DBG_*): results are wrong when non-zero
  long long* trace;  // optional [3][1024] clock64 timeline of CTA 0 (see trace_ev)
};

}  // namespace esrp
