// esrp_conv_common.cu — descriptor -> kernel-parameter plumbing shared by the conv kernel families.
#include <cuda_bf16.h>
#include "esrp_host.h"

namespace esrp {

// descriptor fields that map 1:1 onto kernel parameters (both kernels)
void copy_common(const esrp_conv3x3_t& d, ConvKParams* pp) {
  ConvKParams& p = *pp;
  p.num_chunks = d.num_chunks;
  for (int i = 0; i < d.num_chunks; ++i) {
    p.chunk_src[i] = d.chunk_src[i];
    p.chunk_c0[i] = d.chunk_c0[i];
  }
  p.aux_chunks = d.aux_chunks;
  p.cout = d.cout;
  p.w_packed = static_cast<const uint8_t*>(d.w_packed);
  p.bias = d.bias;
  p.act = d.act; p.s0 = d.s0;
  p.r1 = d.r1; p.r1_is_f32 = d.r1_is_f32; p.r1_ctotal = d.r1_ctotal; p.r1_c0 = d.r1_c0; p.s1 = d.s1;
  p.r2 = d.r2; p.r2_is_f32 = d.r2_is_f32; p.r2_ctotal = d.r2_ctotal; p.r2_c0 = d.r2_c0; p.s2 = d.s2;
  p.noise = d.noise; p.noise_ctotal = d.noise_ctotal; p.noise_c0 = d.noise_c0;
  p.sigma = d.sigma; p.seed = d.seed; p.offset = d.offset; p.seed_ptr = nullptr;
  p.out_bf16 = static_cast<__nv_bfloat16*>(d.out_bf16); p.ob_ctotal = d.ob_ctotal; p.ob_c0 = d.ob_c0;
  p.out_f32 = static_cast<float*>(d.out_f32); p.of_ctotal = d.of_ctotal; p.of_c0 = d.of_c0;
  p.out_nchw = d.out_nchw;
  p.nsl = d.slices > 1 ? d.slices : 1;
  p.sl_stride = d.slices > 1 ? d.slice_stride : 0;
  p.f32_planar = d.f32_planar;
  p.out_lo = d.out_lo; p.ob_lo_c0 = d.ob_lo_c0;
  p.trace = static_cast<long long*>(d.trace);
  p.dbg = d.variant & 0x1F00;
  p.mask_out = static_cast<unsigned short*>(d.mask_out); p.mo_ctotal = d.mask_out_ctotal; p.mo_c0 = d.mask_out_c0;
  p.mask_in = static_cast<const unsigned short*>(d.mask_in); p.mi_ctotal = d.mask_in_ctotal; p.mi_c0 = d.mask_in_c0;
  p.r2_pre = d.r2_pre;
  p.pre_bf16 = static_cast<__nv_bfloat16*>(d.pre_bf16); p.pb_ctotal = d.pb_ctotal; p.pb_c0 = d.pb_c0;
  p.pre_f32 = d.pre_f32 ? static_cast<float*>(d.pre_f32) : nullptr; p.pf_ctotal = d.pf_ctotal; p.pf_c0 = d.pf_c0;
}


}  // namespace esrp
