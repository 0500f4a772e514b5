// plan_tile.inl — launch planning + instantiations of conv3x3_tc_kernel for one value of ESRP_EXT.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>

#include "../../include/esrp.h"
#include "conv3x3_tc.cuh"
#include "esrp_host.h"

namespace esrp {

template <int KC, int BN, bool EXT>
static int plan_conv_t(const esrp_conv3x3_t& d, ConvLaunch* out) {
  constexpr int RB = KC * 2;
  ConvKParams& p = out->params;
  memset(&p, 0, sizeof(p));
  p.n = d.n; p.h = d.h; p.w = d.w;
  const bool has_aux = d.aux_chunks > 0;
  const int nb_rows = (has_aux ? 4 : 3) * BN;
  const int w_chunk_bytes = 3 * nb_rows * RB;
  const int w_all = d.num_chunks * w_chunk_bytes;
  p.nt = nb_rows;

  // M-tile width: the smallest of 16/32/64/128 columns covering the image, else 128 with x-halo blocks
  int cwl = 4;
  while (cwl < 7 && (1 << cwl) < d.w) ++cwl;
  const int force_cwl = (d.variant >> 4) & 15;
  if (force_cwl) {
    if (force_cwl < 4 || force_cwl > 7) return set_error("conv3x3: variant forces cw_log2=%d (4..7)", force_cwl);
    cwl = force_cwl;
  }
  p.cw_log2 = cwl;
  p.cw = 1 << cwl;
  p.rm = 128 >> cwl;
  if (d.w <= p.cw) {
    p.x_tiles = 1;
    p.x_step = p.cw;
  } else {
    p.x_step = p.cw - 2;
    p.x_tiles = (d.w - 1 + p.x_step - 1) / p.x_step;
  }
  p.units_per_col = (d.h + p.rm - 1) / p.rm;
  p.units_total = static_cast<long long>(d.n) * p.x_tiles * p.units_per_col;
  if (p.units_total > 0x7fffffffLL) return set_error("conv3x3: problem too large (%lld M-tiles)", p.units_total);

  // accumulator slots per CTA tile: TMEM (mt * nt <= 512) and shared memory (>= 2 stages) permitting
  const int sms = sm_count();
  if (sms <= 0) return set_error("conv3x3: no CUDA device");
  const long long per_cta = (p.units_total + sms - 1) / sms;
  int mt = 512 / p.nt;
  if (mt > ESRP_MAX_MT) mt = ESRP_MAX_MT;
  if (mt > per_cta) mt = static_cast<int>(per_cta);
  if (mt > p.units_per_col) mt = p.units_per_col;
  const int force_mt = d.variant & 15;
  if (force_mt) {
    if (force_mt > ESRP_MAX_MT || force_mt * p.nt > 512) return set_error("conv3x3: variant forces mt=%d (nt=%d)", force_mt, p.nt);
    mt = force_mt;
  }
  const int avail = kMaxSmem - kSmemFixed - 1024;  // 1 KB alignment slack
  auto a_stage = [&](int m) { return ((m * p.rm + 2) * p.cw * RB + 1023) / 1024 * 1024; };
  // largest tile with resident weights and >= 2 stages; else stream the weights with the chunks
  int best_mt = 0, best_res = 0, best_s = 0;
  for (int pass = 0; pass < 2 && !best_mt; ++pass) {
    for (int m = mt; m >= 1; --m) {
      const int s_res = w_all <= avail ? (avail - w_all) / a_stage(m) : 0;
      const int s_str = avail / (a_stage(m) + w_chunk_bytes);
      if (pass == 0 && s_res >= 2) { best_mt = m; best_res = 1; best_s = s_res; break; }
      if (pass == 1 && s_str >= 2) { best_mt = m; best_res = 0; best_s = s_str; break; }
      if (force_mt) break;
    }
  }
  if (!best_mt) {
    const int s_res = (avail - w_all) / a_stage(mt), s_str = avail / (a_stage(mt) + w_chunk_bytes);
    if (w_all <= avail && s_res >= 1) { best_mt = mt; best_res = 1; best_s = s_res; }
    else if (s_str >= 1) { best_mt = mt; best_res = 0; best_s = s_str; }
    else return set_error("conv3x3: tile does not fit in shared memory (KC=%d BN=%d chunks=%d)", KC, BN, d.num_chunks);
  }
  p.mt = best_mt;
  p.w_resident = best_res;
  p.stages = best_s > kMaxStages ? kMaxStages : best_s;
  p.a_box_bytes = (p.mt * p.rm + 2) * p.cw * RB;
  p.a_stage_bytes = a_stage(p.mt);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.mt * p.nt)) cols <<= 1;
  p.tmem_cols = cols;
  out->smem = kSmemFixed + 1024 + (p.w_resident ? w_all : 0) +
              p.stages * (p.a_stage_bytes + (p.w_resident ? 0 : w_chunk_bytes));

  copy_common(d, &p);

  const int box_rows = p.mt * p.rm + 2;
  if (make_nhwc_tmap(&out->tm0, d.src[0], d.n, d.h, d.w, d.src_ctotal[0], KC, p.cw, box_rows)) return 1;
  if (d.src[1]) {
    if (make_nhwc_tmap(&out->tm1, d.src[1], d.n, d.h, d.w, d.src_ctotal[1], KC, p.cw, box_rows)) return 1;
  } else {
    out->tm1 = out->tm0;
  }

  auto kern = conv3x3_tc_kernel<KC, BN, EXT>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern))) return 1;
  out->kernel = reinterpret_cast<const void*>(kern);
  out->threads = kConvThreads;
  out->grid = p.units_total < sms ? static_cast<int>(p.units_total) : sms;
  return 0;
}


int ESRP_PLAN_TILE_NAME(const esrp_conv3x3_t& d, ConvLaunch* out) {
  constexpr bool X = ESRP_EXT;
  if (d.kc == 64) {
    switch (d.bn) {
      case 16: if constexpr (!X) return plan_conv_t<64, 16, false>(d, out); break;
      case 32: return plan_conv_t<64, 32, X>(d, out);
      case 64: return plan_conv_t<64, 64, X>(d, out);
    }
  } else if (d.kc == 32) {
    switch (d.bn) {
      case 16: if constexpr (!X) return plan_conv_t<32, 16, false>(d, out); break;
      case 32: return plan_conv_t<32, 32, X>(d, out);
      case 64: return plan_conv_t<32, 64, X>(d, out);
    }
  }
  return set_error("conv3x3(tile%s): unsupported kc=%d bn=%d", X ? ", training extensions" : "", d.kc, d.bn);
}

}  // namespace esrp
