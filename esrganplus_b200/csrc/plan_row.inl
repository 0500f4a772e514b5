// plan_row.inl — launch planning + instantiations of conv3x3_row_kernel for one value of ESRP_EXT
// (included by esrp_conv_row.cu with ESRP_EXT = false and esrp_conv_row_ext.cu with ESRP_EXT = true).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

#include "../../include/esrp.h"
#include "conv3x3_row.cuh"
#include "esrp_host.h"

namespace esrp {

// ky-stacked row-streaming kernel (conv3x3_row.cuh)
// PAIR: clusters of two CTAs (cta_group::2 MMAs over two images, half of the weight rows resident per CTA; conv3x3_row.cuh)
template <int KC, int BN, bool AUX, bool EXT, bool PAIR = false>
static int plan_row_t(const esrp_conv3x3_t& d, ConvLaunch* out) {
  constexpr int RB = KC * 2;
  ConvKParams& p = out->params;
  memset(&p, 0, sizeof(p));
  p.n = d.n; p.h = d.h; p.w = d.w;
  const bool has_aux = d.aux_chunks > 0;
  const int nb_rows = (has_aux ? 4 : 3) * BN;
  const int w_chunk_bytes = 3 * (PAIR ? nb_rows / 2 : nb_rows) * RB;  // resident per CTA
  const int w_all = d.num_chunks * w_chunk_bytes;
  p.nt = nb_rows;
  const int nblk = PAIR ? (has_aux ? 7 : 14) : ((has_aux || BN == 64) ? 8 : 16);  // ring positions of the TMEM output-row blocks (conv3x3_row.cuh: NBLK)
  if (has_aux && BN == 64) return set_error("conv3x3(row): bn=64 cannot carry the conv1x1 (TMEM)");
  p.mt = nblk;
  p.cw = kRowTile; p.cw_log2 = 7; p.rm = 1;
  p.x_tiles = (d.w + kRowTile - 1) / kRowTile;
  p.x_step = kRowTile;
  p.units_per_col = d.h;
  p.units_total = static_cast<long long>(PAIR ? d.n / 2 : d.n) * p.x_tiles * d.h;  // (PAIR: rows of the first half of the batch)
  if (p.units_total > 0x7fffffffLL) return set_error("conv3x3: problem too large (%lld rows)", p.units_total);
  p.a_box_bytes = (kRowTile + 2) * RB;
  p.a_stage_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
  // row-buffer ring: D >= 2 buffers of num_chunks tiles each (the producer learns that a buffer is free
  // from the block barrier of the output row its input completed, see conv3x3_row.cuh)
  const int avail = kMaxSmem - kSmemFixed - 1024;
  const int row_bytes = p.a_stage_bytes * d.num_chunks;
  int nbuf = w_all <= avail ? (avail - w_all) / row_bytes : 0;
  if (nbuf > nblk - 2) nbuf = nblk - 2;  // the producer must not be lapped on a block barrier
  // ... which also needs the blocks touched by nbuf consecutive rows to stay below the ring size.  Every segment end adds
  // two blocks; images of a few rows put several segments into that window (tests/test_row_protocol_model.py)
  if (d.h <= nblk - 3 && nbuf > 2) nbuf = 2;
  if (nbuf > kMaxStages) nbuf = kMaxStages;
  const int force = d.variant & 15;
  if (force && force < nbuf) nbuf = force;
  if (nbuf < 2)
    return set_error("conv3x3(row): weights + 2 row buffers do not fit in shared memory (KC=%d BN=%d chunks=%d); split K", KC, BN, d.num_chunks);
  p.w_resident = 1;
  p.stages = nbuf;
  p.tmem_cols = 512;
  out->smem = kSmemFixed + 1024 + w_all + p.stages * row_bytes;
  copy_common(d, &p);
  static const bool no_quad = getenv("ESRP_NO_QUAD") != nullptr;
  p.no_quad = no_quad ? 1 : 0;
  static const int row_alt_env = [] { const char* e = getenv("ESRP_ROW_ALT"); return e ? atoi(e) : -1; }();
  const bool row_alt = row_alt_env >= 0 ? row_alt_env != 0 : (d.variant & ESRP_VARIANT_ROW_ALT) != 0;
  static const bool pair_single = [] { const char* e = getenv("ESRP_PAIR_SINGLE"); return e && atoi(e) != 0; }();
  p.pair_single = (PAIR && pair_single) ? 1 : 0;
  p.row_alt = row_alt ? 2 : 0;  // 2: alternate rows, the idle issuer adds the third arrival on the block barrier
  // One landed-barrier per K-chunk tile (full_bar[b * chunks + c]) where the eight barriers suffice (and, conservatively, the
  // ring depth is even: each issuer warp then meets every phase of the chunk barriers of "its" buffers): the MMAs of a row
  // start when its first chunk is there.  +1.2 .. 2.2 % on the benchmark forward (ESRP_CHUNK_BARS=0 switches it off).
  static const bool no_chunk_bars = [] { const char* e = getenv("ESRP_CHUNK_BARS"); return e && atoi(e) == 0; }();
  p.chunk_bars = (!no_chunk_bars && !p.pair_single && row_alt && d.num_chunks >= 2 && nbuf * d.num_chunks <= kMaxStages && (nbuf % 2) == 0 &&
                  (d.variant & 0x1F00) == 0) ? 1 : 0;   // (not under the ESRP_DBG_* timing variants)
  static const bool no_half = getenv("ESRP_NO_HALF_CHUNK") != nullptr;
  // K-slices of the last chunk beyond k_valid hold zero weights: do not issue them
  p.last_half = (row_alt && !no_half && d.k_valid > 0 && d.k_valid <= d.num_chunks * KC - KC / 2 && KC >= 32) ? 1 : 0;
  if (make_nhwc_tmap(&out->tm0, d.src[0], d.n, d.h, d.w, d.src_ctotal[0], KC, kRowTile + 2, 1)) return 1;
  if (d.src[1]) {
    if (make_nhwc_tmap(&out->tm1, d.src[1], d.n, d.h, d.w, d.src_ctotal[1], KC, kRowTile + 2, 1)) return 1;
  } else {
    out->tm1 = out->tm0;
  }
  auto kern = conv3x3_row_kernel<KC, BN, AUX, EXT, PAIR>;
  if (ensure_max_smem(reinterpret_cast<const void*>(kern))) return 1;
  out->kernel = reinterpret_cast<const void*>(kern);
  out->threads = kRowThreads;
  out->fam = 1; out->kc = KC; out->bn = BN; out->ext = EXT ? 1 : 0;
  const int sms = sm_count();
  if (sms <= 0) return set_error("conv3x3: no CUDA device");
  // co-scheduled slices: groups of nsl CTAs share a row range
  const int per = p.nsl * (PAIR ? 2 : 1);  // CTAs that share a row range
  const int groups = sms / per < 1 ? 1 : sms / per;
  out->grid = (p.units_total < groups ? static_cast<int>(p.units_total) : groups) * per;
  out->cluster = PAIR ? 2 : 1;
  return 0;
}


int ESRP_PLAN_ROW_NAME(const esrp_conv3x3_t& d, ConvLaunch* out) {
  constexpr bool X = ESRP_EXT;
  const bool aux = d.aux_chunks > 0;
  if constexpr (!X) {
    if (d.kc == 64 && d.bn == 16) return aux ? plan_row_t<64, 16, true, false>(d, out) : plan_row_t<64, 16, false, false>(d, out);
    if (d.kc == 32 && d.bn == 16) return aux ? plan_row_t<32, 16, true, false>(d, out) : plan_row_t<32, 16, false, false>(d, out);
  }
  if constexpr (!X) {
    // CTA pairs (ESRP_VARIANT_PAIR, with the row-alternating issuers, on an even batch).  ESRP_PAIR=0 switches them off.
    // Not for co-scheduled slices (conv5): that combination passed its parity cases and memcheck but faulted (illegal
    // address at a multicast commit) under two timing perturbations - a single-issuer experiment and compute-sanitizer
    // racecheck - and the cause was not found; such launches stay on single CTAs.
    static const int pair_env = [] { const char* e = getenv("ESRP_PAIR"); return e ? atoi(e) : 1; }();
    static const int row_alt_env = [] { const char* e = getenv("ESRP_ROW_ALT"); return e ? atoi(e) : -1; }();
    const bool row_alt = row_alt_env >= 0 ? row_alt_env != 0 : (d.variant & ESRP_VARIANT_ROW_ALT) != 0;
    // (experiments: ESRP_PAIR_CHUNKS = bit mask of the K-chunk counts that may run as pairs, e.g. 8 = the three-chunk convs)
    static const int pair_chunks = [] { const char* e = getenv("ESRP_PAIR_CHUNKS"); return e ? atoi(e) : ~0; }();
    if (pair_env && ((pair_chunks >> d.num_chunks) & 1) && (d.variant & ESRP_VARIANT_PAIR) && d.slices <= 1 && row_alt && d.kc == 64 && d.bn == 32 && d.n >= 2 && (d.n % 2) == 0 && (d.variant & 0x1F00) == 0 &&
        sm_count() >= 2)
      return aux ? plan_row_t<64, 32, true, false, true>(d, out) : plan_row_t<64, 32, false, false, true>(d, out);
  }
  if (d.kc == 64 && d.bn == 32) return aux ? plan_row_t<64, 32, true, X>(d, out) : plan_row_t<64, 32, false, X>(d, out);
  if (d.kc == 32 && d.bn == 32) return aux ? plan_row_t<32, 32, true, X>(d, out) : plan_row_t<32, 32, false, X>(d, out);
  if (d.kc == 64 && d.bn == 64) return plan_row_t<64, 64, false, X>(d, out);
  if (d.kc == 32 && d.bn == 64) return plan_row_t<32, 64, false, X>(d, out);
  return set_error("conv3x3(row%s): unsupported kc=%d bn=%d (kc in {32,64}, bn in {%s32,64})", X ? ", training extensions" : "", d.kc, d.bn, X ? "" : "16,");
}

}  // namespace esrp
