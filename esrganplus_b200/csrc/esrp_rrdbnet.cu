// esrp_rrdbnet.cu — native host runtime for the RRDBNet generator forward pass
// (reference: codes/models/modules/architecture.py:47-78, block.py:232-291).
//
// The engine owns (a) the packed bf16 weight cache derived from the reference-format fp32
// state_dict tensors and (b) per-shape launch plans: a flat list of pre-planned kernel launches
// (TMA maps encoded, shared-memory/TMEM budgets fixed) that replays with one cudaLaunchKernel per
// step and no host-side allocation or synchronisation, so the whole forward is capturable in a
// CUDA graph.  Activations live in a caller-provided workspace.
//
// Dataflow per ResidualDenseBlock_5C (block.py:260-268), all NHWC bf16, concat-free:
//   T_in [px, nf] (+ fp32 twin)   G [px, 4*gc]  (x1..x4 written at channel offsets 0, gc, 2gc, 3gc)
//   conv1: K = T_in                         -> G[0:gc]      lrelu
//   conv2: K = T_in | G[0:gc]    (+1x1 aux) -> G[gc:2gc]    lrelu, + conv1x1(x)
//   conv3: K = T_in | G[0:2gc]              -> G[2gc:3gc]   lrelu
//   conv4: K = T_in | G[0:3gc]              -> G[3gc:4gc]   lrelu, + x2
//   conv5: K = T_in | G[0:4gc]              -> T_out        0.2*(.) + x [noise] [RRDB: 0.2*(.) + x_rrdb]
// The residual trunk is carried in fp32 alongside its bf16 MMA-operand copy so that 69 chained
// residual adds do not accumulate bf16 rounding.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/esrp.h"
#include "esrp_host.h"

namespace esrp {

namespace {

struct ConvW {
  // static description
  int cin = 0, cout = 0;       // logical channels
  int kc = 0, bn = 0;
  int num_chunks = 0;
  int lc0[ESRP_MAX_CHUNKS] = {0};  // logical first input channel per chunk
  int aux_chunks = 0;              // leading chunks feeding the fused 1x1
  int w_idx = -1, b_idx = -1, aux_idx = -1;  // indices into the key list
  // packed device storage (offsets into wbuf)
  size_t w_off = 0, b_off = 0, aux_off = 0;
};

struct Step {
  enum Kind { kConv, kPackInput, kUpsample } kind = kConv;
  ConvLaunch conv;
  // elementwise steps
  const void* src = nullptr;
  void* dst = nullptr;
  int n = 0, h = 0, w = 0, c = 0, c_pad = 0;
  bool patch_y = false;      // conv writes the caller's output tensor
  bool is_noise = false;     // conv5 with GaussianNoise (training)
  int noise_index = 0;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int bn_for(int cout) { return cout <= 16 ? 16 : (cout <= 32 ? 32 : 64); }

}  // namespace

struct Rrdbnet {
  int in_nc, out_nc, nf, nb, gc, upscale;
  int in_pad;  // input channels padded to a multiple of 32
  std::vector<std::string> keys;
  std::vector<std::vector<int>> shapes;
  // convs: fea, [nb][3][5] rdb convs, trunk, up1, up2, hr0, hr1
  ConvW fea, trunk, up[2], hr0, hr1;
  std::vector<ConvW> rdb;  // nb*3*5
  uint8_t* wbuf = nullptr;
  size_t wbytes = 0;
  bool weights_loaded = false;
  // plan cache (single entry: the common case is a fixed shape)
  int pn = 0, ph = 0, pw = 0, ptraining = -1;
  void* pws = nullptr;
  std::vector<Step> steps;
};

namespace {

int add_key(Rrdbnet* m, const std::string& k, std::vector<int> shape) {
  m->keys.push_back(k);
  m->shapes.push_back(std::move(shape));
  return static_cast<int>(m->keys.size()) - 1;
}

// Chunking of a conv whose logical input is [x: nf] ++ [growth: ng channels of G].
void set_chunks(ConvW* c, int nf, int ng, int kc_pref) {
  c->kc = kc_pref;
  c->num_chunks = 0;
  for (int ch = 0; ch < nf + ng; ch += c->kc) c->lc0[c->num_chunks++] = ch;
}

void define_conv(Rrdbnet* m, ConvW* c, const std::string& key, int cin, int cout, int nf_part, int ng_part,
                 bool bias = true) {
  c->cin = cin;
  c->cout = cout;
  c->bn = bn_for(cout);
  // 64-wide chunks (128 B swizzle) when both segments are multiples of 64, else 32-wide (64 B swizzle)
  const int kc = (nf_part % 64 == 0 && ng_part % 64 == 0) ? 64 : 32;
  set_chunks(c, nf_part, ng_part, kc);
  c->w_idx = add_key(m, key + ".weight", {cout, cin, 3, 3});
  if (bias) c->b_idx = add_key(m, key + ".bias", {cout});
}

}  // namespace
}  // namespace esrp

using namespace esrp;

extern "C" {

int esrp_rrdbnet_create(int32_t in_nc, int32_t out_nc, int32_t nf, int32_t nb, int32_t gc, int32_t upscale,
                        esrp_rrdbnet_t** out) {
  if (!out) return set_error("rrdbnet_create: null out");
  if (nf % 32 || nf < 32 || nf > 64) return set_error("rrdbnet_create: nf=%d unsupported (32 or 64)", nf);
  if (gc != 32) return set_error("rrdbnet_create: gc=%d unsupported (the reference hard-codes gc=32, architecture.py:56)", gc);
  if (upscale != 4 && upscale != 2 && upscale != 1) return set_error("rrdbnet_create: upscale=%d unsupported (1, 2, 4)", upscale);
  if (in_nc < 1 || in_nc > 64 || out_nc < 1 || out_nc > 64) return set_error("rrdbnet_create: in_nc/out_nc out of range");
  if (nb < 1 || nb > 64) return set_error("rrdbnet_create: nb=%d out of range", nb);
  Rrdbnet* m = new Rrdbnet();
  m->in_nc = in_nc; m->out_nc = out_nc; m->nf = nf; m->nb = nb; m->gc = gc; m->upscale = upscale;
  m->in_pad = (in_nc + 31) / 32 * 32;
  const int n_up = upscale == 4 ? 2 : (upscale == 2 ? 1 : 0);

  // key order == reference state_dict order (sequential() flattening, block.py:95-108)
  define_conv(m, &m->fea, "model.0", in_nc, nf, m->in_pad, 0);
  m->fea.cin = in_nc;
  m->rdb.resize(static_cast<size_t>(nb) * 15);
  for (int i = 0; i < nb; ++i) {
    for (int r = 0; r < 3; ++r) {
      const std::string p = "model.1.sub." + std::to_string(i) + ".RDB" + std::to_string(r + 1) + ".";
      const int aux_key = add_key(m, p + "conv1x1.weight", {gc, nf, 1, 1});
      for (int k = 0; k < 5; ++k) {
        ConvW* c = &m->rdb[(static_cast<size_t>(i) * 3 + r) * 5 + k];
        define_conv(m, c, p + "conv" + std::to_string(k + 1) + ".0", nf + k * gc, k == 4 ? nf : gc, nf, k * gc);
        if (k == 1) {
          c->aux_idx = aux_key;
          c->aux_chunks = nf / c->kc;  // the x chunks come first
        }
      }
    }
  }
  define_conv(m, &m->trunk, "model.1.sub." + std::to_string(nb), nf, nf, nf, 0);
  for (int u = 0; u < n_up; ++u) define_conv(m, &m->up[u], "model." + std::to_string(3 + 3 * u), nf, nf, nf, 0);
  define_conv(m, &m->hr0, "model." + std::to_string(2 + 3 * n_up), nf, nf, nf, 0);
  define_conv(m, &m->hr1, "model." + std::to_string(4 + 3 * n_up), nf, out_nc, nf, 0);

  // packed weight storage
  size_t off = 0;
  auto reserve = [&](ConvW* c) {
    c->w_off = off;
    off = align_up(off + static_cast<size_t>(esrp_packed_conv3x3_bytes(c->num_chunks, c->kc, c->bn)), 1024);
    c->b_off = off;
    off = align_up(off + static_cast<size_t>(c->bn) * 4, 1024);
    if (c->aux_idx >= 0) {
      c->aux_off = off;
      off = align_up(off + static_cast<size_t>(esrp_packed_conv1x1_bytes(c->aux_chunks, c->kc, c->bn)), 1024);
    }
  };
  reserve(&m->fea);
  for (auto& c : m->rdb) reserve(&c);
  reserve(&m->trunk);
  for (int u = 0; u < n_up; ++u) reserve(&m->up[u]);
  reserve(&m->hr0);
  reserve(&m->hr1);
  m->wbytes = off;
  cudaError_t e = cudaMalloc(&m->wbuf, m->wbytes);
  if (e != cudaSuccess) {
    delete m;
    return set_error("rrdbnet_create: cudaMalloc(%zu) failed: %s", off, cudaGetErrorString(e));
  }
  *out = reinterpret_cast<esrp_rrdbnet_t*>(m);
  return 0;
}

void esrp_rrdbnet_destroy(esrp_rrdbnet_t* h) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m) return;
  if (m->wbuf) cudaFree(m->wbuf);
  delete m;
}

int32_t esrp_rrdbnet_num_tensors(const esrp_rrdbnet_t* h) {
  return h ? static_cast<int32_t>(reinterpret_cast<const Rrdbnet*>(h)->keys.size()) : -1;
}

const char* esrp_rrdbnet_tensor_key(const esrp_rrdbnet_t* h, int32_t idx) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  if (!m || idx < 0 || idx >= static_cast<int32_t>(m->keys.size())) return nullptr;
  return m->keys[idx].c_str();
}

int esrp_rrdbnet_tensor_shape(const esrp_rrdbnet_t* h, int32_t idx, int32_t* dims4) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  if (!m || !dims4 || idx < 0 || idx >= static_cast<int32_t>(m->keys.size())) return set_error("tensor_shape: bad index");
  for (int i = 0; i < 4; ++i) dims4[i] = i < static_cast<int>(m->shapes[idx].size()) ? m->shapes[idx][i] : 0;
  return 0;
}

int esrp_rrdbnet_load_weights(esrp_rrdbnet_t* h, const void* const* ptrs, int32_t count, void* stream) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m || !ptrs) return set_error("rrdbnet_load_weights: null argument");
  if (count != static_cast<int32_t>(m->keys.size()))
    return set_error("rrdbnet_load_weights: expected %zu tensors, got %d", m->keys.size(), count);
  for (int i = 0; i < count; ++i)
    if (!ptrs[i]) return set_error("rrdbnet_load_weights: tensor %d (%s) is null", i, m->keys[i].c_str());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto pack = [&](ConvW* c) -> int {
    if (esrp_pack_conv3x3_weights(static_cast<const float*>(ptrs[c->w_idx]), c->cout, c->cin, c->kc, c->bn,
                                  c->num_chunks, c->lc0, m->wbuf + c->w_off, stream))
      return 1;
    ESRP_CUDA_OK(cudaMemsetAsync(m->wbuf + c->b_off, 0, static_cast<size_t>(c->bn) * 4, s));
    if (c->b_idx >= 0)
      ESRP_CUDA_OK(cudaMemcpyAsync(m->wbuf + c->b_off, ptrs[c->b_idx], static_cast<size_t>(c->cout) * 4,
                                   cudaMemcpyDeviceToDevice, s));
    if (c->aux_idx >= 0) {
      if (esrp_pack_conv1x1_weights(static_cast<const float*>(ptrs[c->aux_idx]), c->cout, m->nf, c->kc, c->bn,
                                    c->aux_chunks, c->lc0, m->wbuf + c->aux_off, stream))
        return 1;
    }
    return 0;
  };
  if (pack(&m->fea)) return 1;
  for (auto& c : m->rdb)
    if (pack(&c)) return 1;
  if (pack(&m->trunk)) return 1;
  const int n_up = m->upscale == 4 ? 2 : (m->upscale == 2 ? 1 : 0);
  for (int u = 0; u < n_up; ++u)
    if (pack(&m->up[u])) return 1;
  if (pack(&m->hr0)) return 1;
  if (pack(&m->hr1)) return 1;
  m->weights_loaded = true;
  return 0;
}

}  // extern "C"

namespace esrp {
namespace {

struct Workspace {
  size_t xin, fea_b, fea_f, tb[3], tf[3], g, u0, hr_a, hr_b, hr_c, total;
};

Workspace layout(const Rrdbnet* m, int n, int h, int w) {
  Workspace ws;
  const size_t px = static_cast<size_t>(n) * h * w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  ws.xin = take(px * m->in_pad * 2);
  ws.fea_b = take(px * m->nf * 2);
  ws.fea_f = take(px * m->nf * 4);
  for (int i = 0; i < 3; ++i) {
    ws.tb[i] = take(px * m->nf * 2);
    ws.tf[i] = take(px * m->nf * 4);
  }
  ws.g = take(px * 4 * m->gc * 2);
  ws.u0 = take(px * m->nf * 2);
  const size_t s = static_cast<size_t>(m->upscale) * m->upscale;
  // three rotating full-resolution buffers (upsampled input / conv output), sized for the HR grid
  ws.hr_a = take(px * s * m->nf * 2);
  ws.hr_b = take(px * s * m->nf * 2);
  ws.hr_c = take(px * s * m->nf * 2);
  ws.total = off;
  return ws;
}

// Fill the common part of a conv descriptor from packed weights.
void base_desc(const Rrdbnet* m, const ConvW& c, int n, int h, int w, esrp_conv3x3_t* d) {
  memset(d, 0, sizeof(*d));
  d->n = n; d->h = h; d->w = w;
  d->kc = c.kc;
  d->num_chunks = c.num_chunks;
  d->bn = c.bn;
  d->cout = c.cout;
  d->w_packed = m->wbuf + c.w_off;
  d->bias = reinterpret_cast<const float*>(m->wbuf + c.b_off);
  d->s0 = 1.f; d->s1 = 1.f; d->s2 = 1.f;
  d->sigma = 0.1f;
}

int build_plan(Rrdbnet* m, int n, int h, int w, uint8_t* wsp, int training) {
  m->steps.clear();
  const Workspace ws = layout(m, n, h, w);
  const int nf = m->nf, gc = m->gc;
  auto push_conv = [&](const esrp_conv3x3_t& d, bool patch_y = false, bool is_noise = false, int noise_index = 0) -> int {
    Step st;
    st.kind = Step::kConv;
    if (plan_conv(d, &st.conv)) return 1;
    st.patch_y = patch_y;
    st.is_noise = is_noise;
    st.noise_index = noise_index;
    m->steps.push_back(st);
    return 0;
  };
  // 0. NCHW fp32 -> NHWC bf16 (channels zero-padded to in_pad)
  {
    Step st;
    st.kind = Step::kPackInput;
    st.dst = wsp + ws.xin;
    st.n = n; st.h = h; st.w = w; st.c = m->in_nc; st.c_pad = m->in_pad;
    m->steps.push_back(st);
  }
  esrp_conv3x3_t d;
  // 1. fea_conv (architecture.py:55): no activation; bf16 + fp32 outputs
  base_desc(m, m->fea, n, h, w, &d);
  d.src[0] = wsp + ws.xin; d.src_ctotal[0] = m->in_pad;
  for (int i = 0; i < m->fea.num_chunks; ++i) { d.chunk_src[i] = 0; d.chunk_c0[i] = m->fea.lc0[i]; }
  d.out_bf16 = wsp + ws.fea_b; d.ob_ctotal = nf;
  d.out_f32 = wsp + ws.fea_f; d.of_ctotal = nf;
  if (push_conv(d)) return 1;

  // 2. RRDB trunk
  // trunk buffer rotation: `cur` holds the current RDB input, `rr` the RRDB input (for block.py:291)
  const uint8_t* cur_b = wsp + ws.fea_b;
  const uint8_t* cur_f = wsp + ws.fea_f;
  int noise_index = 0;
  for (int i = 0; i < m->nb; ++i) {
    const uint8_t* rr_f = cur_f;
    for (int r = 0; r < 3; ++r) {
      // pick an output slot that is neither the current input nor the RRDB input
      int slot = -1;
      for (int s = 0; s < 3; ++s) {
        const uint8_t* cand = wsp + ws.tf[s];
        if (cand != cur_f && cand != rr_f) { slot = s; break; }
      }
      uint8_t* out_b = wsp + ws.tb[slot];
      uint8_t* out_f = wsp + ws.tf[slot];
      uint8_t* G = wsp + ws.g;
      for (int k = 0; k < 5; ++k) {
        const ConvW& c = m->rdb[(static_cast<size_t>(i) * 3 + r) * 5 + k];
        base_desc(m, c, n, h, w, &d);
        d.src[0] = cur_b; d.src_ctotal[0] = nf;
        d.src[1] = G; d.src_ctotal[1] = 4 * gc;
        for (int ch = 0; ch < c.num_chunks; ++ch) {
          const int lc = c.lc0[ch];
          d.chunk_src[ch] = lc < nf ? 0 : 1;
          d.chunk_c0[ch] = lc < nf ? lc : lc - nf;
        }
        if (k == 0 && c.num_chunks * c.kc == nf) d.src[1] = nullptr;
        if (k < 4) {
          d.act = 1;
          d.out_bf16 = G; d.ob_ctotal = 4 * gc; d.ob_c0 = k * gc;
          if (k == 1) {  // x2 = lrelu(conv2) + conv1x1(x)   (block.py:262-263)
            d.aux_chunks = c.aux_chunks;
            d.w_aux = m->wbuf + c.aux_off;
          }
          if (k == 3) {  // x4 = lrelu(conv4) + x2            (block.py:265-266)
            d.r1 = G; d.r1_is_f32 = 0; d.r1_ctotal = 4 * gc; d.r1_c0 = gc; d.s1 = 1.f;
          }
          if (push_conv(d)) return 1;
        } else {
          // out = noise(0.2 * x5 + x)                         (block.py:268)
          d.act = 0; d.s0 = 0.2f;
          d.r1 = cur_f; d.r1_is_f32 = 1; d.r1_ctotal = nf; d.r1_c0 = 0; d.s1 = 1.f;
          d.noise = training ? 1 : 0;
          if (r == 2) {  // RRDB: out * 0.2 + x                 (block.py:291)
            d.r2 = rr_f; d.r2_is_f32 = 1; d.r2_ctotal = nf; d.r2_c0 = 0; d.s2 = 0.2f;
          }
          d.out_bf16 = out_b; d.ob_ctotal = nf;
          d.out_f32 = out_f; d.of_ctotal = nf;
          if (push_conv(d, false, training != 0, noise_index++)) return 1;
        }
      }
      cur_b = out_b;
      cur_f = out_f;
    }
  }
  // 3. LR_conv + shortcut (architecture.py:58,73; block.py:84-86)
  base_desc(m, m->trunk, n, h, w, &d);
  d.src[0] = cur_b; d.src_ctotal[0] = nf;
  for (int i = 0; i < m->trunk.num_chunks; ++i) { d.chunk_src[i] = 0; d.chunk_c0[i] = m->trunk.lc0[i]; }
  d.r1 = wsp + ws.fea_f; d.r1_is_f32 = 1; d.r1_ctotal = nf; d.s1 = 1.f;
  d.out_bf16 = wsp + ws.u0; d.ob_ctotal = nf;
  if (push_conv(d)) return 1;

  // 4. upconv blocks: nearest x2 -> conv -> lrelu (block.py:315-322)
  const int n_up = m->upscale == 4 ? 2 : (m->upscale == 2 ? 1 : 0);
  const uint8_t* feat = wsp + ws.u0;
  int ch_ = h, cw_ = w;
  uint8_t* hr[3] = {wsp + ws.hr_a, wsp + ws.hr_b, wsp + ws.hr_c};
  int hr_i = 0;
  for (int u = 0; u < n_up; ++u) {
    Step st;
    st.kind = Step::kUpsample;
    st.src = feat;
    st.dst = hr[hr_i];
    st.n = n; st.h = ch_; st.w = cw_; st.c = nf;
    m->steps.push_back(st);
    ch_ *= 2; cw_ *= 2;
    base_desc(m, m->up[u], n, ch_, cw_, &d);
    d.src[0] = hr[hr_i]; d.src_ctotal[0] = nf;
    for (int i = 0; i < m->up[u].num_chunks; ++i) { d.chunk_src[i] = 0; d.chunk_c0[i] = m->up[u].lc0[i]; }
    d.act = 1;
    d.out_bf16 = hr[(hr_i + 1) % 3]; d.ob_ctotal = nf;
    if (push_conv(d)) return 1;
    feat = hr[(hr_i + 1) % 3];
    hr_i = (hr_i + 2) % 3;
  }
  // 5. HR_conv0 + lrelu (architecture.py:70)
  base_desc(m, m->hr0, n, ch_, cw_, &d);
  d.src[0] = feat; d.src_ctotal[0] = nf;
  for (int i = 0; i < m->hr0.num_chunks; ++i) { d.chunk_src[i] = 0; d.chunk_c0[i] = m->hr0.lc0[i]; }
  d.act = 1;
  uint8_t* hr0_out = hr[hr_i];
  if (hr0_out == feat) hr0_out = hr[(hr_i + 1) % 3];
  d.out_bf16 = hr0_out; d.ob_ctotal = nf;
  if (push_conv(d)) return 1;
  // 6. HR_conv1 (architecture.py:71): NCHW fp32 straight into the caller's output tensor
  base_desc(m, m->hr1, n, ch_, cw_, &d);
  d.src[0] = hr0_out; d.src_ctotal[0] = nf;
  for (int i = 0; i < m->hr1.num_chunks; ++i) { d.chunk_src[i] = 0; d.chunk_c0[i] = m->hr1.lc0[i]; }
  d.out_nchw = reinterpret_cast<float*>(wsp);  // placeholder, patched per call
  if (push_conv(d, true)) return 1;

  m->pn = n; m->ph = h; m->pw = w; m->pws = wsp; m->ptraining = training;
  return 0;
}

}  // namespace
}  // namespace esrp

extern "C" {

int64_t esrp_rrdbnet_workspace_bytes(const esrp_rrdbnet_t* h, int32_t n, int32_t hh, int32_t w) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  if (!m || n < 1 || hh < 1 || w < 1) return -1;
  return static_cast<int64_t>(layout(m, n, hh, w).total);
}

int32_t esrp_rrdbnet_num_launches(const esrp_rrdbnet_t* h) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  return m ? static_cast<int32_t>(m->steps.size()) : -1;
}

int esrp_rrdbnet_forward(esrp_rrdbnet_t* h, const float* x, float* y, int32_t n, int32_t hh, int32_t w,
                         void* workspace, int64_t workspace_bytes, int32_t training, uint64_t seed, void* stream) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m || !x || !y || !workspace) return set_error("rrdbnet_forward: null argument");
  if (!m->weights_loaded) return set_error("rrdbnet_forward: weights not loaded");
  if (n < 1 || hh < 1 || w < 1) return set_error("rrdbnet_forward: bad shape");
  const int64_t need = esrp_rrdbnet_workspace_bytes(h, n, hh, w);
  if (workspace_bytes < need) return set_error("rrdbnet_forward: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return set_error("rrdbnet_forward: workspace must be 1024-byte aligned");
  if (m->pn != n || m->ph != hh || m->pw != w || m->pws != workspace || m->ptraining != (training ? 1 : 0)) {
    if (build_plan(m, n, hh, w, static_cast<uint8_t*>(workspace), training ? 1 : 0)) {
      m->pn = 0;
      return 1;
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (Step& st : m->steps) {
    switch (st.kind) {
      case Step::kPackInput:
        if (esrp_nchw_f32_to_nhwc_bf16(x, st.dst, st.n, st.c, st.h, st.w, st.c_pad, stream)) return 1;
        break;
      case Step::kUpsample:
        if (esrp_upsample2x_nhwc_bf16(st.src, st.dst, st.n, st.h, st.w, st.c, stream)) return 1;
        break;
      case Step::kConv:
        if (st.patch_y) st.conv.params.out_nchw = y;
        if (st.is_noise) {
          st.conv.params.seed = seed;
          st.conv.params.offset = static_cast<unsigned long long>(st.noise_index) << 36;
        }
        if (run_conv(st.conv, s)) return 1;
        break;
    }
  }
  return 0;
}

}  // extern "C"
