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
// Every dense-block launch computes at most 32 output channels (conv5 is two launches over disjoint
// weight rows): they share the N = 3*32 tap-stacked MMA shape and keep their weights resident in shared
// memory.  Single-chunk 64->64 convs (trunk, upconv, HR) run as one N = 3*64 launch.  With nf = 64 the K dimension is walked in 64-channel chunks (128-byte
// swizzle); a chunk may cover growth channels that are not computed yet (e.g. conv2 reads G[0:64]
// but only G[0:32] is x1) — their weights are packed as zeros, and G is zero-initialised so stale
// values are always finite.
// The residual trunk is carried in fp32 alongside its bf16 MMA-operand copy so that 69 chained
// residual adds do not accumulate bf16 rounding.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/esrp.h"
#include "esrp_host.h"
#include "esrp_rrdbnet_int.h"

namespace esrp {

namespace {

int add_key(Rrdbnet* m, const std::string& k, std::vector<int> shape) {
  m->keys.push_back(k);
  m->shapes.push_back(std::move(shape));
  return static_cast<int>(m->keys.size()) - 1;
}

// One logical conv [cout, cin, 3, 3] whose input is [x: nf_part] ++ [growth: ng_part channels of G],
// cut into launches of <= 32 output channels.
void define_conv(Rrdbnet* m, std::vector<ConvW>* out, const std::string& key, int cin, int cout, int nf_part,
                 int ng_part, int aux_key = -1) {
  ConvW c;
  c.cin = cin;
  c.cout = cout;
  c.kc = (nf_part % 64 == 0) ? 64 : 32;
  for (int ch = 0; ch < nf_part + ng_part; ch += c.kc) c.lc0[c.num_chunks++] = ch;
  c.w_idx = add_key(m, key + ".weight", {cout, cin, 3, 3});
  c.b_idx = add_key(m, key + ".bias", {cout});
  if (aux_key >= 0) {
    c.aux_idx = aux_key;
    c.aux_chunks = nf_part / c.kc;  // the x chunks come first
  }
  // <= 32 output channels per launch; a conv whose whole K is one chunk keeps 64 (its 3 x 192-row weight
  // block still fits in shared memory, the MMA runs at N = 192 and the input is read once)
  const int slice = (c.num_chunks == 1 && aux_key < 0) ? 64 : 32;
  for (int r0 = 0; r0 < cout; r0 += slice) {
    c.row0 = r0;
    c.rows = cout - r0 < slice ? cout - r0 : slice;
    c.bn = c.rows <= 16 ? 16 : (c.rows <= 32 ? 32 : 64);
    out->push_back(c);
  }
}

}  // namespace

int pack_one(Rrdbnet* m, ConvW* c, int layout, cudaStream_t s) {
  const void* const* ptrs = m->src_ptrs.data();
  const float* aux = c->aux_idx >= 0 ? static_cast<const float*>(ptrs[c->aux_idx]) : nullptr;
  if (esrp_pack_conv3x3_weights(static_cast<const float*>(ptrs[c->w_idx]), c->cout, c->cin, 0, layout, c->row0,
                                c->rows, c->kc, c->bn, c->num_chunks, c->lc0, aux, m->nf, c->aux_chunks,
                                m->wbuf + c->w_off, s))
    return 1;
  ESRP_CUDA_OK(cudaMemsetAsync(m->wbuf + c->b_off, 0, static_cast<size_t>(c->bn) * 4, s));
  ESRP_CUDA_OK(cudaMemcpyAsync(m->wbuf + c->b_off, static_cast<const float*>(ptrs[c->b_idx]) + c->row0,
                               static_cast<size_t>(c->rows) * 4, cudaMemcpyDeviceToDevice, s));
  c->layout = layout;
  return 0;
}

int ensure_layouts(Rrdbnet* m, int w, cudaStream_t s) {
  auto want = [&](std::vector<ConvW>& cs, int width) -> int {
    const int lay = layout_for_width(width);
    for (auto& c : cs)
      if (c.layout != lay && pack_one(m, &c, lay, s)) return 1;
    return 0;
  };
  if (want(m->fea, w) || want(m->rdb, w) || want(m->trunk, w)) return 1;
  for (auto& c : m->rdb5w)
    if (c.layout != ESRP_LAYOUT_TILE && pack_one(m, &c, ESRP_LAYOUT_TILE, s)) return 1;
  int ww = w;
  for (int u = 0; u < m->n_up; ++u) {
    ww *= 2;
    if (want(m->up[u], ww)) return 1;
  }
  return want(m->hr0, ww) || want(m->hr1, ww);
}

}  // namespace esrp

using namespace esrp;

extern "C" {

int esrp_rrdbnet_create(int32_t in_nc, int32_t out_nc, int32_t nf, int32_t nb, int32_t gc, int32_t upscale,
                        esrp_rrdbnet_t** out) {
  if (!out) return set_error("rrdbnet_create: null out");
  if (nf % 32 || nf < 32 || nf > 64) return set_error("rrdbnet_create: nf=%d unsupported (32 or 64)", nf);
  if (gc != 32) return set_error("rrdbnet_create: gc=%d unsupported (the reference hard-codes gc=32, architecture.py:56)", gc);
  if (upscale != 4 && upscale != 2 && upscale != 1) return set_error("rrdbnet_create: upscale=%d unsupported (1, 2, 4)", upscale);
  if (in_nc < 1 || in_nc > 32 || out_nc < 1 || out_nc > 64) return set_error("rrdbnet_create: in_nc/out_nc out of range");
  if (nb < 1 || nb > 64) return set_error("rrdbnet_create: nb=%d out of range", nb);
  Rrdbnet* m = new Rrdbnet();
  m->in_nc = in_nc; m->out_nc = out_nc; m->nf = nf; m->nb = nb; m->gc = gc; m->upscale = upscale;
  m->in_pad = 32;
  m->n_up = upscale == 4 ? 2 : (upscale == 2 ? 1 : 0);

  // key order == reference state_dict order (sequential() flattening, block.py:95-108)
  define_conv(m, &m->fea, "model.0", in_nc, nf, m->in_pad, 0);
  m->per_rdb = 4 + nf / 32;
  for (int i = 0; i < nb; ++i) {
    for (int r = 0; r < 3; ++r) {
      const std::string p = "model.1.sub." + std::to_string(i) + ".RDB" + std::to_string(r + 1) + ".";
      const int aux_key = add_key(m, p + "conv1x1.weight", {gc, nf, 1, 1});
      for (int k = 0; k < 5; ++k)
        define_conv(m, &m->rdb, p + "conv" + std::to_string(k + 1) + ".0", nf + k * gc, k == 4 ? nf : gc, nf, k * gc,
                    k == 1 ? aux_key : -1);
      if (nf == 64) {
        ConvW wide = m->rdb[m->rdb.size() - 2];  // first conv5 slice: same tensors, all 64 output channels
        wide.row0 = 0; wide.rows = 64; wide.bn = 64;
        wide.layout = ESRP_LAYOUT_TILE;
        m->rdb5w.push_back(wide);
      }
    }
  }
  define_conv(m, &m->trunk, "model.1.sub." + std::to_string(nb), nf, nf, nf, 0);
  for (int u = 0; u < m->n_up; ++u) define_conv(m, &m->up[u], "model." + std::to_string(3 + 3 * u), nf, nf, nf, 0);
  define_conv(m, &m->hr0, "model." + std::to_string(2 + 3 * m->n_up), nf, nf, nf, 0);
  define_conv(m, &m->hr1, "model." + std::to_string(4 + 3 * m->n_up), nf, out_nc, nf, 0);

  // packed weight storage
  size_t off = 0;
  for_each_conv(m, [&](ConvW* c) {
    c->w_off = off;
    off = align_up(off + static_cast<size_t>(esrp_packed_conv3x3_bytes(c->num_chunks, c->kc, c->bn, c->aux_chunks > 0)), 1024);
    c->b_off = off;
    off = align_up(off + static_cast<size_t>(c->bn) * 4, 1024);
    return 0;
  });
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
  destroy_train_state(m);
  for (Step& st : m->steps)
    if (st.kind == Step::kChain) free_chain(&st.chain);
  for (cudaEvent_t e : m->ev0) cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev1) cudaEventDestroy(e);
  if (m->pack_jobs_dev) cudaFree(m->pack_jobs_dev);
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
  m->src_ptrs.assign(ptrs, ptrs + count);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // keep each conv in the layout its last plan asked for (ROW until a plan says otherwise); one batched launch
  m->pack_jobs_host.clear();
  for_each_conv(m, [&](ConvW* c) {
    if (c->layout < 0) c->layout = ESRP_LAYOUT_ROW;
    PackJob j;
    memset(&j, 0, sizeof(j));
    j.type = 0; j.layout = c->layout; j.row0 = c->row0; j.rows = c->rows; j.kc = c->kc; j.bn = c->bn;
    j.num_chunks = c->num_chunks;
    j.out = reinterpret_cast<__nv_bfloat16*>(m->wbuf + c->w_off);
    j.w = static_cast<const float*>(ptrs[c->w_idx]);
    j.w_o = c->cout; j.w_i = c->cin;
    for (int i = 0; i < c->num_chunks; ++i) j.lc0[i] = c->lc0[i];
    j.aux = c->aux_idx >= 0 ? static_cast<const float*>(ptrs[c->aux_idx]) : nullptr;
    j.aux_cin = m->nf; j.aux_chunks = c->aux_chunks;
    j.bias_src = static_cast<const float*>(ptrs[c->b_idx]);
    j.bias_dst = reinterpret_cast<float*>(m->wbuf + c->b_off);
    m->pack_jobs_host.push_back(j);
    return 0;
  });
  if (!m->pack_jobs_dev) ESRP_CUDA_OK(cudaMalloc(&m->pack_jobs_dev, sizeof(PackJob) * m->pack_jobs_host.size()));
  ESRP_CUDA_OK(cudaMemcpyAsync(m->pack_jobs_dev, m->pack_jobs_host.data(), sizeof(PackJob) * m->pack_jobs_host.size(),
                               cudaMemcpyHostToDevice, s));
  if (run_pack_batch(m->pack_jobs_dev, static_cast<int>(m->pack_jobs_host.size()), s)) return 1;
  m->weights_loaded = true;
  ++m->weights_version;
  return 0;
}

}  // extern "C"

namespace esrp {

// Fill the common part of a conv descriptor from packed weights.
void base_desc(const Rrdbnet* m, const ConvW& c, int n, int h, int w, esrp_conv3x3_t* d) {
  memset(d, 0, sizeof(*d));
  d->n = n; d->h = h; d->w = w;
  d->kc = c.kc;
  d->num_chunks = c.num_chunks;
  d->bn = c.bn;
  d->cout = c.rows;
  d->w_packed = m->wbuf + c.w_off;
  d->w_layout = c.layout;
  d->bias = reinterpret_cast<const float*>(m->wbuf + c.b_off);
  d->s0 = 1.f; d->s1 = 1.f; d->s2 = 1.f;
  d->sigma = 0.1f;
}

// Timing experiments: ESRP_NO_COSLICE=1 issues the output slices of a wide conv as separate launches again.
bool no_coslice() {
  static const bool off = getenv("ESRP_NO_COSLICE") != nullptr;
  return off;
}

// Turn the descriptor of slice `c` into ONE launch over `nsl` consecutive slices (esrp_conv3x3_t::slices): their
// packed weights and biases sit at a constant stride in the weight buffer (rrdbnet_create allocates them back to back).
int coslice(const ConvW& c, const ConvW& next, int nsl, esrp_conv3x3_t* d) {
  const long long stride = static_cast<long long>(next.w_off) - static_cast<long long>(c.w_off);
  if (stride <= 0 || static_cast<long long>(next.b_off) - static_cast<long long>(c.b_off) != stride || next.bn != c.bn ||
      next.row0 != c.row0 + c.bn || c.rows != c.bn)
    return set_error("rrdbnet: conv slices are not laid out at a constant stride");
  d->slices = nsl;
  d->slice_stride = stride;
  return 0;
}

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

void clear_steps(Rrdbnet* m) {
  for (Step& st : m->steps)
    if (st.kind == Step::kChain) free_chain(&st.chain);
  m->steps.clear();
}

// Merge every run of >= 2 consecutive chain-compatible conv launches (the dense-block convs of the trunk) into one
// persistent launch — when asked to (esrp_rrdbnet_set_chain, or ESRP_CHAIN=1 for every engine of the process).
int merge_chains(Rrdbnet* m) {
  static const bool env_on = [] { const char* e = getenv("ESRP_CHAIN"); return e && atoi(e) != 0; }();
  if (!m->use_chain && !env_on) return 0;
  static const size_t max_run = [] { const char* e = getenv("ESRP_CHAIN_MAX"); return e ? static_cast<size_t>(atoi(e)) : static_cast<size_t>(kChainMaxPhasesHost); }();
  std::vector<Step> out;
  size_t i = 0;
  while (i < m->steps.size()) {
    size_t j = i;
    auto ok = [&](const Step& s) {
      return s.kind == Step::kConv && !s.patch_y && !s.is_noise && chain_compatible(s.conv) &&
             s.conv.grid == m->steps[i].conv.grid;
    };
    while (j < m->steps.size() && j - i < max_run && ok(m->steps[j])) ++j;
    if (j - i >= 2 || (j - i == 1 && max_run == 1)) {
      std::vector<const ConvLaunch*> cs;
      for (size_t k = i; k < j; ++k) cs.push_back(&m->steps[k].conv);
      Step st;
      st.kind = Step::kChain;
      st.chain_convs = static_cast<int>(j - i);
      st.is_rdb = m->steps[i].is_rdb;
      if (plan_chain(cs.data(), static_cast<int>(cs.size()), &st.chain)) {
        for (Step& o : out)
          if (o.kind == Step::kChain) free_chain(&o.chain);
        return 1;
      }
      out.push_back(st);
      i = j;
    } else {
      out.push_back(m->steps[i]);
      ++i;
    }
  }
  m->steps.swap(out);
  return 0;
}

int build_plan(Rrdbnet* m, int n, int h, int w, uint8_t* wsp, int training, cudaStream_t stream) {
  clear_steps(m);
  const Workspace ws = layout(m, n, h, w);
  const int nf = m->nf, gc = m->gc;
  auto push_conv = [&](const esrp_conv3x3_t& d0, bool patch_y = false, bool is_noise = false, int noise_index = 0) -> int {
    Step st;
    st.kind = Step::kConv;
    esrp_conv3x3_t d = d0;
    d.variant |= ESRP_VARIANT_ROW_ALT;  // inference plans: row-alternating MMA issuers where the row kernel runs
    // CTA pairs for the dense-block convs of an even batch: opt-in (ESRP_PAIR=1), measured SLOWER on config 2 (14.45 vs
    // 13.10 ms); not when the convs are to be merged into a chain
    static const bool env_chain = [] { const char* e = getenv("ESRP_CHAIN"); return e && atoi(e) != 0; }();
    static const bool env_pair = [] { const char* e = getenv("ESRP_PAIR"); return e && atoi(e) != 0; }();
    if (env_pair && !m->use_chain && !env_chain) d.variant |= ESRP_VARIANT_PAIR;
    if (plan_conv(d, &st.conv)) return 1;
    st.patch_y = patch_y;
    st.is_noise = is_noise;
    st.noise_index = noise_index;
    m->steps.push_back(st);
    return 0;
  };
  // (re)pack a conv for the kernel decomposition its image width calls for
  auto want = [&](std::vector<ConvW>& cs, int width) -> int {
    const int lay = layout_for_width(width);
    for (auto& c : cs)
      if (c.layout != lay && pack_one(m, &c, lay, stream)) return 1;
    return 0;
  };
  // The fp32 twin of the trunk is only ever touched by conv epilogues of this plan: where they all run on the row
  // kernel (thread == pixel) it is kept as [n][h][c/4][w][4] so that its reads / writes coalesce (esrp_conv3x3_t::f32_planar).
  static const bool no_planar = getenv("ESRP_NO_PLANAR") != nullptr;  // timing experiments
  const int planar = (layout_for_width(w) == ESRP_LAYOUT_ROW && !no_planar) ? 1 : 0;
  // plain single-source conv (all slices): src -> [bf16 out][f32 out][nchw out], optional fp32 residual
  auto plain = [&](std::vector<ConvW>& cs, int hh, int ww, const void* src, int src_ct, int act, void* out_b,
                   void* out_f, const void* r1_f32, bool to_y) -> int {
    if (want(cs, ww)) return 1;
    for (auto& c : cs) {
      esrp_conv3x3_t d;
      base_desc(m, c, n, hh, ww, &d);
      d.src[0] = src; d.src_ctotal[0] = src_ct;
      for (int i = 0; i < c.num_chunks; ++i) { d.chunk_src[i] = 0; d.chunk_c0[i] = c.lc0[i]; }
      d.act = act;
      if (r1_f32) { d.r1 = r1_f32; d.r1_is_f32 = 1; d.r1_ctotal = c.cout; d.r1_c0 = c.row0; d.s1 = 1.f; }
      if (out_b) { d.out_bf16 = out_b; d.ob_ctotal = c.cout; d.ob_c0 = c.row0; }
      if (out_f) { d.out_f32 = out_f; d.of_ctotal = c.cout; d.of_c0 = c.row0; }
      if (r1_f32 || out_f) d.f32_planar = planar;
      if (to_y) d.out_nchw = reinterpret_cast<float*>(wsp);  // placeholder, patched per call
      if (push_conv(d, to_y)) return 1;
    }
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
  // 1. fea_conv (architecture.py:55): no activation; bf16 + fp32 outputs
  if (plain(m->fea, h, w, wsp + ws.xin, m->in_pad, 0, wsp + ws.fea_b, wsp + ws.fea_f, nullptr, false)) return 1;

  // 2. RRDB trunk
  // trunk buffer rotation: `cur` holds the current RDB input, `rr` the RRDB input (for block.py:291)
  if (want(m->rdb, w)) return 1;
  const uint8_t* cur_b = wsp + ws.fea_b;
  const uint8_t* cur_f = wsp + ws.fea_f;
  int noise_index = 0;
  esrp_conv3x3_t d;
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
      const bool wide5 = !m->rdb5w.empty() && layout_for_width(w) == ESRP_LAYOUT_TILE;
      // row kernel: the conv5 slices (32 output channels each) share one launch, CTA pairs walk the same rows
      const int nsl5 = (!wide5 && m->per_rdb > 5 && !no_coslice()) ? m->per_rdb - 4 : 1;
      for (int k = 0; k < ((wide5 || nsl5 > 1) ? 5 : m->per_rdb); ++k) {
        const ConvW& c = (wide5 && k == 4) ? m->rdb5w[static_cast<size_t>(i) * 3 + r]
                                           : m->rdb[(static_cast<size_t>(i) * 3 + r) * m->per_rdb + k];
        base_desc(m, c, n, h, w, &d);
        if (k == 4 && nsl5 > 1 && coslice(c, (&c)[1], nsl5, &d)) return 1;
        d.src[0] = cur_b; d.src_ctotal[0] = nf;
        d.src[1] = G; d.src_ctotal[1] = 4 * gc;
        for (int ch = 0; ch < c.num_chunks; ++ch) {
          const int lc = c.lc0[ch];
          d.chunk_src[ch] = lc < nf ? 0 : 1;
          d.chunk_c0[ch] = lc < nf ? lc : lc - nf;
        }
        if (k == 0) d.src[1] = nullptr;
        d.k_valid = nf + k * gc;  // conv2 / conv4: the tail of the last 64-channel chunk has zero weights
        if (k < 4) {
          d.act = 1;
          d.out_bf16 = G; d.ob_ctotal = 4 * gc; d.ob_c0 = k * gc;
          if (k == 1) d.aux_chunks = c.aux_chunks;  // x2 = lrelu(conv2) + conv1x1(x)   (block.py:262-263)
          if (k == 3) {  // x4 = lrelu(conv4) + x2            (block.py:265-266)
            d.r1 = G; d.r1_is_f32 = 0; d.r1_ctotal = 4 * gc; d.r1_c0 = gc; d.s1 = 1.f;
          }
          if (push_conv(d)) return 1;
          m->steps.back().is_rdb = true;
        } else {
          // out = noise(0.2 * x5 + x)                         (block.py:268), channels [row0, row0+32)
          const int c0 = c.row0;
          d.act = 0; d.s0 = 0.2f;
          d.r1 = cur_f; d.r1_is_f32 = 1; d.r1_ctotal = nf; d.r1_c0 = c0; d.s1 = 1.f;
          d.noise = training ? 1 : 0; d.noise_ctotal = nf; d.noise_c0 = c0;
          if (r == 2) {  // RRDB: out * 0.2 + x                 (block.py:291)
            d.r2 = rr_f; d.r2_is_f32 = 1; d.r2_ctotal = nf; d.r2_c0 = c0; d.s2 = 0.2f;
          }
          d.out_bf16 = out_b; d.ob_ctotal = nf; d.ob_c0 = c0;
          d.out_f32 = out_f; d.of_ctotal = nf; d.of_c0 = c0;
          d.f32_planar = planar;
          if (push_conv(d, false, training != 0, noise_index)) return 1;
          m->steps.back().is_rdb = true;
        }
      }
      ++noise_index;
      cur_b = out_b;
      cur_f = out_f;
    }
  }
  // 3. LR_conv + shortcut (architecture.py:58,73; block.py:84-86)
  if (plain(m->trunk, h, w, cur_b, nf, 0, wsp + ws.u0, nullptr, wsp + ws.fea_f, false)) return 1;

  // 4. upconv blocks: nearest x2 -> conv -> lrelu (block.py:315-322)
  const uint8_t* feat = wsp + ws.u0;
  int ch_ = h, cw_ = w;
  uint8_t* hr[3] = {wsp + ws.hr_a, wsp + ws.hr_b, wsp + ws.hr_c};
  int hr_i = 0;
  for (int u = 0; u < m->n_up; ++u) {
    Step st;
    st.kind = Step::kUpsample;
    st.src = feat;
    st.dst = hr[hr_i];
    st.n = n; st.h = ch_; st.w = cw_; st.c = nf;
    m->steps.push_back(st);
    ch_ *= 2; cw_ *= 2;
    if (plain(m->up[u], ch_, cw_, hr[hr_i], nf, 1, hr[(hr_i + 1) % 3], nullptr, nullptr, false)) return 1;
    feat = hr[(hr_i + 1) % 3];
    hr_i = (hr_i + 2) % 3;
  }
  // 5. HR_conv0 + lrelu (architecture.py:70)
  uint8_t* hr0_out = hr[hr_i];
  if (hr0_out == feat) hr0_out = hr[(hr_i + 1) % 3];
  if (plain(m->hr0, ch_, cw_, feat, nf, 1, hr0_out, nullptr, nullptr, false)) return 1;
  // 6. HR_conv1 (architecture.py:71): NCHW fp32 straight into the caller's output tensor
  if (plain(m->hr1, ch_, cw_, hr0_out, nf, 0, nullptr, nullptr, nullptr, true)) return 1;

  if (merge_chains(m)) return 1;
  m->pn = n; m->ph = h; m->pw = w; m->pws = wsp; m->ptraining = training;
  m->g_zeroed = false;
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

int esrp_rrdbnet_set_chain(esrp_rrdbnet_t* h, int32_t enable) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m) return set_error("rrdbnet_set_chain: null handle");
  if (m->use_chain != (enable != 0)) {
    m->use_chain = enable != 0;
    m->pn = 0;  // re-plan at the next forward
  }
  return 0;
}

int32_t esrp_rrdbnet_num_chained_convs(const esrp_rrdbnet_t* h) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  if (!m) return -1;
  int32_t n = 0;
  for (const Step& st : m->steps)
    if (st.kind == Step::kChain) n += st.chain_convs;
  return n;
}

int32_t esrp_rrdbnet_num_pair_launches(const esrp_rrdbnet_t* h) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  if (!m) return -1;
  int32_t n = 0;
  for (const Step& st : m->steps)
    if (st.kind == Step::kConv && st.conv.cluster == 2) ++n;
  return n;
}

int esrp_rrdbnet_forward(esrp_rrdbnet_t* h, const float* x, float* y, int32_t n, int32_t hh, int32_t w,
                         void* workspace, int64_t workspace_bytes, int32_t training, uint64_t seed, void* stream) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m || (!x && !m->x_u8) || !y || !workspace) return set_error("rrdbnet_forward: null argument");
  if (!m->weights_loaded) return set_error("rrdbnet_forward: weights not loaded");
  if (n < 1 || hh < 1 || w < 1) return set_error("rrdbnet_forward: bad shape");
  const int64_t need = esrp_rrdbnet_workspace_bytes(h, n, hh, w);
  if (workspace_bytes < need) return set_error("rrdbnet_forward: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
  if (reinterpret_cast<uintptr_t>(workspace) % 1024) return set_error("rrdbnet_forward: workspace must be 1024-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // another plan (e.g. the training plan at a different crop size) may have re-laid the weights out
  if (ensure_layouts(m, w, s)) return 1;
  if (m->pn != n || m->ph != hh || m->pw != w || m->pws != workspace || m->ptraining != (training ? 1 : 0)) {
    if (build_plan(m, n, hh, w, static_cast<uint8_t*>(workspace), training ? 1 : 0, s)) {
      m->pn = 0;
      return 1;
    }
  }
  if (!m->g_zeroed) {
    // growth buffer: chunks may cover not-yet-written channels (zero weights) -> keep them finite
    const Workspace ws = layout(m, n, hh, w);
    ESRP_CUDA_OK(cudaMemsetAsync(static_cast<uint8_t*>(workspace) + ws.g, 0,
                                 static_cast<size_t>(n) * hh * w * 4 * m->gc * 2, s));
    m->g_zeroed = true;
  }
  static const bool debug_sync = getenv("ESRP_DEBUG_SYNC") != nullptr;  // locate a failing launch
  int step_idx = 0;
  int first_rdb = -1, last_rdb = -1;
  if (m->timing) {
    for (int i = 0; i < static_cast<int>(m->steps.size()); ++i)
      if (m->steps[i].is_rdb) { if (first_rdb < 0) first_rdb = i; last_rdb = i; }
  }
  const size_t ev_slot = static_cast<size_t>(m->timed_forwards % 64);
  for (Step& st : m->steps) {
    if (step_idx == first_rdb) ESRP_CUDA_OK(cudaEventRecord(m->ev0[ev_slot], s));
    if (debug_sync) {
      cudaError_t e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) {
        const Step& pr = m->steps[step_idx > 0 ? step_idx - 1 : 0];
        return set_error("step %d (kind %d, n=%d h=%d w=%d chunks=%d aux=%d nt=%d slots=%d bufs=%d grid=%d threads=%d) failed: %s",
                         step_idx - 1, (int)pr.kind, pr.conv.params.n, pr.conv.params.h, pr.conv.params.w,
                         pr.conv.params.num_chunks, pr.conv.params.aux_chunks, pr.conv.params.nt, pr.conv.params.mt,
                         pr.conv.params.stages, pr.conv.grid, pr.conv.threads, cudaGetErrorString(e));
      }
    }
    ++step_idx;
    switch (st.kind) {
      case Step::kPackInput:
        if (m->x_u8) {
          if (esrp_u8hwc_to_nhwc_bf16(m->x_u8, st.dst, st.n, st.h, st.w, st.c, st.c_pad, m->x_bgr, stream)) return 1;
        } else if (esrp_nchw_f32_to_nhwc_bf16(x, st.dst, st.n, st.c, st.h, st.w, st.c_pad, stream)) {
          return 1;
        }
        break;
      case Step::kUpsample:
        if (esrp_upsample2x_nhwc_bf16(st.src, st.dst, st.n, st.h, st.w, st.c, stream)) return 1;
        break;
      case Step::kChain:
        if (run_chain(st.chain, s)) return 1;
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
    if (step_idx - 1 == last_rdb && last_rdb >= 0) {
      ESRP_CUDA_OK(cudaEventRecord(m->ev1[ev_slot], s));
      ++m->timed_forwards;
    }
  }
  return 0;
}

int esrp_rrdbnet_set_timing(esrp_rrdbnet_t* h, int32_t enable) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m) return set_error("rrdbnet_set_timing: null handle");
  if (enable && m->ev0.empty()) {
    m->ev0.resize(64);
    m->ev1.resize(64);
    for (int i = 0; i < 64; ++i) {
      ESRP_CUDA_OK(cudaEventCreate(&m->ev0[i]));
      ESRP_CUDA_OK(cudaEventCreate(&m->ev1[i]));
    }
  }
  m->timing = enable != 0;
  m->timed_forwards = 0;
  return 0;
}

int32_t esrp_rrdbnet_get_timing(esrp_rrdbnet_t* h, float* ms, int32_t max) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m || !ms || max < 1) return -1;
  const long long n = m->timed_forwards < 64 ? m->timed_forwards : 64;
  const long long take = n < max ? n : max;
  for (long long i = 0; i < take; ++i) {
    const size_t slot = static_cast<size_t>((m->timed_forwards - take + i) % 64);
    if (cudaEventElapsedTime(&ms[i], m->ev0[slot], m->ev1[slot]) != cudaSuccess) {
      set_error("rrdbnet_get_timing: events not complete (synchronise the stream first)");
      return -1;
    }
  }
  return static_cast<int32_t>(take);
}

int64_t esrp_rrdbnet_workspace_bytes_u8(const esrp_rrdbnet_t* h, int32_t n, int32_t hh, int32_t w) {
  const Rrdbnet* m = reinterpret_cast<const Rrdbnet*>(h);
  const int64_t base = esrp_rrdbnet_workspace_bytes(h, n, hh, w);
  if (base < 0) return -1;
  return base + static_cast<int64_t>(n) * m->out_nc * hh * m->upscale * w * m->upscale * 4;
}

int esrp_rrdbnet_forward_u8(esrp_rrdbnet_t* h, const uint8_t* x, uint8_t* y, int32_t n, int32_t hh, int32_t w, void* workspace,
                            int64_t workspace_bytes, int32_t bgr, void* stream) {
  Rrdbnet* m = reinterpret_cast<Rrdbnet*>(h);
  if (!m || !x || !y || !workspace) return set_error("rrdbnet_forward_u8: null argument");
  const int64_t base = esrp_rrdbnet_workspace_bytes(h, n, hh, w);
  const int64_t need = esrp_rrdbnet_workspace_bytes_u8(h, n, hh, w);
  if (base < 0 || workspace_bytes < need)
    return set_error("rrdbnet_forward_u8: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need);
  float* y32 = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + base);
  m->x_u8 = x;
  m->x_bgr = bgr;
  const int rc = esrp_rrdbnet_forward(h, nullptr, y32, n, hh, w, workspace, base, 0, 0, stream);
  m->x_u8 = nullptr;
  if (rc) return rc;
  return esrp_nchw_f32_to_u8hwc(y32, y, n, m->out_nc, hh * m->upscale, w * m->upscale, bgr, stream);
}

}  // extern "C"
