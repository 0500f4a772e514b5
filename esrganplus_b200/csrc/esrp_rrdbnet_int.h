// esrp_rrdbnet_int.h — internal state of the RRDBNet engine, shared by the inference runtime
// (esrp_rrdbnet.cu) and the training runtime (esrp_rrdbnet_train.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/esrp.h"
#include "esrp_host.h"
#include "esrp_pack.h"

namespace esrp {

// One launch-sized slice of a logical conv [cout, cin, 3, 3] and where its packed weights live.
struct ConvW {
  int cin = 0, cout = 0;       // logical channels of the reference tensor [cout, cin, 3, 3]
  int row0 = 0, rows = 0;      // output-channel slice this launch computes
  int kc = 0, bn = 0;
  int num_chunks = 0;
  int lc0[ESRP_MAX_CHUNKS] = {0};  // logical first input channel per chunk
  int aux_chunks = 0;              // leading chunks feeding the fused 1x1
  int w_idx = -1, b_idx = -1, aux_idx = -1;  // indices into the key list
  size_t w_off = 0, b_off = 0;     // packed device storage (offsets into wbuf)
  int layout = -1;                 // ESRP_LAYOUT_* currently packed (-1: none)
};

struct Step {
  enum Kind { kConv, kPackInput, kUpsample, kChain } kind = kConv;
  ConvLaunch conv;
  ChainLaunch chain;         // kChain: a run of row-kernel convs merged into one persistent launch (conv3x3_chain.cuh)
  int chain_convs = 0;       //   how many conv launches it replaced
  const void* src = nullptr;
  void* dst = nullptr;
  int n = 0, h = 0, w = 0, c = 0, c_pad = 0;
  bool patch_y = false;      // conv writes the caller's output tensor
  bool is_noise = false;     // conv5 with GaussianNoise (training)
  bool is_rdb = false;       // a dense-block conv of the trunk (what esrp_rrdbnet_set_timing brackets)
  int noise_index = 0;
};

struct TrainState;  // esrp_rrdbnet_train.cu

struct Rrdbnet {
  int in_nc, out_nc, nf, nb, gc, upscale, n_up;
  int in_pad;  // input channels padded to 32
  std::vector<std::string> keys;
  std::vector<std::vector<int>> shapes;
  // every logical conv is a list of <= 32-output-channel launches
  std::vector<ConvW> fea, trunk, up[2], hr0, hr1;
  std::vector<ConvW> rdb;  // nb*3*per_rdb: conv1..4, then conv5 as nf/32 output-channel slices
  std::vector<ConvW> rdb5w;  // nf == 64: conv5 as ONE 64-wide slice, tile layout only (the tile kernel streams its weights,
                             // so narrow images / training crops run conv5 in one launch instead of two)
  int per_rdb = 5;
  uint8_t* wbuf = nullptr;
  size_t wbytes = 0;
  bool weights_loaded = false;
  uint64_t weights_version = 0;       // bumped by every load_weights (derived caches compare against it)
  std::vector<const void*> src_ptrs;  // borrowed fp32 tensors of the last load_weights (for re-layout)
  // plan cache (single entry: the common case is a fixed shape)
  int pn = 0, ph = 0, pw = 0, ptraining = -1;
  void* pws = nullptr;
  bool g_zeroed = false;
  bool use_chain = false;             // merge runs of row-kernel convs into persistent chain launches (esrp_rrdbnet_set_chain;
                                      // off by default: measured slower than one launch per conv)
  std::vector<Step> steps;
  // esrp_rrdbnet_set_timing: a ring of event pairs around the dense-block convs of the trunk, one pair per forward
  bool timing = false;
  std::vector<cudaEvent_t> ev0, ev1;
  long long timed_forwards = 0;
  TrainState* train = nullptr;        // training plan + dgrad weight cache (lazily created)
  const uint8_t* x_u8 = nullptr;     // set for the duration of esrp_rrdbnet_forward_u8: the input is an 8-bit HWC image
  int x_bgr = 0;
  PackJob* pack_jobs_dev = nullptr;   // job table of the batched forward-weight repack (one entry per ConvW)
  std::vector<PackJob> pack_jobs_host;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int layout_for_width(int w) { return w > 64 ? ESRP_LAYOUT_ROW : ESRP_LAYOUT_TILE; }

// (re)pack one conv slice from the borrowed fp32 tensors into `layout`
int pack_one(Rrdbnet* m, ConvW* c, int layout, cudaStream_t s);
// make every conv's packed layout the one its image width calls for (LR width w)
int ensure_layouts(Rrdbnet* m, int w, cudaStream_t s);
// fill the common part of a conv descriptor from packed weights
void base_desc(const Rrdbnet* m, const ConvW& c, int n, int h, int w, esrp_conv3x3_t* d);
void destroy_train_state(Rrdbnet* m);
// co-scheduled output slices of a wide conv (esrp_conv3x3_t::slices), esrp_rrdbnet.cu
bool no_coslice();
int coslice(const ConvW& c, const ConvW& next, int nsl, esrp_conv3x3_t* d);

template <class F>
int for_each_conv(Rrdbnet* m, F&& f) {
  for (auto& c : m->fea) if (f(&c)) return 1;
  for (auto& c : m->rdb) if (f(&c)) return 1;
  for (auto& c : m->rdb5w) if (f(&c)) return 1;
  for (auto& c : m->trunk) if (f(&c)) return 1;
  for (int u = 0; u < m->n_up; ++u)
    for (auto& c : m->up[u]) if (f(&c)) return 1;
  for (auto& c : m->hr0) if (f(&c)) return 1;
  for (auto& c : m->hr1) if (f(&c)) return 1;
  return 0;
}

}  // namespace esrp
