// esrp_api.cu — C ABI (include/esrp.h) over the sm_100a kernels: launch logic, TMA tensor-map
// encoding, weight repacking and the NCHW<->NHWC boundary converters.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/esrp.h"
#include "esrp_philox.cuh"
#include "esrp_host.h"

namespace esrp {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// NHWC bf16 tensor [n,h,w,c_total] viewed as a 4-D TMA tensor (c, x, y, n); box = (kc, bw, bh, 1).
int make_nhwc_tmap(CUtensorMap* tm, const void* ptr, int n, int h, int w, int c_total, int kc,
                   int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {(cuuint64_t)c_total, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c_total * 2, (cuuint64_t)w * c_total * 2,
                           (cuuint64_t)h * w * c_total * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = (kc == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  // L2 promotion: a miss fetches 256 contiguous bytes only where the box covers the whole pixel (its rows are then
  // contiguous in memory).  A 64-channel chunk of the 128-channel growth buffer is 128 of every 256 bytes: with 256-byte
  // promotion conv2 / conv3 pulled the unused half of G from DRAM as well (ncu: 104-107 MB read for 74 MB of operands).
  static const bool promo256 = getenv("ESRP_TMAP_PROMO256") != nullptr;  // timing experiments
  const CUtensorMapL2promotion promo = (kc == c_total || promo256) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (kc * 2 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p nhwc=(%d,%d,%d,%d) kc=%d box=(%d,%d)",
                     (int)r, ptr, n, h, w, c_total, kc, box_w, box_h);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// conv launch
// ------------------------------------------------------------------------------------------------

// Per-device state is keyed by the device ordinal: one process may drive several GPUs (nn.DataParallel, gpu_ids=[0,1]).
static std::mutex g_dev_mu;
int sm_count() {
  static int counts[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (counts[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    counts[dev] = n;
  }
  return counts[dev];
}

int ensure_max_smem(const void* kernel, int bytes) {
  static std::unordered_map<std::string, int>* done = new std::unordered_map<std::string, int>();
  int dev = 0;
  ESRP_CUDA_OK(cudaGetDevice(&dev));
  char key[64];
  snprintf(key, sizeof(key), "%d:%p", dev, kernel);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  auto it = done->find(key);
  if (it != done->end() && it->second >= bytes) return 0;
  ESRP_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  (*done)[key] = bytes;
  return 0;
}

// kernel instantiations live in esrp_conv_{row,tile}{,_ext}.cu (one translation unit per family so they build in parallel)
int run_conv(const ConvLaunch& L, cudaStream_t stream) {
  if (L.grid < 1) return 0;
  void* args[3] = {const_cast<CUtensorMap*>(&L.tm0), const_cast<CUtensorMap*>(&L.tm1),
                   const_cast<ConvKParams*>(&L.params)};
  // Programmatic dependent launch: the kernel's prologue (barrier init, TMEM allocation + zeroing, weight
  // fetch) may overlap the tail of the previous kernel in the stream; every kernel executes
  // griddepcontrol.wait before it touches activations (esrp_ptx.cuh: grid_dep_wait).
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(L.threads);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool no_pdl = getenv("ESRP_NO_PDL") != nullptr;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (L.cluster > 1) {  // CTA pairs of the row kernel (cta_group::2)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = L.cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  ESRP_CUDA_OK(cudaLaunchKernelExC(&cfg, L.kernel, args));
  return 0;
}

int plan_conv(const esrp_conv3x3_t& d, ConvLaunch* out) {
  if (d.n < 1 || d.h < 1 || d.w < 1) return set_error("conv3x3: bad shape n=%d h=%d w=%d", d.n, d.h, d.w);
  if (d.num_chunks < 1 || d.num_chunks > ESRP_MAX_CHUNKS) return set_error("conv3x3: num_chunks=%d out of range", d.num_chunks);
  if (d.aux_chunks < 0 || d.aux_chunks > d.num_chunks) return set_error("conv3x3: aux_chunks=%d out of range", d.aux_chunks);
  if (!d.src[0] || !d.w_packed) return set_error("conv3x3: null src/weights");
  if (d.cout < 1 || d.cout > d.bn) return set_error("conv3x3: cout=%d vs bn=%d", d.cout, d.bn);
  if (d.act < 0 || d.act > 2) return set_error("conv3x3: act=%d (0 none, 1 LeakyReLU(0.2), 2 ReLU)", d.act);
  if (d.out_lo && (d.w_layout != ESRP_LAYOUT_TILE || !d.out_bf16 || d.ob_lo_c0 < 0 || (d.ob_lo_c0 % 8) || d.ob_lo_c0 + d.cout > d.ob_ctotal || d.slices > 1))
    return set_error("conv3x3: out_lo needs ESRP_LAYOUT_TILE weights, an out_bf16 tensor and an 8-aligned ob_lo_c0 inside it");
  for (int i = 0; i < d.num_chunks; ++i) {
    int s_ = d.chunk_src[i];
    if (s_ < 0 || s_ > 1 || !d.src[s_]) return set_error("conv3x3: chunk %d reads missing src %d", i, s_);
    if (d.chunk_c0[i] < 0 || d.chunk_c0[i] + d.kc > d.src_ctotal[s_] || (d.chunk_c0[i] % 8))
      return set_error("conv3x3: chunk %d channel range [%d,%d) invalid for src with %d channels", i,
                       d.chunk_c0[i], d.chunk_c0[i] + d.kc, d.src_ctotal[s_]);
  }
  for (int i = 0; i < 2; ++i)
    if (d.src[i] && (d.src_ctotal[i] % 8)) return set_error("conv3x3: src_ctotal must be a multiple of 8");
  if (d.out_bf16 && ((d.ob_ctotal % 8) || (d.ob_c0 % 8))) return set_error("conv3x3: out_bf16 channel alignment");
  if (d.out_f32 && ((d.of_ctotal % 4) || (d.of_c0 % 4))) return set_error("conv3x3: out_f32 channel alignment");
  if (d.r1 && ((d.r1_ctotal % 8) || (d.r1_c0 % 8))) return set_error("conv3x3: r1 channel alignment");
  if (d.r2 && ((d.r2_ctotal % 8) || (d.r2_c0 % 8))) return set_error("conv3x3: r2 channel alignment");
  if ((d.out_bf16 || d.out_f32 || d.r1 || d.r2) && (d.cout % 16)) return set_error("conv3x3: NHWC outputs/residuals need cout %% 16 == 0");
  if (d.noise && ((d.noise_ctotal % 4) || (d.noise_c0 % 4) || d.noise_ctotal < d.cout)) return set_error("conv3x3: noise_ctotal/noise_c0 must be multiples of 4 and cover cout");
  const bool ext = d.mask_out || d.mask_in || d.pre_bf16 || d.pre_f32 || d.r2_pre;
  if (d.mask_out && ((d.mask_out_ctotal % 16) || (d.mask_out_c0 % 16) || (d.cout % 16))) return set_error("conv3x3: mask_out needs 16-bit aligned channel ranges");
  if (d.mask_in && ((d.mask_in_ctotal % 16) || (d.mask_in_c0 % 16) || (d.cout % 16))) return set_error("conv3x3: mask_in needs 16-bit aligned channel ranges");
  if (d.pre_bf16 && ((d.pb_ctotal % 8) || (d.pb_c0 % 8) || (d.cout % 16))) return set_error("conv3x3: pre_bf16 channel alignment");
  if (d.pre_f32 && ((d.pf_ctotal % 4) || (d.pf_c0 % 4) || (d.cout % 16))) return set_error("conv3x3: pre_f32 channel alignment");
  if (d.slices > 1) {
    if (d.w_layout != ESRP_LAYOUT_ROW) return set_error("conv3x3: slices > 1 needs ESRP_LAYOUT_ROW weights");
    if (d.slices > 4 || d.cout != d.bn) return set_error("conv3x3: slices=%d needs cout == bn and at most 4 slices", d.slices);
    if (d.out_nchw) return set_error("conv3x3: slices > 1 cannot write out_nchw");
    if (d.slice_stride <= 0 || (d.slice_stride % 16)) return set_error("conv3x3: slice_stride must be a positive multiple of 16 bytes");
  }
  if (d.k_valid < 0 || d.k_valid > d.num_chunks * d.kc || (d.k_valid > 0 && d.k_valid <= (d.num_chunks - 1) * d.kc))
    return set_error("conv3x3: k_valid=%d must lie in the last of the %d chunks of %d channels", d.k_valid, d.num_chunks, d.kc);
  if (d.f32_planar) {
    if (d.w_layout != ESRP_LAYOUT_ROW || ext) return set_error("conv3x3: f32_planar needs ESRP_LAYOUT_ROW weights and no training extensions");
    if ((d.r1 && d.r1_is_f32 && ((d.r1_ctotal | d.r1_c0) % 4)) || (d.r2 && d.r2_is_f32 && ((d.r2_ctotal | d.r2_c0) % 4)))
      return set_error("conv3x3: f32_planar channel alignment");
  }
  if (d.w_layout == ESRP_LAYOUT_ROW) return ext ? plan_row_ext(d, out) : plan_row_base(d, out);
  if (d.w_layout != ESRP_LAYOUT_TILE) return set_error("conv3x3: unknown w_layout=%d", d.w_layout);
  return ext ? plan_tile_ext(d, out) : plan_tile_base(d, out);
}


// ------------------------------------------------------------------------------------------------
// weight repack: [w_o, w_i, 3, 3] fp32 -> [chunk][ky][row][kc] bf16, rows pre-swizzled for UMMA/TMA
// ------------------------------------------------------------------------------------------------
struct ChunkTable {
  int lc0[ESRP_MAX_CHUNKS];
};

__device__ __forceinline__ int swizzled_elem(int row, int k, int kc) {
  // 16-byte chunk index XOR row bits, matching CU_TENSOR_MAP_SWIZZLE_{128B,64B} / UMMA layouts.
  const int chunk16 = k >> 3;
  const int x = (kc == 64) ? (row & 7) : ((row >> 1) & 3);
  return row * kc + (((chunk16 ^ x) << 3) | (k & 7));
}

__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int w_o, int w_i, int transpose, int layout, int row0,
                                         int rows, int kc, int bn, int num_chunks, ChunkTable tab,
                                         const float* __restrict__ w_aux, int aux_cin, int aux_chunks,
                                         __nv_bfloat16* __restrict__ out) {
  const int nb_rows = (aux_chunks > 0 ? 4 : 3) * bn;
  const int total = num_chunks * 3 * nb_rows * kc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % kc;
    const int row = (i / kc) % nb_rows;
    const int outer = (i / (kc * nb_rows)) % 3;  // MMA index within a K-slice: ky (TILE) / kx (ROW)
    const int chunk = i / (kc * nb_rows * 3);
    const int blk = row / bn;        // 0..2: tap stacked along N: kx (TILE) / ky (ROW); 3: conv1x1 rows
    const int r = row - blk * bn;    // logical output channel within the packed slice
    const int ci = tab.lc0[chunk] + k;
    const int ky = layout == ESRP_LAYOUT_ROW ? blk : outer;
    float v = 0.f;
    if (r < rows) {
      if (blk < 3) {
        const int kx = layout == ESRP_LAYOUT_ROW ? outer : blk;
        if (!transpose) {
          if (row0 + r < w_o && ci < w_i) v = w[((static_cast<size_t>(row0 + r) * w_i + ci) * 3 + ky) * 3 + kx];
        } else {
          if (ci < w_o && row0 + r < w_i) v = w[((static_cast<size_t>(ci) * w_i + row0 + r) * 3 + (2 - ky)) * 3 + (2 - kx)];
        }
      } else if (outer == 1 && chunk < aux_chunks && w_aux != nullptr) {
        if (row0 + r < w_o && ci < aux_cin) v = w_aux[static_cast<size_t>(row0 + r) * aux_cin + ci];
      }
    }
    out[static_cast<size_t>(chunk * 3 + outer) * nb_rows * kc + swizzled_elem(row, k, kc)] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// boundary converters (HBM-bound, coalesced on the NHWC side, 32x32 smem transpose tiles)
// ------------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 with zero channel padding.  One block handles 32 pixels x all channels.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                    int c, int hw, int c_pad) {
  extern __shared__ float tile[];  // [c_pad][33]
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const float* s = src + static_cast<size_t>(img) * c * hw;
  for (int i = threadIdx.x; i < c_pad * 32; i += blockDim.x) {
    const int ch = i / 32, px = i % 32;
    float v = 0.f;
    if (ch < c && p0 + px < hw) v = s[static_cast<size_t>(ch) * hw + p0 + px];
    tile[ch * 33 + px] = v;
  }
  __syncthreads();
  __nv_bfloat16* d = dst + (static_cast<size_t>(img) * hw + p0) * c_pad;
  for (int i = threadIdx.x; i < c_pad * 32; i += blockDim.x) {
    const int px = i / c_pad, ch = i % c_pad;
    if (p0 + px < hw) d[static_cast<size_t>(px) * c_pad + ch] = __float2bfloat16_rn(tile[ch * 33 + px]);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int c,
                                    int hw, int c_total) {
  extern __shared__ float tile[];  // [c][33]
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const __nv_bfloat16* s = src + (static_cast<size_t>(img) * hw + p0) * c_total;
  for (int i = threadIdx.x; i < c * 32; i += blockDim.x) {
    const int px = i / c, ch = i % c;
    float v = 0.f;
    if (p0 + px < hw) v = __bfloat162float(s[static_cast<size_t>(px) * c_total + ch]);
    tile[ch * 33 + px] = v;
  }
  __syncthreads();
  float* d = dst + static_cast<size_t>(img) * c * hw;
  for (int i = threadIdx.x; i < c * 32; i += blockDim.x) {
    const int ch = i / 32, px = i % 32;
    if (p0 + px < hw) d[static_cast<size_t>(ch) * hw + p0 + px] = tile[ch * 33 + px];
  }
}

// Image plumbing of test_image/test.py on the device.
// :31-35  cv2 image (uint8, HWC, BGR) -> /255 -> RGB -> float CHW : here straight to the NHWC bf16 network input
__global__ void u8hwc_to_nhwc_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long npx, int c,
                                     int c_pad, int swap) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx * c_pad;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c_pad);
    const long long px = i / c_pad;
    float v = 0.f;
    if (ch < c) v = static_cast<float>(src[px * c + ((swap && c == 3) ? 2 - ch : ch)]) / 255.0f;
    dst[i] = __float2bfloat16_rn(v);
  }
}
// :37-39  output.clamp_(0, 1) -> RGB->BGR, CHW->HWC -> (output * 255.0).round() : NCHW fp32 -> uint8 HWC
__global__ void nchw_to_u8hwc_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int c, long long hw, int swap) {
  const long long img = blockIdx.y;
  const float* s = src + img * c * hw;
  uint8_t* d = dst + img * c * hw;
  for (long long px = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; px < hw;
       px += static_cast<long long>(gridDim.x) * blockDim.x) {
    for (int ch = 0; ch < c; ++ch) {
      float v = s[static_cast<long long>((swap && c == 3) ? 2 - ch : ch) * hw + px];
      v = fminf(fmaxf(v, 0.f), 1.f);            // clamp_ (NaN -> 0 like fmaxf; the reference would propagate it)
      d[px * c + ch] = static_cast<uint8_t>(rintf(v * 255.0f));   // numpy round: half to even
    }
  }
}

// nearest x2 upsample, NHWC bf16, 16-byte vectors: each thread copies one 8-channel vector to 4 outputs.
__global__ void upsample2x_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n, int h,
                                  int w, int cv) {
  const size_t total = static_cast<size_t>(n) * h * w * cv;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t t = i / cv;
    const int x = static_cast<int>(t % w); t /= w;
    const int y = static_cast<int>(t % h);
    const int img = static_cast<int>(t / h);
    const uint4 val = __ldg(src + i);
    const size_t ow = static_cast<size_t>(2) * w;
    const size_t base = ((static_cast<size_t>(img) * 2 * h + 2 * y) * ow + 2 * x) * cv + v;
    dst[base] = val;
    dst[base + cv] = val;
    dst[base + ow * cv] = val;
    dst[base + ow * cv + cv] = val;
  }
}

}  // namespace esrp

// ================================================================================================
// extern "C"
// ================================================================================================
using namespace esrp;

extern "C" {

const char* esrp_last_error(void) { return g_err; }
int esrp_version(void) { return 100; }
int esrp_sm_count(void) { return sm_count(); }
int32_t esrp_sizeof_conv3x3(void) { return static_cast<int32_t>(sizeof(esrp_conv3x3_t)); }

int esrp_philox_normal_host(uint64_t seed, uint64_t offset, int64_t count, float* out_host) {
  if (!out_host || count < 0) return set_error("philox_normal_host: bad arguments");
  for (int64_t e = 0; e < count; e += 4) {
    float z[4];
    philox_normal4(seed, offset + static_cast<unsigned long long>(e / 4), z);
    for (int j = 0; j < 4 && e + j < count; ++j) out_host[e + j] = z[j];
  }
  return 0;
}

// Planning a launch (argument validation, two cuTensorMapEncodeTiled calls, shared-memory / TMEM budgeting) costs more
// host time than the launch itself.  A caller that issues the same descriptor again — a training loop whose allocator
// hands out the same addresses every step — hits this small cache keyed on the descriptor bytes (pointers included).
int esrp_conv3x3_nhwc(const esrp_conv3x3_t* desc, void* stream) {
  if (!desc) return set_error("esrp_conv3x3_nhwc: null desc");
  static std::mutex mu;
  static std::unordered_map<std::string, ConvLaunch>* cache = new std::unordered_map<std::string, ConvLaunch>();
  esrp_conv3x3_t d = *desc;
  // Philox key / offset and the debug trace pointer do not influence planning: keep them out of the key
  const unsigned long long seed = d.seed, offset = d.offset;
  d.seed = d.offset = 0;
  if (d.trace != nullptr || d.variant != 0) {
    ConvLaunch L;
    if (plan_conv(*desc, &L)) return 1;
    return run_conv(L, static_cast<cudaStream_t>(stream));
  }
  std::string key(reinterpret_cast<const char*>(&d), sizeof(d));
  ConvLaunch L;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache->find(key);
    if (it == cache->end()) {
      if (plan_conv(d, &L)) return 1;
      if (cache->size() >= 8192) cache->clear();
      cache->emplace(std::move(key), L);
    } else {
      L = it->second;
    }
  }
  L.params.seed = seed;
  L.params.offset = offset;
  return run_conv(L, static_cast<cudaStream_t>(stream));
}

int64_t esrp_packed_conv3x3_bytes(int32_t num_chunks, int32_t kc, int32_t bn, int32_t has_aux) {
  return static_cast<int64_t>(num_chunks) * 3 * (has_aux ? 4 : 3) * bn * kc * 2;
}

int esrp_pack_conv3x3_weights(const float* w_oihw, int32_t w_o, int32_t w_i, int32_t transpose, int32_t layout,
                              int32_t row0, int32_t rows, int32_t kc, int32_t bn, int32_t num_chunks,
                              const int32_t* chunk_lc0_host, const float* w_aux_oi, int32_t aux_cin,
                              int32_t aux_chunks, void* out, void* stream) {
  if (!w_oihw || !out || !chunk_lc0_host) return set_error("pack_weights: null pointer");
  if (kc != 32 && kc != 64) return set_error("pack_weights: kc must be 32 or 64");
  if (layout != ESRP_LAYOUT_TILE && layout != ESRP_LAYOUT_ROW) return set_error("pack_weights: unknown layout %d", layout);
  if (bn != 16 && bn != 32 && bn != 64) return set_error("pack_weights: bn must be 16, 32 or 64");
  if (rows < 1 || rows > bn || row0 < 0) return set_error("pack_weights: bad row slice row0=%d rows=%d bn=%d", row0, rows, bn);
  if (num_chunks < 1 || num_chunks > ESRP_MAX_CHUNKS) return set_error("pack_weights: num_chunks=%d", num_chunks);
  if (aux_chunks < 0 || aux_chunks > num_chunks || (aux_chunks > 0 && (!w_aux_oi || transpose)))
    return set_error("pack_weights: bad conv1x1 arguments");
  ChunkTable tab;
  for (int i = 0; i < ESRP_MAX_CHUNKS; ++i) tab.lc0[i] = i < num_chunks ? chunk_lc0_host[i] : 0;
  const int total = num_chunks * 3 * (aux_chunks > 0 ? 4 : 3) * bn * kc;
  const int threads = 256;
  int blocks = (total + threads - 1) / threads;
  if (blocks > 1024) blocks = 1024;
  pack_conv_weights_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, w_o, w_i, transpose, layout, row0, rows, kc, bn, num_chunks, tab, w_aux_oi, aux_cin, aux_chunks,
      static_cast<__nv_bfloat16*>(out));
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t c_pad, void* stream) {
  if (!src || !dst || c_pad < c || c_pad > 256) return set_error("nchw_to_nhwc: bad arguments");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, n);
  nchw_to_nhwc_kernel<<<grid, 256, c_pad * 33 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), c, hw, c_pad);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t c_total, void* stream) {
  if (!src || !dst || c_total < c || c > 256) return set_error("nhwc_to_nchw: bad arguments");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, n);
  nhwc_to_nchw_kernel<<<grid, 256, c * 33 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, c, hw, c_total);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_u8hwc_to_nhwc_bf16(const uint8_t* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c, int32_t c_pad,
                            int32_t bgr, void* stream) {
  if (!src || !dst || c < 1 || c_pad < c) return set_error("u8hwc_to_nhwc: bad arguments");
  const long long npx = static_cast<long long>(n) * h * w;
  long long blocks = (npx * c_pad + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  u8hwc_to_nhwc_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), npx, c, c_pad, bgr);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_nchw_f32_to_u8hwc(const float* src, uint8_t* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t bgr,
                           void* stream) {
  if (!src || !dst || c < 1 || n < 1) return set_error("nchw_to_u8hwc: bad arguments");
  const long long hw = static_cast<long long>(h) * w;
  long long bx = (hw + 255) / 256;
  if (bx > 1024) bx = 1024;
  nchw_to_u8hwc_kernel<<<dim3(static_cast<unsigned>(bx), n), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, c, hw, bgr);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

int esrp_upsample2x_nhwc_bf16(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c,
                              void* stream) {
  if (!src || !dst || (c % 8)) return set_error("upsample2x: c must be a multiple of 8");
  const int cv = c / 8;
  const size_t total = static_cast<size_t>(n) * h * w * cv;
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 0;
  upsample2x_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), n, h, w, cv);
  ESRP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
