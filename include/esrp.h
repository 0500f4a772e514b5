/* esrp.h — C ABI of libesrp.so, the B200-native (sm_100a) hot path of ESRGAN+/nESRGAN+.
 *
 * The reference (ncarraz/ESRGANplus) has no FFI: its hot path is the Python nn.Module surface
 *   codes/models/modules/architecture.py:47-78   (RRDBNet)
 *   codes/models/modules/architecture.py:87-129  (Discriminator_VGG_128)
 *   codes/models/modules/block.py:232-291        (ResidualDenseBlock_5C, RRDB, GaussianNoise)
 *   test_image/architecture.py:7-38              (RRDB_Net)
 * whose forward() dispatches to torch.nn.Conv2d / LeakyReLU / cat / add.  This library is what a
 * maintainer binds *beneath* those classes (ctypes stub in INTEGRATION.md): plain pointers and
 * sizes, no torch types.  All pointers are DEVICE pointers unless a name says host; all tensors
 * are borrowed (caller owns memory); every call is asynchronous on `stream` (a cudaStream_t
 * passed as void*).  Every function returns 0 on success, non-zero on error;
 * esrp_last_error() returns a thread-local message.
 *
 * Data layout in HBM: activations NHWC bf16 (optionally with an fp32 NHWC twin for the residual
 * trunk), weights repacked from the reference's OIHW fp32 wire format into K-major, pre-swizzled
 * bf16 tiles (derived cache; never stored in a state_dict).
 */
#ifndef ESRP_H_
#define ESRP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESRP_MAX_CHUNKS 32

/* Tiling overrides for esrp_conv3x3_t.variant (0 = library default): bits 0-3 force the number
 * of M-tile accumulator slots per CTA tile (1..5), bits 4-7 force log2 of the M-tile width in
 * pixels (4..7).  Used by tests to exercise every tiling on small inputs. */
#define ESRP_VARIANT_MT(mt) ((mt) & 15)
#define ESRP_VARIANT_CWLOG2(l) (((l) & 15) << 4)
/* Row kernel (ESRP_LAYOUT_ROW): the two MMA-issuer warps alternate whole input rows instead of splitting the taps
 * of every row (7 % faster on the benchmark forward; also lets k_valid skip zero-weight K-slices).  Opt-in per
 * launch: the engine sets it for its inference plans, where it has been stress-tested.
 * The environment variable ESRP_ROW_ALT=0 / 2 overrides every launch (off / on). */
#define ESRP_VARIANT_ROW_ALT 0x2000
/* Row kernel, together with ESRP_VARIANT_ROW_ALT, kc = 64, bn = 32, an even batch: launch clusters of two CTAs that
 * stream the same rows of images i and i + n/2 and share every MMA (tcgen05 cta_group::2, M = 256; each CTA keeps half
 * of the weight rows resident).  Experimental: measured slower than the single-CTA launches on the benchmark forward
 * (DESIGN.md section 4.9), so the engine only sets it when ESRP_PAIR=1; ESRP_PAIR=0 switches it off everywhere. */
#define ESRP_VARIANT_PAIR 0x4000
/* Timing experiments (row kernel; the conv result is WRONG with any of these set). */
#define ESRP_DBG_NO_XHALO 0x100 /* load boxes at x0 instead of x0-1 (no out-of-bounds on the left)  */
#define ESRP_DBG_NO_MMA 0x200   /* TMA only: stages are released without issuing MMAs              */
#define ESRP_DBG_NO_TMA 0x400   /* MMA only: stages are marked full without loading                 */
#define ESRP_DBG_NO_EPI 0x800   /* epilogue releases the accumulators without reading / storing     */
#define ESRP_DBG_EMPTY 0x1000   /* kernel prologue + teardown only                                  */

/* Packed-weight layouts == kernel decompositions (esrp_conv3x3_t.w_layout, esrp_pack_conv3x3_weights):
 *   ROW : one 128-pixel image row per M-tile, kernel rows ky stacked along N, column shift by
 *         shifted shared-memory operands (conv3x3_row.cuh).  Best for images wider than ~64 px;
 *         bn must be 16 or 32.
 *   TILE: RM x CW pixel M-tiles, kernel columns kx stacked along N, column shift by warp shuffles
 *         (conv3x3_tc.cuh).  Any width; the choice for narrow images (training crops). */
#define ESRP_LAYOUT_TILE 0
#define ESRP_LAYOUT_ROW 1

/* One fused 3x3 / stride-1 / zero-pad-1 convolution over NHWC bf16 activations.
 * Replaces one `conv_block` call plus the torch.cat that feeds it and the elementwise ops that
 * follow it in block.py:260-268 / 287-291 and architecture.py:55-71:
 *
 *   acc  = sum_{chunk,tap} W[chunk][tap] . src[chunk_src][.., chunk_c0 : chunk_c0+kc]
 *   aux  = sum_{chunk<aux_chunks} Waux[chunk] . src (centre tap only)            (conv1x1, :263)
 *   v    = s0 * act(acc + bias) + aux + s1 * r1                                   (:262-268)
 *   v    = v * (1 + sigma * N(0,1))           if noise                            (:117-121)
 *   v    = s2 * v + r2                        if r2                               (:291)
 *   out_* <- v
 *
 * Training extensions (the autograd backward of the same lines, SRRaGAN_model.py:140,167):
 *   mask_out : bit (mask_out_c0 + ch) of pixel's mask word row <- (acc + bias > 0), the LeakyReLU
 *              derivative selector saved by the forward pass (1 bit per activation);
 *   r2_pre   : r2 (if any) is added BEFORE the pre_* stores (v += r2) and s2 becomes a plain final scale;
 *   pre_bf16 / pre_f32 : copies of v taken before mask_in / noise / s2 (the gradient of a tensor
 *              that is both an operand of a later launch and the input of an activation);
 *   mask_in  : v *= bit ? 1 : 0.2  (LeakyReLU backward) — applied after the pre_* stores.
 * The data-gradient of a conv is this same operator over weights packed by
 * esrp_pack_dgrad_weights; the gradient of a dense block is again a dense block.
 */
typedef struct esrp_conv3x3 {
  int32_t n, h, w;              /* batch, height, width (output == input spatial size)       */
  const void* src[2];           /* up to two NHWC bf16 source tensors [n,h,w,src_ctotal[i]]  */
  int32_t src_ctotal[2];        /* channels of each allocation (multiple of 8)               */
  int32_t kc;                   /* K-chunk width in channels: 32 or 64                       */
  int32_t num_chunks;           /* K = 9 * kc * num_chunks                                   */
  int32_t chunk_src[ESRP_MAX_CHUNKS]; /* which src each chunk reads                         */
  int32_t chunk_c0[ESRP_MAX_CHUNKS];  /* first channel of the chunk in that src             */
  int32_t aux_chunks;           /* leading chunks that also feed the 1x1 aux accumulator     */
  int32_t bn;                   /* padded Cout = UMMA N: 16, 32 or 64                        */
  int32_t cout;                 /* real Cout <= bn                                           */
  const void* w_packed;         /* from esrp_pack_conv3x3_weights (incl. the conv1x1 rows)   */
  int32_t w_layout;             /* ESRP_LAYOUT_* the weights were packed for                 */
  const float* bias;            /* [bn] fp32 (zero padded), or NULL                          */
  int32_t act;                  /* 0 none, 1 LeakyReLU(0.2), 2 ReLU                          */
  float s0;
  const void* r1;               /* residual 1, NHWC [n,h,w,r1_ctotal], read at r1_c0..       */
  int32_t r1_is_f32, r1_ctotal, r1_c0;
  float s1;
  const void* r2;               /* residual 2 (RRDB level)                                   */
  int32_t r2_is_f32, r2_ctotal, r2_c0;
  float s2;
  int32_t noise;                /* 1: multiplicative Gaussian noise (train mode)             */
  int32_t noise_ctotal, noise_c0; /* Philox element index = pixel*noise_ctotal + noise_c0 + ch */
  float sigma;
  uint64_t seed, offset;        /* Philox key / per-call counter offset                      */
  void* out_bf16;               /* NHWC bf16 [n,h,w,ob_ctotal] written at ob_c0.. or NULL    */
  int32_t ob_ctotal, ob_c0;
  void* out_f32;                /* NHWC fp32 twin, or NULL                                   */
  int32_t of_ctotal, of_c0;
  float* out_nchw;              /* NCHW fp32 [n,cout,h,w], or NULL                           */
  int32_t variant;              /* ESRP_VARIANT_* bits                                       */
  void* trace;                  /* NULL, or device int64[3*1024]: clock64 timeline of CTA 0  */
  /* ---- training extensions (all optional; zero = off) ---- */
  void* mask_out;               /* uint16 words, [n*h*w][mask_out_ctotal/16]; bit = pre-act > 0 */
  int32_t mask_out_ctotal, mask_out_c0; /* bits per pixel / first bit (multiples of 16)       */
  const void* mask_in;          /* same format; v *= bit ? 1 : 0.2                            */
  int32_t mask_in_ctotal, mask_in_c0;
  int32_t r2_pre;               /* 1: v += r2 before the pre_* stores; s2 scales at the end   */
  void* pre_bf16;               /* NHWC bf16 copy of v before mask_in/noise/s2, or NULL       */
  int32_t pb_ctotal, pb_c0;
  void* pre_f32;                /* NHWC fp32 copy of v before mask_in/noise/s2, or NULL       */
  int32_t pf_ctotal, pf_c0;
  /* ---- co-scheduled output slices (optional; 0 or 1 = off; ESRP_LAYOUT_ROW only) ----
   * slices = S > 1: ONE launch computes S consecutive slices of bn output channels of the same conv
   * (a 64-channel conv whose K is too large for one resident weight set, e.g. conv5 of a dense block,
   * block.py:258).  Slice s uses w_packed + s*slice_stride and bias + s*slice_stride (bytes) and adds
   * s*bn to every channel offset (ob_c0, of_c0, r1_c0, r2_c0, noise_c0, mask_*_c0, pb_c0, pf_c0).
   * CTAs S*i..S*i+S-1 walk the same image rows, so the input is fetched from HBM once. */
  int32_t slices;
  int64_t slice_stride;
  /* ---- layout of the fp32 operands (optional; ESRP_LAYOUT_ROW, no training extensions) ----
   * 0: r1 / r2 / out_f32 (where fp32) are NHWC [n,h,w,c].  1: they are [n,h,c/4,w,4] ("planar"): the
   * row kernel maps one thread to one pixel, so a warp then reads / writes 512 contiguous bytes per
   * instruction.  Private to a caller that owns both producer and consumer of the tensor (the
   * engine's fp32 residual trunk); channel counts / offsets must be multiples of 4. */
  int32_t f32_planar;
  /* ---- K padding (optional) ----
   * k_valid > 0: only the first k_valid of the num_chunks*kc input channels carry non-zero weights
   * (a dense-block conv whose Cin is not a multiple of kc, block.py:253-257: conv2 96, conv4 160).
   * The kernel may skip MMAs over the zero-weight tail; results are unchanged. */
  int32_t k_valid;
  /* ---- split-precision output (optional; ESRP_LAYOUT_TILE only) ----
   * out_lo = 1: besides hi = bf16(v) at ob_c0 the kernel stores lo = bf16(v - hi) at ob_lo_c0 of the SAME out_bf16 tensor.
   * A caller that feeds [hi | lo] activations and [W_hi | W_hi | W_lo] weights through three chunk groups
   * (A_hi, A_lo, A_hi) gets products accurate to ~16 mantissa bits out of the bf16 tensor pipe: the fp32-parity mode of
   * esrganplus_b200/precise.py. */
  int32_t out_lo;
  int32_t ob_lo_c0;
} esrp_conv3x3_t;

const char* esrp_last_error(void);
int esrp_version(void);
/* sizeof(esrp_conv3x3_t) as compiled into the library (binding self-check). */
int32_t esrp_sizeof_conv3x3(void);
/* HOST helper: the N(0,1) samples the noise epilogue draws, for element indices [0,count) of a
 * conv whose Philox key/offset are (seed, offset): element e = pixel*cout + channel (NHWC order)
 * uses counter offset + e/4, lane e%4.  out_host is a host pointer.  Lets a caller reproduce the
 * GaussianNoise draw (block.py:120) outside the kernel; the engine uses offset = rdb_index << 36. */
int esrp_philox_normal_host(uint64_t seed, uint64_t offset, int64_t count, float* out_host);

/* Device/SM query: returns the SM count of the current device, or -1. */
int esrp_sm_count(void);

int esrp_conv3x3_nhwc(const esrp_conv3x3_t* desc, void* stream);

/* Packed size in bytes of the weights for (num_chunks, kc, bn); has_aux != 0 reserves the conv1x1
 * rows (aux_chunks > 0 in the conv descriptor). */
int64_t esrp_packed_conv3x3_bytes(int32_t num_chunks, int32_t kc, int32_t bn, int32_t has_aux);

/* Repack reference-format weights (fp32 [w_o, w_i, 3, 3], device) into UMMA B tiles
 * [chunk][ky][row = kx*bn + r (, 3*bn + r: conv1x1)][kc] (layout TILE; ROW swaps ky and kx),
 * bf16, rows pre-swizzled.
 *   transpose == 0 (forward operator):  row r, k  <-  W[row0 + r][lc0[chunk] + k][ky][kx]
 *   transpose != 0 (data-gradient op):  row r, k  <-  W[lc0[chunk] + k][row0 + r][2-ky][2-kx]
 * `rows` (<= bn) logical output channels starting at row0 are packed; indices outside the tensor
 * are zero-filled.  chunk_lc0_host[i] = first logical input channel of chunk i.
 * w_aux_oi: the bias-free 1x1 conv [w_o, aux_cin] (block.py:244,263) feeding the first aux_chunks
 * chunks (forward only), or NULL with aux_chunks = 0. */
int esrp_pack_conv3x3_weights(const float* w_oihw, int32_t w_o, int32_t w_i, int32_t transpose,
                              int32_t layout, int32_t row0, int32_t rows, int32_t kc, int32_t bn,
                              int32_t num_chunks,
                              const int32_t* chunk_lc0_host, const float* w_aux_oi, int32_t aux_cin,
                              int32_t aux_chunks, void* out, void* stream);

/* Boundary layout converters (reference tensors are NCHW fp32, test_image/test.py:31-35). */
int esrp_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t n, int32_t c, int32_t h,
                               int32_t w, int32_t c_pad, void* stream);
int esrp_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int32_t n, int32_t c, int32_t h,
                               int32_t w, int32_t c_total, void* stream);
/* Image plumbing of the caller, on the device.
 * test_image/test.py:31-35: uint8 HWC image (cv2: BGR, bgr != 0 swaps to RGB) -> /255 -> NHWC bf16 network input;
 * test_image/test.py:37-39: NCHW fp32 output -> clamp(0,1) -> x255 -> round half to even -> uint8 HWC (RGB -> BGR if bgr). */
int esrp_u8hwc_to_nhwc_bf16(const uint8_t* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c, int32_t c_pad,
                            int32_t bgr, void* stream);
int esrp_nchw_f32_to_u8hwc(const float* src, uint8_t* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t bgr,
                           void* stream);
/* nn.Upsample(scale_factor=2, mode='nearest') on NHWC bf16 (block.py:319). */
int esrp_upsample2x_nhwc_bf16(const void* src, void* dst, int32_t n, int32_t h, int32_t w,
                              int32_t c, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Discriminator_VGG_128 pieces (architecture.py:87-129).  The convolutions run on esrp_conv3x3_nhwc:
 * a 4x4 / stride-2 / pad-1 conv is the 2x2 conv (3x3 with five zero taps) over the space-to-depth
 * tensor produced here; BatchNorm2d (block.py:32) is statistics + a per-channel affine + LeakyReLU.
 * ------------------------------------------------------------------------------------------- */
/* dst[n, Y, X, (a*2+b)*c + ch] = src[n, 2Y+a-1, 2X+b-1, ch] (0 outside), Y <= h/2, X <= w/2; NHWC bf16. */
int esrp_s2d_pad_nhwc_bf16(const void* src, void* dst, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
/* sums2c[ch] = sum, sums2c[c+ch] = sum of squares over the valid [h,w] region of x [n,hp,wp,c] fp32. */
int esrp_bn_stats_nhwc_f32(const float* x, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, int32_t c,
                           double* sums2c, void* stream);
/* y = [lrelu](x*scale[ch] + shift[ch]) over the valid region -> NHWC bf16 [n,h,w,c] and/or NCHW fp32. */
int esrp_bn_apply_nhwc(const float* x, int32_t n, int32_t h, int32_t w, int32_t hp, int32_t wp, int32_t c,
                       const float* scale, const float* shift, int32_t act, void* out_bf16,
                       float* out_nchw_f32, void* stream);
/* Backward of BatchNorm2d (batch statistics) + LeakyReLU(0.2) over the valid [h,w] region of the conv output z
 * [n,hp,wp,c] fp32.  dout is the gradient of the activated output: bf16 NHWC [n,h,w,c], or (when dout_bf16 is
 * NULL) fp32 in NCHW-flattened order [n, c*h*w] (the classifier's input gradient).  coef7c = fp32 [7][c]:
 * mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, g_rs = gamma*rstd, a, b.  With dzb = dout * lrelu'(z*scale+shift):
 *   reduce: sums2c[ch] = sum dzb (= d beta), sums2c[c+ch] = sum dzb * xhat (= d gamma), xhat = (z - mean) * rstd
 *   apply : dz = g_rs * (dzb - a - xhat * b)  (a = sum dzb / N, b = sum dzb*xhat / N; a = b = 0 for eval-mode BN),
 *           written as bf16 on the whole [n,hp,wp,c] grid, zero outside the valid region. */
int esrp_bn_bwd_reduce(const float* z, const void* dout_bf16, const float* dout_nchw_f32, int32_t n, int32_t h, int32_t w,
                       int32_t hp, int32_t wp, int32_t c, const float* coef7c, double* sums2c, void* stream);
int esrp_bn_bwd_apply(const float* z, const void* dout_bf16, const float* dout_nchw_f32, int32_t n, int32_t h, int32_t w,
                      int32_t hp, int32_t wp, int32_t c, const float* coef7c, void* dz_bf16, void* stream);
/* BatchNorm2d bookkeeping on [c]-sized vectors in one launch.  esrp_bn_finalize: from the sums of
 * esrp_bn_stats_nhwc_f32 over `count` samples per channel (training != 0: batch statistics, biased variance for the
 * normalisation and the momentum update of running_mean / running_var with the unbiased variance — nn.BatchNorm2d
 * semantics, block.py:32; training == 0: the running statistics) to coef7c rows mean, rstd, scale, shift, g_rs
 * (rows a, b zeroed).  esrp_bn_bwd_finalize: from the sums of esrp_bn_bwd_reduce to d_gamma / d_beta and, for
 * batch-statistics layers, the a / b rows. */
int esrp_bn_finalize(const double* sums2c, double count, const float* gamma, const float* beta, float eps, float momentum,
                     int32_t training, float* running_mean, float* running_var, int32_t c, float* coef7c, void* stream);
int esrp_bn_bwd_finalize(const double* sums2c, double count, int32_t batch_stats, int32_t c, float* coef7c, float* dgamma,
                         float* dbeta, void* stream);
/* Inverse of esrp_s2d_pad_nhwc_bf16 (gradient of the rearrangement): ds [n,h/2+1,w/2+1,4c] -> din [n,h,w,c];
 * lrelu_ref (optional, bf16 [n,h,w,c]): din *= (ref > 0 ? 1 : 0.2), the LeakyReLU derivative of the activation fed forward. */
int esrp_s2d_pad_bwd_nhwc_bf16(const void* ds, void* din, const void* lrelu_ref, int32_t n, int32_t h, int32_t w, int32_t c,
                               void* stream);
/* nn.Linear backward; yout_act (optional) = the layer's LeakyReLU-activated output (dy is masked by its sign).
 * dx [b,k], dw [o,k], db [o]: each optional. */
int esrp_linear_bwd_f32(const float* dy, const float* yout_act, const float* x, const float* w, float* dx, float* dw, float* db,
                        int32_t b, int32_t k, int32_t o, void* stream);
/* Concurrency helper for layers whose independent launches each fill a fraction of the GPU (the discriminator's deep
 * layers: 7-70 M-tiles per output-channel slice): n non-blocking side streams owned by the library; fork makes them wait
 * for everything enqueued on `main` so far, join makes `main` wait for everything enqueued on them. */
int esrp_streams_create(int32_t n, void** out_streams);
int esrp_streams_fork(void* main_stream, void* const* sides, int32_t n);
int esrp_streams_join(void* main_stream, void* const* sides, int32_t n);
/* ---------------------------------------------------------------------------------------------
 * Solver arithmetic between the passes (csrc/esrp_solver.cu)
 * ------------------------------------------------------------------------------------------- */
/* torch.optim.Adam (SRRaGAN_model.py:82-89: lr, betas, weight_decay as L2, eps 1e-8, no amsgrad) over flat storage, one
 * kernel: p, g, m (exp_avg), v (exp_avg_sq) are device arrays of n floats (n % 4 == 0, 16-byte aligned), `step` counts
 * from 1 (bias corrections 1 - beta^step, evaluated in double like torch; hyper-parameters are doubles for the same
 * reason: 1 - beta2 must round to float once).  The learning rate of the step comes from the caller (MultiStepLR,
 * SRRaGAN_model.py:91-95, is a host-side schedule). */
int esrp_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int64_t step, void* stream);
/* Relativistic-average GAN loss terms (SRRaGAN_model.py:133-136,151-154 with GANLoss('vanilla') = BCEWithLogits,
 * loss.py:11-38): A = BCE(pred_real - mean(pred_fake), t_real), B = BCE(pred_fake - mean(pred_real), t_fake) over n
 * logits each.  out4 = {A, B, mean(pred_real), mean(pred_fake)}; dA_* / dB_* [n]: gradients of the two terms w.r.t. the
 * raw logits (through the means).  One single-block launch. */
int esrp_ragan_bce(const float* pred_real, const float* pred_fake, int32_t n, float t_real, float t_fake, float* out4,
                   float* dA_dreal, float* dA_dfake, float* dB_dreal, float* dB_dfake, void* stream);
/* nn.L1Loss (SRRaGAN_model.py:31-39,123): *loss = mean |a - b| over n floats (n % 4 == 0), grad (optional) = d loss / d a
 * = sign(a - b) / n.  scratch: one device double. */
int esrp_l1_loss_grad(const float* a, const float* b, int64_t n, float* grad, float* loss, double* scratch, void* stream);
/* Capture everything enqueued on `stream` (and on streams forked from it with esrp_streams_fork / joined with
 * esrp_streams_join) between begin and end into an executable CUDA graph; launch replays it with one call.  The
 * discriminator's passes are lists of ~150 recorded C-ABI calls each: replayed as graphs they cost one launch of host
 * time.  abort ends a capture after a failed call. */
int esrp_graph_begin(void* stream);   /* (not the legacy default stream: it cannot be captured) */
int esrp_graph_end(void* stream, void** out_exec);
int esrp_graph_abort(void* stream);
int esrp_graph_launch(void* exec, void* stream);
void esrp_graph_destroy(void* exec);
/* cudaMemsetAsync(p, 0, bytes) as a recordable / capturable C-ABI call. */
int esrp_memset_zero(void* p, int64_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Perceptual branch: VGGFeatureExtractor (architecture.py:279-307; csrc/esrp_vgg.cu).
 * Its sixteen 3x3 convs run on esrp_conv3x3_nhwc with act = 2 (ReLU); these are the pieces between them.
 * ------------------------------------------------------------------------------------------- */
/* (x * scale[c] + shift[c]) -> NHWC bf16 padded to c_pad channels: `(x - mean) / std` of architecture.py:304-305 with
 * scale = 1 / std, shift = -mean / std (both fp32 [c], or both NULL for a plain re-layout). */
int esrp_nchw_f32_to_nhwc_bf16_affine(const float* src, const float* scale, const float* shift, void* dst, int32_t n, int32_t c,
                                      int32_t h, int32_t w, int32_t c_pad, void* stream);
/* nn.MaxPool2d(kernel_size=2, stride=2) on NHWC bf16 [n,h,w,c] -> [n,h/2,w/2,c] (h, w even, c % 8 == 0). */
int esrp_maxpool2x2_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
/* dz = dy where y > 0 else 0; y = the ReLU's OUTPUT (bf16, count elements, count % 8 == 0). */
int esrp_relu_bwd_nhwc_bf16(const void* y, const void* dy, void* dz, int64_t count, void* stream);
/* Gradient through [ReLU, MaxPool2d(2,2)]: y [n,h,w,c] = the ReLU output that was pooled, dpool [n,h/2,w/2,c]; the first
 * maximum of each window in torch's scan order receives dpool if it is positive, every other element gets 0. */
int esrp_maxpool2x2_relu_bwd_nhwc_bf16(const void* y, const void* dpool, void* dz, int32_t n, int32_t h, int32_t w, int32_t c,
                                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training data path on the device (csrc/esrp_data.cu)
 * ------------------------------------------------------------------------------------------- */
/* One training sample: an HR image resident on the device as the reference reads it (cv2: uint8, HWC, BGR) plus the
 * decisions the reference draws on the host (LRHR_dataset.py:99-100 crop offsets in LR pixels, util.py:96-98 flips /
 * rotation) and the MATLAB-bicubic tables of its height and width (util.py:219-274 calculate_weights_indices for
 * scale 1/s: wh [h/s][ph] weights, ih [h/s] first index of each window in the symmetric-padded image, sym_hs = rows
 * mirrored in front; same for the width).  All pointers are device pointers. */
typedef struct {
  const uint8_t* img;
  int32_t h, w;
  int32_t rnd_h, rnd_w;
  int32_t hflip, vflip, rot90;
  const float* wh;
  const int32_t* ih;
  const float* ww;
  const int32_t* iw;
  int32_t ph, pw, sym_hs, sym_ws;
} esrp_lrhr_job_t;
int32_t esrp_sizeof_lrhr_job(void);
/* LRHR_dataset.py:83-121 with on-the-fly LR for `count` samples (jobs_dev: device array): lr_out [count,3,hr_size/s,hr_size/s],
 * hr_out [count,3,hr_size,hr_size] fp32 RGB CHW in [0,1] — util.imresize_np(img/255, 1/s) restricted to the crop, the HR
 * crop, util.augment, the BGR->RGB swap and the HWC->CHW transpose in one launch (one CTA per sample).  xw_max: an upper
 * bound of the padded columns one LR crop row reads: (hr_size/s - 1) * s + pw + 2. */
int esrp_lrhr_batch(const esrp_lrhr_job_t* jobs_dev, int32_t count, int32_t scale, int32_t hr_size, int32_t xw_max, float* lr_out,
                    float* hr_out, void* stream);

/* nn.Linear (+ optional LeakyReLU 0.2): y[b,o] = sum_k x[b,k] w[o,k] + bias[o]  (architecture.py:122-123). */
int esrp_linear_f32(const float* x, const float* w, const float* bias, float* y, int32_t b, int32_t k, int32_t o,
                    int32_t act, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward pass (autograd of block.py:260-268 / architecture.py:47-78 as driven by
 * SRRaGAN_model.py:140,167).  Data gradients run on esrp_conv3x3_nhwc over weights packed here;
 * weight gradients are accumulated per "unit" (32 input channels x 64 output-gradient channels x 9
 * taps) and scattered into the reference's OIHW fp32 layout.
 * ------------------------------------------------------------------------------------------- */
/* One 32-channel group of the K dimension of a data-gradient conv: K channel k of the group is the
 * gradient of output channel co0 + k of the source conv w [w_o, w_i, 3, 3] (fp32, device); w == NULL
 * packs zeros (a group that is not an operand of this slice). */
typedef struct esrp_dgrad_group {
  const float* w;
  int32_t w_o, w_i;
  int32_t co0;
  float scale;
} esrp_dgrad_group_t;
/* Packs rows [row0, row0+rows) (= INPUT channels of the source convs = output channels of the data
 * gradient) over num_groups K groups (chunks of kc = 32 or 64 channels: ceil(num_groups*32/kc) chunks,
 * size esrp_packed_conv3x3_bytes(chunks, kc, bn, 0)); taps are flipped (conv_transpose). */
int esrp_pack_dgrad_weights(const esrp_dgrad_group_t* groups_host, int32_t num_groups, int32_t layout,
                            int32_t row0, int32_t rows, int32_t kc, int32_t bn, void* out, void* stream);

#define ESRP_WGRAD_MAX_UNITS 256 /* per call; the units of one call may span at most 4 tensors and 64 (chunk, slab) jobs */
/* acc[tap = ky*3+kx][c < 64][i < 32] += sum_px dy[px][dy_c0 + c] * x[px + (ky-1, kx-1)][x_c0 + i]
 * (zero padding; dy channels beyond dy_ctotal read as zero); bias_acc[c] += sum_px dy[px][dy_c0 + c]. */
typedef struct esrp_wgrad_unit {
  const void* x;                /* NHWC bf16 [n,h,w,x_ctotal]  */
  int32_t x_ctotal, x_c0;
  const void* dy;               /* NHWC bf16 [n,h,w,dy_ctotal] */
  int32_t dy_ctotal, dy_c0;
  float* acc;                   /* fp32 [9][64][32], accumulated atomically (zero it first) */
  float* bias_acc;              /* fp32 [64] or NULL */
} esrp_wgrad_unit_t;
/* All units share the spatial shape; splits = CTAs per unit over the pixel dimension (0 = fill the GPU). */
int esrp_conv3x3_wgrad(const esrp_wgrad_unit_t* units_host, int32_t num_units, int32_t n, int32_t h, int32_t w,
                       int32_t splits, void* stream);
/* kind 0: dst[dst_off + ((co0 + c) * w_i + ci0 + i) * 9 + tap] = scale * acc[tap][col0 + c][i], c < ncols, i < nci
 * kind 1: dst[dst_off + i] = scale * acc[i], i < ncols.  dst_index is used by the engines only.
 * kind 2: a 4x4 / stride-2 conv run as a 3x3 conv over the space-to-depth tensor: unit channel ci0 + i =
 *         (a*2+b) * w_i + ci and tap (A+1, B+1) go to dst[((co0 + c) * w_i + ci) * 16 + (2A+a) * 4 + 2B+b]. */
typedef struct esrp_scatter_entry {
  const float* acc;
  float* dst;
  int64_t dst_off;
  int32_t dst_index;
  int32_t kind, col0, ncols, nci, co0, ci0, w_i;
  float scale;
} esrp_scatter_entry_t;
int esrp_wgrad_scatter(const esrp_scatter_entry_t* entries_host, int32_t num, void* stream);
/* Bias-free 1x1 conv of block.py:244,263, both gradients: g[px][0:nf] (fp32, in place) += u^T dx2[px]
 * (+ extra[px] if extra != NULL); du_acc[32][nf] (fp32, atomics; NULL to skip) += dx2 (x) x.
 * x: NHWC bf16 (first nf of x_ctotal channels), dx2: NHWC bf16 32 channels at d_c0, u: fp32 [32][nf]. */
int esrp_conv1x1_bwd(int32_t nf, const void* x, int32_t x_ctotal, const void* dx2, int32_t d_ctotal, int32_t d_c0,
                     const float* u, float* g, const float* extra, float* du_acc, int64_t npx, void* stream);
/* nn.Upsample(x2, nearest) backward: [n,2h,2w,c] bf16 -> 2x2 sums [n,h,w,c] (fp32 out_f32, unmasked) and
 * bf16 out_bf16 after the optional LeakyReLU mask (bit mask_c0 + ch of the pixel's mask_ctotal bits). */
int esrp_upsample2x_bwd_nhwc_bf16(const void* dup, int32_t n, int32_t h, int32_t w, int32_t c, const void* mask,
                                  int32_t mask_ctotal, int32_t mask_c0, void* out_bf16, float* out_f32, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RRDBNet generator engine: what RRDBNet.forward / RRDB_Net.forward binds to
 * (architecture.py:76-78, test_image/architecture.py:36-38).  The handle owns the packed bf16
 * weight cache and per-shape launch plans; activations live in the caller's workspace.
 * ------------------------------------------------------------------------------------------- */
typedef struct esrp_rrdbnet esrp_rrdbnet_t;

/* Mirrors RRDBNet(in_nc, out_nc, nf, nb, gc=32, upscale=4, norm_type=None, act_type='leakyrelu',
 * mode='CNA', upsample_mode='upconv') (architecture.py:48-49).  Supported: nf in {32,64}, gc=32,
 * upscale in {1,2,4}; anything else returns an error (no fallback). */
int esrp_rrdbnet_create(int32_t in_nc, int32_t out_nc, int32_t nf, int32_t nb, int32_t gc,
                        int32_t upscale, esrp_rrdbnet_t** out);
void esrp_rrdbnet_destroy(esrp_rrdbnet_t* h);
/* The state_dict tensors the engine consumes, in reference order, with their key names
 * ("model.0.weight", "model.1.sub.0.RDB1.conv1x1.weight", ...) and OIHW shapes. */
int32_t esrp_rrdbnet_num_tensors(const esrp_rrdbnet_t* h);
const char* esrp_rrdbnet_tensor_key(const esrp_rrdbnet_t* h, int32_t idx);
int esrp_rrdbnet_tensor_shape(const esrp_rrdbnet_t* h, int32_t idx, int32_t* dims4);
/* (Re)build the packed weight cache from fp32 device tensors; ptrs is a HOST array of `count`
 * DEVICE pointers ordered like esrp_rrdbnet_tensor_key.  Call after load_state_dict and after
 * every optimizer step. */
int esrp_rrdbnet_load_weights(esrp_rrdbnet_t* h, const void* const* ptrs, int32_t count, void* stream);
int64_t esrp_rrdbnet_workspace_bytes(const esrp_rrdbnet_t* h, int32_t n, int32_t hgt, int32_t w);
/* Kernel launches in the most recently planned forward (for bench.py's gpu_launches). */
int32_t esrp_rrdbnet_num_launches(const esrp_rrdbnet_t* h);
/* Persistent conv chain (opt-in): consecutive row-kernel convs of the inference plan — the five convs of every
 * ResidualDenseBlock_5C of the trunk (block.py:260-268) — run as phases of ONE launch that synchronises row
 * neighbours through flags in device memory (csrc/conv3x3_chain.cuh) instead of one launch per conv: 10 launches
 * per forward instead of 354.  Off by default: on B200 it measures 8 % SLOWER than the launch-per-conv path at
 * 16 x 128 x 128 (13.7 vs 12.7 ms) and 35 % slower for a single tile, because a phase boundary (epilogue tail ->
 * flag -> first TMA row, ~6 us) costs what a programmatic dependent launch does, and its sums are reproducible up
 * to fp32 addition order only.  enable != 0 selects it; the next forward re-plans. */
int esrp_rrdbnet_set_chain(esrp_rrdbnet_t* h, int32_t enable);
/* Device-time the dense-block convs of the trunk — the "RRDB 3x3-conv stack" BASELINE.json quotes the tensor-pipe
 * fraction on — of every following inference forward: one pair of CUDA events on the launching stream per forward
 * (ring of 64).  After synchronising the stream, esrp_rrdbnet_get_timing copies the milliseconds of the last
 * (at most `max`) timed forwards to ms[] (oldest first) and returns how many; -1 on error. */
int esrp_rrdbnet_set_timing(esrp_rrdbnet_t* h, int32_t enable);
int32_t esrp_rrdbnet_get_timing(esrp_rrdbnet_t* h, float* ms, int32_t max);
/* Conv launches the chains of the most recently planned forward replaced (0 when none was built). */
int32_t esrp_rrdbnet_num_chained_convs(const esrp_rrdbnet_t* h);
/* Conv launches of the most recently planned forward that run as clusters of two CTAs (ESRP_VARIANT_PAIR; 0 unless ESRP_PAIR=1). */
int32_t esrp_rrdbnet_num_pair_launches(const esrp_rrdbnet_t* h);
/* x: NCHW fp32 [n,in_nc,h,w] -> y: NCHW fp32 [n,out_nc,upscale*h,upscale*w] (unclamped, like the
 * reference).  training!=0 enables the per-RDB multiplicative Gaussian noise (block.py:117-121)
 * drawn from Philox(seed).  workspace: 1024-byte aligned device memory of at least
 * esrp_rrdbnet_workspace_bytes(). */
int esrp_rrdbnet_forward(esrp_rrdbnet_t* h, const float* x, float* y, int32_t n, int32_t hgt, int32_t w,
                         void* workspace, int64_t workspace_bytes, int32_t training, uint64_t seed,
                         void* stream);

/* The same forward fed and drained as 8-bit images (what test_image/test.py:31-40 does around model(img_LR) on the host):
 * x: uint8 [n,h,w,in_nc] HWC, y: uint8 [n,upscale*h,upscale*w,out_nc] HWC, both BGR when bgr != 0 (cv2 order).  The
 * workspace must hold esrp_rrdbnet_workspace_bytes_u8() bytes (the fp32 result lives at its end). */
int64_t esrp_rrdbnet_workspace_bytes_u8(const esrp_rrdbnet_t* h, int32_t n, int32_t hgt, int32_t w);
int esrp_rrdbnet_forward_u8(esrp_rrdbnet_t* h, const uint8_t* x, uint8_t* y, int32_t n, int32_t hgt, int32_t w,
                            void* workspace, int64_t workspace_bytes, int32_t bgr, void* stream);

/* ---- training (the autograd graph of architecture.py:76-78 as driven by SRRaGAN_model.py:120,140) ----
 * esrp_rrdbnet_train_forward computes the same y as esrp_rrdbnet_forward and keeps, in `workspace`, what the
 * backward needs (per dense block: its bf16 input, x1..x4, one LeakyReLU sign bit per activation; the
 * upsampled / HR activations of the tail); the GaussianNoise draws are regenerated from (seed, block index).
 * esrp_rrdbnet_backward must follow on the SAME workspace: dy is NCHW fp32 [n,out_nc,upscale*h,upscale*w];
 * grads is a HOST array of esrp_rrdbnet_num_tensors() DEVICE pointers (fp32, the shapes of the state_dict
 * tensors, ordered like esrp_rrdbnet_tensor_key), each OVERWRITTEN with dL/dtensor (NULL entries are skipped).
 * The gradient w.r.t. x is not produced: the LR input never requires one (SRRaGAN_model.py:103-111). */
int64_t esrp_rrdbnet_train_workspace_bytes(const esrp_rrdbnet_t* h, int32_t n, int32_t hgt, int32_t w);
int esrp_rrdbnet_train_forward(esrp_rrdbnet_t* h, const float* x, float* y, int32_t n, int32_t hgt, int32_t w,
                               void* workspace, int64_t workspace_bytes, int32_t noise, uint64_t seed, void* stream);
int esrp_rrdbnet_backward(esrp_rrdbnet_t* h, const float* dy, float* const* grads, int32_t count, void* workspace,
                          void* stream);
/* Kernel launches of the planned training forward (which = 0) / backward (which = 1). */
int32_t esrp_rrdbnet_train_num_launches(const esrp_rrdbnet_t* h, int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* ESRP_H_ */
