/* libzns_sm100 -- C ABI of the B200-native ZeroNS hot path.
 *
 * The reference (deezer/zeroNoteSamba) has no FFI: its boundary for this path is the Python API
 *   IR.generate_XQT(y, sr, mode)                       zeroNoteSamba/processing/input_rep.py:11-57
 *   _CNN / DS_CNN / Pretext_CNN / Down_CNN .forward    zeroNoteSamba/models/models.py:7-150
 *   NTXent(batch_len, temperature).forward             zeroNoteSamba/models/loss_functions.py:7-55
 *   train_epoch / val_epoch, Adam(lr=1e-6), crops      zeroNoteSamba/pretext.py:202,308-321,453-592
 * Each entry point below names the reference code it replaces.  The Python package
 * zeronotesamba_b200 binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions: every function returns 0 on success or a zns_status (text via zns_last_error(),
 * thread-local).  Pointers are raw device pointers unless the name ends in _host.  `stream` is
 * a cudaStream_t passed as void*.  Nothing here allocates, frees or synchronises caller memory
 * except where stated (plans own their scratch; *_host helpers copy and synchronise).
 * One process per GPU.  There is no CPU fallback: without a CUDA device calls fail with ZNS_ERR_CUDA.
 *
 * Activation layout used between encoder layers ("act"): 16-bit [G][H][W][8][C], G = ceil(B/8);
 * clip b lives at group b/8, slot b%8; slots >= B hold zeros.  Element type: the FORWARD activations and the
 * forward weight pack are fp16 or bf16 (the caller says which: `*_f16` arguments, zns_conv_desc.fmt; the product path
 * uses fp16 -- 11 significant bits instead of 8 at the same tensor-core rate, values are far inside its range),
 * GRADIENT activations and the data-gradient weight pack are always bf16 (range).
 */
#ifndef ZNS_H_
#define ZNS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZNS_VERSION 100

typedef enum {
  ZNS_OK = 0,
  ZNS_ERR_INVALID = 1, /* bad argument / unsupported shape */
  ZNS_ERR_CUDA = 2,    /* CUDA runtime / driver error (incl. no device) */
  ZNS_ERR_ALLOC = 3
} zns_status;

int zns_version(void);
const char* zns_last_error(void);
/* 0 if a CUDA device of compute capability 10.x is current, else an error. */
int zns_device_check(void);

/* ------------------------------------------------------------------------------------------
 * VQT / CQT front-end.  Replaces librosa.vqt / librosa.cqt as called at input_rep.py:27-34,42-49
 * plus the log-magnitude of input_rep.py:36-37,51-52.
 * ---------------------------------------------------------------------------------------- */
typedef struct zns_vqt_plan zns_vqt_plan;

/* Host-only: time-domain filter kernels of one octave (0 = top) of the librosa 0.8.1 basis
 * (constant_q -> FFT -> sparsify_rows(0.01) -> * sqrt(2^octave)), transformed back so that
 *   C[k,t] = sum_n (re[k*n_fft+n] + i im[k*n_fft+n]) * ypad[t*hop_octave + n].
 * re/im hold bins_per_octave*n_fft floats (n_fft <= 1024).  gamma < 0 selects librosa's VQT
 * default 24.7*alpha/0.108; gamma = 0 is the CQT.  Needs no GPU. */
int zns_vqt_basis_host(int sr, int n_bins, int bins_per_octave, double fmin, double gamma, int octave,
                       float* re, float* im, int* n_fft);
/* Host-only: the 32 decimator taps h[0..31] of resampy 0.4.2 "kaiser_fast" at ratio 1/2
 * (y[t] = sqrt(2) * sum_{|j|<=31} h[|j|] x[2t+j]). */
int zns_vqt_decimator_taps_host(double* taps32);

int zns_vqt_plan_create(int sr, int hop, int n_bins, int bins_per_octave, double fmin, double gamma,
                        int max_batch, int max_samples, zns_vqt_plan** plan);
int zns_vqt_plan_destroy(zns_vqt_plan* plan);
/* 1 + n_samples / hop -- the frame count generate_XQT returns (pretext.py:255-256 relies on 626). */
int zns_vqt_num_frames(int n_samples, int hop);
/* y [batch][n_samples] fp32 -> out [batch][n_bins][frames] fp32 = log(|V| + 1e-9). */
int zns_vqt_forward(zns_vqt_plan* plan, const float* y, int batch, int n_samples, float* out, void* stream);
/* Host buffers in and out (the numpy contract of generate_XQT): H2D, transform, D2H, synchronise. */
int zns_vqt_forward_host(zns_vqt_plan* plan, const float* y_host, int batch, int n_samples, float* out_host,
                         void* stream);

/* Crop sampler of pretext.py:308-318: out[i] = vqt[:, :, starts[i] : starts[i]+T].
 * vqt [C][bins][F] fp32 (C = 2 stems), starts int32 [n_crops] on device, out [n_crops][C][bins][T]. */
int zns_crop_gather(const float* vqt, int channels, int bins, int frames, const int32_t* starts, int n_crops,
                    int crop_frames, float* out, void* stream);

/* RMS stem gate, check_CL_clips of processing/stem_check.py:21-51 (used at pretext.py:66-81):
 * counts[clip] = number of librosa.feature.rms frames (2048 / hop 512, centred, reflect) with
 * rms(ros)/2 < rms(stem) < 4 rms(ros); there are 1 + n_samples/512 frames.  stem/ros fp32 [batch][n]. */
int zns_rms_gate(const float* stem, const float* ros, int batch, int n_samples, int32_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------
 * Encoder layers (models.py:16-74).  "Same" convolutions, stride 1.
 * ---------------------------------------------------------------------------------------- */
/* cv1 (models.py:16,37-39): x fp32, clip b / row h / frame w at x[b*x_clip_stride + h*x_row_stride + w]
 * (the (B,1,96,T) NCHW input; the strides let channel 0 / 1 of a (B,2,96,T) crop batch be read in
 * place, pretext.py:476-477, and let overlapping time segments of ONE long clip act as the "clips" of a
 * group: x_clip_stride = segment hop, x_row_stride = T)
 * -> act [G][H][W][8][64] = dropout(relu(conv + bias)), fp16 if out_f16 else bf16; out_act_bf16 (may be NULL)
 * receives the same values as bf16 (the x operand of cv2's weight gradient, whose two operands must share a type). */
int zns_conv1_fwd(const float* x, long long x_clip_stride, long long x_row_stride, const float* weight, const float* bias,
                  void* out_act,
                  int batch, int H, int W, float dropout_p, uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream,
                  int out_f16, void* out_act_bf16, void* stream);
/* cv1 weight/bias gradient: dy act bf16 [G][H][W][8][64]; dw fp32 [64][1][3][11] and db [64] are
 * accumulated into (+=). */
int zns_conv1_wgrad(const void* dy_act, const float* x, long long x_clip_stride, long long x_row_stride, float* dw,
                    float* db, int batch, int H, int W, void* stream);

typedef struct {
  int batch;      /* B clips (act tensors hold ceil(B/8) groups) */
  int H, W;       /* frequency rows, time frames of input and output */
  int c_in;       /* channels of the input act (multiple of 64) */
  int c_out;      /* channels of the output act: 64, 128 or 256 */
  int kh, kw;     /* odd filter extents; padding is (kh/2, kw/2) */
  int relu;       /* epilogue ReLU */
  float dropout_p;/* epilogue dropout (train mode), 0 disables */
  uint32_t seed, rng_stream;
  const uint32_t* seed_dev; /* optional device word XORed into seed (a step counter: lets a captured
                               CUDA graph draw a fresh mask on every replay); NULL = unused */
  float out_scale;/* epilogue multiplies by this after masking (dgrad: 1/(1-p)) */
  int fmt;        /* ZNS_FMT_* bits: which operands are fp16 (0 = all bf16) */
} zns_conv_desc;
#define ZNS_FMT_IN_F16 1  /* zns_conv_fwd: `in` is fp16 */
#define ZNS_FMT_W_F16 2   /* zns_conv_fwd: `wpk` is fp16 */
#define ZNS_FMT_OUT_F16 4 /* zns_conv_fwd: `out` is written as fp16 */
/* forward convolution of the product path / its data gradient (dy bf16, flipped weights bf16, mask = fp16 forward act;
 * the mask test "> 0" reads sign and magnitude bits and is the same for both types) */
#define ZNS_FMT_FORWARD_F16 (ZNS_FMT_IN_F16 | ZNS_FMT_W_F16 | ZNS_FMT_OUT_F16)

/* cv2..cv8 forward (models.py:17-23,41-70) and, with transposed/flipped packed weights, the data
 * gradient.  Implicit GEMM on tcgen05/TMEM, operands staged by TMA.  `n_br` (1 or 2) encoders of
 * identical geometry are processed by one launch (anchor and postve, models.py:114-124).
 *   in[br]    act bf16 [G][H][W][8][c_in]
 *   wpk[br]   bf16 [kh*kw][c_out][c_in]   (zns_pack_weights)
 *   bias[br]  fp32 [c_out] or NULL
 *   mask[br]  act bf16 [G][H][W][8][c_out] or NULL: output is zeroed where mask <= 0
 *             (dgrad through ReLU/dropout of the layer below)
 *   out[br]   act bf16 [G][H][W][8][c_out] = scale * mask(dropout(relu?(conv + bias)))
 *   out_bf16  NULL, or per branch a second act tensor that receives the same values as bf16 (used when `out` is
 *             fp16: the weight gradient of the next layer needs its x operand in dy's type; tcgen05.mma kind::f16
 *             with one fp16 and one bf16 operand is an illegal instruction on sm_100a)
 * d->fmt says which of in / wpk / out are fp16. */
int zns_conv_fwd(const zns_conv_desc* d, int n_br, const void* const* in, const void* const* wpk,
                 const float* const* bias, const void* const* mask, void* const* out, void* const* out_bf16,
                 void* stream);

/* Weight gradient of cv2..cv8: x act bf16 [G][H][W][8][c_in], dy act bf16 [G][H][W][8][c_out],
 * dwpk[br] fp32 [kh*kw][c_out][c_in] accumulated (+=, atomics).  Only d->batch,H,W,c_in,c_out,kh,kw
 * are read. */
int zns_conv_wgrad(const zns_conv_desc* d, int n_br, const void* const* x, const void* const* dy,
                   float* const* dwpk, void* stream);

/* db[c] += sum over positions of dy act[...][c]. */
int zns_bias_grad(const void* dy_act, int batch, int H, int W, int C, float* db, void* stream);

/* fp32 [c_out][c_in][kh][kw] (state_dict layout, models.py:16-23) -> forward pack
 * wf [kh*kw][c_out][c_in] (fp16 if wf_f16 else bf16) and bf16 data-gradient pack wd [kh*kw][c_in][c_out] with taps
 * flipped.  Either output may be NULL. */
int zns_pack_weights(const float* w, int c_out, int c_in, int kh, int kw, void* wf, void* wd, int wf_f16, void* stream);
/* fp32 [kh*kw][c_out][c_in] -> g[c_out][c_in][kh][kw] (= or +=) scale * packed. */
int zns_unpack_grads(const float* gpk, int c_out, int c_in, int kh, int kw, float scale, int accumulate, float* g,
                     void* stream);

/* MaxPool2d((pool,1)) -> ReLU -> Dropout (models.py:41-44,50-53,59-62):
 * y act [G][H][W][8][C] -> out act [G][H/pool][W][8][C] (both fp16 if act_f16 else bf16); out_act_bf16 (may be NULL)
 * receives a bf16 copy of out. */
int zns_pool_fwd(const void* y_act, void* out_act, int batch, int H, int W, int C, int pool, float dropout_p,
                 uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream, int act_f16, void* out_act_bf16,
                 void* stream);
/* Backward of the above: dpool act [G][H/pool][W][8][C] (already masked and scaled by the dgrad
 * epilogue) is routed to the first arg-max row of each window of y; dy act [G][H][W][8][C]. */
int zns_pool_bwd(const void* y_act, const void* dpool_act, void* dy_act, int batch, int H, int W, int C, int pool,
                 int y_f16, void* stream);

/* fc1 + sigmoid + flatten (models.py:99-101): x act [G][1][T][8][128] -> emb fp32 [B][T]. */
int zns_head_fwd(const void* x_act, const float* w128, const float* bias1, float* emb, int batch, int T,
                 int x_f16, void* stream);
/* Backward: d_emb [B][T] -> dw128 += , dbias1 +=, dy act [G][1][T][8][128] = gradient at cv8's
 * pre-activation (x is cv8's ReLU/dropout output, so dy = dz * w * (x > 0) * out_scale). */
int zns_head_bwd(const void* x_act, const float* emb, const float* d_emb, const float* w128, float* dw128,
                 float* dbias1, void* dy_act, int batch, int T, float out_scale, int x_f16, void* stream);

/* Down_CNN merge (models.py:144-148): mode 0 = maximum, 1 = mean. */
int zns_merge(const float* a, const float* b, float* out, long long n, int mode, void* stream);

/* act (fp16 if act_f16 else bf16) [G][H][W][8][C] <-> fp32 NCHW [B][C][H][W] (module boundaries, tests). */
int zns_act_from_nchw(const float* x, void* act, int batch, int C, int H, int W, int act_f16, void* stream);
int zns_act_to_nchw(const void* act, float* x, int batch, int C, int H, int W, int act_f16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched variants: the anchor and the positive encoder (models.py:114-124) have identical geometry, so one launch
 * serves both (`n_br` = 1 or 2, arrays of per-branch pointers; branch b draws its dropout mask from rng_stream + b), and
 * one launch packs / unpacks every convolution weight of both encoders.  The single-tensor entry points above are the
 * n_br = 1 / n = 1 cases of these.
 * ---------------------------------------------------------------------------------------- */
int zns_conv1_fwd_nbr(int n_br, const float* const* x, long long x_clip_stride, long long x_row_stride,
                      const float* const* weight, const float* const* bias, void* const* out_act, int batch, int H, int W,
                      float dropout_p, uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream, int out_f16,
                      void* const* out_act_bf16, void* stream);
int zns_conv1_wgrad_nbr(int n_br, const void* const* dy_act, const float* const* x, long long x_clip_stride,
                        long long x_row_stride, float* const* dw, float* const* db, int batch, int H, int W, void* stream);
int zns_pool_fwd_nbr(int n_br, const void* const* y_act, void* const* out_act, int batch, int H, int W, int C, int pool,
                     float dropout_p, uint32_t seed, const uint32_t* seed_dev, uint32_t rng_stream, int act_f16,
                     void* const* out_act_bf16, void* stream);
int zns_pool_bwd_nbr(int n_br, const void* const* y_act, const void* const* dpool_act, void* const* dy_act, int batch, int H,
                     int W, int C, int pool, int y_f16, void* stream);
int zns_head_fwd_nbr(int n_br, const void* const* x_act, const float* const* w128, const float* const* bias1,
                     float* const* emb, int batch, int T, int x_f16, void* stream);
int zns_head_bwd_nbr(int n_br, const void* const* x_act, const float* const* emb, const float* const* d_emb,
                     const float* const* w128, float* const* dw128, float* const* dbias1, void* const* dy_act, int batch, int T,
                     float out_scale, int x_f16, void* stream);
/* Convolution with the pooling block that follows it (models.py:41-44,50-53) fused into the epilogue:
 * conv + bias -> MaxPool2d((pool, 1)) -> ReLU -> Dropout(d->dropout_p, d->rng_stream + branch).  d describes the convolution
 * (d->relu must be 0: the reference applies ReLU after the pool); the pre-pool tensor is never written.
 *   out_pooled[br]       act [G][H/pool][W][8][c_out] (fp16 / bf16 per d->fmt)
 *   out_pooled_bf16      NULL, or per branch a bf16 copy (x operand of the next weight gradient)
 *   argmax               NULL, or per branch uint8 [G][H/pool][W][8][c_out]: row (0..pool-1) of the first maximum of each
 *                        window, the routing table zns_pool_bwd_arg_nbr consumes
 * A tile holds whole pool windows in its TMEM accumulators: c_out = 64 (stacked kernel, pool 3 -> six rows, pool 4 / 8 ->
 * four / eight rows), c_out = 128 (pool <= 4); other shapes return ZNS_ERR_INVALID (use zns_conv_fwd + zns_pool_fwd). */
int zns_conv_pool_fwd(const zns_conv_desc* d, int pool, int n_br, const void* const* in, const void* const* wpk,
                      const float* const* bias, void* const* out_pooled, void* const* out_pooled_bf16, void* const* argmax,
                      void* stream);
/* Backward of the fused pooling: dy act bf16 [G][H][W][8][C] = dpool routed to the arg-max row of each window, zeros
 * elsewhere (dpool is already masked and scaled by the data-gradient epilogue of the layer above). */
int zns_pool_bwd_arg_nbr(int n_br, const void* const* argmax, const void* const* dpool_act, void* const* dy_act, int batch,
                         int H, int W, int C, int pool, void* stream);
int zns_bias_grad_nbr(int n_br, const void* const* dy_act, int batch, int H, int W, int C, float* const* db, void* stream);
/* n <= 16 weights: w[i] fp32 [c_out[i]][c_in[i]][kh[i]][kw[i]] -> wf[i] / wd[i] as zns_pack_weights (entries or whole
 * arrays may be NULL). */
int zns_pack_weights_multi(int n, const float* const* w, const int* c_out, const int* c_in, const int* kh, const int* kw,
                           void* const* wf, void* const* wd, int wf_f16, void* stream);
/* n <= 32 packed gradients -> state_dict layout as zns_unpack_grads; with zero_packed the packed accumulators are
 * cleared behind the read, so the next step's atomics start from zero without a memset launch. */
int zns_unpack_grads_multi(int n, float* const* gpk, const int* c_out, const int* c_in, const int* kh, const int* kw,
                           float scale, int accumulate, int zero_packed, float* const* g, void* stream);
/* cudaMemsetAsync(p, 0, bytes): a memset node in a captured graph instead of a fill kernel. */
int zns_zero(void* p, long long bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * NT-Xent (loss_functions.py:24-55), forward and backward in one launch.
 * anchors/poss fp32 [n_rows][dim]; result[0..2] = loss, mean cos(a_i,p_i), mean_i mean_{j!=i}
 * cos(a_i,p_j), all divided by batch_len as the reference does (rows beyond n_rows count as zero
 * loss, loss_functions.py:30).  d_anchors/d_poss (NULL to skip) = d loss / d input.
 * ---------------------------------------------------------------------------------------- */
int zns_ntxent_fwd_bwd(const float* anchors, const float* poss, int n_rows, int dim, int batch_len,
                       float temperature, float* result3, float* d_anchors, float* d_poss, void* stream);

/* torch.nn.BCELoss (mean reduction) of the downstream beat head (loader.py:20, epochs.py:52-54), forward and backward in one
 * launch: out / target fp32 [n] in (0, 1) / {0, 1}; *loss = mean BCE with torch's clamps; d_out (NULL to skip) = d loss / d out. */
int zns_bce_fwd_bwd(const float* out, const float* target, long long n, float* loss, float* d_out, void* stream);

/* torch.optim.Adam defaults (pretext.py:202) over flat fp32 buffers; `step` counts from 1 and is
 * read from the device word step_dev when that is not NULL (CUDA-graph replays); gradients are
 * multiplied by grad_scale first (1/world_size after a sum all-reduce). */
int zns_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int step, const uint32_t* step_dev, float grad_scale, void* stream);
/* Data-parallel variant: fused gradient reduce-scatter + Adam + parameter all-gather over NVLink
 * peer memory.  g_peers[r] / p_peers[r] (HOST arrays of `world` device pointers) address rank r's flat
 * gradient / parameter buffer (symmetric memory mapped into this process); this rank reduces and
 * updates its own 1/world shard and stores the new parameters into every rank's buffer.  m, v are
 * local (only the owned shard is touched).  Gradients are averaged (1/world).  The caller must issue a
 * cross-rank barrier before (all gradients complete) and after (all parameters visible). */
int zns_adam_p2p(int world, int rank, const void* const* g_peers, void* const* p_peers, float* m, float* v, long long n,
                 float lr, float beta1, float beta2, float eps, int step, const uint32_t* step_dev, void* stream);
/* *ctr += inc on the stream (the device step counter used above). */
int zns_counter_add(uint32_t* ctr, uint32_t inc, void* stream);

/* ------------------------------------------------------------------------------------------
 * Test-only SIMT reference convolution on the same act layout (fp32 accumulate, no tensor
 * cores): used by tests to localise errors of the tcgen05 kernels at sizes the CPU oracle
 * cannot reach.  Not called by the product path.
 * ---------------------------------------------------------------------------------------- */
int zns_dbg_conv_fwd_simt(const zns_conv_desc* d, const void* in, const void* wpk, const float* bias,
                          const void* mask, void* out, void* stream);
int zns_dbg_conv_wgrad_simt(const zns_conv_desc* d, const void* x, const void* dy, float* dwpk, void* stream);
/* Raw tcgen05 probe (csrc/dbg_probe.cu): a verbatim shared-memory image plus host-built matrix descriptors; returns the
 * accumulator columns (reps == 1) and the cycles the MMA list took.  tools/umma_view_probe.py drives it. */
int zns_dbg_umma_raw(const void* image, int image_bytes, const uint64_t* a_desc, const uint64_t* b_desc,
                     const uint32_t* d_col, const uint32_t* acc, const uint32_t* idescs, int n_mma, int n_cols_out,
                     float* d_out, int reps, long long* cycles, void* stream);
/* Host-only (no CUDA call): the tcgen05 level plan of one VQT pyramid level (struct VqtLevelDev of csrc/vqt_plan.h, raw
 * bytes) and its fp16 coefficient image, so that the MMA list can be replayed in numpy against the oracle
 * (tests/test_vqt_level_plan.py). */
int zns_dbg_vqt_level_plan(int sr, int hop, int n_bins, int bins_per_octave, double fmin, double gamma, int level,
                           void* level_struct, int struct_bytes, uint16_t* coef_image, int coef_halfwords);
/* Diagnostic: device buffer (16 int64 per pyramid level) that CTA 0 of every VQT level kernel fills with per-role cycle
 * counters (issuer / epilogue / loader: total and waiting); NULL switches it off.  tools/vqt_bench.py --timing. */
int zns_dbg_vqt_timing(long long* buf);
/* tcgen05 GEMM probe with hand-swizzled operands (descriptor self-test); csrc/conv_umma.cu. */
int zns_dbg_umma_probe(int variant, const void* a, const void* b, float* d, int n, int k, void* stream);
/* tcgen05 issue/throughput microbenchmark (diagnostic): cycles[n_ctas] per CTA. */
int zns_dbg_umma_rate(int n, int iters, int per_group, int mode, int n_ctas, long long* cycles, void* stream);

/* Host-only (no CUDA call): the launch geometry zns_conv_fwd / zns_conv_wgrad would choose for a layer, so that tile
 * plans and the CTA-pair work-item table can be checked on a machine without a GPU (tests/test_lib_host.py).
 *   fwd  out[14]: kernel (0 direct, 1 stacked), N, CTAs per cluster, hb, nb, hs, ns, n_cols, n_total, A-row slots,
 *                 weight stages, grid x, dynamic shared memory bytes, units per column
 *   wgrad out[16]: NB, CTAs per cluster, position slices, accumulators per CTA, tap groups per row, grp_base, grp_rem,
 *                 row items, stack_dy, fold, cin blocks, cout blocks, item pairs, stages, grid x, dynamic smem bytes;
 *         items[2 * item pairs] (may be NULL): r | s0 << 4 | n_acc << 10 | cin block << 14 | valid << 18 */
int zns_dbg_conv_fwd_plan(const zns_conv_desc* d, int n_br, int* out);
int zns_dbg_conv_wgrad_plan(const zns_conv_desc* d, int n_br, int* out, unsigned int* items);

#ifdef __cplusplus
}
#endif
#endif /* ZNS_H_ */
