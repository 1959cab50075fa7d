/*
 * avcer_b200 -- C ABI of the B200-native (sm_100a) kernels behind AVCER's batched
 * inference-and-fusion path.
 *
 * The reference (ElenaRyumina/AVCER) is pure Python/PyTorch and has no FFI layer; its boundary
 * is the Python function surface.  The Python package `avcer_b200`
 * keeps those entry points and calls the functions below through ctypes.  Every entry point
 * names the reference call site (file:line under src/) whose device work it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - the caller owns every buffer (no allocation, no hidden synchronisation inside);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = ok, non-zero = error; the message is available from avcer_last_error();
 *   - activations are channels-last ("NHWC" / [B, T, C]); dtype codes: AVCER_BF16 or AVCER_F32.
 */
#ifndef AVCER_B200_H_
#define AVCER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVCER_F32 0
#define AVCER_BF16 1

#define AVCER_ACT_NONE 0
#define AVCER_ACT_RELU 1
#define AVCER_ACT_GELU 2 /* exact erf GELU (HF wav2vec2 "gelu") */

const char* avcer_last_error(void);
int avcer_version(void);
/* "bf16" (libavcer_b200.so) or "fp16" (libavcer_b200_fp16.so): the 16-bit storage type this build of the library gives to
 * dtype code AVCER_BF16 -- same kernels, same entry points; the fp16 build keeps 11 mantissa bits instead of 8. */
const char* avcer_storage_type(void);
/* 0 when a CUDA device of compute capability 10.x is usable; error otherwise (no CPU fallback). */
int avcer_device_check(void);
int avcer_num_sms(void);
/* Grid share of the persistent kernels launched (or captured into a CUDA graph) by the calling thread from now on:
 * n_sms SMs (even, > 0) instead of the whole device, 0 = whole device.  Lets the VS branch (HBM-bound early layers) and
 * the audio branch (tensor-bound) of one pipeline step run side by side on two streams without one kernel's CTAs
 * occupying every SM.  The reference runs the two branches one after the other (run.py:224-268). */
int avcer_set_sm_limit(int n_sms);
/* Development aid: a device buffer of 4 CTAs x 64 tiles x 16 slots of uint64 that the two-SM contraction kernel fills
 * with per-tile clock64 stamps of its first CTAs (producer / MMA / epilogue phases; scripts/trace_gemm2.py renders them);
 * NULL (the default) disables tracing.  Not used on the product path. */
int avcer_debug_set_trace(void* buf);

/* ------------------------------------------------------------------------------------------
 * K1  face-crop preprocessing.
 * Replaces data/utils.py:19-39 (pth_processing: PIL NEAREST resize to 224x224, RGB->BGR flip,
 * per-channel mean subtraction, no /255) + the H2D copy at get_prob_video.py:108.
 * src: u8 crops, HWC with 3 channels in the order cv2.imread delivers (BGR); crop i starts at
 *      src + src_offsets[i] and is src_h[i] x src_w[i] (row pitch src_w*3).
 * dst layout 0: fp32 NCHW [n,3,224,224] (bit-exact restatement of the reference tensor);
 * dst layout 1: bf16 zero-bordered NHWC4 [n, 232, 240, 4] (row pitch 240 px) with the image at rows/cols 2..225
 *               (TF-"same" padding 2|3 of the stem, architectures/video.py:63-90, materialised
 *               once; channel 3 is zero) -- the layout the tensor-core stem consumes;
 * dst layout 2: fp32, same geometry as layout 1 (fp32 mode).
 * The border of layouts 1/2 must have been zeroed once by the caller.
 */
int avcer_preprocess_u8(const uint8_t* src, const int64_t* src_offsets, const int32_t* src_h,
                        const int32_t* src_w, const int16_t* maps, int n, void* dst, int dst_layout,
                        void* stream);
/* Nearest-neighbour source-index tables of Pillow's resize for every crop: maps[n][2][224]
 * (y table, then x table).  src_offsets/src_h/src_w/maps may all be NULL in avcer_preprocess_u8
 * when every crop is a packed 224x224 image (resize = identity). */
int avcer_preprocess_maps(const int32_t* src_h, const int32_t* src_w, int n, int16_t* maps,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Generic dense contraction (implicit GEMM): every convolution and Linear on the path.
 * Replaces the cuDNN/cuBLAS calls issued by architectures/video.py:46-58,116,124-133,
 * HF Wav2Vec2 conv/linear layers and attention_layers.py:92-97,45-55.
 *
 *   out[w,h,n, co] = act( sum_{ty,tx,c} A[c + g(co), w+off_w+tx, h+off_h+ty, n] * Wt[co, (ty*taps_w+tx)*cin + c]
 *                         + bias[co] + residual[w,h,n,co] )
 *
 * A is described as a rank-5 strided view (c, w, h, n, t) with c contiguous; reads outside
 * [0,a_dim) return 0 (this is the zero padding).  If tap_h_in_dim4 != 0 the ty tap indexes
 * dimension 4 instead of shifting h (strided / dilated row taps).
 * dtype AVCER_BF16: A, Wt, residual bf16; bias fp32; tcgen05 tensor-core kernel (TMA + TMEM).
 *                   cin must be a multiple of 32 and Cout a multiple of 32.
 * dtype AVCER_F32 : everything fp32; SIMT kernel (the "fp32 mode" of the north star).
 */
typedef struct {
  const void* a;
  int64_t a_dim[5];          /* extents of (c, w, h, n, t) */
  int64_t a_stride[5];       /* element strides; a_stride[0] must be 1 */
  const void* wt;            /* [cout, taps_h*taps_w*cin] row-major */
  const float* bias;         /* [cout] or NULL */
  const void* residual;      /* or NULL */
  void* out;
  int64_t out_stride[3];     /* element strides of (w, h, n) in out; channels contiguous */
  int64_t res_stride[3];
  int32_t W, H, NB;          /* output extents */
  int32_t cin, cout;         /* cin = contraction channels per tap */
  int32_t taps_w, taps_h, off_w, off_h, tap_h_in_dim4;
  int32_t group_cin_shift;   /* grouped conv: channel shift per 64 output channels, else 0 */
  int32_t a_strip;           /* bf16 only. 1: a_dim = (e, chunks, h, n, taps_h) describes, for every (h, n, tap), ONE
                              * contiguous strip of e*chunks elements; output position w reads the cin elements that
                              * start 16 bytes * w into the strip (overlapping windows: a strided conv row).  The
                              * strip is fetched once per tap; the overlap lives in the MMA descriptor (no im2col). */
  const void* wt_packed;     /* strip mode: weights per K chunk in UMMA core-matrix order [k/8][cout/8][8][8] */
  int32_t act;
  int32_t res_after_act;     /* 0: act(acc+bias+res); 1: act(acc+bias)+res */
  int32_t dtype;             /* AVCER_BF16 / AVCER_F32 (operands) */
  int32_t out_f32;           /* bf16 path only: write fp32 output */
  int32_t a_step;            /* bf16 path only. s > 1: a strided spatial convolution -- a_dim / a_stride describe the
                              * FULL-resolution input (c, w_in, h_in, n, 1) and output pixel (w, h) reads input pixel
                              * (s*w + off_w + tap_w, s*h + off_h + tap_h): the TMA box walks the input with traversal
                              * stride s (a 3x3 "same" conv evaluated only at every s-th pixel).  0 / 1: unit step. */
  int32_t reverse_tiles;     /* bf16 path only. 1: the persistent CTAs walk the output tiles from the LAST to the first.  A
                              * consumer that runs opposite to its producer starts on the rows the producer wrote last, i.e.
                              * the ones still in L2 (tensors larger than the 126 MB L2 are otherwise re-read from HBM in
                              * full).  Same tiles, same arithmetic: results are bit-identical. */
} avcer_contract_desc;

int avcer_contract(const avcer_contract_desc* d, void* stream);
/* Kernel the calling thread's last avcer_contract dispatched to (measurement aid: per-kernel attribution of launch timings).
 * 1 SIMT fp32; 2 / 3 / 4 single-CTA tcgen05 tiles 128x64 / 128x128 / 128x256; 5 fp32-output direct epilogue; 6 / 7 two-SM
 * 256x256 ring epilogue without / with residual; 8 / 9 two-SM 256x256 FLAT epilogue without / with residual; 10 two-SM
 * 256x128; 11 / 12 halo 3x3 conv with 64 / 128 output channels; 13 strip-mode stem. */
int avcer_last_contract_kernel(void);

/* ------------------------------------------------------------------------------------------
 * K4  probability fusion + compound-expression rule + argmax, one warp-cooperative pass.
 * Replaces run.py:105-165 and data/utils.py:222-241 (numpy float64, CPU).
 * p_vs: [n,7] f32 VS probabilities (audio emotion order), p_vd / p_a: [n,7] f32 probabilities.
 * w1: [3][7] f64 per-class weights or NULL (=> mean of the three, run.py:115-116),
 * w2: [3] f64 per-model weights.  ce_weights_type / ce_mask as in run.py:31-32.
 * labels: [4][label_pitch] int64 (AV, VS, VD, A), frames 0..n-1 of each row written; label_pitch = 0 means n
 * (a dense [4][n] array).  A pitch > n lets a shard's labels land directly in its slot of a [4][all frames] buffer
 * (multi-GPU: every rank fuses the all-gathered per-frame rows block by block, no concatenation pass).
 * Arithmetic is IEEE double, left-to-right, no FMA contraction; argmax returns the first maximum (numpy
 * semantics, NaN counts as maximum).
 */
int avcer_fuse_compound(const float* p_vs, const float* p_vd, const float* p_a, int64_t n,
                        const double* w1_host, const double* w2_host, int ce_weights_type,
                        int ce_mask, int64_t* labels, int64_t label_pitch, void* stream);
/* Same with float64 probability inputs (the reference DataFrames become float64 when a zero row
 * was appended, get_prob_video.py:89,160-178); arithmetic is then float64 in every branch. */
int avcer_fuse_compound_f64(const double* p_vs, const double* p_vd, const double* p_a, int64_t n,
                            const double* w1_host, const double* w2_host, int ce_weights_type,
                            int ce_mask, int64_t* labels, int64_t label_pitch, void* stream);

/* data/utils.py:222-241 (get_compound_expression) as a stand-alone op: pred [n,ncols] (f32 or f64),
 * k pairs (i1,i2) with weights (w1,w2) given on the host, optional 1/7 mask; out [n,k] f64. */
int avcer_compound_scores(const void* pred, int64_t n, int ncols, int pred_f64,
                          const int32_t* pairs_host, const double* w_host, int k, int ce_mask,
                          double* out, void* stream);

/* Weight search (data/utils.py:138-209: get_weights_prob_model / get_weights_v_model / get_weights_av_model).
 * For each of n_weights candidate weight sets weights[w][m][c] the fused prediction
 * argmax_c sum_m preds[m][f][c] * weights[w][m][c] (binary64, left to right, first maximum) is compared with
 * gt[f]; cm[w][gt][pred] (7x7, caller-zeroed, uint64) receives the confusion counts from which the host
 * derives sklearn's precision / recall / F1 exactly.  preds: [n_models][n][7] f64, n_models 2 or 3. */
int avcer_weight_search_confusion(const double* preds, int n_models, int64_t n, const int32_t* gt,
                                  const double* weights, int64_t n_weights, uint64_t* cm, void* stream);

/* Weighted fusion + arg-max over the 7 basic emotions (get_pred_av.py:34-40, get_metrics): labels[f] =
 * argmax_c sum_m preds[m][f][c] * w1[m][c] * w2[m], float64, products and sum left to right, numpy arg-max semantics.
 * preds: [n_models, n, 7] float64; w1: [n_models, 7] and w2: [n_models] DEVICE float64; labels: [n] int32. */
int avcer_fused_argmax(const double* preds, int n_models, int64_t n, const double* w1, const double* w2,
                       int32_t* labels, void* stream);

/* Row softmax over 7 classes in fp32, exactly data/utils.py:125-127 (max-subtract, exp, sum, div).
 * `ld` = row pitch of the input in floats (8 for the 8-class audio logits: "Other" is dropped
 * before the softmax, run.py:96). */
int avcer_softmax7(const float* x, int64_t n, int ld, float* y, void* stream);
int avcer_softmax7_f64(const double* x, int64_t n, int ld, double* y, void* stream);

/* Frame alignment (run.py:90-103 + get_prob_audio_8_cl.py:94-101): per-frame mean over all audio
 * windows whose frame range [f_lo[w], f_hi[w]) covers the frame, NaN windows skipped (pandas
 * groupby.mean skipna), accumulated in window order in fp32... see DESIGN.md for the rounding
 * contract.  logits: [n_win, ncls]; out: [n_frames, ncls]; frames never covered get NaN. */
int avcer_window_to_frame_mean(const float* logits, int n_win, int ncls, const int32_t* f_lo,
                               const int32_t* f_hi, int64_t n_frames, float* out, void* stream);
/* Same in float64: pandas accumulates in the column dtype, and the tables get_pred_av.py:246-249 reads back from CSV
 * are float64. */
int avcer_window_to_frame_mean_f64(const double* logits, int n_win, int ncls, const int32_t* f_lo,
                                   const int32_t* f_hi, int64_t n_frames, double* out, void* stream);

/* Gather rows: out[i,:] = src_index[i] >= 0 ? src[src_index[i], :] : 0   (carry-forward / gap
 * expansion of get_prob_video.py:157-178, and column permutation to audio order run.py:85-88
 * when `perm` is non-NULL: out[i, j] = src[idx, perm[j]]). */
int avcer_gather_rows(const float* src, const int32_t* src_index, int64_t n_out, int ncols,
                      const int32_t* perm, float* out, void* stream);
int avcer_gather_rows_f64(const double* src, const int32_t* src_index, int64_t n_out, int ncols,
                          const int32_t* perm, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small layers of the VS / VD / A networks (channels-last).  `dtype` selects bf16 or fp32 storage.
 */
/* ResNet-50 stem fused with its max-pool (bf16 only): Conv2dSame 7x7/2 with TF-"same" padding 2|3
 * (architectures/video.py:63-90, :98-100) + folded BatchNorm (eps 1e-3) + ReLU (:116) + MaxPool2d(3, 2) without
 * padding (:103, :117).  x_padded: zero-bordered NHWC4 crops [n, 232, 240, 4] (layout 1 of avcer_preprocess_u8);
 * w_packed: the folded filter bank per filter row in UMMA core-matrix order [7][32/8][64/8][8][8] bf16 (28 KB);
 * bias: [64] fp32; out: [n, 55, 55, 64] bf16.  Bit-identical to avcer_contract (stem geometry) followed by
 * avcer_maxpool3x3s2; the 112x112x64 stem activation is never written to global memory. */
int avcer_stem_pool(const void* x_padded, const void* w_packed, const float* bias, int n, void* out, void* stream);
/* Same with `out_pitch` (>= 64, multiple of 8) bf16 elements between output pixels: the pooled activation can be
 * written into the first 64 columns of a wider [n*55*55, out_pitch] matrix (used to place the block input next to
 * conv2's output so that conv3 and the projection shortcut, video.py:46-58, become one K-concatenated GEMM). */
int avcer_stem_pool_ld(const void* x_padded, const void* w_packed, const float* bias, int n, void* out,
                       int64_t out_pitch, void* stream);
/* K1 + stem + pool in ONE kernel for packed 224x224 crops (resize = identity; BASELINE configs): crops uint8 [n,224,224,3]
 * in cv2's BGR order are converted on the way into shared memory with exactly K1's arithmetic (float(px) - mean ->
 * bf16, data/utils.py:19-39), so the 442 KB per crop of avcer_preprocess_u8 layout 1 are neither written nor re-read.
 * Output identical to avcer_preprocess_u8(layout 1) + avcer_stem_pool_ld, bit for bit. */
int avcer_stem_pool_u8(const uint8_t* crops, const void* w_packed, const float* bias, int n, void* out,
                       int64_t out_pitch, void* stream);
/* Every stride-th pixel of an NHWC tensor as rows of a [n*ho*wo, y_pitch] matrix (ho = (h-1)/stride+1): the input
 * sampling of the stride-2 1x1 convolutions of layer2-4's first blocks (architectures/video.py:13-15, 141-148), done once
 * for conv1 and the projection shortcut; y_pitch >= c leaves room for conv2's output next to it (K-concatenated conv3). */
int avcer_subsample_rows(const void* x, int n, int h, int w, int c, int stride, void* y, int64_t y_pitch,
                         int dtype, void* stream);
/* 3x3 stride-2 un-padded max pool, NHWC (architectures/video.py:103,117). */
int avcer_maxpool3x3s2(const void* x, int n, int h, int w, int c, void* y, int dtype, void* stream);
/* Global average pool NHWC -> [n, c] (video.py:124). */
int avcer_avgpool(const void* x, int n, int hw, int c, void* y, int dtype, void* stream);
/* Tiny Linear (+ optional softmax) with fp32 weights and fp32 output: fc2 + F.softmax
 * (video.py:133 + get_prob_video.py:107), LSTM fc (video.py:184), feature_downsample
 * (audio_8_cl.py:189).  x: [n, k] (dtype, row pitch ldx), w: [m, k] f32, b: [m] f32, y: [n, m] f32; m <= 8. */
int avcer_small_linear(const void* x, int64_t n, int k, int64_t ldx, const float* w, const float* b,
                       int m, int softmax, float* y, int dtype, void* stream);
/* LSTM cell pointwise step (PyTorch gate order i,f,g,o; video.py:169-185):
 * gates = xproj[xidx[r]] + hproj[r] (both [*, 4H] fp32, biases already folded into xproj);
 * c,h updated in place; c is fp32 [n,H]; h_out (dtype) feeds the next recurrent GEMM.
 * hproj may be NULL for the first step (h_{-1} = 0), xproj may be NULL when the input
 * projection is folded into hproj; ldh = row pitch of h_out in elements.
 * split != 0 (bf16 only): h is written as the bf16x3 operand [hi | lo | hi] (3*H columns; lo = bf16(h - hi)) that,
 * against weights packed [w_hi | w_hi | w_lo], gives hi*w_hi + lo*w_hi + hi*w_lo: the recurrence keeps 16 mantissa
 * bits through the bf16 tensor cores.  h_f32 (optional, [n,H]) also receives h in fp32 (input of the final Linear). */
int avcer_lstm_cell(const float* xproj, const int32_t* xidx, const float* hproj, float* c,
                    void* h_out, int64_t ldh, int64_t n, int hidden, int first, int split, float* h_f32,
                    int dtype, void* stream);
/* GRU cell pointwise step (PyTorch gate order r,z,n): the recurrent layers of ExprModelV1 (architectures/audio_8_cl.py:23-29,
 * forward :63).  xg: [*, 3H] fp32 input projections W_ih x + b_ih of ALL time steps, the row of sequence r at this step is
 * x_row0 + r * x_row_stride; hg: [n, 3H] fp32 = W_hh h_{t-1} + b_hh; h_state [n, H] fp32 is updated in place;
 * h_out (dtype, row pitch ldh) feeds the next recurrent GEMM (split != 0: bf16x3 operand [hi | lo | hi], as avcer_lstm_cell);
 * y (optional, dtype, [*, H]) receives h_t at row y_row0 + r * y_row_stride (the layer's output sequence). */
int avcer_gru_cell(const float* xg, int64_t x_row0, int64_t x_row_stride, const float* hg, float* h_state,
                   void* h_out, int64_t ldh, int split, void* y, int64_t y_row0, int64_t y_row_stride,
                   int64_t n, int hidden, int dtype, void* stream);
/* x fp32 [rows, k] (row pitch ldx) -> bf16 [rows, 3k] (row pitch ldo) = [hi | lo | hi]: the bf16x3 split of the
 * relu(fc1) features feeding the LSTM input projection (get_prob_video.py:115-122). */
int avcer_split_bf16x3(const float* x, int64_t rows, int k, int64_t ldx, void* out, int64_t ldo, void* stream);

/* Audio decode seam (data/utils.py:49-60, convert_mp4_to_mp3 after its ffmpeg call): interleaved int16 PCM
 * [n, channels] -> x/32768 -> channel mean -> torchaudio.transforms.Resample (default "sinc_interp_hann" polyphase
 * filter: out[f*nnew + j] = sum_k bank[j][k] * x[f*orig + k - width], zero outside the signal).
 * bank_tap_major: [2*width + orig][nnew] fp32 (the transposed filter bank), orig/nnew = the two rates divided by
 * their gcd; n_out = ceil(nnew*n/orig).  bank_tap_major == NULL: same rate, only the scale + channel mean (n_out == n). */
int avcer_pcm16_resample(const int16_t* pcm, int64_t n, int channels, const float* bank_tap_major,
                         int orig, int nnew, int width, float* out, int64_t n_out, void* stream);
/* K5a: gather audio chunks wav[starts[i], ends[i]) (at most `win` samples; chunks of several clips
 * may live in one concatenated buffer), pad the tail to `win` with the chunk mean ("mean"), zeros
 * ("constant") or by tiling ("repeat") (data/utils.py:63-89), then HF
 * zero-mean/unit-variance normalisation over all `win` samples (population variance, eps 1e-7).
 * out: [n_win, win] fp32.  pad_mode: 0 mean, 1 constant, 2 repeat. */
int avcer_audio_normalize_windows(const float* wav, const int64_t* starts, const int64_t* ends,
                                  int n_win, int win, int pad_mode, float* out, void* stream);
/* K5b: wav2vec2 conv layer 0 (Cin=1, k=10, s=5, bias) + LayerNorm(512) + GELU fused.
 * x: [n, t_in] fp32; y: [n, t_out_pitch, 512] (dtype), t_out = (t_in-10)/5+1 rows written. */
int avcer_w2v_conv0_ln_gelu(const float* x, int n, int t_in, const float* w, const float* b,
                            const float* ln_g, const float* ln_b, void* y, int64_t y_pitch_rows,
                            int dtype, void* stream);
/* K5b on the tensor cores (16-bit storage only): the same layer as a 128 x 512 x 32 tcgen05 contraction per 128 time steps
 * (16-bit x3 split of samples, filters and bias; fp32 accumulation) with LayerNorm + GELU applied straight from TMEM.
 * w_packed: [4][64][8][8] 16-bit core matrices of rows [w_hi | w_hi | w_lo | b_hi | b_lo] (weights.pack_conv0_tc).
 * Replaces the same reference call as avcer_w2v_conv0_ln_gelu (HF Wav2Vec2LayerNormConvLayer #0 behind
 * architectures/audio_8_cl.py:135,180). */
int avcer_w2v_conv0_tc(const float* x, int n, int t_in, const void* w_packed, const float* ln_g,
                       const float* ln_b, float eps, void* y, int64_t y_pitch_rows, void* stream);
/* Row LayerNorm over `c` channels (eps given) with optional fused GELU and optional additive
 * term (positional encoding / residual) applied BEFORE the norm: y = act(LN(x + add)).
 * Rows are addressed as x + r*ldx; add row index = r % add_rows (add_rows = 0: no add). */
int avcer_layernorm(const void* x, int64_t rows, int c, int64_t ldx, const void* add,
                    int64_t add_rows, const float* g, const float* b, float eps, int act, void* y,
                    int64_t ldy, int dtype, void* stream);
/* y = x + add[r % add_rows]  (sinusoidal positional encoding, attention_layers.py:216). */
int avcer_add_rows(const void* x, int64_t rows, int c, const void* add, int64_t add_rows, void* y,
                   int dtype, void* stream);
/* Multi-head self-attention over T tokens (no mask): softmax(Q K^T * scale) V.
 * qkv: [n, T, 3, heads, dh] (dtype) packed projections; out: [n, T, heads*dh]. */
int avcer_attention(const void* qkv, int n, int t, int heads, int dh, float scale, void* out,
                    int dtype, void* stream);
/* Audio head tail (audio_8_cl.py:146-159): MaxPool1d(5)+ReLU after conv/BN, and
 * AdaptiveAvgPool1d(1)+ReLU; x: [n, t, c] -> y: [n, t/5, c] resp. [n, c]. */
int avcer_maxpool1d5_relu(const void* x, int n, int t, int c, void* y, int dtype, void* stream);
int avcer_avgpool1d_relu(const void* x, int n, int t, int c, void* y, int dtype, void* stream);
/* dtype conversion helpers (fp32 <-> bf16). */
int avcer_cast(const void* x, int64_t n, int src_dtype, void* y, int dst_dtype, void* stream);


/* ------------------------------------------------------------------------------------------
 * Baseline-JPEG decode of the face crops: replaces cv2.imread at get_prob_video.py:95 for
 * the files data/get_face_images.py:60 writes with cv2.imwrite defaults (baseline sequential DCT, 8 bit, YCbCr 4:2:0 or
 * 4:4:4, no restart markers).  Bit-identical to cv2.imread: T.81 Huffman decoding, libjpeg-turbo's jpeg_idct_islow,
 * h2v2 fancy up-sampling and ycc_rgb_convert restated as integer kernels.
 * The HOST walks the marker segments (SOF0 / DQT / DHT / SOS) and fills, per image: */
typedef struct {
  int64_t data_off;   /* the image's entropy-coded segment starts at raw + data_off + src_shift; data_off is a multiple of 4
                       * (its un-stuffed copy is written to data + data_off) */
  int64_t data_len;   /* its length in bytes */
  int64_t coef_off;   /* first 8x8 block of the image in `coefs` (all Y blocks row-major, then Cb, then Cr) */
  int64_t plane_off;  /* byte offset of the image's Y sample plane in `planes` (Cb and Cr planes follow); multiple of 8 */
  int64_t out_off;    /* byte offset of the decoded BGR image in `out` */
  int32_t width, height;
  int32_t mcus_w, mcus_h;  /* MCU grid: ceil(width / (8*hs)), ceil(height / (8*hs)) */
  int32_t hs;         /* luma sampling factor on both axes: 2 = 4:2:0, 1 = 4:4:4 */
  int32_t qt_y, qt_c; /* indices into qtables */
  int32_t src_shift;  /* 0..3, see data_off: lets whole files be uploaded as they are, header included */
} avcer_jpeg_image;
/* raw: all entropy-coded segments as they sit in the files (still byte-stuffed), image i at data_off / data_len;
 * huff_bits [4][16] / huff_vals [4][256]: the tables DC0, AC0, DC1, AC1 (luma, chroma) shared by the batch; qtables [*][64]
 * uint16 in natural (row-major) order; pixel_prefix [n]: index of every image's first pixel in the batch; data (same size
 * as raw), lens [n], coefs [total_blocks][64] int16 and planes (sum of MCU-padded component planes) are scratch;
 * out receives height x width x 3 bytes per image in B, G, R order; *status (device) is 0 on success, 1 + the index of an
 * image with a corrupt bit stream, or -(1 + index) for an image with restart markers (not covered). */
int avcer_jpeg_decode(const uint8_t* raw, const avcer_jpeg_image* images, int n, const uint8_t* huff_bits,
                      const uint8_t* huff_vals, const uint16_t* qtables, const int64_t* pixel_prefix,
                      int64_t total_blocks, int64_t total_pixels, uint8_t* data, int64_t* lens, int16_t* coefs,
                      uint8_t* planes, uint8_t* out, int32_t* status, void* stream);

/* Host-side staging of the crop files (no device work): reads n files with `threads` workers into one caller-owned
 * (normally pinned) buffer, file i at dst + offsets[i] (written here, 16-byte aligned), sizes[i] = its length; *needed =
 * total bytes required (also set when capacity is too small, so the caller can grow the buffer and call again).
 * Replaces the per-file read inside cv2.imread (get_prob_video.py:95) for the GPU-decoded path. */
int avcer_read_files(const char* const* paths, int n, uint8_t* dst, int64_t capacity, int64_t* offsets,
                     int64_t* sizes, int64_t* needed, int threads);

/* ---- Face detector: RetinaFace-ResNet50 behind data/get_face_images.py:38-63.  The 1x1 / 3x3
 * convolutions of body, FPN, SSH and heads are avcer_contract calls; these are the layers it does not cover. ---- */
/* Stem on raw video frames (retina_face_predictor.py:61-67 + torchvision resnet50 conv1/bn1/relu): frames [n,h,w,3] uint8
 * (B,G,R; rgb != 0: R,G,B), minus (104,117,123), conv 7x7 / 2 pad 3 with wt [147][64] fp32 (row = (ky*7+kx)*3 + c in BGR
 * order, BatchNorm folded) + bias [64], ReLU -> out [n, ceil(h/2), ceil(w/2), 64] (dtype). */
int avcer_det_stem(const uint8_t* frames, int n, int h, int w, int rgb, const float* wt, const float* bias,
                   void* out, int dtype, void* stream);
/* The same stem's input for the tensor-core path: frames -> zero-bordered NHWC4 [n, hp, wp, 4] in the library's 16-bit
 * storage type, pixel (y, x) at (y + 3, x + 3) as (B - 104, G - 117, R - 123, 0) (exact integers), zeros elsewhere; the
 * convolution itself is then a strip-mode avcer_contract (a_strip) per band of 128 output columns. hp >= h + 6, wp >= w + 6. */
int avcer_det_prepare(const uint8_t* frames, int n, int h, int w, int rgb, int hp, int wp, void* out, void* stream);
/* Max pool 3x3 / 2 pad 1 over NHWC (torchvision resnet50.maxpool): [n,h,w,c] -> [n, ceil(h/2), ceil(w/2), c]. */
int avcer_maxpool3x3s2p1(const void* x, int n, int h, int w, int c, void* y, int dtype, void* stream);
/* FPN top-down merge (retina_face_net.py:88-94): out[n,y,x,:] = a[n,y,x,:] + b[n, ymap[y], xmap[x], :] with a, out
 * [n,h,w,c], b [n,hb,wb,c]; ymap [h] / xmap [w] int32 = the source indices of F.interpolate(mode="nearest"). */
int avcer_upsample_add(const void* a, const void* b, int n, int h, int w, int hb, int wb, int c,
                       const int32_t* ymap, const int32_t* xmap, void* out, int dtype, void* stream);
/* Anchors (prior_box.py:17-33), 2-class softmax, box and landmark decoding (box_utils.py:210-249) and scaling to pixels
 * (retina_face_predictor.py:75-84) of the three head maps: heads_k [n * ceil(height/s_k) * ceil(width/s_k), head_pitch]
 * fp32 with s = 8, 16, 32 and columns [cls a0 (bg, face), cls a1 | box a0 (4), a1 | landmarks a0 (10), a1] ->
 * dets [n, P, 15] fp32 rows (x1, y1, x2, y2, score, 5 x (x, y)), priors in the reference's order. */
int avcer_det_decode(const float* heads0, const float* heads1, const float* heads2, int64_t head_pitch, int n,
                     int height, int width, float* dets, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVCER_B200_H_ */
