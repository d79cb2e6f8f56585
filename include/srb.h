/*
 * srb.h -- C ABI of libsrb.so, the sm_100a kernels behind speech_resynth_b200.
 *
 * The reference (misternasty/speech_resynth) is pure Python and has no FFI; every entry point below replaces a
 * PyTorch call site of the reference's unit-to-speech path (file:line relative to /root/reference, "HF:" =
 * transformers/models/fastspeech2_conformer/modeling_fastspeech2_conformer.py).  INTEGRATION.md shows the ctypes
 * binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - all pointers are DEVICE pointers borrowed from the caller (torch owns the memory); nothing is allocated
 *     (only exception: the tiny `kernel` / `dilation` int32 arrays of srb_hifigan_conv are HOST arrays);
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous and CUDA-graph capturable;
 *   - return value 0 = success, negative = error (srb_last_error() gives the text); nothing throws;
 *   - activations are channel-last: (batch, rows, channels) with explicit element strides, bf16 unless noted;
 *   - "packed" weights are bf16 [n][k] with k contiguous, k ordered (tap, channel) and the channel count padded to
 *     a multiple of the K block the op uses (see speech_resynth_b200/packing.py for the exact recipe);
 *   - `lengths` is int32[batch]: number of valid (non-pad) rows per utterance;
 *   - ordering: kernels are launched with programmatic dependent launch; each waits in-kernel for its predecessor on
 *     the stream before it touches activations, but may read weight / bias / table operands before that, so those
 *     must stay constant while calls are in flight (environment: SRB_PDL=0, SRB_WEIGHTS_EARLY=0 turn this off).
 */
#ifndef SRB_H_
#define SRB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRB_VERSION 100

int srb_version(void);
/* 1 = the product library (libsrb.so: plain bf16 activations).  3 = the tight-precision build of the same sources
 * (libsrb_tight.so, -DSRB_SPLIT): every bf16 ACTIVATION tensor named below is stored with three column blocks
 * [hi | lo | hi] (hi = bf16(x), lo = bf16(x - hi); 3 C columns for a logical width C, blocks C columns apart) and every
 * GEMM weight must be packed as [Wh | Wh | Wl] along K, so the unchanged bf16 tensor-core loop accumulates
 * xh Wh + xl Wh + xh Wl in fp32: results within ~1e-5 of the reference's fp32 path ("tight in fp32/TF32",
 * BASELINE.json north_star).  fp32 tensors are unaffected.  Entries without a tight form return an error there. */
int srb_split_factor(void);
const char* srb_last_error(void);
/* compute capability major*10+minor of the current device, or negative error */
int srb_device_arch(void);

/* ---- unit embedding gather: nn.Embedding forward, src/flow_matching/models.py:50-52,154 ----------------------
 * out[m, :] = table[ids[m], :]  (fp32 rows, bit exact).  ids int64 in [0, vocab]; dim % 4 == 0. */
int srb_embed_gather(const float* table, const int64_t* ids, float* out, int64_t m, int32_t vocab_rows, int32_t dim,
                     void* stream);

/* valid-frame count per utterance: mask = input_ids.ne(0) (models.py:152); lengths[b] = sum(mask[b]) */
int srb_unit_lengths(const int64_t* ids, int32_t* lengths, int32_t batch, int32_t frames, void* stream);
/* same, plus extents[b] = index of the last non-pad id + 1.  The kernels mask by prefix length, the reference by position
 * (models.py:152, transformer.py:115-127): both agree exactly when lengths[b] == extents[b] (right-padded rows, the
 * reference's input convention, synthesize.py:39-42); the host raises otherwise.  Either output may be pinned host memory. */
int srb_unit_extents(const int64_t* ids, int32_t* lengths, int32_t* extents, int32_t batch, int32_t frames, void* stream);

/* ---- time conditioning table: models.py:47-49,179 + norm.py:42 --------------------------------------------
 * For each ODE time t_s (s < nfe): c = SiLU(Linear([t, sin(2*pi*t*w), cos(2*pi*t*w)])) and, for each of the
 * n_norm AdaptiveRMSNorm layers, g[s][j][:] = sqrt(H) * (W_j c + 1).  All fp32, H = 256.
 * four_w[128], lin_w[256*257], lin_b[256], gamma_w[n_norm][256*256] -> time_emb[nfe][256], g[nfe][n_norm][256]. */
int srb_time_cond_table(const float* times, int32_t nfe, const float* four_w, const float* lin_w, const float* lin_b,
                        const float* gamma_w, int32_t n_norm, float* time_emb, float* g, void* stream);

/* rotary table: transformer.py:56-63.  cos/sin[(pos, f)] for pos < rows, f < 64 (fp32, freq = pos * inv_freq[f]) */
int srb_rotary_table(const float* inv_freq, int32_t rows, float* cos_out, float* sin_out, void* stream);

/* noise truncation + bf16 copy: models.py:168-170.  xt (fp32, in place) = torch.clamp(xt, -tv, tv) when has_truncation
 * (any tv, including 0 and negative values: clamp's min > max rule applies); xt_bf16 = bf16(xt).  n = batch*frames*80 */
int srb_prior_prepare(float* xt, void* xt_bf16, int64_t n, float truncation, int32_t has_truncation, void* stream);

/* Per-call input staging ahead of the (graph-captured) loop: caller tensors -> the plan's static buffers whose rows are
 * padded from frames_in to frames (multiple of 8).  ids (B, frames) = ids_in (B, frames_in) right-padded with 0;
 * xt (B, frames, 80) fp32 = clamp(prior_in (B, frames_in, 80)) (models.py:168-170) with zero pad rows; xt_bf16 = bf16(xt);
 * [zero_a, +zero_a_bytes) and [zero_b, +zero_b_bytes) are cleared (nullable; the rows of xn beyond B*frames that the v^T
 * GEMM reads, and the attention norm bounds).  Replaces four framework copy / fill kernels per call. */
int srb_stage_inputs(const int64_t* ids_in, const float* prior_in, int64_t* ids, float* xt, void* xt_bf16, void* zero_a,
                     int64_t zero_a_bytes, void* zero_b, int64_t zero_b_bytes, int32_t batch, int32_t frames_in,
                     int32_t frames, float truncation, int32_t has_truncation, void* stream);

/* Log-mel front end = mel_spectrogram of src/hifigan/data.py:17-53 (n_fft 400, hop 320, hann window, center=False,
 * 80 slaney mel bands 0-8 kHz, log(max(., 1e-5))):  wav (B, samples) fp32 with row pitch wav_stride ->
 * out (B, 80, frames) fp32, frames <= 1 + (samples - 400) / 320.  window[400] = hann; tw_cos / tw_sin[400] =
 * cos / sin(2 pi i / 400); mel_basis (80, 201) = librosa.filters.mel(sr=16000, n_fft=400, n_mels=80, fmin=0, fmax=8000). */
int srb_log_mel(const float* wav, int64_t wav_stride, int32_t batch, int32_t samples, const float* window, const float* tw_cos,
                const float* tw_sin, const float* mel_basis, float* out, int32_t frames, void* stream);

/* ---- unit quantiser: the k-means half of the upstream unit encoder (src/flow_matching/utils/textless.py:9-21 ->
 * textless SpeechEncoder: units = kmeans_model.predict(dense features), optional de-duplication) ----------------------
 * The codebook is the decoder's own embedding table without its pad row (utils/textless.py:33-35:
 * to_cond_emb.weight = [0; kmeans.cluster_centers_]), so units = argmin_j |x - c_j|^2 = argmax_j (x . c_j - |c_j|^2 / 2).
 *   srb_kmeans_assign: feats (rows, dim) fp32 -> units (rows) int64 = nearest centroid + id_offset (1 = the decoder's
 *     input ids, 0 = sklearn's labels).  Ties go to the smallest index (numpy / sklearn argmin).  The score GEMM runs on the
 *     tensor cores over SPLIT bf16 operands (fp32-grade products, see srb_split_factor) in both libraries:
 *       centroids_packed (n_padded, k_pad) bf16 = [Ch | Ch | Cl] per row, k_pad = 3 dim rounded up to 64, zero filled,
 *       n_padded = centroid count rounded up to 256;  neg_half_norm2 (n_padded) fp32 = -|c_j|^2 / 2, -inf for padding rows;
 *       split_ws (rows, 3 dim) bf16 and keys_ws (rows) u64 are caller-provided scratch.  dim % 8 == 0.
 *     It is the three launches below in sequence.
 *   srb_split_bf16: out (rows, 3 width) bf16 = [bf16(x) | bf16(x - bf16(x)) | bf16(x)]; keys_to_clear (rows u64, or NULL) zeroed.
 *   srb_kmeans_scores_argmax: the GEMM with the arg-max epilogue: keys[row] = max_j ((ordered score bits << 32) | ~j).
 *   srb_kmeans_decode: units[row] = column of keys[row] + id_offset; with lengths (rows = batch * frames) positions at or
 *     beyond lengths[b] become 0, the pad id (srb_kmeans_assign passes its lengths / frames through; NULL = no padding).
 *   srb_unique_consecutive: torch.unique_consecutive(return_counts=True) per utterance over its first lengths[b] ids
 *     (lengths NULL = all `frames`): out_ids / out_counts (batch, frames) right-padded with 0, out_lengths[b] = runs. */
int srb_kmeans_assign(const float* feats, const void* centroids_packed, const float* neg_half_norm2, void* split_ws,
                      uint64_t* keys_ws, int64_t* units, int64_t rows, int32_t dim, int32_t n_padded, int32_t id_offset,
                      const int32_t* lengths, int32_t frames, void* stream);
int srb_split_bf16(const float* x, void* out_bf16, int64_t rows, int32_t width, uint64_t* keys_to_clear, void* stream);
int srb_kmeans_scores_argmax(const void* feats_split_bf16, const void* centroids_packed, const float* neg_half_norm2,
                             void* keys_u64, int64_t rows, int32_t dim, int32_t n_padded, void* stream);
int srb_kmeans_decode(const uint64_t* keys, int64_t* units, int64_t rows, int32_t id_offset, const int32_t* lengths,
                      int32_t frames, void* stream);
int srb_unique_consecutive(const int64_t* ids, const int32_t* lengths, int64_t* out_ids, int32_t* out_counts,
                           int32_t* out_lengths, int32_t batch, int32_t frames, void* stream);

/* Duration-prediction variant (models.py:157-164; fastspeech/modules.py:76-107; transformers length_regulator HF:88-134):
 *   srb_duration_predict : durations[b, n] = clamp(round(exp(conv_k3(E[ids])[b, n]) - 1), 0), 0 at pads; totals[b] = sum.
 *                          dur_table (3, vocab_rows) fp32 = per-unit dot products of the Conv1d(768 -> 1, k 3) taps with the
 *                          embedding rows (row 0 = pad = 0), bias = conv bias.
 *   srb_length_regulate  : out_ids (B, frames_out) = every id repeated durations[b, n] times, right-padded with 0;
 *                          all_one != 0: treat every duration as 1 (the regulator's all-zero rule, HF:113-114). */
int srb_duration_predict(const int64_t* ids, const float* dur_table, float bias, int32_t* durations, int32_t* totals,
                         int32_t batch, int32_t frames, int32_t vocab_rows, void* stream);
int srb_length_regulate(const int64_t* ids, const int32_t* durations, int64_t* out_ids, int32_t batch, int32_t frames,
                        int32_t frames_out, int32_t all_one, void* stream);

/* ---- ODE loop body ------------------------------------------------------------------------------------------
 * to_embed: x0 = xt @ W[:, :80]^T + cond_proj       (models.py:175-176; cond_proj = hoisted cond part + bias, fp32)
 *   xt_bf16 (B, N, 80) bf16; w_packed [256][128]; cond_proj, x0: (B, N, 256) fp32 */
int srb_cfm_embed(const void* xt_bf16, const void* w_packed, const float* cond_proj, float* x0, int32_t batch,
                  int32_t frames, void* stream);

/* ConvPositionEmbed + residual (transformer.py:84-96, models.py:177) fused with the first AdaptiveRMSNorm
 * (norm.py:41-43):  x = gelu(dwconv31(mask(x0)) + b) * mask + x0 ;  xn = bf16(x / max(|x|,1e-12) * g) * mask
 *   dw_w [31][256] fp32 (tap-major), dw_b[256], g[256] (= sqrt(H)(gamma+1) for this step) */
int srb_cfm_posconv_norm(const float* x0, const float* dw_w, const float* dw_b, const float* g, const int32_t* lengths,
                         float* x, void* xn_bf16, int32_t batch, int32_t frames, void* stream);

/* to_qkv + rotary on q,k (transformer.py:109-113): qkv (B, N, 768) bf16 = [rope(q) | rope(k) | v] = xn @ Wqkv^T.
 * rot_cos / rot_sin: (rows, 64) fp32 tables of srb_rotary_table with rows >= max(frames, 32).  qk_norm2_max /
 * qk_norm2_clear (both nullable): the norm bounds described at srb_cfm_qk_rope below. */
int srb_cfm_qkv_rope(const void* xn_bf16, const void* w_packed, const float* rot_cos, const float* rot_sin,
                     void* qkv_bf16, float* qk_norm2_max, float* qk_norm2_clear, int32_t batch, int32_t frames,
                     void* stream);

/* tcgen05 attention path (to_qkv, rotary, masked softmax attention: transformer.py:109-127):
 *   srb_cfm_qk_rope      : qk (B, N, 512) bf16 = rope(xn @ Wqk^T), Wqk = first 512 rows of to_qkv.weight.  rot_cos /
 *                          rot_sin: (rows, 64) fp32 tables of srb_rotary_table with rows >= max(frames, 32).
 *                          qk_norm2_max (nullable): (B, 2 [q|k], 2 [head], 2 [frequency half]) fp32, zeroed by the
 *                          caller; the kernel raises each entry to the largest partial squared row norm (over the
 *                          head's columns i, i + 64 with i in that half of [0, 64)) it produced (atomic max); the
 *                          sum of the two halves bounds the largest squared row norm.
 *                          qk_norm2_clear (nullable, != qk_norm2_max): a second such buffer this launch zeroes, so that
 *                          two buffers used alternately by successive projections need no separate clearing
 *   srb_cfm_v_transposed : vt [256][m_pad] bf16, vt[h*128 + d][b*N + n] = v[b, n, h, d]  (swapped-operand GEMM; xn must
 *                          be readable for m_pad rows, rows >= B*N zero; m_pad multiple of 256)
 *   srb_cfm_attention_tc : o (B, N, 256) bf16 = softmax(q k^T / sqrt(128) + key mask) v, exact softmax, S and O
 *                          accumulators in TMEM.  With qk_norm2_max (the bounds srb_cfm_qk_rope recorded) and
 *                          |q|max |k|max log2(e) / sqrt(128) <= 100 the softmax is a single pass with shift 0;
 *                          otherwise (or with NULL) two passes (row maxima first).  ld = row pitch of qk (>= 512);
 *                          frames % 8 == 0 (TMA box origins inside v^T must be 16-byte aligned; the host pads with
 *                          masked frames). */
int srb_cfm_qk_rope(const void* xn_bf16, const void* w_packed, const float* rot_cos, const float* rot_sin,
                    void* qk_bf16, float* qk_norm2_max, float* qk_norm2_clear, int32_t batch, int32_t frames,
                    void* stream);
int srb_cfm_v_transposed(const void* xn_bf16, const void* wv_bf16, void* vt_bf16, int64_t m_pad, void* stream);
/* the two projections above as ONE launch (the whole to_qkv GEMM, N = 768): w_packed = all 768 rows of to_qkv.weight;
 * q | k with rotary into qk (B, N, 512) and the norm bounds as srb_cfm_qk_rope, v stored transposed into vt [256][m_pad]
 * by the epilogue (columns beyond batch * frames are not written: srb_cfm_attention_tc never reads them) */
int srb_cfm_qk_rope_vt(const void* xn_bf16, const void* w_packed, const float* rot_cos, const float* rot_sin, void* qk_bf16,
                       void* vt_bf16, int64_t m_pad, float* qk_norm2_max, float* qk_norm2_clear, int32_t batch,
                       int32_t frames, void* stream);
int srb_cfm_attention_tc(const void* qk_bf16, int32_t ld, const void* vt_bf16, int64_t m_pad, const int32_t* lengths,
                         const float* qk_norm2_max, void* o_bf16, int32_t batch, int32_t frames, void* stream);
/* to_out + residual (transformer.py:129-130,203) fused with the following AdaptiveRMSNorm:
 *   x += o @ Wout^T ; xn = bf16(adanorm(x, g)) * mask */
int srb_cfm_attn_out_norm(const void* o_bf16, const void* w_packed, const float* g, const int32_t* lengths, float* x,
                          void* xn_bf16, int32_t batch, int32_t frames, void* stream);

/* FeedForward conv1 (k=3) + SIGLU + pad mask (fastspeech/modules.py:58-69):
 *   h (B, N, 896) bf16 = mask * silu(gate) * value, (value | gate) = conv1(xn) + b.
 *   w_packed [1792][768] with rows permuted so every 256-row block holds 128 value rows then their 128 gate rows;
 *   bias_packed[1792] in the same row order.
 *   pad_separated != 0 asserts that every utterance ends in at least one pad row (lengths[b] < frames; pad rows of
 *   xn are zero): the rows are then processed as one sequence, row tiles spanning utterances (same results). */
int srb_cfm_ffn_glu(const void* xn_bf16, const void* w_packed, const float* bias_packed, const int32_t* lengths,
                    void* h_bf16, int32_t batch, int32_t frames, int32_t pad_separated, void* stream);

/* FeedForward conv2 (k=3) + bias + residual (fastspeech/modules.py:71-73, transformer.py:206) fused with the NEXT
 * norm: norm_mode 1 = AdaptiveRMSNorm with g (next layer), 2 = final nn.RMSNorm (transformer.py:208; g = weight,
 * eps = FLT_EPSILON).  x += conv2(h) + b ; xn = bf16(norm(x)) * mask */
int srb_cfm_ffn_out_norm(const void* h_bf16, const void* w_packed, const float* bias, const float* g, int32_t norm_mode,
                         const int32_t* lengths, float* x, void* xn_bf16, int32_t batch, int32_t frames, void* stream);

/* to_pred + Euler update (models.py:183-184), and on the last step the de-normalisation and pad fill
 * (models.py:186-187): xt += dt * (xn @ Wpred^T); xt_bf16 = bf16(xt);
 * if mel != NULL: mel = xt*std + mean, pad rows = log(1e-5) (fp32 and bf16 copies), both COMPACT (B, mel_rows, 80) with
 * mel_rows <= frames: the caller's frame count, which is what the vocoder must see (rows mel_rows..frames are the
 * alignment padding of the loop's buffers and are dropped here) */
int srb_cfm_pred_euler(const void* xn_bf16, const void* w_packed, float dt, float* xt, void* xt_bf16, float* mel,
                       void* mel_bf16, int32_t mel_rows, float std, float mean, float pad_value, const int32_t* lengths,
                       int32_t batch, int32_t frames, void* stream);

/* ---- HiFi-GAN generator (HF:1308-1367, 1451-1491) ------------------------------------------------------------
 * "same"-padded dilated Conv1d as an implicit GEMM with a fused epilogue:
 *   y = (sum_i conv_{k_i, dil_i}(x_i) + bias + sum_j res_j) * scale ; out_raw = bf16(y) ; out_act = bf16(lrelu(y, slope))
 * n_src = 1 for conv_pre / resblock convs; n_src = 3 fuses the last conv2 of the three MRF resblocks, their
 * residuals and the /3 mean (HF:1475-1478) into one launch.  Any of res*, out_raw, out_act may be NULL.
 * x_i: (B, L, c_in) bf16 contiguous; outputs/residuals (B, L, c_out) bf16 contiguous. */
int srb_hifigan_conv(const void* x0, const void* x1, const void* x2, int32_t n_src, const int32_t* kernel,
                     const int32_t* dilation, const void* w_packed, const float* bias, const void* res0,
                     const void* res1, const void* res2, void* out_raw, void* out_act, int32_t batch, int32_t rows,
                     int32_t c_in, int32_t c_out, float scale, float slope, void* stream);

/* Same launch with residuals that carry a leaky_relu: res_slope > 0 says res* hold bf16(leaky_relu(r, res_slope)) instead
 * of bf16(r), and the epilogue adds the raw r recovered from them (negative values divided by res_slope: as accurate as a
 * bf16 copy of r itself).  The resblock chains then keep ONE copy of every tensor -- the activated one the next conv reads
 * -- instead of a raw and an activated one: a third less HBM traffic for the memory-bound convs.  res_slope = 0 is
 * srb_hifigan_conv. */
int srb_hifigan_conv_res_act(const void* x0, const void* x1, const void* x2, int32_t n_src, const int32_t* kernel,
                             const int32_t* dilation, const void* w_packed, const float* bias, const void* res0,
                             const void* res1, const void* res2, float res_slope, void* out_raw, void* out_act, int32_t batch,
                             int32_t rows, int32_t c_in, int32_t c_out, float scale, float slope, void* stream);

/* One resblock pair in ONE launch for the C = 64 stage, single-copy form (HF:1359-1367):
 *   out_act = leaky_relu( r + conv2_{k,1}( leaky_relu( conv1_{k,dilation}(x_act) + b1 ) ) + b2 ),  r = the raw value of x_act
 * x_act (B, L, 64) bf16 = leaky_relu(x, slope) is both the conv input and the residual (recovered in the epilogue);
 * out_act (B, L, 64) != x_act.  kernel 3 or 7 (the k = 11 resblock is tensor-bound unfused and keeps its two launches);
 * w1 / w2 packed as for srb_hifigan_conv ([64][k * 64]).  Two passes over the tensors instead of five. */
int srb_hifigan_pair_fused(const void* x_act, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                           void* out_act, int32_t batch, int32_t rows, int32_t channels, int32_t kernel, int32_t dilation,
                           float slope, void* stream);

/* ConvTranspose1d (HF:1392-1402,1473) in polyphase form: for phase r < stride, output rows q*stride + r are a
 * 2-3 tap conv of the input.  x (B, L_in, c_in) bf16 (already leaky-relu'ed by its producer);
 * out rows L_out = (L_in-1)*stride - 2*pad + k; writes raw and lrelu(slope) copies. */
int srb_hifigan_upsample(const void* x, const void* w_packed, const float* bias, void* out_raw, void* out_act,
                         int32_t batch, int32_t rows_in, int32_t c_in, int32_t c_out, int32_t kernel, int32_t stride,
                         float slope, void* stream);

/* Number of time phases D the fused-MRF kernel uses for a channel width (host query, no launch): the packed weights
 * of its dilation-1 convs depend on it (speech_resynth_b200/packing.py: pack_mrf_conv). */
int srb_hifigan_mrf_phases(int32_t channels);

/* Whole multi-receptive-field stage for the narrow stages (channels = 16 or 32), fused in one kernel:
 *   out_act = leaky_relu((resblock_3(u) + resblock_7(u) + resblock_11(u)) / 3, slope_next)      (HF:1359-1367, 1475-1480)
 * 18 convolutions with every intermediate kept in shared memory / TMEM.  u_raw, out_act: (B, L, C) bf16.
 * w_packed: all 18 conv weights in tcgen05 operand layout (speech_resynth_b200/packing.py:pack_mrf_weights),
 * bias: fp32 [18][C], conv order (resblock 3/7/11) x (pair 0..2) x (convs1, convs2). */
int srb_hifigan_mrf_fused(const void* u_raw, const void* w_packed, const float* bias, void* out_act, int32_t batch,
                          int32_t rows, int32_t channels, float slope, float slope_next, void* stream);

/* conv_post (16 -> 1, k = 7) + tanh (HF:1480-1482); x is the leaky_relu(0.01)'ed stage-5 output (B, L, 16) bf16;
 * w[7][16] fp32 (tap-major).  Dense form (lengths == NULL): wav (B, L) fp32.  Ragged form (lengths = int32[B] valid
 * frames per utterance): utterance b keeps its first n_b = min(L, 320 * lengths[b] + 80) samples
 * (_get_waveform_lengths, models.py:211-221), stored back to back at wav[sum_{i<b} n_i ...] -- the per-utterance crop
 * loop of the reference (models.py:252-256) folded into the store; wav must hold sum_b n_b floats. */
int srb_hifigan_post(const void* x_act, const float* w, float bias, float* wav, int32_t batch, int32_t rows,
                     const int32_t* lengths, void* stream);

/* ---- reference-style forms (fp32 CUDA-core arithmetic; the tight-precision engine and tests) ---------------------
 * Key-padding-masked softmax attention, 2 heads x 128, online softmax in fp32 (transformer.py:115-127):
 *   qkv (B, N, 768) bf16 = [q | k | v] after rotary (srb_cfm_qkv_rope), o (B, N, 256) bf16. */
int srb_cfm_attention_simt(const void* qkv_bf16, const int32_t* lengths, void* o_bf16, int32_t batch, int32_t frames,
                           void* stream);
/* out = leaky_relu((x0 + x1 + x2) * scale, slope) over (rows, channels) bf16 tensors: the MRF mean (HF:1475-1480) when
 * the fused tail of srb_hifigan_conv is not used. */
int srb_hifigan_mean3(const void* x0, const void* x1, const void* x2, void* out, int64_t rows, int32_t channels, float scale,
                      float slope, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRB_H_ */
