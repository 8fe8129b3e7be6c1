/*
 * lcasr_b200 — C ABI of the B200-native (sm_100a) encoder + CTC hot path of
 * robflynnyh/long-context-asr.
 *
 * The reference has no FFI / plugin registry: its boundary is the Python class contract
 * (SURVEY.md §8b).  This header is what a native binding for that contract binds to.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types;
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises the device or the
 *     stream (exceptions are documented per function), so calls are CUDA-graph capturable;
 *   - return value: 0 = ok; <0 = LCASR_E_* ; the message is available from lcasr_last_error()
 *     (thread-local);  nothing is ever thrown across the boundary and there is NO CPU fallback;
 *   - matrices are row-major; "W[N,K]" weights use the torch.nn.Linear layout (out, in);
 *   - dtype codes: LCASR_F32 / LCASR_BF16.
 */
#ifndef LCASR_B200_H_
#define LCASR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCASR_ABI_VERSION 3

enum { LCASR_F32 = 0, LCASR_BF16 = 1 };
enum { LCASR_OK = 0, LCASR_E_BADARG = -1, LCASR_E_UNSUPPORTED = -2, LCASR_E_CUDA = -3, LCASR_E_NOMEM = -4 };
enum { LCASR_ACT_NONE = 0, LCASR_ACT_GELU_TANH = 1, LCASR_ACT_SILU = 2 };
enum { LCASR_NORM_LAYERNORM = 0, LCASR_NORM_RMSNORM = 1 };
enum { LCASR_GEMM_AUTO = 0, LCASR_GEMM_SIMT = 1, LCASR_GEMM_TCGEN05 = 2 };
enum { LCASR_ATTN_AUTO = 0, LCASR_ATTN_SIMT = 1, LCASR_ATTN_TCGEN05 = 2 };

int lcasr_abi_version(void);
const char* lcasr_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
int64_t lcasr_launch_count(void);
void lcasr_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Unit operators (one per reference op on the path; used by the parity tests and by ncu)
 * ---------------------------------------------------------------------------------------- */

/* torch.nn.LayerNorm(d, eps=1e-5) / lcasr RMSNorm (lcasr/components/normalisation.py:30-47,
 * eps added OUTSIDE the sqrt) over the last dim of x[M,d] (fp32).  Writes any of: out_f32[M,d],
 * out_lo[M,d] in `lo_dtype` (the GEMM-operand copy).  Replaces PreNorm.norm
 * (lcasr/components/wrappers.py:14-15), ConformerLayer.norm_out (sconformer_xl.py:371) and
 * decoder.norm (decoder.py:23). `bias` may be NULL (RMSNorm). In-place (out_f32 == x) is allowed. */
int lcasr_layernorm(const float* x, const float* weight, const float* bias, int64_t M, int d,
                    float eps, int kind, float* out_f32, void* out_lo, int lo_dtype, void* stream);

/* Up to three of those norms back to back on a row that stays in registers (d a multiple of 128, <= 2048):
 * ConformerLayer.norm_out followed by decoder.norm for the self-conditioning branch (sconformer_xl.py:371 + decoder.py:23),
 * and norm_out -> decoder.norm -> decoder.norm at the end with legasee_double_norm (:245-247).  weights / biases: HOST
 * arrays of n_stages device pointers (biases or its entries may be NULL).  out_f32 (may be NULL) receives the result of
 * stage f32_stage, out_lo (may be NULL) the result of the last stage.  In-place (out_f32 == x) is allowed. */
int lcasr_layernorm_chain(const float* x, int n_stages, const float* const* weights, const float* const* biases,
                          int64_t M, int d, float eps, int kind, int f32_stage, float* out_f32, void* out_lo,
                          int lo_dtype, void* stream);

/* ConvSubsampling.conv[0] + SiLU: Conv2d(1->C, 3x3, stride 2, pad 1) on the [B,1,T,F] image
 * (lcasr/components/subsampling.py:277-290, run at :418).  spec is the model input [B,F,T] fp32
 * (sconformer_xl.py:185 transposes it); w [C,9] (= weight[C,1,3,3], taps (time,freq) row-major),
 * out is channels-last [B,T1,F1,C] with T1=(T-1)/2+1, F1=(F-1)/2+1. */
int lcasr_subsample_conv0(const float* spec, const float* w, const float* b, int B, int F, int64_t T,
                          int C, void* out, int out_dtype, void* stream);

/* ConvSubsampling depthwise Conv2d(C,C,3x3,stride 2,pad 1,groups=C) (subsampling.py:296-312), no
 * activation; channels-last in [B,Tin,Fin,C] -> out [B,Tout,Fout,C], same dtype. */
int lcasr_subsample_dwconv(const void* in, int dtype, const float* w, const float* b, int B,
                           int64_t Tin, int Fin, int C, void* out, void* stream);

/* conv[0] + SiLU + conv[2] (the first depthwise level) in one kernel, bf16 output [B,T2,F2,C]: the
 * 160x-expanded conv0 activation never leaves shared memory.  C must be a multiple of 64. */
int lcasr_subsample_conv0_dw(const float* spec, const float* w0, const float* b0, const float* w1,
                             const float* b1, int B, int F, int64_t T, int C, void* out, void* stream);

/* out = epilogue(A[M,K] . W[N,K]^T): every dense contraction of the path — nn.Linear / 1x1 Conv
 * (attention.py:483,487; fused_dense.py:464-470; convolution.py:62-86 pointwise; decoder.py:18-19;
 * subsampling.py:314-323,374).  A and W share `ab_dtype`; fp32 accumulation.
 *   y = acc + bias (bias[N] fp32 or NULL);  y = act(y);
 *   if resid != NULL: out(fp32) = resid[M,N] + alpha * y   (resid may alias out)
 *   else            : out = y in `out_dtype`.
 * impl: LCASR_GEMM_TCGEN05 needs ab_dtype == BF16, K % 8 == 0 and 16-byte aligned A/W;
 *       LCASR_GEMM_AUTO picks tcgen05 for bf16 operands and the SIMT fp32 kernel otherwise. */
int lcasr_gemm(const void* A, const void* W, int ab_dtype, int64_t M, int N, int K,
               const float* bias, int act, const float* resid, float alpha, void* out,
               int out_dtype, int impl, void* stream);

/* The qkv projection with the rotary embedding applied in the GEMM epilogue (attention.py:483 + :499-507 in one kernel;
 * bf16, tcgen05): W = qkv weights with interleaved q / k head rows (see lcasr_layer_weights.qkv_w_il); out [M, N] bf16 =
 * [q | k | v] with columns < rope_cols rotated by the PAIR-MAJOR tables cos/sin [Dh/2, rope_n] of lcasr_rope_table_t
 * (position = row % rope_n) in fp32 before the store.
 * q.k^T is invariant under the common permutation of the head dimension, so lcasr_attention_qkv on this output equals
 * lcasr_rope_split + lcasr_attention on the plain projection. */
int lcasr_gemm_rope(const void* A, const void* W, int64_t M, int N, int K, const float* cos_t, const float* sin_t,
                    int64_t rope_n, int rope_cols, int Dh, void* out, void* stream);

/* pointwise_conv1 + GLU in one kernel (convolution.py:105-107; bf16, tcgen05): W [2d, K] / bias [2d] packed in 64-row blocks
 * (32 value channels, then their 32 gate channels); out [M, d] bf16 = value * sigmoid(gate). */
int lcasr_gemm_glu(const void* A, const void* W, int64_t M, int N, int K, const float* bias, void* out, void* stream);

/* Attention reading q, k, v as the three column blocks of ONE row-major [B*N, 3*H*Dh] bf16 matrix (the qkv projection):
 * no split / copy pass.  kv_len may be NULL. out [B,N,H*Dh] bf16. */
int lcasr_attention_qkv(const void* qkv, int B, int64_t N, const int32_t* kv_len, int H, int Dh, void* out, void* stream);

/* fp32 -> compute-dtype copy of n elements (the implicit autocast cast in front of a Linear when
 * decoder_norm=False, decoder.py:23-24). */
int lcasr_cast_f32(const float* in, int64_t n, void* out, int out_dtype, void* stream);

/* torch.nn.functional.glu(x, dim=channels) (convolution.py:107): in [M,2d] -> out [M,d],
 * out[:,j] = in[:,j] * sigmoid(in[:,d+j]). */
int lcasr_glu(const void* in, int dtype, int64_t M, int d, void* out, void* stream);

/* RotaryPositionalEmbedding.forward (rotary_emb.py:44-57): cos/sin [N, Dh/2] fp32 of
 * angle = fp32(pos_offset + n) / interp * inv_freq[j]. */
int lcasr_rope_table(const float* inv_freq, float interp, int64_t pos_offset, int64_t N, int half,
                     float* cos_out, float* sin_out, void* stream);

/* The same tables written pair-major, cos/sin [Dh/2, N] — the layout lcasr_gemm_rope reads (one pair index of 32
 * consecutive token positions is one coalesced 128-byte line for the epilogue warp that owns those 32 rows). */
int lcasr_rope_table_t(const float* inv_freq, float interp, int64_t pos_offset, int64_t N, int half,
                       float* cos_out, float* sin_out, void* stream);

/* The qkv split + rotary of Attention.forward (attention.py:485 "b n (h d qkv)", :499-506,
 * rotary_emb.py:61-73).  qkv [B*N, 3*H*Dh] holds the projection computed with DE-INTERLEAVED
 * weight rows, i.e. columns are [q(h,dh) | k(h,dh) | v(h,dh)].  Writes q,k [B,N,H,Dh] (rotated;
 * cos/sin may be NULL = no rotary) and v either [B,N,H,Dh] (v_transposed=0) or [B,H,Dh,Npad]
 * (v_transposed=1, row pitch Npad >= N, the K-major layout the tcgen05 PV product wants). */
int lcasr_rope_split(const void* qkv, int dtype, int B, int64_t N, int H, int Dh, const float* cos_t,
                     const float* sin_t, void* q, void* k, void* v, int v_transposed, int64_t Npad,
                     void* stream);

/* softmax(q k^T / sqrt(Dh)) v, non-causal, no mask (attention.py:532 flash path / :541 SDPA).
 * q,k [B,N,H,Dh]; v [B,N,H,Dh] (v_transposed=0) or [B,H,Dh,Npad]; out [B,N,H*Dh]. */
int lcasr_attention(const void* q, const void* k, const void* v, int dtype, int B, int64_t N, int H,
                    int Dh, int v_transposed, int64_t Npad, void* out, int impl, void* stream);

/* Same attention with Nq query rows against Nk keys/values per batch entry (q [B,Nq,H,Dh]; k,v
 * [B,Nk,H,Dh], natural layout): the sequence-parallel form of Attention.forward, where a rank owns
 * a block of Nq query tokens and attends to the all-gathered K/V of the whole recording. */
int lcasr_attention_cross(const void* q, const void* k, const void* v, int dtype, int B, int64_t Nq,
                          int64_t Nk, int H, int Dh, void* out, int impl, void* stream);

/* Attention with a key-padding mask: kv_len[b] (device int32[B]) keys of batch entry b are valid, the rest
 * are masked (the ragged-batch path of Attention.forward, attention.py:511,541,547 with the att_mask built at
 * sconformer_xl.py:207-213).  Rows of padded queries are unspecified (the reference zeroes them). */
int lcasr_attention_masked(const void* q, const void* k, const void* v, int dtype, int B, int64_t Nq,
                           int64_t Nk, const int32_t* kv_len, int H, int Dh, void* out, int impl,
                           void* stream);

/* Local (windowed) self-attention, flash-attn window_size=(win_left, win_right) semantics (attention.py:466,
 * 527-530; the 'windowed_attention' evaluation mode of eval/run.py:38-43): query i attends to keys
 * [i - win_left, i + win_right]; -1 = unlimited on that side.  kv_len may be NULL.  Key tiles outside a CTA's band
 * are skipped, so the cost is O(N * window) instead of O(N^2). */
int lcasr_attention_window(const void* q, const void* k, const void* v, int dtype, int B, int64_t N,
                           const int32_t* kv_len, int H, int Dh, int win_left, int win_right, void* out,
                           int impl, void* stream);

/* GLU that writes zeros for tokens n >= lengths[b] (convolution.py:107-110: glu then masked_fill). */
int lcasr_glu_masked(const void* in, int dtype, int B, int64_t N, int d, const int32_t* lengths, void* out,
                     void* stream);

/* depthwise Conv1d(k, pad (k-1)/2, groups=d) + bias -> BatchRenorm1d eval affine
 * ((y-mean)/std*weight+bias, no eps; batchrenorm.py:86-91) -> SiLU  (convolution.py:112-121).
 * in/out channels-last [B,N,d]; w [d,ksize] fp32. */
int lcasr_dwconv_brn_silu(const void* in, int dtype, int B, int64_t N, int d, int ksize, const float* w,
                          const float* b, const float* brn_mean, const float* brn_std,
                          const float* brn_w, const float* brn_b, void* out, int out_dtype, void* stream);

/* row softmax over V classes (sconformer_xl.py:242): in [M,V] -> out [M,V]. */
int lcasr_softmax(const void* in, int in_dtype, int64_t M, int V, void* out, int out_dtype, void* stream);

/* F.log_softmax(logits, -1) (decoder.py:25), in place on fp32 logits [M,V], fused with the
 * per-frame argmax of GreedyCTCDecoder (decoding/greedy.py:19); argmax (int32[M]) may be NULL. */
int lcasr_log_softmax_argmax(float* logits, int64_t M, int V, int32_t* argmax, void* stream);

/* torch.argmax(emission, dim=-1) (decoding/greedy.py:19) on an fp32 [M,V] tensor; first max wins. */
int lcasr_argmax_rows(const float* x, int64_t M, int V, int32_t* argmax, void* stream);

/* GreedyCTCDecoder.forward (decoding/greedy.py:19-21): unique_consecutive + drop blank on
 * per-frame argmax ids [B,N]; lengths[B] (int32, NULL = N). tokens [B,N] int32 (compacted prefix),
 * n_tokens [B] int32. */
int lcasr_greedy_collapse(const int32_t* argmax, int B, int64_t N, const int32_t* lengths, int blank,
                          int32_t* tokens, int32_t* n_tokens, void* stream);

/* torch.nn.CTCLoss(blank, reduction='none') forward (exp/train.py:104,249).  log_probs is
 * BATCH-major [B,N,V] fp32 (the tensor the model returns; the reference passes its transpose
 * view).  targets [B,S_max] int64 (padded), input_lengths int32[B], target_lengths int64[B].
 * nll [B] fp32.  alpha_ws: NULL, or fp32 [B,N,2*S_max+1] to keep alpha for the backward. */
int lcasr_ctc_loss_fwd(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                       int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                       int blank, float* nll, float* alpha_ws, void* stream);

/* gradient of sum_b grad_nll[b]*nll[b] w.r.t. log_probs [B,N,V] (ATen ctc_loss_backward semantics;
 * frames >= input_length get 0).  alpha_ws from the forward; beta_ws fp32 [B,N,2*S_max+1] scratch. */
int lcasr_ctc_loss_bwd(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                       int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                       int blank, const float* nll, const float* grad_nll, const float* alpha_ws,
                       float* beta_ws, float* grad, void* stream);

/* Front-end (lcasr/utils/audio_tools.py:44-57 to_spectogram): waveform [B, n_samples] fp32 (16 kHz) -> mel power
 * spectrogram out [B, n_mels, n_frames], n_frames = lcasr_melspec_frames(n_samples) = 1 + n_samples/160
 * (torchaudio MelSpectrogram: n_fft 512, Hann(400) centred in the frame, hop 160, center + reflect padding, power 2),
 * optionally standardised per (recording, mel bin) over time with the unbiased std.  cos_tab / sin_tab [512, 257]: the
 * DFT twiddles multiplied by the window; fb [257, n_mels]: the mel filterbank; sums: fp64 scratch [B, n_mels, 2]. */
int64_t lcasr_melspec_frames(int64_t n_samples);
int lcasr_melspec(const float* wave, int B, int64_t n_samples, const float* cos_tab, const float* sin_tab,
                  const float* fb, int n_mels, float* out, double* sums, int normalise, void* stream);

/* SpecAugment (lcasr/utils/augmentation.py:61-104; exp/train.py:227) in one pass over spec [B, F, T] (fp32).
 * lcasr_specaug_mean leaves {sum, count} of the un-padded region t < lengths[b] (lengths NULL: everything) in acc[2]
 * (fp64, device) — the reference's fill value when zero_masking is off.  lcasr_specaug_apply rebuilds every mask
 * interval from its two uniform draws as torchaudio's mask_along_axis(_iid) does (value = u1*param; start =
 * trunc(u2*(size - value)); end = start + trunc(value)) and writes out = fill where any mask covers (b, f, t), spec
 * elsewhere.  u_time [n_time, 2, draws_per_mask], u_freq [n_freq, 2, draws_per_mask]: draws_per_mask = B (iid masks) or
 * 1 (one interval shared by the batch); *_param already limited by max_p; mean_acc NULL = zero masking. */
int lcasr_specaug_mean(const float* spec, int B, int F, int64_t T, const int32_t* lengths, double* acc, void* stream);
int lcasr_specaug_apply(const float* spec, int B, int F, int64_t T, int n_time, int time_param, const float* u_time,
                        int n_freq, int freq_param, const float* u_freq, int draws_per_mask,
                        const double* mean_acc, float* out, void* stream);

/* Long-form moving-window merge (lcasr/eval/utils.py:45-111 fetch_logits): window k occupies rows
 * [win_row0[k], win_row0[k] + win_len[k]) of logp [*, V] (fp32 log-probs) and covers merged frames
 * [win_pos[k], win_pos[k] + win_len[k]); windows sorted by win_pos, max_len = max win_len.  For every merged frame:
 * out = log(mean over covering windows of exp(logp)) (optional, [n_total, V]) and argmax (optional, [n_total]). */
int lcasr_window_merge(const float* logp, int V, int K, const int64_t* win_row0, const int32_t* win_len,
                       const int32_t* win_pos, int max_len, int64_t n_total, float* out, int32_t* argmax,
                       void* stream);

/* Buffered long-form mode (lcasr/eval/buffered_transcription.py:74-90): same arguments, but the windows' row ranges
 * (here: the central chunk of every buffer) tile the merged frames without overlap and the log-probabilities are
 * copied unchanged (bit-exact) together with the per-frame argmax. */
int lcasr_window_concat(const float* logp, int V, int K, const int64_t* win_row0, const int32_t* win_len,
                        const int32_t* win_pos, int max_len, int64_t n_total, float* out, int32_t* argmax,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step (cfg 5): the operators behind loss.backward() of exp/train.py:249-262.  The
 * reference gets these from torch.autograd over the modules of SURVEY §8 a3-a11; here each is
 * a hand-written kernel.  Activations bf16 (the reference trains under bf16 autocast,
 * exp/train.py:225), residual stream / its gradient / parameter gradients fp32.
 * fp32 gradient outputs ACCUMULATE (+=): the caller zeroes them (or passes an existing .grad).
 * ---------------------------------------------------------------------------------------- */

enum { LCASR_EPI_SCALE = 0, LCASR_EPI_EXP2 = 1, LCASR_EPI_DS = 2, LCASR_EPI_GELU_BWD = 3, LCASR_EPI_SILU_BWD = 4 };

/* General tcgen05 GEMM of the backward pass: for every batch entry (b1 < nb1, b2 < nb2)
 *     acc[M,N] = sum_k A(m,k) * B(n,k)
 * A is stored [M,K] row-major (a_mn=0, pitch lda) or [K,M] row-major (a_mn=1, "MN-major"); B is
 * stored [N,K] (b_mn=0) or [K,N] (b_mn=1); batch entry (b1,b2) of a tensor starts s?1*b1 + s?2*b2
 * elements after its base.  bf16 operands, fp32 accumulation.  Epilogue `epi`:
 *   SCALE     out = alpha*acc
 *   EXP2      out = exp2(alpha*acc - rowvec[row])                (attention backward: P from the saved lse)
 *   DS        out = aux * (acc - rowvec[row]) * alpha            (attention backward: dS = P o (dP - D) * scale)
 *   GELU_BWD  out = alpha*acc * gelu_tanh'(aux)                  (fused_dense.py:466 backward)
 *   SILU_BWD  out = alpha*acc * silu'(aux)
 * out_dtype BF16: stored; F32: out += alpha*acc with fp32 atomics (SCALE only), which is what split-K
 * (ksplit > 1, or 0 = choose) and gradient accumulation need.  Pitches and strides are multiples of 8; so is
 * N, except for bf16 outputs whose pitch leaves room for N rounded up to 8 (padding columns are unspecified).
 * Replaces the cuBLAS calls autograd issues for nn.Linear / 1x1 Conv backward and the flash-attn
 * backward (attention.py:532) of the reference. */
typedef struct lcasr_gemm_ex_args {
  const void* A; const void* B; void* out; const void* aux; const float* rowvec;
  int64_t M; int32_t N, K;
  int32_t a_mn, b_mn;
  int64_t lda, ldb, ldo, ldaux;
  int32_t nb1, nb2;
  int64_t sa1, sa2, sb1, sb2, so1, so2, sx1, sx2, sr1, sr2;
  float alpha; int32_t epi; int32_t out_dtype; int32_t ksplit;
} lcasr_gemm_ex_args;
int lcasr_gemm_ex(const lcasr_gemm_ex_args* args, void* stream);

/* First half of the attention backward in ONE pass per (recording, head): S = q k^T and dP = dO v^T as two
 * accumulators of the same tile, P = exp2(S*scale*log2e - lse2) and dS = P o (dP - dvec) * scale written as
 * bf16 [nb,H,N,Np] (Np = N rounded up to 8).  q,k,v,d_out bf16 [nb,N,H,Dh]; lse2, dvec fp32 [nb,H,N]. */
int lcasr_attention_bwd_pds(const void* q, const void* k, const void* v, const void* d_out, const float* lse2,
                            const float* dvec, int nb, int64_t N, int H, int Dh, void* P, void* dS, void* stream);

/* The whole attention backward, flash style (head_dim 64 or 128): dq, dk, dv from q, k, v, the saved output o, its
 * gradient d_out and the forward's lse2, WITHOUT materialising P or dS — O(N) memory, so a 20-minute context trains
 * (replaces flash-attn's backward behind attention.py:509-551 / F.scaled_dot_product_attention autograd).  One launch
 * computes dk, dv (a CTA owns 128 keys and streams the queries), a second dq (a CTA owns 128 queries and streams the
 * keys); the score tiles are recomputed on the tensor cores in both.  All tensors bf16 [nb, n_pitch, H, Dh], the first
 * N tokens of every recording valid (gradient rows beyond N are not written); lse2 fp32 [nb, H, lse_pitch];
 * workspace: lcasr_attention_bwd_flash_workspace_bytes(nb, N, H) bytes, 16-byte aligned. */
int64_t lcasr_attention_bwd_flash_workspace_bytes(int nb, int64_t N, int H);
int lcasr_attention_bwd_flash(const void* q, const void* k, const void* v, const void* o, const void* d_out,
                              const float* lse2, int nb, int64_t N, int64_t n_pitch, int64_t lse_pitch, int H, int Dh,
                              void* dq, void* dk, void* dv, void* workspace, int64_t workspace_bytes, void* stream);

/* Attention forward that also returns lse [B,H,N] fp32 = log2-domain log-sum-exp of the scaled scores
 * (bf16, natural-layout V, tcgen05 kernel). */
int lcasr_attention_train(const void* q, const void* k, const void* v, int B, int64_t N, int H, int Dh,
                          void* out, float* lse, void* stream);

/* Padded training batches (exp/train.py:236-241 passes length=a_lengths): the forward with a key-padding mask, and the
 * in-place zeroing of the rows of padded tokens (x [B,N,d], rows n >= lengths[b]) that attention.py:511,541 applies to
 * the attention input / output and that autograd applies to the matching gradients. */
int lcasr_attention_train_masked(const void* q, const void* k, const void* v, int B, int64_t N, const int32_t* kv_len,
                                 int H, int Dh, void* out, float* lse, void* stream);
int lcasr_mask_rows(void* x, int dtype, int B, int64_t N, int d, const int32_t* lengths, void* stream);

/* Training forward of Linear + activation: pre = A.W^T + bias and out = act(pre), both bf16 [M,N] (tcgen05 GEMM
 * with a two-output epilogue; replaces lcasr_gemm + lcasr_act_fwd). */
int lcasr_gemm_act_pre(const void* A, const void* W, int64_t M, int N, int K, const float* bias, int act,
                       void* out, void* pre_out, void* stream);
/* out(bf16) = scale * in(fp32), n % 8 == 0 */
int lcasr_scale_cast(const float* in, int64_t n, float scale, void* out, void* stream);
/* out = gelu_tanh(in) / silu(in), bf16 (the pre-activation is kept for the backward) */
int lcasr_act_fwd(const void* in, int64_t n, int act, void* out, void* stream);
/* out = dy * act'(pre), bf16 (pre = the saved pre-activation) */
int lcasr_act_bwd(const void* pre, const void* dy, int64_t n, int act, void* out, void* stream);
/* x(fp32) += a(bf16) */
int lcasr_add_bf16(float* x, const void* a, int64_t n, void* stream);
/* GLU backward: u [M,2d] (the GLU input), dg [M,d] -> du [M,2d] */
int lcasr_glu_bwd(const void* u, const void* dg, int64_t M, int d, void* du, void* stream);
/* rotary backward (rotation by -theta of dq, dk) + merge with dv into dqkv [B*N, 3*H*Dh] = [dq|dk|dv] */
int lcasr_rope_bwd_merge(const void* dq, const void* dk, const void* dv, int B, int64_t N, int H, int Dh,
                         const float* cos_t, const float* sin_t, void* dqkv, void* stream);
/* out[b,h,n] = sum_dh a[b,n,h,dh]*b[b,n,h,dh] (the D term of the attention backward) */
int lcasr_rowdot(const void* a, const void* b, int B, int64_t N, int H, int Dh, float* out, void* stream);
/* dl = scale * p o (dp - sum(p o dp)), rows of V (self-conditioning softmax, sconformer_xl.py:242) */
int lcasr_softmax_bwd(const void* p, const void* dp, int64_t M, int V, float scale, void* dl, void* stream);
/* dl(bf16) = scale * (dlp - exp(lp) * sum(dlp)) (decoder.py:29 log_softmax; lp, dlp fp32) */
int lcasr_log_softmax_bwd(const float* lp, const float* dlp, int64_t M, int V, float scale, void* dl,
                          void* stream);
/* out[d] += scale * column sums of in [M,d] (bias gradients) */
int lcasr_colsum(const void* in, int dtype, int64_t M, int d, float scale, float* out, void* stream);
/* LayerNorm / lcasr-RMSNorm backward from the saved input x (statistics recomputed): dx = or += (accumulate),
 * dweight += , dbias += (LayerNorm only). dy in dy_dtype. */
int lcasr_layernorm_bwd(const float* x, const void* dy, int dy_dtype, const float* weight, int64_t M, int d,
                        float eps, int kind, int accumulate, float* dx, float* dweight, float* dbias,
                        void* stream);
/* same, and cast_out (bf16 [M,d], may be NULL) = cast_scale * (the new dx): the dY operand of the next sub-layer's
 * backward GEMMs, produced here instead of by a separate cast pass */
int lcasr_layernorm_bwd_cast(const float* x, const void* dy, int dy_dtype, const float* weight, int64_t M, int d,
                             float eps, int kind, int accumulate, float* dx, float* dweight, float* dbias,
                             void* cast_out, float cast_scale, void* stream);
/* depthwise Conv1d(k, pad (k-1)/2, groups=d)+bias, channels-last bf16 [B,N,d] (convolution.py:112), with
 * optional per-channel sum / sum-of-squares of the output (BatchRenorm batch statistics); its data gradient
 * and its weight / bias gradients (dw [d,k], db [d], +=). */
int lcasr_dwconv1d_fwd(const void* in, int B, int64_t N, int d, int ksize, const float* w, const float* b,
                       void* out, double* sum, double* sumsq, void* stream);
int lcasr_dwconv1d_bwd_data(const void* dout, int B, int64_t N, int d, int ksize, const float* w, void* din,
                            void* stream);
int lcasr_dwconv1d_bwd_weight(const void* x, const void* dout, int B, int64_t N, int d, int ksize, float* dw,
                              float* db, void* stream);
/* BatchRenorm1d training forward (batchrenorm.py:52-84) from the channel sums over `count` tokens:
 * writes the affine z = c*A + Bc, saves stats [5,d] = (mu, sigma, r, d, std) and updates running_mean /
 * running_std in place with `momentum`.  rmax, dmax: the module's current clamps (:41-50). */
int lcasr_brn_train_stats(const double* sum, const double* sumsq, int64_t count, int d, float* running_mean,
                          float* running_std, float eps, float rmax, float dmax, float momentum,
                          const float* weight, const float* bias, float* A, float* Bc, float* stats,
                          void* stream);
/* y = silu(c*A + Bc) */
int lcasr_affine_silu(const void* c, int64_t M, int d, const float* A, const float* Bc, void* out, void* stream);
/* dz = dy * silu'(c*A+Bc) (bf16) and S1 += sum dz, S2 += sum dz*xhat per channel.  The per-channel sums of the
 * BatchRenorm forward (lcasr_dwconv1d_fwd) and backward are accumulated across CTAs in fp64 so that the fp32 statistics
 * and everything downstream of them are reproducible run to run. */
int lcasr_affine_silu_bwd(const void* c, const void* dy, int64_t M, int d, const float* A, const float* Bc,
                          const float* stats, void* dz, double* S1, double* S2, void* stream);
/* BatchRenorm backward, per channel: dweight +=, dbias +=, and coef [3,d] of dc = k1*dz + k2*c + k3 */
int lcasr_brn_bwd_finalize(const double* S1, const double* S2, int64_t count, int d, const float* weight,
                           const float* stats, float* dweight, float* dbias, float* coef, void* stream);
/* out = coef[0]*a + coef[1]*b + coef[2] per channel, bf16 [M,d] */
int lcasr_affine3(const void* a, const void* b, int64_t M, int d, const float* coef, void* out, void* stream);
/* backward of the subsampling stencils (subsampling.py:277-312): depthwise 3x3 s2 data / weight gradients
 * (channels-last bf16) and the conv0 (+SiLU) weight gradient (pre-activation recomputed from the spectrogram) */
int lcasr_subsample_dwconv_bwd_data(const void* dout, const float* w, int B, int64_t Tin, int Fin, int C,
                                    void* din, void* stream);
int lcasr_subsample_dwconv_bwd_weight(const void* in, const void* dout, int B, int64_t Tin, int Fin, int C,
                                      float* dw, float* db, void* stream);
int lcasr_subsample_conv0_bwd(const float* spec, const float* w, const float* b, const void* ds1, int B, int F,
                              int64_t T, int C, float* dw, float* db, void* stream);
/* Fused backward of conv0 + SiLU + the first depthwise level (subsampling.py:277-296: Conv2d(1->C,3x3,s2), SiLU,
 * depthwise Conv2d(3x3,s2); autograd in the reference), C % 64 == 0: from the spectrogram and
 * dd1 = dL/d(depthwise-1 output) [B,T2,F2,C] (bf16) accumulate (+=) the gradients of conv0 w [C,9] / b [C] and of the
 * depthwise w [C,9] / b [C].  The conv0 activation and its gradient are recomputed / consumed on chip: together with
 * lcasr_subsample_conv0_dw in the forward, the 160x-expanded [B,T/2,40,C] tensor never exists in HBM. */
int lcasr_subsample_l1_bwd(const float* spec, const float* w0, const float* b0, const float* w1, const void* dd1,
                           int B, int F, int64_t T, int C, float* dw0, float* db0, float* dw1, float* db1,
                           void* stream);

/* Training form of the CTC loss (up to 4096 extended states): the alpha and beta recursions are independent, so
 * one launch runs both concurrently (2*B CTAs) into alpha_ws / beta_ws [B,N,2*S_max+1]; the backward is then
 * only lcasr_ctc_loss_grad (the class scatter).  Same results as lcasr_ctc_loss_fwd + lcasr_ctc_loss_bwd. */
int lcasr_ctc_loss_fwd_ab(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                          int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                          int blank, float* nll, float* alpha_ws, float* beta_ws, void* stream);
int lcasr_ctc_loss_grad(const float* log_probs, int B, int64_t N, int V, const int64_t* targets,
                        int64_t S_max, const int32_t* input_lengths, const int64_t* target_lengths,
                        int blank, const float* nll, const float* grad_nll, const float* alpha_ws,
                        const float* beta_ws, float* grad, void* stream);

/* Workspace forms of the three CTC entry points.  With a workspace of lcasr_ctc_workspace_bytes() bytes (16-byte aligned;
 * 0 = not applicable) the recursions run as a TIME-SKEWED WAVEFRONT over the whole GPU: the extended states of a lattice are cut
 * into contiguous chunks, one persistent CTA each (cooperative launch), chunk g running 16 frames behind chunk g-1 and receiving
 * its neighbour's two boundary states through the workspace — no per-frame barrier wider than one CTA.  Results are bit-identical
 * to the plain forms (same arithmetic per state); a 1-hour lattice (45000 frames x 27001 states) takes milliseconds instead of
 * 94 ms.  workspace == NULL: identical to the plain entry points. */
int64_t lcasr_ctc_workspace_bytes(int B, int64_t N, int64_t S_max, int both_directions);
int lcasr_ctc_loss_fwd_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets, int64_t S_max,
                          const int32_t* input_lengths, const int64_t* target_lengths, int blank, float* nll,
                          float* alpha_ws, void* workspace, int64_t workspace_bytes, void* stream);
int lcasr_ctc_loss_bwd_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets, int64_t S_max,
                          const int32_t* input_lengths, const int64_t* target_lengths, int blank, const float* nll,
                          const float* grad_nll, const float* alpha_ws, float* beta_ws, float* grad, void* workspace,
                          int64_t workspace_bytes, void* stream);
int lcasr_ctc_loss_fwd_ab_ws(const float* log_probs, int B, int64_t N, int V, const int64_t* targets, int64_t S_max,
                             const int32_t* input_lengths, const int64_t* target_lengths, int blank, float* nll,
                             float* alpha_ws, float* beta_ws, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer step of the training loop: exp/train.py:46-61 (clip_grad_norm_ + optimizer.step) with
 * MADGRAD (lcasr/optim/madgrad.py:81-212, dense fp32 branch), as multi-tensor kernels.
 * `tensors` is a DEVICE array; work is cut into chunks of lcasr_opt_chunk_elems() elements:
 * chunk c covers elements [chunk_index[c]*E, +E) of tensor chunk_tensor[c].  g == NULL: skipped.
 * ---------------------------------------------------------------------------------------- */
typedef struct lcasr_opt_tensor {
  float* p;          /* parameter, updated in place */
  const float* g;    /* gradient (or NULL) */
  float* gss;        /* state: grad_sum_sq */
  float* s;          /* state: s */
  float* x0;         /* state: x0 (momentum != 0) */
  int64_t n;
} lcasr_opt_tensor;
int lcasr_opt_chunk_elems(void);
/* *sumsq = sum over all gradients of g^2 (the square of clip_grad_norm_'s total norm) */
int lcasr_grad_sumsq(const lcasr_opt_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                     int n_chunks, float* sumsq, void* stream);
/* g *= min(1, max_norm / (sqrt(*sumsq) + 1e-6)) in place (torch.nn.utils.clip_grad_norm_) */
int lcasr_grad_scale(const lcasr_opt_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                     int n_chunks, const float* sumsq, float max_norm, void* stream);
/* one MADGRAD step on every tensor; lr = the group's lr (+eps if non-zero, madgrad.py:104), lamb = lr*sqrt(k+1).
 * sumsq != NULL and max_norm > 0: the clip coefficient is applied to the gradients on the fly (read from device
 * memory: no host synchronisation); the gradients themselves are left untouched. */
int lcasr_madgrad_step(const lcasr_opt_tensor* tensors, const int32_t* chunk_tensor, const int32_t* chunk_index,
                       int n_chunks, const float* sumsq, float max_norm, float lr, float lamb, float eps,
                       float weight_decay, float momentum, int decouple_decay, void* stream);

/* ------------------------------------------------------------------------------------------
 * The model: SCConformerXL.forward (lcasr/models/sconformer_xl.py:162-252), equal-length path
 * ---------------------------------------------------------------------------------------- */

typedef struct lcasr_config {
  int32_t abi_version;        /* LCASR_ABI_VERSION */
  int32_t n_layers, d_model, n_heads, head_dim;
  int32_t feat_in;            /* 80 */
  int32_t conv_channels;      /* subsampling_conv_channels */
  int32_t conv_kernel_size;   /* 9 */
  int32_t num_classes;        /* vocab_size + 1 */
  int32_t norm_kind;          /* LCASR_NORM_* (default_norm) */
  int32_t decoder_norm;       /* decoder_norm kwarg */
  int32_t use_rotary, self_conditioning, legasee_double_norm, bias_in_ff;
  int32_t compute_dtype;      /* LCASR_BF16 (tcgen05 path) or LCASR_F32 (SIMT fp32 parity mode) */
  float   rotary_interp;      /* rotary_interpolation_factor */
  float   norm_eps;           /* 1e-5 (LayerNorm) / 1e-8 (RMSNorm) */
  int32_t attn_window_left;   /* attention_window_size(_left): keys [i - left, i + right]; -1 = unlimited (ABI 2) */
  int32_t attn_window_right;
} lcasr_config;

/* Matrices ("*_w" of rank 2) are in `compute_dtype`; every vector and every depthwise filter is
 * fp32.  qkv_w rows are de-interleaved to [q | k | v]; sub_out_w columns are permuted from the
 * reference's (c*F3+f) to channels-last (f*C+c).  Unused pointers (e.g. biases when bias_in_ff=0,
 * norm biases for RMSNorm) are NULL. */
typedef struct lcasr_layer_weights {
  const float *ff1_norm_w, *ff1_norm_b; const void *ff1_fc1_w; const float *ff1_fc1_b; const void *ff1_fc2_w; const float *ff1_fc2_b;
  const float *attn_norm_w, *attn_norm_b; const void *qkv_w; const void *out_w;
  const float *conv_norm_w, *conv_norm_b; const void *pw1_w; const float *pw1_b;
  const float *dw_w, *dw_b, *brn_mean, *brn_std, *brn_w, *brn_b; const void *pw2_w; const float *pw2_b;
  const float *ff2_norm_w, *ff2_norm_b; const void *ff2_fc1_w; const float *ff2_fc1_b; const void *ff2_fc2_w; const float *ff2_fc2_b;
  const float *norm_out_w, *norm_out_b;
  /* ABI 3 — optional (NULL = use the unfused kernels), compute dtype, for the fused epilogues of the bf16 path:
   * qkv_w_il : qkv_w with the rows of every q and k head interleaved (new row 2i <- old i, 2i+1 <- old i + Dh/2) so that a
   *            rotary pair is two adjacent output columns (rotary_emb.py:61-73 rotate_half pairs (i, i + Dh/2));
   * pw1_w_glu / pw1_b_glu : pointwise_conv1 rows / bias in 64-row blocks [32 value channels | their 32 gate channels]
   *            (convolution.py:105-107: glu pairs channel c with channel d + c). */
  const void *qkv_w_il; const void *pw1_w_glu; const float *pw1_b_glu;
} lcasr_layer_weights;

typedef struct lcasr_weights {
  const float *conv0_w, *conv0_b;                          /* [C,9], [C] */
  const float *dw1_w, *dw1_b; const void *pw1_w; const float *pw1_b;   /* conv.2 / conv.3 */
  const float *dw2_w, *dw2_b; const void *pw2_w; const float *pw2_b;   /* conv.5 / conv.6 */
  const void *sub_out_w;                                   /* [d, F3*C] permuted */
  const float *inv_freq;                                   /* [Dh/2] or NULL */
  const float *dec_norm_w, *dec_norm_b; const void *dec_ff_w; const float *dec_ff_b;
  const void *dec_rep_w; const float *dec_rep_b;
  const lcasr_layer_weights* layers_host;                  /* HOST array [n_layers] of device ptrs */
} lcasr_weights;

typedef struct lcasr_model lcasr_model;

/* Builds a model handle.  Weights are BORROWED (the caller — PyTorch — owns them and must keep
 * them alive and unchanged; call again after load_state_dict).  Replaces SCConformerXL.__init__ +
 * load_state_dict (sconformer_xl.py:32-160, lcasr/utils/general.py:57-59). */
int lcasr_model_create(const lcasr_config* cfg, const lcasr_weights* w, lcasr_model** out);
void lcasr_model_destroy(lcasr_model* m);

/* test / profiling hook: force the GEMM (LCASR_GEMM_*) and attention (LCASR_ATTN_*) kernels */
int lcasr_model_set_impl(lcasr_model* m, int gemm_impl, int attn_impl);

/* Tuning hook of the dense attention launch (bf16 tensor-core path).  A CTA owns 256 queries of one head and walks all
 * keys, one CTA per SM, so a launch takes ceil(units / SMs) unit-times; by default (tail_pairs < 0) the forward computes
 * the last `tail_pairs` query-tile pairs of the last recording as `key_pieces` key-range partial results on side streams
 * (merged exactly) whenever a list-schedule model says the last wave would otherwise be mostly idle.  tail_pairs = 0
 * turns that off; tail_pairs > 0 forces a split (key_pieces in 2..4). */
int lcasr_model_set_attention_tail(lcasr_model* m, int tail_pairs, int key_pieces);
/* host-only (no GPU needed): the split the automatic mode chooses for B recordings of N tokens, H heads, `sms` SMs
 * (tail_pairs = 0: none). */
int lcasr_attention_tail_plan(int B, int64_t N, int H, int sms, int* tail_pairs, int* key_pieces);

/* In-step kernel timing for the roofline report: when enabled, lcasr_model_forward brackets every
 * kernel launch with CUDA events on the launch stream; lcasr_model_get_timing synchronises, returns
 * the summed milliseconds / launch counts per category and clears the recorder.
 * Categories: 0 subsampling stencils, 1 norms/casts, 2 GEMMs, 3 attention, 4 rotary, 5 conv module
 * (GLU, dwconv+BRN+SiLU), 6 softmax / log-softmax+argmax.  ncat >= 7. */
int lcasr_model_set_timing(lcasr_model* m, int enable);
int lcasr_model_get_timing(lcasr_model* m, float* ms_by_cat, int32_t* launches_by_cat, int ncat);

/* tokens after 8x subsampling: calc_length (subsampling.py:557-567) applied three times */
int64_t lcasr_out_length(int64_t T);

/* bytes of device workspace lcasr_model_forward needs for a [B,feat_in,T] input */
int64_t lcasr_model_workspace_bytes(const lcasr_model* m, int B, int64_t T);

/* SCConformerXL.forward(audio_signal[B,feat_in,T]) with all lengths == T
 * (sconformer_xl.py:162-252).  Writes out[B,N,num_classes] fp32: log-probs, or logits when
 * return_logits != 0.  argmax (int32 [B,N], may be NULL) is the fused per-frame argmax that
 * GreedyCTCDecoder (decoding/greedy.py:19) would compute.  workspace: device scratch of at least
 * lcasr_model_workspace_bytes().  Asynchronous on `stream`. */
int lcasr_model_forward(lcasr_model* m, const float* spec, int B, int64_t T, float* out,
                        int32_t* argmax, int return_logits, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* Same with a ragged batch: tok_len (device int32[B]) = valid tokens per recording after subsampling
 * (lcasr_out_length of each item's frame count).  NULL = all equal (identical to lcasr_model_forward). */
int lcasr_model_forward_lengths(lcasr_model* m, const float* spec, int B, int64_t T, const int32_t* tok_len,
                                float* out, int32_t* argmax, int return_logits, void* workspace,
                                int64_t workspace_bytes, void* stream);

/* End-to-end convenience used by bench.py's e2e leg and by a non-PyTorch host: pinned-host
 * spectrogram in, token ids out (H2D copy, forward, greedy collapse, D2H copy of tokens, stream
 * synchronise).  tokens_host [B,N] int32, n_tokens_host [B] int32. */
int lcasr_model_transcribe_host(lcasr_model* m, const float* spec_host, int B, int64_t T,
                                int32_t* tokens_host, int32_t* n_tokens_host, float* logp_dev_or_null,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* workspace bytes lcasr_model_transcribe_host needs (forward workspace + input/token staging) */
int64_t lcasr_model_transcribe_workspace_bytes(const lcasr_model* m, int B, int64_t T);

/* ------------------------------------------------------------------------------------------
 * Sequence-parallel forward (BASELINE configs 3 and 4): ONE recording, its tokens split into
 * `world` contiguous blocks, one per rank (one process per GPU).  The reference has no multi-GPU path
 * (single GPU, exp/gn.sh:5); what these replace is SCConformerXL.forward (sconformer_xl.py:162-252) for
 * contexts whose attention (attention.py:509-551) is worth spreading over an NVLink domain.
 * ---------------------------------------------------------------------------------------- */

/* One (query block x key block) term of attention: out32 [B,Nq,H*Dh] fp32 = softmax over THIS key block only,
 * lse [B,H,Nq] fp32 = log2-domain log-sum-exp of its scaled scores (bf16 operands, tcgen05 kernel). */
int lcasr_attention_partial(const void* q, const void* k, const void* v, int B, int64_t Nq, int64_t Nk, int H, int Dh,
                            float* out32, float* lse, void* stream);
/* Exact combination of P such terms (parts [P][rows,H*Dh], lses [P][H,rows], rows = B*Nq with B == 1):
 * out = sum_s 2^(lse_s - L) parts_s, L = log2 sum_s 2^lse_s — softmax over the union of the key blocks. */
int lcasr_attention_merge(const float* parts, const float* lses, int P, int64_t rows, int H, int Dh, void* out,
                          int out_dtype, void* stream);

/* NCCL communicator owned by this library (NCCL is bound at run time: dlopen of the libnccl.so.2 already mapped
 * by the host process, or `nccl_path`).  Rank 0 calls lcasr_comm_unique_id (128 bytes, host) and distributes the
 * id by its own means (torch.distributed broadcast in the Python host); every rank then calls lcasr_comm_create
 * (collective).  The communicator also owns the side streams of the sequence-parallel forward. */
typedef struct lcasr_comm lcasr_comm;
int lcasr_comm_unique_id(char* id_out_host, const char* nccl_path_or_null);
int lcasr_comm_create(const char* id_host, int rank, int world, const char* nccl_path_or_null, lcasr_comm** out);
void lcasr_comm_destroy(lcasr_comm* c);

/* token block [start, start+n) of `rank` for a T-frame recording (T % 8 == 0), and the per-rank workspace */
int lcasr_model_seqpar_block(const lcasr_model* m, int world, int rank, int64_t T, int64_t* start_tok, int64_t* n_tok);
int64_t lcasr_model_seqpar_workspace_bytes(const lcasr_model* m, int world, int rank, int64_t T);

/* One rank of the forward.  spec_full [1,feat_in,T] fp32 (only this rank's slice + 8 frames of left context are
 * read); out_local [n_rank, num_classes] fp32; argmax_full [T/8] int32 or NULL (ids of the WHOLE recording on
 * every rank, for the greedy collapse).  K/V blocks move as ncclSend/ncclRecv pairs in ring order on a side
 * stream while attention runs on the blocks already present; the conv module exchanges (k-1)/2 halo rows with
 * the two neighbours.  Asynchronous on `stream`. */
int lcasr_model_forward_seqpar(lcasr_model* m, lcasr_comm* comm, const float* spec_full, int64_t T, float* out_local,
                               int32_t* argmax_full, int return_logits, void* workspace, int64_t workspace_bytes,
                               void* stream);

/* The same per-rank phases for all `world` ranks in ONE process on one GPU (device-to-device copies instead of
 * NCCL transfers): the parity check that needs no second GPU.  out_full [T/8, num_classes]. */
int64_t lcasr_model_seqpar_emulated_workspace_bytes(const lcasr_model* m, int world, int64_t T);
int lcasr_model_forward_seqpar_emulated(lcasr_model* m, int world, const float* spec_full, int64_t T, float* out_full,
                                        int32_t* argmax_full, int return_logits, void* workspace,
                                        int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCASR_B200_H_ */
