/*
 * whisper_b200 - C ABI of the B200-native Whisper inference hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (passed as void* so the
 * header needs no CUDA include), returns 0 on success and a non-zero status otherwise
 * (wb200_last_error() gives the text).  No C++ exceptions cross this boundary, the library never
 * frees caller memory, and it keeps no reference to caller buffers after a call returns except
 * through the explicit handles (wb200_model, wb200_decoder) documented below.
 *
 * The reference (openai/whisper) has no FFI of its own - its operator seam is Python duck typing
 *  - so each function below names the reference call site it replaces
 * (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB200_DTYPE_BF16 0
#define WB200_DTYPE_F16 1

#define WB200_OK 0

/* ---------------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------------- */
const char* wb200_version(void);
const char* wb200_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t wb200_launch_count(void);

/* Per-kernel device timing for roofline reports: after wb200_profile_enable(id) every launch of
 * that kernel is bracketed by CUDA events on its own stream; wb200_profile_read() waits for them,
 * returns the summed duration and launch count, and clears the record.  id 0 disables. */
#define WB200_KERNEL_CROSS_ATTENTION 1
#define WB200_KERNEL_SELF_ATTENTION 2
#define WB200_KERNEL_GEMM 3
#define WB200_KERNEL_ENCODER_ATTENTION 4
#define WB200_KERNEL_LAYERNORM 5
#define WB200_KERNEL_SELECT 6
#define WB200_KERNEL_LOG_MEL 7
#define WB200_KERNEL_DECODER_LAYER 8   /* fused decoder-layer GEMM chain (LayerNorm folded in), csrc/dec_layer.cu */
int wb200_profile_enable(int kernel_id);
int wb200_profile_read(double* total_ms, int64_t* launches);

/* ---------------------------------------------------------------------------------------------
 * primitive operators (each is a hand-written sm_100a kernel; exposed for parity tests)
 * ------------------------------------------------------------------------------------------- */

/* Linear: C[M,N] = A[M,K] @ W[N,K]^T (+bias) (GELU) (+residual); replaces whisper/model.py:44-50
 * (Linear.forward -> F.linear) and the MLP GELU at model.py:155-157.  A, W, bias, residual, C are
 * 16-bit `dtype`; C is fp32 when out_f32 != 0 (the `.float()` logits of model.py:245-247).
 * lda/ldw/ldr/ldc are row strides in elements.  residual may alias C. */
int wb200_linear(int dtype, int M, int N, int K, const void* A, int64_t lda, const void* W,
                 int64_t ldw, const void* bias, const void* residual, int64_t ldr, void* C,
                 int64_t ldc, int gelu, int out_f32, void* stream);

/* Split-K for the skinny GEMMs is opt-in (WB200_SPLITK=1 in the environment, or this call). */
int wb200_set_splitk(int enabled);
/* 64-row (UMMA M = 64) tiles for skinny GEMMs are on by default; 0 forces 128-row tiles. */
int wb200_set_bm64(int enabled);
/* Programmatic dependent launch for the decoder-layer kernels (layer norm, GEMM, self- / cross-attention): on by
 * default (WB200_PDL=0 in the environment or this call turns it off).  Each of those kernels may be scheduled while
 * its predecessor drains and waits (griddepcontrol.wait) before it touches global memory. */
int wb200_set_pdl(int enabled);
/* Fused decoder-layer kernel for the autoregressive step (ResidualAttentionBlock, whisper/model.py:142-171, as driven
 * per token by decoding.py:680-710): the six Linears and three LayerNorms of a layer run as three persistent launches
 * (QKV | out-proj + cross-query | cross-out + MLP + next layer's QKV) around the two attention kernels, LayerNorm folded
 * into the consuming Linear.  On by default for sessions created AFTER the call (WB200_FUSED_LAYER=0 in the environment
 * or this call turns it off; the unfused kernels then run, as they always do for the prefill). */
int wb200_set_fused_decoder_layer(int enabled);
/* Few-rows form of the fused decoder-layer kernel (sessions with n_audio * n_group <= 8 rows: one audio decoded
 * greedily as in whisper/transcribe.py:272-508, one audio with 5 beams): same phases and LayerNorm folding, the main
 * loop a weight-stationary matrix-vector product (every SM bulk-copies its own slice of the weight rows into shared
 * memory ahead of the grid barrier and runs mma.sync with the weight rows as the M operand).  On by default for
 * sessions created AFTER the call (WB200_FUSED_ROWS=0 in the environment or this call: the 64-row tile form runs). */
int wb200_set_fused_decoder_rows(int enabled);
/* Few-rows sessions with head-major kv caches: ONE launch per decoder iteration for the whole stack - the kernel above
 * walks a phase table that strings every layer's Linear chains together with its self-attention (kv append + attention
 * over the rows' lineages through the parent table, whisper/model.py:124-127,327-333 + decoding.py:172-176) and its
 * cross-attention (key slices + merge) as further grid-barrier phases.  mode 1 (default) for sessions created AFTER the
 * call; 0 (or WB200_FUSED_STACK=0): three few-rows launches per layer around the two attention kernels; 2: the table
 * also ends with the decoder's final LayerNorm and the logits (measured: identical tokens, no gain over the two
 * separate launches). */
int wb200_set_fused_decoder_stack(int mode);
/* Layout of the decoder's kv caches for sessions created AFTER the call (default 1, or WB200_KV_HEAD_MAJOR=0 in
 * the environment).  1: head-major - cross-attention K/V [n_audio, 2H, 1500, 64] (written that way by the K/V
 * projection's epilogue), self-attention caches [rows, H, 448, 64] - or, in sessions that run the beam-window
 * self-attention kernel (wb200_set_self_attention_tma), [n_audio, H, 448, n_group, 64]: the n_group rows of an audio
 * interleaved per position - so every (audio, head) streams one contiguous block, which is what the TMA attention
 * kernels need.  0: cross K/V [n_audio, 1500, 2d] and self caches [rows, 448, d], one head's 128 bytes per position
 * strided by the model width.  With the cp.async attention kernels the layouts give bit-identical results (measured
 * on B200: head-major +1 %, profiles/r2_ab_switches.txt). */
int wb200_set_kv_head_major(int enabled);
/* Decoder-step cross attention (whisper/model.py:101-109 + SDPA, one query per beam against the audio's 1500 cached
 * keys): 1 (default, WB200_XATTN_TMA=0 to disable) = persistent kernel, one CTA per SM, K/V tiles streamed by TMA
 * through an mbarrier ring that stays full across (audio, head) work items; needs the head-major layout.  0 = the
 * cp.async kernel (always used for the prefill).  Takes effect at the next launch. */
int wb200_set_cross_attention_tma(int enabled);
/* Decoder-step self attention under beam search / best_of (2 <= n_group <= 8): 1 (opt-in, or WB200_SATTN_TMA=1 in the
 * environment; measured 10 % slower than the default at the headline shape because it cannot profit from beams that
 * share ancestors, profiles/r2_summary.md) = the G rows of an audio are processed together by a persistent TMA-fed kernel that streams the audio's
 * whole (position x beam-slot) history of a head as one contiguous block of the head-major ("beam window") cache and
 * masks each row by the beam's parent table - the device form of PyTorchInference.rearrange_kv_cache
 * (whisper/decoding.py:172-176) + the kv-cache hooks' torch.cat (whisper/model.py:327-333).  0 (default) = one warp per
 * (row, head) gathering 128-byte pieces through the parent table (always used for greedy decoding, the prefill, the
 * position-major layout and batches with fewer (audio, head) pairs than half the SMs).  The cache layout goes with
 * the kernel, so the switch applies to sessions created AFTER the call. */
int wb200_set_self_attention_tma(int enabled);

/* Same operator with split-K enabled for skinny problems (the 320-row decode-step GEMMs): `workspace`
 * holds fp32 partial slabs (up to 8 * M * N floats are used), `tickets` is an int32 array of n_tickets
 * entries that must be zero on entry and is zero again on exit. */
int wb200_linear_splitk(int dtype, int M, int N, int K, const void* A, int64_t lda, const void* W,
                        int64_t ldw, const void* bias, const void* residual, int64_t ldr, void* C,
                        int64_t ldc, int gelu, int out_f32, void* workspace, size_t workspace_bytes,
                        int32_t* tickets, int n_tickets, void* stream);

/* Conv1d(kernel=3, padding=1, stride in {1,2}) + GELU on time-major activations; replaces
 * whisper/model.py:53-59 + F.gelu at model.py:193-194.  x: [B, T_in, C_in] 16-bit, w: [C_out, 3*C_in]
 * tap-major (w[o, k*C_in + c] = weight[o, c, k]), bias [C_out], pos (optional fp32 [T_out, C_out],
 * the sinusoid table added at model.py:198), y: [B, T_out, C_out], T_out = T_in / stride. */
int wb200_conv1d_k3_gelu(int dtype, int B, int T_in, int C_in, int C_out, int stride, const void* x,
                         const void* w, const void* bias, const float* pos, void* y, void* stream);

/* LayerNorm over the last dim in fp32, eps 1e-5; replaces whisper/model.py:39-41. */
int wb200_layernorm(int dtype, const void* x, void* y, const float* gamma, const float* beta,
                    int rows, int d, void* stream);

/* (B, C, T) fp32 -> (B, T, C) 16-bit: mel.half() of decoding.py:645-646 fused with the layout
 * change the conv GEMM wants. */
int wb200_transpose_to16(int dtype, const float* x, void* y, int B, int C, int T, void* stream);

/* Non-causal multi-head self-attention of the audio encoder; replaces qkv_attention at
 * whisper/model.py:114-139 (SDPA branch, scale 1/sqrt(64)).  qkv: packed [B*T, 3*d] (q | k | v),
 * out: [B*T, d]; head dim is 64. */
int wb200_encoder_attention(int dtype, const void* qkv, void* out, int B, int T, int n_head,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * audio front-end
 * ------------------------------------------------------------------------------------------- */

/* log_mel_spectrogram, whisper/audio.py:110-157 (after any right-padding, which the caller does by
 * allocating zeros): audio [n_audio, n_samples] fp32 -> out [n_audio, n_mels, n_samples / 160] fp32.
 * filters: the dense (n_mels x 201) fp32 mel matrix of audio.py:91-107 (device).  per_row_max = 0
 * reproduces the reference exactly (ONE max over the whole call, audio.py:155); 1 clamps each
 * waveform against its own max (== calling the reference once per waveform). */
size_t wb200_log_mel_workspace_bytes(int n_audio);
int wb200_log_mel(const float* audio, int n_audio, int64_t n_samples, int n_mels, const float* filters,
                  float* out, void* workspace, size_t workspace_bytes, int per_row_max, void* stream);

/* ---------------------------------------------------------------------------------------------
 * model handle: a table of caller-owned device tensors (weights already converted to `dtype`).
 * dims order = ModelDimensions of whisper/model.py:25-36:
 *   n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer,
 *   n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer
 * Tensor slots (T = 16-bit `dtype`, F = fp32), reference state-dict names in brackets:
 *   global (12): conv1.w T[d,3*n_mels] tap-major, conv1.b T, conv2.w T[d,3*d] tap-major, conv2.b T,
 *                enc.pos F[1500,d], ln_post.w F, ln_post.b F, tok_emb T[V,d], tok_emb F[V,d],
 *                dec.pos F[448,d], dec.ln.w F, dec.ln.b F
 *   per encoder layer (12): attn_ln.w F, attn_ln.b F, qkv.w T[3d,d] (query|key|value), qkv.b T[3d]
 *                (key part zero: model.py:88), out.w T, out.b T, mlp_ln.w F, mlp_ln.b F,
 *                fc1.w T[4d,d], fc1.b T, fc2.w T[d,4d], fc2.b T
 *   per decoder layer (29): attn_ln.w/b F, qkv.w T, qkv.b T, out.w T, out.b T, cross_ln.w/b F,
 *                cq.w T[d,d], cq.b T, ckv.w T[2d,d] (key|value), ckv.b T[2d] (key part zero),
 *                cout.w T, cout.b T, mlp_ln.w/b F, fc1.w T, fc1.b T, fc2.w T, fc2.b T,
 *                then the LayerNorm-folded forms used by the fused decoder-layer kernel, for each of
 *                (attn_ln -> qkv), (cross_ln -> cq), (mlp_ln -> fc1):  wf T = W * gamma (per input column),
 *                c1 F[out] = row sums of wf (as stored in T), c2 F[out] = W beta + bias:
 *                qkv.wf, qkv.c1, qkv.c2, cq.wf, cq.c1, cq.c2, fc1.wf, fc1.c1, fc1.c2
 * The handle stores the pointers only; the caller keeps the tensors alive until destroy.
 * Replaces the nn.Module state of whisper/model.py:252-276 (and the per-call weight casts of
 * model.py:44-59). */
typedef struct wb200_model wb200_model;
int wb200_model_num_tensors(const int32_t dims[10]);
int wb200_model_create(const int32_t dims[10], int dtype, const void* const* tensors, int n_tensors,
                       wb200_model** out);
void wb200_model_destroy(wb200_model* model);

/* AudioEncoder.forward, whisper/model.py:188-204: mel [n_audio, n_mels, 3000] fp32 ->
 * features [n_audio, 1500, d] 16-bit.  workspace: wb200_encoder_workspace_bytes(). */
size_t wb200_encoder_workspace_bytes(const wb200_model* model, int n_audio);
int wb200_encoder_forward(const wb200_model* model, const float* mel, int n_audio, void* features,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * decoder session: kv-cache, logit filters and token selection of whisper/decoding.py:144-176
 * (PyTorchInference), :272-404 (GreedyDecoder / BeamSearchDecoder), :423-505 (logit filters) and the
 * body of DecodingTask._main_loop (:680-710), resident on the device.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_audio;                      /* segments decoded together */
  int32_t n_group;                      /* beam_size (beam search) or 1 (greedy)   decoding.py:527 */
  int32_t beam_search;                  /* 1: BeamSearchDecoder, 0: GreedyDecoder (temperature 0) */
  int32_t max_candidates;               /* round(beam_size * patience)             decoding.py:313 */
  int32_t n_init;                       /* len(initial_tokens)                     decoding.py:535 */
  int32_t sample_begin;                 /* == n_init                               decoding.py:536 */
  int32_t sot_index;                    /* position of <|startoftranscript|>       decoding.py:537 */
  int32_t eot, no_speech, no_timestamps, timestamp_begin;   /* ids; no_speech < 0: none */
  int32_t suppress_blank;               /* SuppressBlank installed                 decoding.py:555 */
  int32_t timestamp_rules;              /* ApplyTimestampRules installed           decoding.py:559 */
  int32_t max_initial_timestamp_index;  /* < 0: none                               decoding.py:561 */
  int32_t n_suppress, n_blank;
  int32_t all_logits;                   /* 1: the prefill keeps logits of ALL n_init positions (the
                                           un-cached forward of model.py:293-296); such a session is
                                           forward-only (no select / step) */
  const int32_t* suppress_ids;          /* host: sorted SuppressTokens ids         decoding.py:615 */
  const int32_t* blank_ids;             /* host: tokenizer.encode(" ")             decoding.py:430 */
} wb200_decode_config;

typedef struct wb200_decoder wb200_decoder;
size_t wb200_decoder_workspace_bytes(const wb200_model* model, const wb200_decode_config* cfg);
int wb200_decoder_create(const wb200_model* model, const wb200_decode_config* cfg, void* workspace,
                         size_t workspace_bytes, wb200_decoder** out, void* stream);
void wb200_decoder_destroy(wb200_decoder* dec);
/* cross-attention K/V of every layer from features [n_audio, 1500, d] (model.py:104-109) */
int wb200_decoder_set_audio(wb200_decoder* dec, const void* features, void* stream);
/* reset the session and run the n_init-token first pass (decoding.py:687-696 at i == 0):
 * initial_tokens: HOST int32 [n_audio, n_init].  Leaves the logits of the last prompt position
 * current and stores no_speech_prob per audio. */
int wb200_decoder_prefill(wb200_decoder* dec, const int32_t* initial_tokens, void* stream);
/* logit filters + GreedyDecoder.update / BeamSearchDecoder.update on the current logits
 * (decoding.py:699-703), including the beam kv-cache reorder (a parent-table update). */
int wb200_decoder_select(wb200_decoder* dec, void* stream);
/* GreedyDecoder with a temperature (decoding.py:283: Categorical(logits / temperature).sample(); n_group =
 * best_of independent samples per audio, decoding.py:524-526).  temperature 0 (the default) is argmax.  The
 * reference draws from torch's global generator, which no other implementation can replay; this library draws
 * by Gumbel-max with a counter-based generator so that a (seed, row, step) triple fully determines the sample:
 *     words  = Philox4x32-10(counter = (v >> 2, row, L, 0), key = (seed & 0xffffffff, seed >> 32))
 *     u_v    = ((words[v & 3] >> 8) + 0.5) * 2^-24,   g_v = -log(-log(u_v))            (fp32)
 *     token  = argmax_v( logit_v / temperature + g_v )   over the tokens the filters leave, ties to the lower id
 * with L the number of tokens already in the row.  sum_logprobs accumulates the UN-tempered log-probability of
 * the drawn token (decoding.py:285-287).  Only for greedy (non-beam) sessions; call before select / run. */
int wb200_decoder_set_sampling(wb200_decoder* dec, float temperature, uint64_t seed);
/* one TextDecoder step on the last token of every row with kv-cache append (decoding.py:687) */
int wb200_decoder_step(wb200_decoder* dec, void* stream);
/* up to max_steps x (step, select); stops early once the device-side completion flag is seen
 * (decoding.py:705).  steps_issued (host, optional) receives the number of iterations launched. */
int wb200_decoder_run(wb200_decoder* dec, int max_steps, int32_t* steps_issued, void* stream);
/* teacher forcing for parity tests: append HOST int32 tokens [n_audio*n_group] instead of selecting */
int wb200_decoder_force_tokens(wb200_decoder* dec, const int32_t* next_tokens, void* stream);

/* find_alignment support (timing.py:185-197): before wb200_decoder_prefill, ask for the PRE-softmax
 * cross-attention scores (q k^T / 8, fp32) of selected heads.  heads: HOST int32 [n_heads][2] =
 * (layer, head); qk_out: device fp32 [n_heads, n_init, 1500] of audio 0, filled by the next prefill.
 * n_heads = 0 cancels. */
int wb200_decoder_set_alignment(wb200_decoder* dec, const int32_t* heads, int n_heads, float* qk_out);

/* timing.py:207-214 on the exported scores: softmax over the first n_frames frames of qk * qk_scale,
 * z-score over the token axis (population std), median filter of odd `medfilt_width` along frames,
 * mean over heads.  qk: [n_heads, n_tokens, t_stride]; out: [n_tokens, n_frames] (negated when
 * `negate` != 0, ready for DTW); scratch: 2 * n_heads * n_tokens * n_frames floats. */
int wb200_alignment_weights(const float* qk, int n_heads, int n_tokens, int t_stride, int n_frames,
                            float qk_scale, int medfilt_width, int negate, float* out, float* scratch,
                            void* stream);

/* state access: copies between the session and caller memory (host or device), asynchronously on
 * `stream`.  Element types: int32 except LOGITS / SUM_LOGPROBS / NO_SPEECH / TOP_VAL / FIN_SCORE (fp32). */
#define WB200_STATE_TOKENS 0        /* [R, n_text_ctx]                                   */
#define WB200_STATE_LENGTH 1        /* [1]                                               */
#define WB200_STATE_SUM_LOGPROBS 2  /* [R]                                               */
#define WB200_STATE_NO_SPEECH 3     /* [n_audio]                                         */
#define WB200_STATE_LOGITS 4        /* [R (or n_audio after prefill), ld]; ld = logits row stride */
#define WB200_STATE_TOP_VAL 5       /* [R, K]                                            */
#define WB200_STATE_TOP_IDX 6       /* [R, K]                                            */
#define WB200_STATE_SOURCES 7       /* [R] beam parents of the last select                */
#define WB200_STATE_FIN_TOKENS 8    /* [n_audio, max_candidates, n_text_ctx]              */
#define WB200_STATE_FIN_LEN 9       /* [n_audio, max_candidates]                          */
#define WB200_STATE_FIN_SCORE 10    /* [n_audio, max_candidates]                          */
#define WB200_STATE_FIN_COUNT 11    /* [n_audio]                                          */
#define WB200_STATE_DONE 12         /* [1]                                               */
int64_t wb200_decoder_logits_ld(const wb200_decoder* dec);
int wb200_decoder_get_state(wb200_decoder* dec, int what, void* dst, size_t bytes, void* stream);
int wb200_decoder_set_state(wb200_decoder* dec, int what, const void* src, size_t bytes, void* stream);

/* softmax over the token range [first, first + n) of each of `rows` fp32 logit rows (row stride ld), everything outside
 * the range counting as masked.  Replaces the mask / argmax / softmax of detect_language (whisper/decoding.py:60-66,
 * range = the language tokens) and the softmax + gather of find_alignment (whisper/timing.py:198-201, range = [0, eot)).
 * Optional outputs (null to skip): probs [rows, n]; argmax [rows] token ids (ties to the lower id); gather_probs [rows] =
 * probability of gather_tokens[row]. */
int wb200_range_softmax(const float* logits, int64_t ld, int first, int n, int rows, float* probs, int32_t* argmax,
                        const int32_t* gather_tokens, float* gather_probs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * word timing
 * ------------------------------------------------------------------------------------------- */
/* median_filter, whisper/timing.py:19-54 (+ triton_ops.py:43-117): x, y [rows, T] fp32, reflect
 * padding of width/2, odd width <= 21, T > width/2 (shorter inputs are returned unchanged by the
 * caller, timing.py:22-24). */
int wb200_median_filter(const float* x, float* y, int64_t rows, int T, int width, void* stream);
/* dtw, whisper/timing.py:82-151 (+ triton_ops.py:13-40): x [N, M] fp32 cost matrix (device) ->
 * path [2, N + M + 1] int32 (device; row 0 text indices, row 1 time indices, first *path_len
 * entries valid).  tie_mode 0 = the rule the reference applies to CUDA tensors (triton_ops.py:38-40),
 * 1 = its CPU rule (timing.py:95-100). */
size_t wb200_dtw_workspace_bytes(int N, int M);
int wb200_dtw(const float* x, int N, int M, int32_t* path, int32_t* path_len, void* workspace,
              size_t workspace_bytes, int tie_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
