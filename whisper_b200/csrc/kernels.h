// Internal C++ launch interface shared by the .cu translation units and api.cu.
// Nothing here is exported; the C-ABI lives in include/whisper_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

namespace wb {

enum { DT_BF16 = 0, DT_F16 = 1 };

// kernels launched by this library since load (reported by bench.py as gpu_launches)
extern unsigned long long g_launch_count;
// Launch accounting (wb200_launch_count).  The process-wide counter is updated atomically because several host
// threads may drive decoder sessions at once; the thread-local one lets a stream capture
// subtract exactly the launches IT recorded (a capture records kernels, it does not run them).
extern thread_local unsigned long long t_launch_count;
inline void count_launch(int n = 1) {
  __atomic_fetch_add(&g_launch_count, static_cast<unsigned long long>(n), __ATOMIC_RELAXED);
  t_launch_count += static_cast<unsigned long long>(n);
}

// Optional per-kernel timing for bench.py's roofline: when profiling is enabled for a kernel id,
// every launch of that kernel is bracketed by CUDA events on its own stream (api.cu owns the pool).
enum { PROF_NONE = 0, PROF_CROSS_ATTN = 1, PROF_SELF_ATTN = 2, PROF_GEMM = 3, PROF_ENC_ATTN = 4,
       PROF_LAYERNORM = 5, PROF_SELECT = 6, PROF_MEL = 7, PROF_DEC_LAYER = 8 };
extern int g_profile_kernel;
void profile_mark(cudaStream_t s, bool begin);
struct ProfileScope {
  cudaStream_t s;
  bool on;
  ProfileScope(int id, cudaStream_t st) : s(st), on(g_profile_kernel == id) {
    if (on) profile_mark(s, true);
  }
  ~ProfileScope() {
    if (on) profile_mark(s, false);
  }
};

// C[b*rows_per_batch + t, n] = epilogue( sum_tap sum_k A_tap[b, t + a_row_off[tap], k] * W[n, tap*K_tap + k] )
// A_tap is the 3-D view {K_tap, a_rows_per_batch, batch} at A + a_base_off[tap] with row stride lda and
// batch stride a_batch_stride (elements).  Rows outside [0, a_rows_per_batch) read as zero (TMA OOB fill),
// which is exactly Conv1d's zero padding (reference model.py:53-59, 193-194).
struct LinearArgs {
  int dtype = DT_BF16;
  int batch = 1;
  int rows_per_batch = 0;
  int a_rows_per_batch = 0;
  long long a_batch_stride = 0;
  long long lda = 0;
  int N = 0, K_tap = 0, taps = 1;
  int a_row_off[3] = {0, 0, 0};
  long long a_base_off[3] = {0, 0, 0};
  const void* A = nullptr;
  const void* W = nullptr;
  long long ldw = 0;
  const void* bias = nullptr;      // T[N] or null
  const void* residual = nullptr;  // T[rows, ldr] or null (may alias C)
  long long ldr = 0;
  const float* pos = nullptr;      // fp32 [rows_per_batch, N] added per batch row, or null
  // head-major output (16-bit, N % 64 == 0): element (row, n) of the [rows, N] result is stored at
  // C[((row / hm_T) * (N / 64) + n / 64) * hm_T + row % hm_T][n % 64] - i.e. [batch][head][t][64] for rows = batch * hm_T.
  // Used for the cross-attention K/V of the decoder when the head-major kv layout is on.  0 = row-major.
  int head_major_T = 0;
  void* C = nullptr;
  long long ldc = 0;
  int gelu = 0;
  int out_f32 = 0;
  int block_n = 0;  // 0 = choose (256 for large M, 64 for skinny M)
  int block_m = 0;  // 0 = choose (128; 64 for skinny plain GEMMs)
  const int* skip_flag = nullptr;  // optional device int: non-zero -> the kernel is a no-op
  // optional split-K scratch (decode-step GEMMs): fp32 slabs + per-tile tickets (zero between launches)
  float* splitk_ws = nullptr;
  size_t splitk_ws_bytes = 0;
  int* splitk_counters = nullptr;
  int splitk_max_tiles = 0;
};
int launch_linear(const LinearArgs& a, cudaStream_t s);
extern int g_splitk_on;
extern int g_bm64_on;
extern int g_xattn_tma;       // step-mode cross attention through the persistent TMA kernel (wb200_set_cross_attention_tma)
extern int g_kv_head_major;   // kv caches stored [.., head, position, 64] instead of [.., position, d] (wb200_set_kv_head_major)
extern int g_pdl_on;   // programmatic dependent launch for the decoder-layer kernels (wb200_set_pdl / WB200_PDL)

// Per-DEVICE caches of launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device
// only and SM counts differ between devices, so "done once" flags are indexed by the device ordinal: a second model
// on cuda:1 in the same process makes its own opt-ins.  Races between host threads write identical values.
constexpr int kMaxDevices = 64;
inline int device_ordinal() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
struct SmemOptIn {
  int bytes[kMaxDevices] = {};
  template <typename K>
  bool ensure(K kern, int want) {      // false on failure
    const int d = device_ordinal();
    if (bytes[d] >= want) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, want) != cudaSuccess) return false;
    bytes[d] = want;
    return true;
  }
};
inline int sm_count() {
  static int n[kMaxDevices] = {};
  const int d = device_ordinal();
  if (!n[d]) cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
  return n[d];
}

// Launch with the programmatic-stream-serialization attribute (when enabled): inside a stream or a
// captured graph the kernel may be scheduled while the previous kernel drains; the kernels that are
// launched this way all call pdl_wait() before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  if (g_pdl_on < 0) {
    const char* e = getenv("WB200_PDL");
    g_pdl_on = (e && e[0] == '0') ? 0 : 1;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl_on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Row LayerNorm in fp32 (reference model.py:39-41): y = (x - mean) / sqrt(var + 1e-5) * g + b
int launch_layernorm(int dtype, const void* x, long long ldx, void* y, long long ldy, const float* g,
                     const float* b, int rows, int d, cudaStream_t s, const int* skip_flag = nullptr);

// (B, C, T) fp32 -> (B, T, C) 16-bit.  Used to put the mel spectrogram time-major for the conv GEMM.
int launch_transpose_to16(int dtype, const float* x, void* y, int B, int C, int T, cudaStream_t s);

// Encoder (non-causal) multi-head attention over packed qkv [B*T, 3*d] -> out [B*T, d]
int launch_enc_attention(int dtype, const void* qkv, void* out, int B, int T, int n_head,
                         cudaStream_t s);

// ---- decoder attention (dec_attention.cu)
int cross_attention_splits(int T, int n_groups);
size_t cross_attention_partial_floats(int n_audio, int n_q, int n_head, int T);
// q: [n_audio*n_q, d]; k/v: [n_audio, T, kv_ld] row stride kv_ld elements; out: [n_audio*n_q, d]
int launch_cross_attention(int dtype, const void* q, const void* k, const void* v, void* out,
                           float* partial, int* counters, const int* skip_flag, int n_audio, int n_q,
                           int T, int n_head, int kv_ld, cudaStream_t s, int head_major = 0);
// Step-mode cross attention fed by TMA from ONE layer's head-major K/V block [n_audio][2H][T][64] (n_q <= 16);
// -1: shape not covered, use launch_cross_attention.
int launch_cross_attention_tma(int dtype, const void* q, const void* kv, void* out, float* partial, int* counters,
                               const int* skip_flag, int n_audio, int n_q, int T, int n_head, cudaStream_t s);
// Step-mode self attention of all beams of an audio together (beam-window kv layout, 2 <= group <= 8, enough
// (audio, head) items to fill the SMs - self_attention_tma_covers, which the session consults when it fixes the cache
// layout); appends the new K / V.
bool self_attention_tma_covers(int n_audio, int G, int n_head, int max_ctx);
int launch_self_attention_tma(int dtype, const void* qkv, void* kcache, void* vcache, void* out, const int* indir,
                              const int* len_ptr, const int* skip_flag, int n_audio, int G, int n_head, int max_ctx,
                              cudaStream_t s);
extern int g_sattn_tma;       // wb200_set_self_attention_tma / WB200_SATTN_TMA=1 (default off: measured slower)
// step mode (indir != null): one new position per row, appended to the cache; prefill mode
// (indir == null): n_init positions per audio, causal, cache rows a*group.
int launch_self_attention(int dtype, const void* qkv, void* kcache, void* vcache, void* out,
                          const int* indir, const int* len_ptr, const int* skip_flag, int n_rows,
                          int n_head, int max_ctx, int n_init, int group, cudaStream_t s, int head_major = 0);   // 2: beam window

// ---- token selection (select.cu)
struct FilterParams {
  const float* logits;      // [n_logit_rows, ld]
  long long ld;
  int V;
  int row_div;              // row r reads logits row r / row_div (G on the first step, else 1)
  const int* tokens;        // [R, max_ctx] current sequences
  int max_ctx;
  const int* len_ptr;       // device: current length L
  const int* skip_flag;
  const uint32_t* suppress_mask;  // V bits: SuppressTokens set (+ no_timestamps when rules are on)
  const uint32_t* blank_mask;     // V bits: SuppressBlank set (" " ids + eot)
  int sample_begin, eot, timestamp_begin, max_initial_ts;  // max_initial_ts < 0: none
  int suppress_blank, ts_rules;
  int K;
  float* top_val;           // [R, K] log-probabilities, best first
  int* top_idx;             // [R, K]
  // temperature sampling (K == 1 only): inv_temp = 1 / temperature, 0 = argmax.  The sample is the Gumbel-max
  // argmax_v( logit_v * inv_temp + g_v ), g_v = -log(-log(u_v)), u_v from Philox4x32-10 (see select.cu).
  float inv_temp;
  uint32_t seed_lo, seed_hi;
};
struct GreedyParams {
  int* tokens;            // [R, max_ctx], appended in place
  int max_ctx, R, eot;
  int* len_ptr;           // incremented
  float* sum_logprobs;    // [R]
  const float* top_val;   // [R, 1]
  const int* top_idx;
  int* done_flag;         // set when every row's last token is EOT
  const int* skip_flag;
};
struct BeamParams {
  const int* tokens_in;   // [R, max_ctx]
  int* tokens_out;        // [R, max_ctx]
  const int* indir_in;    // [R, max_ctx]
  int* indir_out;
  int max_ctx, n_audio, G, eot, max_candidates;
  int* len_ptr;
  float* sum_logprobs;    // [R] in/out
  const float* top_val;   // [R, G+1]
  const int* top_idx;
  int* fin_tokens;        // [n_audio, max_candidates, max_ctx]
  int* fin_len;           // [n_audio, max_candidates]
  float* fin_score;       // [n_audio, max_candidates]
  int* fin_count;         // [n_audio]
  int* source_out;        // [R] parent row of each new beam (diagnostics / parity tests)
  int* done_flag;
  const int* skip_flag;
  int* cur_out_ptr;       // device int: receives out_index (which ping-pong buffer is current)
  int out_index;
  int n_init;             // prompt length: positions < n_init were written by the prefill
  int* tickets;           // device int[2], zero between launches: arrival ticket, count of full audios
  // [n_audio][16][16] bytes: same_in[a][j1][j2] != 0 iff beams j1, j2 of audio a hold the same token prefix (the keys
  // of the reference's candidate dict, decoding.py:344); same_out receives the matrix of the new beams
  const unsigned char* same_in;
  unsigned char* same_out;
};
int launch_filter_topk(const FilterParams& p, int R, cudaStream_t s);
int launch_no_speech(const float* logits, long long ld, int V, int no_speech, float* out, int rows,
                     int row_stride, int row_offset, cudaStream_t s);
int launch_range_softmax(const float* logits, long long ld, int first, int n, int rows, float* probs, int* argmax,
                         const int* gather_tok, float* gather_out, cudaStream_t s);
int launch_greedy_update(const GreedyParams& p, cudaStream_t s);
int launch_beam_update(const BeamParams& p, cudaStream_t s);

// ---- audio front-end (mel.cu)
size_t log_mel_workspace_bytes(int n_audio);
int launch_log_mel(const float* audio, int n_audio, long long n_samples, int n_mels, const float* filters,
                   float* out, void* workspace, int per_row_max, cudaStream_t s);

// ---- word timing (timing.cu)
// qk[i][t] = (q_i . k_t) / 8 for one head: q rows [n_q, ldq], k rows [T, ldk] (16-bit), out fp32 [n_q, T]
int launch_qk_export(int dtype, const void* q, long long ldq, const void* k, long long ldk, float* out,
                     int n_q, int T, cudaStream_t s);
int launch_alignment_weights(const float* qk, int n_heads, int n_tokens, int t_stride, int n_frames,
                             float qk_scale, int medfilt_width, int negate, float* out, float* scratch,
                             cudaStream_t s);
int launch_median_filter(const float* x, float* y, long long rows, int T, int width, cudaStream_t s);
size_t dtw_workspace_bytes(int N, int M);
int launch_dtw(const float* x, int N, int M, int* path, int* path_len, void* workspace, int tie_mode,
               cudaStream_t s);

}  // namespace wb
