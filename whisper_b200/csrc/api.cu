// C-ABI entry points (include/whisper_b200.h).  Thin argument checking + dispatch to the
// kernels; no torch types, no exceptions, status codes only.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/whisper_b200.h"
#include "engine.h"
#include "kernels.h"

#include <vector>

namespace wb {
unsigned long long g_launch_count = 0;
thread_local unsigned long long t_launch_count = 0;
int g_profile_kernel = 0;
static std::vector<cudaEvent_t> g_prof_events;   // begin/end pairs
static size_t g_prof_used = 0;
void profile_mark(cudaStream_t s, bool begin) {
  (void)begin;
  if (g_prof_used == g_prof_events.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof_events.push_back(e);
  }
  cudaEventRecord(g_prof_events[g_prof_used++], s);
}
static thread_local char g_err[512] = "";
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace wb

using namespace wb;

#define WB_CHECK_DTYPE(dt) \
  if ((dt) != WB200_DTYPE_BF16 && (dt) != WB200_DTYPE_F16) return set_error(100, "bad dtype %d", (dt))

extern "C" {

const char* wb200_version(void) { return "whisper_b200 0.1 (sm_100a)"; }
const char* wb200_last_error(void) { return g_err; }
uint64_t wb200_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }

int wb200_profile_enable(int kernel_id) {
  g_profile_kernel = kernel_id;
  g_prof_used = 0;
  return 0;
}

int wb200_profile_read(double* total_ms, int64_t* launches) {
  double tot = 0.0;
  int64_t n = 0;
  for (size_t i = 0; i + 1 < g_prof_used; i += 2) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_prof_events[i + 1]) != cudaSuccess) return set_error(140, "profile: event sync failed");
    if (cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]) != cudaSuccess)
      return set_error(141, "profile: elapsed time failed");
    tot += ms;
    ++n;
  }
  g_prof_used = 0;
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return 0;
}

int wb200_linear(int dtype, int M, int N, int K, const void* A, int64_t lda, const void* W,
                 int64_t ldw, const void* bias, const void* residual, int64_t ldr, void* C,
                 int64_t ldc, int gelu, int out_f32, void* stream) {
  WB_CHECK_DTYPE(dtype);
  if (M < 0 || N <= 0 || K <= 0) return set_error(101, "wb200_linear: bad shape M=%d N=%d K=%d", M, N, K);
  if (M == 0) return 0;
  LinearArgs a;
  a.dtype = dtype;
  a.batch = 1;
  a.rows_per_batch = M;
  a.a_rows_per_batch = M;
  a.lda = lda;
  a.N = N;
  a.K_tap = K;
  a.taps = 1;
  a.A = A;
  a.W = W;
  a.ldw = ldw;
  a.bias = bias;
  a.residual = residual;
  a.ldr = ldr;
  a.C = C;
  a.ldc = ldc;
  a.gelu = gelu;
  a.out_f32 = out_f32;
  int r = launch_linear(a, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_linear: launch failed (%d): %s", r, cudaGetErrorString(cudaGetLastError())) : 0;
}

int wb200_set_splitk(int enabled) {
  g_splitk_on = enabled ? 1 : 0;
  return 0;
}

int wb200_set_bm64(int enabled) {
  g_bm64_on = enabled ? 1 : 0;
  return 0;
}

int wb200_set_kv_head_major(int enabled) {
  g_kv_head_major = enabled ? 1 : 0;
  return 0;
}

int wb200_set_self_attention_tma(int enabled) {
  g_sattn_tma = enabled ? 1 : 0;
  return 0;
}

int wb200_set_cross_attention_tma(int enabled) {
  g_xattn_tma = enabled ? 1 : 0;
  return 0;
}

int wb200_set_fused_decoder_layer(int enabled) {
  g_fused_layer = enabled ? 1 : 0;
  return 0;
}

int wb200_set_fused_decoder_rows(int enabled) {
  g_fused_rows = enabled ? 1 : 0;
  return 0;
}

int wb200_set_fused_decoder_stack(int mode) {
  g_fused_stack = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return 0;
}

int wb200_set_pdl(int enabled) {
  g_pdl_on = enabled ? 1 : 0;
  return 0;
}

int wb200_linear_splitk(int dtype, int M, int N, int K, const void* A, int64_t lda, const void* W,
                        int64_t ldw, const void* bias, const void* residual, int64_t ldr, void* C,
                        int64_t ldc, int gelu, int out_f32, void* workspace, size_t workspace_bytes,
                        int32_t* tickets, int n_tickets, void* stream) {
  WB_CHECK_DTYPE(dtype);
  if (M <= 0 || N <= 0 || K <= 0) return set_error(101, "wb200_linear_splitk: bad shape M=%d N=%d K=%d", M, N, K);
  LinearArgs a;
  a.dtype = dtype;
  a.batch = 1;
  a.rows_per_batch = M;
  a.a_rows_per_batch = M;
  a.lda = lda;
  a.N = N;
  a.K_tap = K;
  a.taps = 1;
  a.A = A;
  a.W = W;
  a.ldw = ldw;
  a.bias = bias;
  a.residual = residual;
  a.ldr = ldr;
  a.C = C;
  a.ldc = ldc;
  a.gelu = gelu;
  a.out_f32 = out_f32;
  a.splitk_ws = static_cast<float*>(workspace);
  a.splitk_ws_bytes = workspace_bytes;
  a.splitk_counters = tickets;
  a.splitk_max_tiles = n_tickets;
  int r = launch_linear(a, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_linear_splitk: launch failed (%d): %s", r, cudaGetErrorString(cudaGetLastError())) : 0;
}

int wb200_conv1d_k3_gelu(int dtype, int B, int T_in, int C_in, int C_out, int stride, const void* x,
                         const void* w, const void* bias, const float* pos, void* y, void* stream) {
  WB_CHECK_DTYPE(dtype);
  if (stride != 1 && stride != 2) return set_error(102, "wb200_conv1d_k3_gelu: stride must be 1 or 2");
  if (stride == 2 && (T_in & 1)) return set_error(102, "wb200_conv1d_k3_gelu: odd T_in with stride 2");
  if (B <= 0) return 0;
  LinearArgs a;
  a.dtype = dtype;
  a.batch = B;
  a.N = C_out;
  a.K_tap = C_in;
  a.taps = 3;
  a.A = x;
  a.W = w;
  a.ldw = 3LL * C_in;
  a.bias = bias;
  a.pos = pos;
  a.C = y;
  a.ldc = C_out;
  a.gelu = 1;
  a.a_batch_stride = static_cast<long long>(T_in) * C_in;
  if (stride == 1) {
    // out[t] = sum_k W_k x[t + k - 1]
    a.rows_per_batch = T_in;
    a.a_rows_per_batch = T_in;
    a.lda = C_in;
    a.a_row_off[0] = -1;
    a.a_row_off[1] = 0;
    a.a_row_off[2] = 1;
  } else {
    // out[t] = sum_k W_k x[2t + k - 1]; view x as rows of pairs (even | odd), row stride 2*C_in:
    //   k=0 -> odd row t-1, k=1 -> even row t, k=2 -> odd row t
    a.rows_per_batch = T_in / 2;
    a.a_rows_per_batch = T_in / 2;
    a.lda = 2LL * C_in;
    a.a_base_off[0] = C_in;
    a.a_row_off[0] = -1;
    a.a_base_off[1] = 0;
    a.a_row_off[1] = 0;
    a.a_base_off[2] = C_in;
    a.a_row_off[2] = 0;
  }
  int r = launch_linear(a, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_conv1d_k3_gelu: launch failed (%d): %s", r, cudaGetErrorString(cudaGetLastError())) : 0;
}

int wb200_layernorm(int dtype, const void* x, void* y, const float* gamma, const float* beta,
                    int rows, int d, void* stream) {
  WB_CHECK_DTYPE(dtype);
  int r = launch_layernorm(dtype, x, d, y, d, gamma, beta, rows, d, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_layernorm: failed (%d)", r) : 0;
}

int wb200_transpose_to16(int dtype, const float* x, void* y, int B, int C, int T, void* stream) {
  WB_CHECK_DTYPE(dtype);
  int r = launch_transpose_to16(dtype, x, y, B, C, T, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_transpose_to16: failed (%d)", r) : 0;
}

int wb200_encoder_attention(int dtype, const void* qkv, void* out, int B, int T, int n_head,
                            void* stream) {
  WB_CHECK_DTYPE(dtype);
  int r = launch_enc_attention(dtype, qkv, out, B, T, n_head, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_encoder_attention: failed (%d): %s", r, cudaGetErrorString(cudaGetLastError())) : 0;
}

size_t wb200_log_mel_workspace_bytes(int n_audio) { return log_mel_workspace_bytes(n_audio); }

int wb200_log_mel(const float* audio, int n_audio, int64_t n_samples, int n_mels, const float* filters,
                  float* out, void* workspace, size_t workspace_bytes, int per_row_max, void* stream) {
  if (n_mels != 80 && n_mels != 128) return set_error(110, "Unsupported n_mels: %d", n_mels);  // audio.py:103
  if (workspace_bytes < log_mel_workspace_bytes(n_audio)) return set_error(111, "wb200_log_mel: workspace too small");
  int r = launch_log_mel(audio, n_audio, n_samples, n_mels, filters, out, workspace, per_row_max,
                         static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_log_mel: failed (%d): %s", r, cudaGetErrorString(cudaGetLastError())) : 0;
}

struct wb200_model { Model m; };
struct wb200_decoder { Decoder* d; };

static Dims dims_from(const int32_t d[10]) {
  Dims o;
  o.n_mels = d[0]; o.n_audio_ctx = d[1]; o.n_audio_state = d[2]; o.n_audio_head = d[3]; o.n_audio_layer = d[4];
  o.n_vocab = d[5]; o.n_text_ctx = d[6]; o.n_text_state = d[7]; o.n_text_head = d[8]; o.n_text_layer = d[9];
  return o;
}

int wb200_model_num_tensors(const int32_t dims[10]) { return Model::num_tensors(dims_from(dims)); }

int wb200_model_create(const int32_t dims[10], int dtype, const void* const* tensors, int n_tensors,
                       wb200_model** out) {
  WB_CHECK_DTYPE(dtype);
  Dims d = dims_from(dims);
  if (d.n_audio_state != d.n_audio_head * 64 || d.n_text_state != d.n_text_head * 64)
    return set_error(120, "model: head dim must be 64 (state %d/%d heads %d/%d)", d.n_audio_state, d.n_text_state,
                     d.n_audio_head, d.n_text_head);
  if (d.n_audio_state != d.n_text_state) return set_error(121, "model: audio/text widths must match");
  if (d.n_audio_state % 64 || d.n_audio_state > 2048) return set_error(122, "model: unsupported width %d", d.n_audio_state);
  if (n_tensors != Model::num_tensors(d)) return set_error(123, "model: expected %d tensors, got %d", Model::num_tensors(d), n_tensors);
  for (int i = 0; i < n_tensors; ++i)
    if (!tensors[i]) return set_error(124, "model: tensor slot %d is null", i);
  wb200_model* h = new wb200_model();
  h->m.dims = d;
  h->m.dtype = dtype;
  h->m.t.assign(tensors, tensors + n_tensors);
  *out = h;
  return 0;
}

void wb200_model_destroy(wb200_model* model) { delete model; }

size_t wb200_encoder_workspace_bytes(const wb200_model* model, int n_audio) {
  return encoder_workspace_bytes(&model->m, n_audio);
}

int wb200_encoder_forward(const wb200_model* model, const float* mel, int n_audio, void* features,
                          void* workspace, size_t workspace_bytes, void* stream) {
  return encoder_forward(&model->m, mel, n_audio, features, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

size_t wb200_decoder_workspace_bytes(const wb200_model* model, const wb200_decode_config* cfg) {
  return decoder_workspace_bytes(&model->m, cfg);
}

int wb200_decoder_create(const wb200_model* model, const wb200_decode_config* cfg, void* workspace,
                         size_t workspace_bytes, wb200_decoder** out, void* stream) {
  Decoder* d = nullptr;
  int r = decoder_create(&model->m, cfg, workspace, workspace_bytes, &d, static_cast<cudaStream_t>(stream));
  if (r) return r;
  wb200_decoder* h = new wb200_decoder();
  h->d = d;
  *out = h;
  return 0;
}

void wb200_decoder_destroy(wb200_decoder* dec) {
  if (!dec) return;
  if (dec->d) {
    if (dec->d->pinned) cudaFreeHost(dec->d->pinned);
    if (dec->d->pair_graph) cudaGraphExecDestroy(dec->d->pair_graph);
    if (dec->d->ev_in) cudaEventDestroy(dec->d->ev_in);
    if (dec->d->ev_out) cudaEventDestroy(dec->d->ev_out);
    if (dec->d->gstream) cudaStreamDestroy(dec->d->gstream);
    delete dec->d;
  }
  delete dec;
}

int wb200_decoder_set_audio(wb200_decoder* dec, const void* features, void* stream) {
  return decoder_set_audio(dec->d, features, static_cast<cudaStream_t>(stream));
}
int wb200_decoder_prefill(wb200_decoder* dec, const int32_t* initial_tokens, void* stream) {
  return decoder_prefill(dec->d, initial_tokens, static_cast<cudaStream_t>(stream));
}
int wb200_decoder_set_sampling(wb200_decoder* dec, float temperature, uint64_t seed) {
  if (!dec) return set_error(2, "wb200_decoder_set_sampling: null decoder");
  return decoder_set_sampling(dec->d, temperature, seed);
}

int wb200_decoder_select(wb200_decoder* dec, void* stream) {
  return decoder_select(dec->d, static_cast<cudaStream_t>(stream));
}
int wb200_decoder_step(wb200_decoder* dec, void* stream) {
  return decoder_step(dec->d, static_cast<cudaStream_t>(stream));
}
int wb200_decoder_run(wb200_decoder* dec, int max_steps, int32_t* steps_issued, void* stream) {
  int n = 0;
  int r = decoder_run(dec->d, max_steps, &n, static_cast<cudaStream_t>(stream));
  if (steps_issued) *steps_issued = n;
  return r;
}
int wb200_decoder_force_tokens(wb200_decoder* dec, const int32_t* next_tokens, void* stream) {
  return decoder_append(dec->d, next_tokens, static_cast<cudaStream_t>(stream));
}
int wb200_decoder_set_alignment(wb200_decoder* dec, const int32_t* heads, int n_heads, float* qk_out) {
  Decoder* D = dec->d;
  D->align_heads.clear();
  D->align_qk = nullptr;
  if (n_heads <= 0) return 0;
  if (!heads || !qk_out) return set_error(150, "set_alignment: null argument");
  for (int i = 0; i < n_heads; ++i) {
    const int l = heads[2 * i], h = heads[2 * i + 1];
    if (l < 0 || l >= D->m->dims.n_text_layer || h < 0 || h >= D->m->dims.n_text_head)
      return set_error(151, "set_alignment: head (%d, %d) out of range", l, h);
    D->align_heads.push_back(l);
    D->align_heads.push_back(h);
  }
  D->align_qk = qk_out;
  return 0;
}

int wb200_alignment_weights(const float* qk, int n_heads, int n_tokens, int t_stride, int n_frames,
                            float qk_scale, int medfilt_width, int negate, float* out, float* scratch,
                            void* stream) {
  if (medfilt_width < 1 || (medfilt_width & 1) == 0) return set_error(152, "alignment_weights: filter width must be odd");
  int r = launch_alignment_weights(qk, n_heads, n_tokens, t_stride, n_frames, qk_scale, medfilt_width, negate, out,
                                   scratch, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_alignment_weights: failed (%d)", r) : 0;
}

int64_t wb200_decoder_logits_ld(const wb200_decoder* dec) { return dec->d->ldv; }

int wb200_decoder_get_state(wb200_decoder* dec, int what, void* dst, size_t bytes, void* stream) {
  void* p = nullptr;
  size_t n = 0;
  int r = decoder_state_ptr(dec->d, what, &p, &n, static_cast<cudaStream_t>(stream));
  if (r) return r;
  if (bytes > n) return set_error(271, "get_state(%d): asked for %zu bytes, have %zu", what, bytes, n);
  if (cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return set_error(272, "get_state(%d): copy failed", what);
  return 0;
}
int wb200_decoder_set_state(wb200_decoder* dec, int what, const void* src, size_t bytes, void* stream) {
  void* p = nullptr;
  size_t n = 0;
  int r = decoder_state_ptr(dec->d, what, &p, &n, static_cast<cudaStream_t>(stream));
  if (r) return r;
  if (bytes > n) return set_error(273, "set_state(%d): %zu bytes given, room for %zu", what, bytes, n);
  if (cudaMemcpyAsync(p, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return set_error(274, "set_state(%d): copy failed", what);
  return 0;
}

int wb200_range_softmax(const float* logits, int64_t ld, int first, int n, int rows, float* probs, int32_t* argmax,
                        const int32_t* gather_tokens, float* gather_probs, void* stream) {
  if (first < 0 || n <= 0 || rows < 0 || ld < first + n) return set_error(160, "wb200_range_softmax: bad range [%d, %d) of %lld", first, first + n, (long long)ld);
  int r = launch_range_softmax(logits, ld, first, n, rows, probs, argmax, gather_tokens, gather_probs, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_range_softmax: failed (%d)", r) : 0;
}

int wb200_median_filter(const float* x, float* y, int64_t rows, int T, int width, void* stream) {
  int r = launch_median_filter(x, y, rows, T, width, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_median_filter: failed (%d) rows=%lld T=%d width=%d", r, (long long)rows, T, width) : 0;
}

size_t wb200_dtw_workspace_bytes(int N, int M) { return dtw_workspace_bytes(N, M); }

int wb200_dtw(const float* x, int N, int M, int32_t* path, int32_t* path_len, void* workspace,
              size_t workspace_bytes, int tie_mode, void* stream) {
  if (workspace_bytes < dtw_workspace_bytes(N, M)) return set_error(130, "wb200_dtw: workspace too small");
  int r = launch_dtw(x, N, M, path, path_len, workspace, tie_mode, static_cast<cudaStream_t>(stream));
  return r ? set_error(r, "wb200_dtw: failed (%d) N=%d M=%d", r, N, M) : 0;
}

}  // extern "C"
