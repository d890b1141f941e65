// Model / encoder / decoder orchestration behind the C ABI: which kernels run, in what order, on
// which slices of the caller-provided workspace.  No allocation happens here except the small
// host-side handle structs; every device byte belongs to the caller (PyTorch).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/whisper_b200.h"
#include "engine.h"
#include "kernels.h"
#include "ptx.cuh"

namespace wb {

int set_error(int code, const char* fmt, ...);

// -------------------------------------------------------------------------------------------------
// small kernels: embedding gather, row gather, state initialisation
// -------------------------------------------------------------------------------------------------
// x[row] = T(token_embedding[tok] + positional_embedding[pos])   (reference model.py:235-239)
// ln_part (optional): also leave the row's LayerNorm statistics (count, mean, M2 of the STORED values) as partial
// slot 0 for the fused decoder-layer kernel, which never runs LayerNorm as a pass of its own (dec_layer.cu).
template <typename T>
__global__ void __launch_bounds__(128) embed_kernel(const int* __restrict__ tokens, int max_ctx, const int* __restrict__ len_ptr,
                                                    int n_init, int group, const float* __restrict__ emb, const float* __restrict__ pos,
                                                    T* __restrict__ x, int d, const int* skip_flag, float4* ln_part) {
  if (skip_flag && *skip_flag) return;
  const int row = blockIdx.x;
  int tok, p;
  if (len_ptr) {            // step: the last token of every row
    p = *len_ptr - 1;
    tok = tokens[static_cast<long long>(row) * max_ctx + p];
  } else {                  // prefill: token i of audio a (read from the audio's first beam row)
    const int a = row / n_init;
    p = row % n_init;
    tok = tokens[static_cast<long long>(a) * group * max_ctx + p];
  }
  const float* e = emb + static_cast<long long>(tok) * d;
  const float* pp = pos + static_cast<long long>(p) * d;
  T* xr = x + static_cast<long long>(row) * d;
  float2 kept[8];           // d <= 2048: at most 8 pairs per thread
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x * 2 + i * 256;
    kept[i] = make_float2(0.f, 0.f);
    if (c < d) {
      const float2 a = *reinterpret_cast<const float2*>(e + c);
      const float2 b = *reinterpret_cast<const float2*>(pp + c);
      const uint32_t pk = Cvt<T>::pack2(a.x + b.x, a.y + b.y);
      *reinterpret_cast<uint32_t*>(xr + c) = pk;
      kept[i] = Cvt<T>::unpack2(pk);
      sum += kept[i].x + kept[i].y;
    }
  }
  if (!ln_part) return;
  __shared__ float red[4];
  __shared__ float s_mean;
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) s_mean = (red[0] + red[1] + red[2] + red[3]) / static_cast<float>(d);
  __syncthreads();
  const float mean = s_mean;
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (threadIdx.x * 2 + i * 256 < d) {
      const float t0 = kept[i].x - mean, t1 = kept[i].y - mean;
      m2 += t0 * t0 + t1 * t1;
    }
  }
  m2 = warp_sum(m2);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m2;
  __syncthreads();
  if (threadIdx.x == 0) ln_part[row] = make_float4(static_cast<float>(d), mean, red[0] + red[1] + red[2] + red[3], 0.f);
}

// dst[i] = src[idx(i)] for the two prefill positions whose logits are needed (decoding.py:692,696)
template <typename T>
__global__ void gather_prefill_rows_kernel(const T* __restrict__ x, T* __restrict__ dst, int n_audio, int n_init,
                                           int sot_index, int d) {
  const int i = blockIdx.x;  // [0, 2*n_audio)
  const int a = i % n_audio;
  const int src = a * n_init + (i < n_audio ? sot_index : n_init - 1);
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8)
    *reinterpret_cast<uint4*>(dst + static_cast<long long>(i) * d + c) =
        *reinterpret_cast<const uint4*>(x + static_cast<long long>(src) * d + c);
}

__global__ void decoder_init_state_kernel(int* tokens0, int* tokens1, int* indir0, int* indir1, int max_ctx, int R,
                                          int group, int n_init, const int* __restrict__ init_tokens /*[n_audio,n_init]*/,
                                          float* sum_lp, int* len_ptr, int* done, int* cur, int* fin_count, int* fin_len,
                                          int n_audio, int max_cand, int* counters, int n_counters, unsigned int* dl_sync,
                                          unsigned char* same0, unsigned char* same1) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (long long i = tid; i < static_cast<long long>(R) * max_ctx; i += stride) {
    const int r = static_cast<int>(i / max_ctx), p = static_cast<int>(i % max_ctx);
    const int a = r / group;
    const int t = p < n_init ? init_tokens[a * n_init + p] : 0;
    tokens0[i] = t;
    tokens1[i] = t;
    const int ph = p < n_init ? a * group : r;
    indir0[i] = ph;
    indir1[i] = ph;
  }
  for (int i = tid; i < R; i += stride) sum_lp[i] = 0.f;
  for (int i = tid; i < n_audio; i += stride) fin_count[i] = 0;
  for (int i = tid; i < n_audio * max_cand; i += stride) fin_len[i] = 0;
  for (int i = tid; i < n_counters; i += stride) counters[i] = 0;
  for (int i = tid; i < n_audio * 256; i += stride) {      // every beam starts from the same prompt
    same0[i] = 1;
    same1[i] = 1;
  }
  if (tid == 0) {
    *len_ptr = n_init;
    *done = 0;
    *cur = 0;
    cur[8] = 0;      // beam tickets (scalars + 24, + 25)
    cur[9] = 0;
    dl_sync[0] = 0;  // grid-barrier / exit counters of the fused decoder-layer kernel
    dl_sync[1] = 0;
  }
}

// teacher forcing: append given tokens (tests)
__global__ void append_tokens_kernel(int* tokens, int max_ctx, int R, const int* __restrict__ next, int* len_ptr,
                                     unsigned char* same, int G) {
  const int L = *len_ptr;
  for (int r = threadIdx.x; r < R; r += blockDim.x) tokens[static_cast<long long>(r) * max_ctx + L] = next[r];
  // beams keep their rows: two prefixes stay equal only if the appended tokens are equal too
  if (same && G > 1)
    for (int i = threadIdx.x; i < (R / G) * G * G; i += blockDim.x) {
      const int a = i / (G * G), j1 = (i / G) % G, j2 = i % G;
      if (next[a * G + j1] != next[a * G + j2]) same[a * 256 + j1 * 16 + j2] = 0;
    }
  __syncthreads();
  if (threadIdx.x == 0) *len_ptr = L + 1;
}

// -------------------------------------------------------------------------------------------------
// workspace carving
// -------------------------------------------------------------------------------------------------
struct Arena {
  uint8_t* base;
  size_t off, cap;
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

#define WB_TRY(expr)                                  \
  do {                                                \
    int _r = (expr);                                  \
    if (_r) return set_error(_r, "%s failed (%d) at %s:%d", #expr, _r, __FILE__, __LINE__); \
  } while (0)

static int linear(const Model* m, const void* A, long long lda, int M, const void* W, int N, int K, const void* bias,
                  const void* residual, void* C, long long ldc, int gelu, int out_f32, cudaStream_t s,
                  const int* skip = nullptr, const Decoder* D = nullptr, int head_major_T = 0) {
  LinearArgs a;
  a.dtype = m->dtype;
  a.batch = 1;
  a.rows_per_batch = M;
  a.a_rows_per_batch = M;
  a.lda = lda;
  a.N = N;
  a.K_tap = K;
  a.taps = 1;
  a.A = A;
  a.W = W;
  a.ldw = K;
  a.bias = bias;
  a.residual = residual;
  a.ldr = ldc;
  a.C = C;
  a.ldc = ldc;
  a.gelu = gelu;
  a.out_f32 = out_f32;
  a.skip_flag = skip;
  a.head_major_T = head_major_T;
  if (D) {
    a.splitk_ws = D->gemm_ws;
    a.splitk_ws_bytes = D->gemm_ws_bytes;
    a.splitk_counters = D->gemm_counters;
    a.splitk_max_tiles = 256;
  }
  return launch_linear(a, s);
}

// -------------------------------------------------------------------------------------------------
// encoder (reference model.py:174-204)
// -------------------------------------------------------------------------------------------------
struct EncBufs {
  void *x, *ln, *qkv, *att, *hid;
};
static size_t enc_carve(const Model* m, int B, Arena& ar, EncBufs& b) {
  const size_t es = 2;
  const size_t d = m->dims.n_audio_state, T = m->dims.n_audio_ctx;
  b.x = ar.take(B * T * d * es);
  b.ln = ar.take(B * T * d * es);
  b.att = ar.take(B * T * d * es);
  b.qkv = ar.take(B * T * 3 * d * es);            // also holds the time-major mel (B*2T*n_mels)
  b.hid = ar.take(B * T * 4 * d * es);            // also holds conv1's output (B*2T*d)
  return ar.off;
}

size_t encoder_workspace_bytes(const Model* m, int B) {
  Arena ar{nullptr, 0, 0};
  EncBufs b;
  return enc_carve(m, B, ar, b) + 256;
}

int encoder_forward(const Model* m, const float* mel, int B, void* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (B <= 0) return 0;
  if (ws_bytes < encoder_workspace_bytes(m, B)) return set_error(200, "encoder workspace too small");
  Arena ar{static_cast<uint8_t*>(ws), 0, ws_bytes};
  EncBufs b;
  enc_carve(m, B, ar, b);
  const int d = m->dims.n_audio_state, T = m->dims.n_audio_ctx, H = m->dims.n_audio_head;
  const int n_mels = m->dims.n_mels, T2 = 2 * T;
  const int dt = m->dtype;
  void* melT = b.qkv;
  void* h1 = b.hid;
  WB_TRY(launch_transpose_to16(dt, mel, melT, B, n_mels, T2, s));
  {
    LinearArgs a;   // conv1 + GELU (model.py:193)
    a.dtype = dt; a.batch = B; a.rows_per_batch = T2; a.a_rows_per_batch = T2;
    a.a_batch_stride = static_cast<long long>(T2) * n_mels; a.lda = n_mels;
    a.N = d; a.K_tap = n_mels; a.taps = 3; a.a_row_off[0] = -1; a.a_row_off[1] = 0; a.a_row_off[2] = 1;
    a.A = melT; a.W = m->t[G_CONV1_W]; a.ldw = 3LL * n_mels; a.bias = m->t[G_CONV1_B];
    a.C = h1; a.ldc = d; a.gelu = 1;
    WB_TRY(launch_linear(a, s));
  }
  {
    LinearArgs a;   // conv2 (stride 2) + GELU + positional embedding (model.py:194-198)
    a.dtype = dt; a.batch = B; a.rows_per_batch = T; a.a_rows_per_batch = T;
    a.a_batch_stride = static_cast<long long>(T2) * d; a.lda = 2LL * d;
    a.N = d; a.K_tap = d; a.taps = 3;
    a.a_base_off[0] = d; a.a_row_off[0] = -1; a.a_base_off[1] = 0; a.a_row_off[1] = 0;
    a.a_base_off[2] = d; a.a_row_off[2] = 0;
    a.A = h1; a.W = m->t[G_CONV2_W]; a.ldw = 3LL * d; a.bias = m->t[G_CONV2_B];
    a.pos = static_cast<const float*>(m->t[G_ENC_POS]);
    a.C = b.x; a.ldc = d; a.gelu = 1;
    WB_TRY(launch_linear(a, s));
  }
  const int rows = B * T;
  for (int l = 0; l < m->dims.n_audio_layer; ++l) {
    const void* const* L = m->enc_layer(l);
    WB_TRY(launch_layernorm(dt, b.x, d, b.ln, d, (const float*)L[E_ATTN_LN_W], (const float*)L[E_ATTN_LN_B], rows, d, s));
    WB_TRY(linear(m, b.ln, d, rows, L[E_QKV_W], 3 * d, d, L[E_QKV_B], nullptr, b.qkv, 3 * d, 0, 0, s));
    WB_TRY(launch_enc_attention(dt, b.qkv, b.att, B, T, H, s));
    WB_TRY(linear(m, b.att, d, rows, L[E_OUT_W], d, d, L[E_OUT_B], b.x, b.x, d, 0, 0, s));
    WB_TRY(launch_layernorm(dt, b.x, d, b.ln, d, (const float*)L[E_MLP_LN_W], (const float*)L[E_MLP_LN_B], rows, d, s));
    WB_TRY(linear(m, b.ln, d, rows, L[E_FC1_W], 4 * d, d, L[E_FC1_B], nullptr, b.hid, 4 * d, 1, 0, s));
    WB_TRY(linear(m, b.hid, 4 * d, rows, L[E_FC2_W], d, 4 * d, L[E_FC2_B], b.x, b.x, d, 0, 0, s));
  }
  WB_TRY(launch_layernorm(dt, b.x, d, out, d, (const float*)m->t[G_ENC_LN_POST_W], (const float*)m->t[G_ENC_LN_POST_B],
                          rows, d, s));
  return 0;
}

// -------------------------------------------------------------------------------------------------
// decoder
// -------------------------------------------------------------------------------------------------
static void dec_carve(const Model* m, const wb200_decode_config& c, Arena& ar, Decoder* D) {
  const size_t es = 2;
  const size_t d = m->dims.n_text_state, V = m->dims.n_vocab, ctx = m->dims.n_text_ctx;
  const size_t Ta = m->dims.n_audio_ctx, NL = m->dims.n_text_layer, H = m->dims.n_text_head;
  const size_t B = c.n_audio, G = c.n_group, R = B * G;
  const size_t P = B * c.n_init;
  const size_t rows = R > P ? R : P;
  const size_t ldv = (V + 31) / 32 * 32;
  Decoder dd;
  Decoder* o = D ? D : &dd;
  o->ldv = static_cast<long long>(ldv);
  o->cross_kv = ar.take(NL * B * Ta * 2 * d * es);
  o->self_k = ar.take(NL * R * ctx * d * es);
  o->self_v = ar.take(NL * R * ctx * d * es);
  o->x = ar.take(rows * d * es);
  o->ln = ar.take(rows * d * es);
  o->qkv = ar.take(rows * 3 * d * es);
  o->att = ar.take(rows * d * es);
  o->q = ar.take(rows * d * es);
  o->hid = ar.take(rows * 4 * d * es);
  o->sel = ar.take(2 * B * d * es);
  size_t logit_rows = R > 2 * B ? R : 2 * B;
  if (c.all_logits && P > logit_rows) logit_rows = P;
  o->logits = static_cast<float*>(ar.take(logit_rows * ldv * 4));
  const size_t nq = c.n_init > (int)G ? c.n_init : G;
  o->partial = static_cast<float*>(ar.take(cross_attention_partial_floats((int)B, (int)nq, (int)H, (int)Ta) * 4));
  o->n_counters = static_cast<int>(B * ((nq + 15) / 16) * H);
  o->counters = static_cast<int*>(ar.take((o->n_counters + 256) * 4));
  o->gemm_counters = o->counters ? o->counters + o->n_counters : nullptr;
  o->gemm_ws_bytes = 40u << 20;
  o->gemm_ws = static_cast<float*>(ar.take(o->gemm_ws_bytes));
  for (int i = 0; i < 2; ++i) {
    o->tokens[i] = static_cast<int*>(ar.take(R * ctx * 4));
    o->indir[i] = static_cast<int*>(ar.take(R * ctx * 4));
  }
  o->sum_lp = static_cast<float*>(ar.take(R * 4));
  o->no_speech = static_cast<float*>(ar.take(B * 4));
  const size_t K = c.beam_search ? G + 1 : 1;
  o->top_val = static_cast<float*>(ar.take(R * K * 4));
  o->top_idx = static_cast<int*>(ar.take(R * K * 4));
  o->sources = static_cast<int*>(ar.take(R * 4));
  const size_t mc = c.max_candidates > 0 ? c.max_candidates : 1;
  o->fin_tokens = static_cast<int*>(ar.take(B * mc * ctx * 4));
  o->fin_len = static_cast<int*>(ar.take(B * mc * 4));
  o->fin_score = static_cast<float*>(ar.take(B * mc * 4));
  o->fin_count = static_cast<int*>(ar.take(B * 4));
  o->suppress_mask = static_cast<uint32_t*>(ar.take(ldv / 8 + 64));
  o->blank_mask = static_cast<uint32_t*>(ar.take(ldv / 8 + 64));
  o->init_tokens = static_cast<int*>(ar.take(P * 4 + 64));
  o->scalars = static_cast<int*>(ar.take(256));
  for (int i = 0; i < 2; ++i) o->beam_same[i] = static_cast<unsigned char*>(ar.take(B * 256));
  // fused decoder-layer kernel: LN partial statistics [slots <= SMs][rows padded to 64] and its two counters
  o->ln_ld = static_cast<int>((R + 63) / 64 * 64);
  o->ln_part = static_cast<float4*>(ar.take(static_cast<size_t>(256) * o->ln_ld * sizeof(float4)));
  o->dl_sync = static_cast<unsigned int*>(ar.take(256));
  o->stack_table = static_cast<DLPhase*>(ar.take((static_cast<size_t>(9) * NL + 4) * sizeof(DLPhase)));
}

size_t decoder_workspace_bytes(const Model* m, const wb200_decode_config* c) {
  Arena ar{nullptr, 0, 0};
  dec_carve(m, *c, ar, nullptr);
  return ar.off + 512;
}

// Launch plans of the fused decoder-layer kernel for the step path (R = n_audio * n_group new positions per step):
//   head[l] = {QKV}                                   (used for layer 0 only; other layers get it from the previous tail)
//   mid[l]  = {out-proj + residual, cross-query}
//   tail[l] = {cross-out + residual, fc1 + GELU, fc2 + residual, QKV of layer l + 1}
// Few-rows sessions (every chain of the plan in the few-rows form, head-major caches): string the chains and the two
// attentions of every layer into ONE phase table - [QKV0] then per layer {self-attention, out-proj, cross-query,
// cross-attention (+ merge of its key slices), cross-out, fc1, fc2, next layer's QKV} - for a single launch per iteration.
static int build_stack_plan(Decoder* D, cudaStream_t s) {
  const Model* m = D->m;
  D->stack_ready = false;
  if (g_fused_stack < 0) {
    const char* e = getenv("WB200_FUSED_STACK");
    g_fused_stack = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  const int NL = m->dims.n_text_layer, H = m->dims.n_text_head, d = m->dims.n_text_state, ctx = m->dims.n_text_ctx;
  const int Ta = m->dims.n_audio_ctx, B = D->cfg.n_audio, G = D->cfg.n_group, R = B * G;
  if (!g_fused_stack || !D->kv_head_major || D->kv_window || d != H * 64) return 0;
  for (int l = 0; l < NL; ++l)
    if (D->dl_mid[l].rows_smem <= 0 || D->dl_tail[l].rows_smem <= 0 || (l == 0 && D->dl_head[0].rows_smem <= 0)) return 0;
  const int grid = D->dl_head[0].grid;
  const int pairs = R * H;
  int splits = grid / pairs;                       // key slices per (row, head) of the cross-attention
  if (splits < 1) splits = 1;
  if (splits > Ta / 64) splits = Ta / 64 > 0 ? Ta / 64 : 1;
  const size_t nq = D->cfg.n_init > G ? D->cfg.n_init : G;
  if (static_cast<size_t>(pairs) * splits * 66 > cross_attention_partial_floats(B, static_cast<int>(nq), H, Ta)) return 0;
  const size_t cross_per_layer = static_cast<size_t>(B) * Ta * 2 * d * 2;
  const size_t self_per_layer = static_cast<size_t>(R) * ctx * d * 2;
  std::vector<DLPhase>& tab = D->stack_host;
  tab.clear();
  auto attention = [&](int type, void* kc, void* vc) {
    DLPhase ph = {};
    ph.type = type;
    ph.kc = kc;
    ph.vc = vc;
    tab.push_back(ph);
  };
  tab.push_back(D->dl_head[0].p.ph[0]);
  for (int l = 0; l < NL; ++l) {
    attention(DS_SELF, static_cast<uint8_t*>(D->self_k) + l * self_per_layer, static_cast<uint8_t*>(D->self_v) + l * self_per_layer);
    for (int i = 0; i < D->dl_mid[l].p.n_phases; ++i) tab.push_back(D->dl_mid[l].p.ph[i]);
    attention(DS_CROSS, static_cast<uint8_t*>(D->cross_kv) + l * cross_per_layer, nullptr);
    if (splits > 1) attention(DS_COMBINE, nullptr, nullptr);
    for (int i = 0; i < D->dl_tail[l].p.n_phases; ++i) tab.push_back(D->dl_tail[l].p.ph[i]);
  }
  const size_t n_layers_phases = tab.size();
  if (g_fused_stack >= 2) {
    // opt-in (mode 2): the decoder's final LayerNorm and the logits (model.py:243-247) close the table - the rows are
    // normalised by one warp each, then every CTA runs its 1/148 of the vocabulary through the slab buffer in ~20 slabs
    // that alternate between the two halves of the buffer.  Measured on the turbo one-audio decode
    // (tools/time_stack_modes.py): 300-308 us per iteration against 302 us with the two separate launches (LayerNorm
    // 6 us + tcgen05 GEMM 42 us) - identical tokens, no gain, hence not the default.
    DLPhase ln = {};
    ln.type = DS_LN;
    ln.N = d;
    ln.a = D->x;
    ln.lda = d;
    ln.out = D->ln;
    ln.ldo = d;
    ln.c1 = static_cast<const float*>(m->t[G_DEC_LN_W]);
    ln.c2 = static_cast<const float*>(m->t[G_DEC_LN_B]);
    tab.push_back(ln);
    DLPhase lg = {};
    lg.type = DS_LINEAR;
    lg.N = m->dims.n_vocab;
    lg.K = d;
    lg.flags = DL_OUTF32;
    lg.a = D->ln;
    lg.lda = d;
    lg.w = m->t[G_TOK_EMB16];
    lg.out = D->logits;
    lg.ldo = D->ldv;
    tab.push_back(lg);
  }
  if (tab.size() > static_cast<size_t>(9) * NL + 4) return 80;
  DLLaunch& S = D->dl_stack;
  dl_init_launch(S, m->dtype, R, grid, D->ln_part, D->ln_ld, D->dl_sync, D->done_ptr, 1);
  D->stack_has_logits = tab.size() > n_layers_phases;
  if (!dl_plan_stack(S, tab.data(), static_cast<int>(tab.size()), D->stack_table)) {
    if (!D->stack_has_logits) return 0;
    tab.resize(n_layers_phases);             // the logits slabs do not fit next to the input rows: the layers alone
    D->stack_has_logits = false;
    if (!dl_plan_stack(S, tab.data(), static_cast<int>(tab.size()), D->stack_table)) return 0;
  }
  if (cudaMemcpyAsync(D->stack_table, tab.data(), tab.size() * sizeof(DLPhase), cudaMemcpyHostToDevice, s) != cudaSuccess) return 81;
  S.p.qkv = D->qkv;
  S.p.q = D->q;
  S.p.att = D->att;
  S.p.indir = D->indir[0];                // set per launch (the parent tables ping-pong)
  S.p.len_ptr = D->len_ptr;
  S.p.xpart = D->partial;
  S.p.n_head = H;
  S.p.ctx = ctx;
  S.p.T = Ta;
  S.p.G = G;
  S.p.splits = splits;
  S.p.d = d;
  D->stack_ready = true;
  return 0;
}

static int build_fused_plan(Decoder* D, cudaStream_t s) {
  const Model* m = D->m;
  D->fused = false;
  if (g_fused_layer < 0) {
    const char* e = getenv("WB200_FUSED_LAYER");
    g_fused_layer = (e && e[0] == '0') ? 0 : 1;
  }
  const int R = D->cfg.n_audio * D->cfg.n_group, d = m->dims.n_text_state, dt = m->dtype;
  const int grid = dl_grid_size();
  if (!g_fused_layer || D->cfg.all_logits || !dl_supported(R, d, grid)) return 0;
  const int NL = m->dims.n_text_layer;
  D->dl_head.assign(NL, DLLaunch());
  D->dl_mid.assign(NL, DLLaunch());
  D->dl_tail.assign(NL, DLLaunch());
  const int* skip = D->done_ptr;
  for (int l = 0; l < NL; ++l) {
    const void* const* L = m->dec_layer(l);
    auto qkv_phase = [&](DLLaunch& X, int idx, const void* const* LL) {
      return dl_fill_phase(X, idx, dt, R, grid, D->x, d, LL[D_QKV_WF], 3 * d, d, nullptr, (const float*)LL[D_QKV_C1],
                           (const float*)LL[D_QKV_C2], DL_FOLD, D->qkv, 3LL * d);
    };
    DLLaunch& H = D->dl_head[l];
    dl_init_launch(H, dt, R, grid, D->ln_part, D->ln_ld, D->dl_sync, skip, l == 0 ? 1 : 0);
    if (int r = qkv_phase(H, 0, L)) return 10 + r;
    H.p.n_phases = 1;
    dl_use_rows_form(H);
    DLLaunch& Mi = D->dl_mid[l];
    dl_init_launch(Mi, dt, R, grid, D->ln_part, D->ln_ld, D->dl_sync, skip, 0);
    if (int r = dl_fill_phase(Mi, 0, dt, R, grid, D->att, d, L[D_OUT_W], d, d, L[D_OUT_B], nullptr, nullptr,
                              DL_RESID | DL_STATS, D->x, d)) return 20 + r;
    if (int r = dl_fill_phase(Mi, 1, dt, R, grid, D->x, d, L[D_CQ_WF], d, d, nullptr, (const float*)L[D_CQ_C1],
                              (const float*)L[D_CQ_C2], DL_FOLD, D->q, d)) return 30 + r;
    Mi.p.n_phases = 2;
    dl_use_rows_form(Mi);
    DLLaunch& Ta = D->dl_tail[l];
    dl_init_launch(Ta, dt, R, grid, D->ln_part, D->ln_ld, D->dl_sync, skip, 0);
    if (int r = dl_fill_phase(Ta, 0, dt, R, grid, D->att, d, L[D_COUT_W], d, d, L[D_COUT_B], nullptr, nullptr,
                              DL_RESID | DL_STATS, D->x, d)) return 40 + r;
    if (int r = dl_fill_phase(Ta, 1, dt, R, grid, D->x, d, L[D_FC1_WF], 4 * d, d, nullptr, (const float*)L[D_FC1_C1],
                              (const float*)L[D_FC1_C2], DL_FOLD | DL_GELU, D->hid, 4LL * d)) return 50 + r;
    if (int r = dl_fill_phase(Ta, 2, dt, R, grid, D->hid, 4LL * d, L[D_FC2_W], d, 4 * d, L[D_FC2_B], nullptr, nullptr,
                              DL_RESID | DL_STATS, D->x, d)) return 60 + r;
    Ta.p.n_phases = 3;
    if (l + 1 < NL) {
      if (int r = qkv_phase(Ta, 3, m->dec_layer(l + 1))) return 70 + r;
      Ta.p.n_phases = 4;
    }
    dl_use_rows_form(Ta);
  }
  D->fused = true;
  return build_stack_plan(D, s);
}

int decoder_create(const Model* m, const wb200_decode_config* c, void* ws, size_t ws_bytes, Decoder** out,
                   cudaStream_t s) {
  if (c->n_audio <= 0 || c->n_group <= 0) return set_error(210, "decoder: n_audio/n_group must be positive");
  if (c->n_group > 16) return set_error(211, "decoder: at most 16 beams / samples per audio");
  if (c->n_init < 1 || c->n_init + 1 > m->dims.n_text_ctx) return set_error(212, "decoder: bad n_init %d", c->n_init);
  if (c->sot_index < 0 || c->sot_index >= c->n_init) return set_error(213, "decoder: bad sot_index");
  if (ws_bytes < decoder_workspace_bytes(m, c)) return set_error(214, "decoder workspace too small");
  Decoder* D = new Decoder();
  D->m = m;
  D->cfg = *c;
  if (g_kv_head_major < 0) {
    const char* e = getenv("WB200_KV_HEAD_MAJOR");
    g_kv_head_major = (e && e[0] == '0') ? 0 : 1;        // default: head-major (contiguous (audio, head) streams, TMA-able)
  }
  if (g_xattn_tma < 0) {
    const char* e = getenv("WB200_XATTN_TMA");
    g_xattn_tma = (e && e[0] == '0') ? 0 : 1;
  }
  if (g_sattn_tma < 0) {
    // opt-in: measured slower than the gather kernel at the headline shape (profiles/r2_summary.md section 4)
    const char* e = getenv("WB200_SATTN_TMA");
    g_sattn_tma = (e && e[0] == '1') ? 1 : 0;
  }
  D->kv_head_major = g_kv_head_major != 0;
  // beam-window self caches only where the kernel that wants them runs (fixed per session: the layout cannot change later)
  D->kv_window = D->kv_head_major && g_sattn_tma &&
                 self_attention_tma_covers(c->n_audio, c->n_group, m->dims.n_text_head, m->dims.n_text_ctx);
  D->cfg.suppress_ids = nullptr;
  D->cfg.blank_ids = nullptr;
  Arena ar{static_cast<uint8_t*>(ws), 0, ws_bytes};
  dec_carve(m, *c, ar, D);
  D->len_ptr = D->scalars;
  D->done_ptr = D->scalars + 8;
  D->cur_ptr = D->scalars + 16;
  D->cur = 0;
  if (cudaMallocHost(reinterpret_cast<void**>(&D->pinned), 64) != cudaSuccess) {
    delete D;
    return set_error(216, "decoder: pinned scratch allocation failed");
  }
  D->host_len = 0;
  // suppress / blank bit masks (decoding.py:423-438, 454-455)
  const size_t words = static_cast<size_t>(D->ldv) / 32 + 1;
  std::vector<uint32_t> sup(words, 0u), blank(words, 0u);
  const int V = m->dims.n_vocab;
  for (int i = 0; i < c->n_suppress; ++i) {
    const int t = c->suppress_ids[i];
    if (t >= 0 && t < V) sup[t >> 5] |= 1u << (t & 31);
  }
  if (c->timestamp_rules && c->no_timestamps >= 0 && c->no_timestamps < V)
    sup[c->no_timestamps >> 5] |= 1u << (c->no_timestamps & 31);
  for (int i = 0; i < c->n_blank; ++i) {
    const int t = c->blank_ids[i];
    if (t >= 0 && t < V) blank[t >> 5] |= 1u << (t & 31);
  }
  if (c->eot >= 0 && c->eot < V) blank[c->eot >> 5] |= 1u << (c->eot & 31);
  if (cudaMemcpyAsync(D->suppress_mask, sup.data(), words * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(D->blank_mask, blank.data(), words * 4, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    cudaFreeHost(D->pinned);
    delete D;
    return set_error(215, "decoder: mask upload failed");
  }
  {
    // the beam-window self-attention kernel reads whole tiles, including rows no beam refers to (masked to zero
    // probability): those bytes must be finite, so the arenas start from zero instead of whatever the allocator left
    const size_t self_bytes = static_cast<size_t>(m->dims.n_text_layer) * c->n_audio * c->n_group * m->dims.n_text_ctx *
                              m->dims.n_text_state * 2;
    if (cudaMemsetAsync(D->self_k, 0, self_bytes, s) != cudaSuccess || cudaMemsetAsync(D->self_v, 0, self_bytes, s) != cudaSuccess) {
      cudaFreeHost(D->pinned);
      delete D;
      return set_error(218, "decoder: kv arena initialisation failed");
    }
  }
  {
    int r = build_fused_plan(D, s);
    if (r) {
      cudaFreeHost(D->pinned);
      delete D;
      return set_error(217, "decoder: fused decoder-layer plan failed (%d)", r);
    }
  }
  *out = D;
  return 0;
}

// cross-attention K/V for every layer, once per segment (model.py:104-109 first-call branch)
int decoder_set_audio(Decoder* D, const void* features, cudaStream_t s) {
  const Model* m = D->m;
  const int d = m->dims.n_text_state, Ta = m->dims.n_audio_ctx, B = D->cfg.n_audio;
  if (m->dims.n_audio_state != d) return set_error(220, "audio/text widths differ");
  const size_t per_layer = static_cast<size_t>(B) * Ta * 2 * d * 2;
  for (int l = 0; l < m->dims.n_text_layer; ++l) {
    const void* const* L = m->dec_layer(l);
    void* kv = static_cast<uint8_t*>(D->cross_kv) + l * per_layer;
    // row-major: [B * Ta, 2d] (K | V per position); head-major: [B][2H][Ta][64] (K heads then V heads)
    WB_TRY(linear(m, features, d, B * Ta, L[D_CKV_W], 2 * d, d, L[D_CKV_B], nullptr, kv, 2 * d, 0, 0, s, nullptr, nullptr,
                  D->kv_head_major ? Ta : 0));
  }
  return 0;
}

// decoder self-attention of one layer in the step: all beams of an audio together through the beam-window TMA kernel
// when the shape allows, else one warp per (row, head)
static int self_attention_step(Decoder* D, void* kc, void* vc, cudaStream_t s) {
  const Model* m = D->m;
  const int H = m->dims.n_text_head, ctx = m->dims.n_text_ctx, B = D->cfg.n_audio, G = D->cfg.n_group;
  if (D->kv_window)
    return launch_self_attention_tma(m->dtype, D->qkv, kc, vc, D->att, D->indir[D->cur], D->len_ptr, D->done_ptr, B, G, H, ctx, s);
  return launch_self_attention(m->dtype, D->qkv, kc, vc, D->att, D->indir[D->cur], D->len_ptr, D->done_ptr, B * G, H, ctx,
                               D->cfg.n_init, G, s, D->kv_head_major ? 1 : 0);
}

// decoder cross-attention of one layer: the persistent TMA kernel for the step (head-major K/V, <= 16 queries per
// audio), the cp.async kernel for the prefill and every other shape
static int cross_attention(Decoder* D, const uint8_t* ckv, bool step, int n_q, cudaStream_t s) {
  const Model* m = D->m;
  const int d = m->dims.n_text_state, H = m->dims.n_text_head, Ta = m->dims.n_audio_ctx, B = D->cfg.n_audio;
  const int* skip = step ? D->done_ptr : nullptr;
  if (step && D->kv_head_major && g_xattn_tma) {
    const int r = launch_cross_attention_tma(m->dtype, D->q, ckv, D->att, D->partial, D->counters, skip, B, n_q, Ta, H, s);
    if (r >= 0) return r;
  }
  return launch_cross_attention(m->dtype, D->q, ckv, ckv + static_cast<size_t>(d) * 2, D->att, D->partial, D->counters, skip, B,
                                n_q, Ta, H, 2 * d, s, D->kv_head_major ? 1 : 0);
}

// the transformer stack for `rows` new positions; step mode when `step` is true
static int decoder_stack(Decoder* D, int rows, bool step, cudaStream_t s) {
  const Model* m = D->m;
  const int d = m->dims.n_text_state, H = m->dims.n_text_head, ctx = m->dims.n_text_ctx;
  const int Ta = m->dims.n_audio_ctx, B = D->cfg.n_audio, G = D->cfg.n_group, dt = m->dtype;
  const int R = B * G;
  const int* skip = step ? D->done_ptr : nullptr;
  const size_t cross_per_layer = static_cast<size_t>(B) * Ta * 2 * d * 2;
  const size_t self_per_layer = static_cast<size_t>(R) * ctx * d * 2;
  const int n_q = step ? G : D->cfg.n_init;
  if (step && D->fused && D->stack_ready) {
    // few rows: the whole stack - Linear chains and both attentions of every layer - in one persistent launch
    D->dl_stack.p.indir = D->indir[D->cur];
    return dl_launch(D->dl_stack, s);
  }
  if (step && D->fused) {
    // fused GEMM chains (dec_layer.cu) around the two attention kernels: 4 launches per layer instead of 11
    for (int l = 0; l < m->dims.n_text_layer; ++l) {
      void* kc = static_cast<uint8_t*>(D->self_k) + l * self_per_layer;
      void* vc = static_cast<uint8_t*>(D->self_v) + l * self_per_layer;
      const uint8_t* ckv = static_cast<const uint8_t*>(D->cross_kv) + l * cross_per_layer;
      if (l == 0) WB_TRY(dl_launch(D->dl_head[0], s));
      WB_TRY(self_attention_step(D, kc, vc, s));
      WB_TRY(dl_launch(D->dl_mid[l], s));
      WB_TRY(cross_attention(D, ckv, true, n_q, s));
      WB_TRY(dl_launch(D->dl_tail[l], s));
    }
    return 0;
  }
  for (int l = 0; l < m->dims.n_text_layer; ++l) {
    const void* const* L = m->dec_layer(l);
    void* kc = static_cast<uint8_t*>(D->self_k) + l * self_per_layer;
    void* vc = static_cast<uint8_t*>(D->self_v) + l * self_per_layer;
    const uint8_t* ckv = static_cast<const uint8_t*>(D->cross_kv) + l * cross_per_layer;
    WB_TRY(launch_layernorm(dt, D->x, d, D->ln, d, (const float*)L[D_ATTN_LN_W], (const float*)L[D_ATTN_LN_B], rows, d, s, skip));
    WB_TRY(linear(m, D->ln, d, rows, L[D_QKV_W], 3 * d, d, L[D_QKV_B], nullptr, D->qkv, 3 * d, 0, 0, s, skip, D));
    if (step)
      WB_TRY(self_attention_step(D, kc, vc, s));
    else
      WB_TRY(launch_self_attention(dt, D->qkv, kc, vc, D->att, nullptr, D->len_ptr, skip, rows, H, ctx, D->cfg.n_init, G, s,
                                   D->kv_window ? 2 : D->kv_head_major ? 1 : 0));
    WB_TRY(linear(m, D->att, d, rows, L[D_OUT_W], d, d, L[D_OUT_B], D->x, D->x, d, 0, 0, s, skip, D));
    WB_TRY(launch_layernorm(dt, D->x, d, D->ln, d, (const float*)L[D_CROSS_LN_W], (const float*)L[D_CROSS_LN_B], rows, d, s, skip));
    WB_TRY(linear(m, D->ln, d, rows, L[D_CQ_W], d, d, L[D_CQ_B], nullptr, D->q, d, 0, 0, s, skip, D));
    if (!step && D->align_qk) {
      // timing.py:186-197: pre-softmax cross-attention scores of the alignment heads (audio 0 only)
      for (size_t i = 0; i + 1 < D->align_heads.size(); i += 2) {
        if (D->align_heads[i] != l) continue;
        const int hh = D->align_heads[i + 1];
        float* dst = D->align_qk + (i / 2) * static_cast<size_t>(D->cfg.n_init) * Ta;
        // K of audio 0, head hh: strided rows of the [Ta, 2d] block, or the contiguous [Ta, 64] block of that head
        const uint8_t* k_head = D->kv_head_major ? ckv + static_cast<size_t>(hh) * Ta * 128 : ckv + static_cast<size_t>(hh) * 128;
        WB_TRY(launch_qk_export(dt, static_cast<const uint8_t*>(D->q) + static_cast<size_t>(hh) * 128, d,
                                k_head, D->kv_head_major ? 64 : 2 * d, dst, D->cfg.n_init, Ta, s));
      }
    }
    WB_TRY(cross_attention(D, ckv, step, n_q, s));
    WB_TRY(linear(m, D->att, d, rows, L[D_COUT_W], d, d, L[D_COUT_B], D->x, D->x, d, 0, 0, s, skip, D));
    WB_TRY(launch_layernorm(dt, D->x, d, D->ln, d, (const float*)L[D_MLP_LN_W], (const float*)L[D_MLP_LN_B], rows, d, s, skip));
    WB_TRY(linear(m, D->ln, d, rows, L[D_FC1_W], 4 * d, d, L[D_FC1_B], nullptr, D->hid, 4 * d, 1, 0, s, skip, D));
    WB_TRY(linear(m, D->hid, 4 * d, rows, L[D_FC2_W], d, 4 * d, L[D_FC2_B], D->x, D->x, d, 0, 0, s, skip, D));
  }
  return 0;
}

template <typename T>
static void launch_embed(Decoder* D, int rows, bool step, cudaStream_t s) {
  const Model* m = D->m;
  embed_kernel<T><<<rows, 128, 0, s>>>(D->tokens[D->cur], m->dims.n_text_ctx, step ? D->len_ptr : nullptr, D->cfg.n_init,
                                       D->cfg.n_group, (const float*)m->t[G_TOK_EMB32], (const float*)m->t[G_DEC_POS],
                                       static_cast<T*>(D->x), m->dims.n_text_state, step ? D->done_ptr : nullptr,
                                       (step && D->fused) ? D->ln_part : nullptr);
  count_launch();
}

int decoder_prefill(Decoder* D, const int32_t* init_tokens_host, cudaStream_t s) {
  const Model* m = D->m;
  const wb200_decode_config& c = D->cfg;
  const int B = c.n_audio, G = c.n_group, R = B * G, P = B * c.n_init;
  const int d = m->dims.n_text_state, V = m->dims.n_vocab, ctx = m->dims.n_text_ctx, dt = m->dtype;
  if (cudaMemcpyAsync(D->init_tokens, init_tokens_host, static_cast<size_t>(P) * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
    return set_error(230, "prefill: token upload failed");
  D->cur = 0;
  decoder_init_state_kernel<<<256, 256, 0, s>>>(D->tokens[0], D->tokens[1], D->indir[0], D->indir[1], ctx, R, G, c.n_init,
                                                D->init_tokens, D->sum_lp, D->len_ptr, D->done_ptr, D->cur_ptr, D->fin_count, D->fin_len,
                                                B, c.max_candidates > 0 ? c.max_candidates : 1, D->counters, D->n_counters + 256,
                                                D->dl_sync, D->beam_same[0], D->beam_same[1]);
  count_launch();
  if (dt == DT_BF16) launch_embed<__nv_bfloat16>(D, P, false, s); else launch_embed<__half>(D, P, false, s);
  WB_TRY(decoder_stack(D, P, false, s));
  if (c.all_logits) {
    // un-cached full forward (model.py:293-296): logits of every prompt position
    WB_TRY(launch_layernorm(dt, D->x, d, D->ln, d, (const float*)m->t[G_DEC_LN_W], (const float*)m->t[G_DEC_LN_B], P, d, s, nullptr));
    WB_TRY(linear(m, D->ln, d, P, m->t[G_TOK_EMB16], V, d, nullptr, nullptr, D->logits, D->ldv, 0, 1, s));
    if (c.no_speech >= 0) WB_TRY(launch_no_speech(D->logits, D->ldv, V, c.no_speech, D->no_speech, B, c.n_init, c.sot_index, s));
    D->logits_cur = D->logits;
    D->logits_row_div = 1;
    D->logits_rows = P;
    D->host_len = c.n_init;
    D->forward_only = true;
    return 0;
  }
  D->forward_only = false;
  D->logits_rows = B;
  if (dt == DT_BF16)
    gather_prefill_rows_kernel<__nv_bfloat16><<<2 * B, 128, 0, s>>>((const __nv_bfloat16*)D->x, (__nv_bfloat16*)D->sel, B, c.n_init, c.sot_index, d);
  else
    gather_prefill_rows_kernel<__half><<<2 * B, 128, 0, s>>>((const __half*)D->x, (__half*)D->sel, B, c.n_init, c.sot_index, d);
  count_launch();
  WB_TRY(launch_layernorm(dt, D->sel, d, D->ln, d, (const float*)m->t[G_DEC_LN_W], (const float*)m->t[G_DEC_LN_B], 2 * B, d, s, nullptr));
  WB_TRY(linear(m, D->ln, d, 2 * B, m->t[G_TOK_EMB16], V, d, nullptr, nullptr, D->logits, D->ldv, 0, 1, s));
  if (c.no_speech >= 0) WB_TRY(launch_no_speech(D->logits, D->ldv, V, c.no_speech, D->no_speech, B, 1, 0, s));
  D->logits_cur = D->logits + static_cast<size_t>(B) * D->ldv;   // rows [B, 2B): last prompt position
  D->logits_row_div = G;
  D->host_len = c.n_init;
  if (cudaGetLastError() != cudaSuccess) return set_error(231, "prefill: launch error");
  return 0;
}

int decoder_step(Decoder* D, cudaStream_t s) {
  const Model* m = D->m;
  const int R = D->cfg.n_audio * D->cfg.n_group;
  const int d = m->dims.n_text_state, V = m->dims.n_vocab, dt = m->dtype;
  if (D->host_len >= m->dims.n_text_ctx) return set_error(240, "step: context full");
  if (D->forward_only) return set_error(241, "step: this session was created with all_logits (forward-only)");
  if (dt == DT_BF16) launch_embed<__nv_bfloat16>(D, R, true, s); else launch_embed<__half>(D, R, true, s);
  WB_TRY(decoder_stack(D, R, true, s));
  if (!(D->fused && D->stack_ready && D->stack_has_logits)) {      // (mode 2 of the one-launch stack ends with both)
    WB_TRY(launch_layernorm(dt, D->x, d, D->ln, d, (const float*)m->t[G_DEC_LN_W], (const float*)m->t[G_DEC_LN_B], R, d, s, D->done_ptr));
    WB_TRY(linear(m, D->ln, d, R, m->t[G_TOK_EMB16], V, d, nullptr, nullptr, D->logits, D->ldv, 0, 1, s, D->done_ptr, D));
  }
  D->logits_cur = D->logits;
  D->logits_row_div = 1;
  D->logits_rows = R;
  return 0;
}

int decoder_set_sampling(Decoder* D, float temperature, unsigned long long seed) {
  if (!(temperature >= 0.f)) return set_error(243, "set_sampling: temperature must be >= 0");
  if (temperature > 0.f && D->cfg.beam_search)
    return set_error(244, "set_sampling: temperature > 0 needs a greedy session (decoding.py:548-552 builds a GreedyDecoder)");
  D->temperature = temperature;
  D->seed = seed;
  if (D->pair_graph) {                 // the captured decode loop has the old parameters baked in
    cudaGraphExecDestroy(D->pair_graph);
    D->pair_graph = nullptr;
    D->pair_graph_cur = -1;
  }
  return 0;
}

int decoder_select(Decoder* D, cudaStream_t s) {
  const Model* m = D->m;
  const wb200_decode_config& c = D->cfg;
  const int B = c.n_audio, G = c.n_group, R = B * G, ctx = m->dims.n_text_ctx;
  if (D->forward_only) return set_error(242, "select: this session was created with all_logits (forward-only)");
  FilterParams f;
  f.logits = D->logits_cur;
  f.ld = D->ldv;
  f.V = m->dims.n_vocab;
  f.row_div = D->logits_row_div;
  f.tokens = D->tokens[D->cur];
  f.max_ctx = ctx;
  f.len_ptr = D->len_ptr;
  f.skip_flag = D->done_ptr;
  f.suppress_mask = D->suppress_mask;
  f.blank_mask = D->blank_mask;
  f.sample_begin = c.sample_begin;
  f.eot = c.eot;
  f.timestamp_begin = c.timestamp_begin;
  f.max_initial_ts = c.max_initial_timestamp_index;
  f.suppress_blank = c.suppress_blank;
  f.ts_rules = c.timestamp_rules;
  f.K = c.beam_search ? G + 1 : 1;
  f.inv_temp = (!c.beam_search && D->temperature > 0.f) ? 1.0f / D->temperature : 0.f;
  f.seed_lo = static_cast<uint32_t>(D->seed & 0xffffffffull);
  f.seed_hi = static_cast<uint32_t>(D->seed >> 32);
  f.top_val = D->top_val;
  f.top_idx = D->top_idx;
  WB_TRY(launch_filter_topk(f, R, s));
  if (!c.beam_search) {
    GreedyParams g;
    g.tokens = D->tokens[D->cur];
    g.max_ctx = ctx;
    g.R = R;
    g.eot = c.eot;
    g.len_ptr = D->len_ptr;
    g.sum_logprobs = D->sum_lp;
    g.top_val = D->top_val;
    g.top_idx = D->top_idx;
    g.done_flag = D->done_ptr;
    g.skip_flag = D->done_ptr;
    WB_TRY(launch_greedy_update(g, s));
  } else {
    BeamParams b;
    b.tokens_in = D->tokens[D->cur];
    b.tokens_out = D->tokens[D->cur ^ 1];
    b.indir_in = D->indir[D->cur];
    b.indir_out = D->indir[D->cur ^ 1];
    b.max_ctx = ctx;
    b.n_audio = B;
    b.G = G;
    b.eot = c.eot;
    b.max_candidates = c.max_candidates;
    b.len_ptr = D->len_ptr;
    b.sum_logprobs = D->sum_lp;
    b.top_val = D->top_val;
    b.top_idx = D->top_idx;
    b.fin_tokens = D->fin_tokens;
    b.fin_len = D->fin_len;
    b.fin_score = D->fin_score;
    b.fin_count = D->fin_count;
    b.source_out = D->sources;
    b.done_flag = D->done_ptr;
    b.skip_flag = D->done_ptr;
    b.cur_out_ptr = D->cur_ptr;       // the device records which buffer is current: once the done
    b.out_index = D->cur ^ 1;         // flag is up later launches are no-ops and the host view goes stale
    b.n_init = c.n_init;
    b.tickets = D->scalars + 24;
    b.same_in = D->beam_same[D->cur];
    b.same_out = D->beam_same[D->cur ^ 1];
    WB_TRY(launch_beam_update(b, s));
    D->cur ^= 1;
  }
  D->host_len += 1;
  return 0;
}

int decoder_append(Decoder* D, const int32_t* next_host, cudaStream_t s) {
  const int R = D->cfg.n_audio * D->cfg.n_group;
  if (cudaMemcpyAsync(D->sources, next_host, static_cast<size_t>(R) * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
    return set_error(250, "append: upload failed");
  append_tokens_kernel<<<1, 256, 0, s>>>(D->tokens[D->cur], D->m->dims.n_text_ctx, R, D->sources, D->len_ptr,
                                         D->cfg.beam_search ? D->beam_same[D->cur] : nullptr, D->cfg.n_group);
  count_launch();
  D->host_len += 1;
  return 0;
}

static int poll_int(Decoder* D, const int* dev, int* value, cudaStream_t s) {
  if (cudaMemcpyAsync(D->pinned, dev, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return set_error(260, "decoder: flag read failed: %s", cudaGetErrorString(cudaGetLastError()));
  *value = D->pinned[0];
  return 0;
}

// DecodingTask._main_loop for i >= 1 (decoding.py:686-706): step, filters, update; the completion
// flag lives on the device and is polled every 8 iterations (kernels become no-ops once it is set,
// so overshooting leaves the state exactly as it was when the flag went up).
static int run_direct(Decoder* D, int max_steps, int* issued_out, cudaStream_t s) {
  const int ctx = D->m->dims.n_text_ctx;
  int issued = 0;
  for (int i = 0; i < max_steps; ++i) {
    if (D->host_len >= ctx) break;
    int r = decoder_step(D, s);
    if (r) return r;
    r = decoder_select(D, s);
    if (r) return r;
    ++issued;
    if ((i & 7) == 7) {
      int done = 0;
      r = poll_int(D, D->done_ptr, &done, s);
      if (r) return r;
      if (done) break;
    }
  }
  *issued_out = issued;
  return 0;
}

static bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("WB200_NO_GRAPH");
    v = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return v == 1 && g_profile_kernel == 0;   // per-kernel event timing needs real launches
}

int decoder_run(Decoder* D, int max_steps, int* steps_issued, cudaStream_t s) {
  const int ctx = D->m->dims.n_text_ctx;
  int issued = 0;
  if (!graphs_enabled() || max_steps < 6) {
    int r = run_direct(D, max_steps, &issued, s);
    if (steps_issued) *steps_issued = issued;
    return r;
  }
  // ---- graph mode: everything runs on the session's private stream, fenced against the caller's
  if (!D->gstream) {
    if (cudaStreamCreateWithFlags(&D->gstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&D->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&D->ev_out, cudaEventDisableTiming) != cudaSuccess)
      return set_error(280, "decoder: stream/event creation failed");
  }
  cudaStream_t g = D->gstream;
  cudaEventRecord(D->ev_in, s);
  cudaStreamWaitEvent(g, D->ev_in, 0);
  int remaining = max_steps;
  // first iteration directly: warms every kernel (function attributes) before capture
  {
    int n = 0;
    int r = run_direct(D, 1, &n, g);
    if (r) return r;
    issued += n;
    remaining -= n;
    if (n == 0) remaining = 0;
  }
  if (remaining >= 2 && D->host_len + 2 <= ctx && (!D->pair_graph || D->pair_graph_cur != D->cur)) {
    if (D->pair_graph) {
      cudaGraphExecDestroy(D->pair_graph);
      D->pair_graph = nullptr;
    }
    const int cur0 = D->cur, len0 = D->host_len;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(g, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
      return set_error(281, "decoder: stream capture failed to start");
    const unsigned long long launches0 = t_launch_count;
    int r = 0;
    for (int k = 0; k < 2 && !r; ++k) {
      r = decoder_step(D, g);
      if (!r) r = decoder_select(D, g);
    }
    cudaError_t ce = cudaStreamEndCapture(g, &graph);
    D->launches_per_pair = static_cast<int>(t_launch_count - launches0);
    // capture records launches, it does not run them: take this thread's recordings back out of the global count
    __atomic_fetch_sub(&g_launch_count, t_launch_count - launches0, __ATOMIC_RELAXED);
    t_launch_count = launches0;
    D->host_len = len0;
    D->cur = cur0;
    if (r || ce != cudaSuccess || !graph) return set_error(282, "decoder: graph capture failed (%d, %s)", r, cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&D->pair_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return set_error(283, "decoder: graph instantiate failed: %s", cudaGetErrorString(ce));
    D->pair_graph_cur = cur0;
  }
  int pairs_since_poll = 0;
  bool done_seen = false;
  while (remaining >= 2 && D->host_len + 2 <= ctx && D->pair_graph) {
    if (cudaGraphLaunch(D->pair_graph, g) != cudaSuccess) return set_error(284, "decoder: graph launch failed");
    count_launch(D->launches_per_pair);
    D->host_len += 2;
    issued += 2;
    remaining -= 2;
    if (++pairs_since_poll == 4) {
      pairs_since_poll = 0;
      int done = 0;
      int r = poll_int(D, D->done_ptr, &done, g);
      if (r) return r;
      if (done) {
        done_seen = true;
        break;
      }
    }
  }
  if (!done_seen && remaining > 0) {
    int n = 0;
    int r = run_direct(D, remaining, &n, g);
    if (r) return r;
    issued += n;
  }
  cudaEventRecord(D->ev_out, g);
  cudaStreamWaitEvent(s, D->ev_out, 0);
  if (steps_issued) *steps_issued = issued;
  return 0;
}

int decoder_state_ptr(Decoder* D, int what, void** ptr, size_t* bytes, cudaStream_t s) {
  const Model* m = D->m;
  const wb200_decode_config& c = D->cfg;
  const size_t B = c.n_audio, R = B * c.n_group, ctx = m->dims.n_text_ctx;
  const size_t K = c.beam_search ? c.n_group + 1 : 1;
  const size_t mc = c.max_candidates > 0 ? c.max_candidates : 1;
  int cur = 0;
  if (what == WB200_STATE_TOKENS && c.beam_search) {
    int r = poll_int(D, D->cur_ptr, &cur, s);
    if (r) return r;
  }
  switch (what) {
    case WB200_STATE_TOKENS: *ptr = D->tokens[cur]; *bytes = R * ctx * 4; break;
    case WB200_STATE_LENGTH: *ptr = D->len_ptr; *bytes = 4; break;
    case WB200_STATE_SUM_LOGPROBS: *ptr = D->sum_lp; *bytes = R * 4; break;
    case WB200_STATE_NO_SPEECH: *ptr = D->no_speech; *bytes = B * 4; break;
    case WB200_STATE_LOGITS: *ptr = const_cast<float*>(D->logits_cur);
      *bytes = static_cast<size_t>(D->logits_rows) * static_cast<size_t>(D->ldv) * 4; break;
    case WB200_STATE_TOP_VAL: *ptr = D->top_val; *bytes = R * K * 4; break;
    case WB200_STATE_TOP_IDX: *ptr = D->top_idx; *bytes = R * K * 4; break;
    case WB200_STATE_SOURCES: *ptr = D->sources; *bytes = R * 4; break;
    case WB200_STATE_FIN_TOKENS: *ptr = D->fin_tokens; *bytes = B * mc * ctx * 4; break;
    case WB200_STATE_FIN_LEN: *ptr = D->fin_len; *bytes = B * mc * 4; break;
    case WB200_STATE_FIN_SCORE: *ptr = D->fin_score; *bytes = B * mc * 4; break;
    case WB200_STATE_FIN_COUNT: *ptr = D->fin_count; *bytes = B * 4; break;
    case WB200_STATE_DONE: *ptr = D->done_ptr; *bytes = 4; break;
    default: return set_error(270, "decoder: unknown state id %d", what);
  }
  return 0;
}

}  // namespace wb
