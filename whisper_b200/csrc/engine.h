// Host-side handle structs behind the opaque wb200_model / wb200_decoder pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/whisper_b200.h"
#include "dec_layer.h"

namespace wb {

// tensor slot order of wb200_model_create (documented in include/whisper_b200.h)
enum GlobalSlot {
  G_CONV1_W, G_CONV1_B, G_CONV2_W, G_CONV2_B, G_ENC_POS, G_ENC_LN_POST_W, G_ENC_LN_POST_B,
  G_TOK_EMB16, G_TOK_EMB32, G_DEC_POS, G_DEC_LN_W, G_DEC_LN_B, G_COUNT
};
enum EncSlot {
  E_ATTN_LN_W, E_ATTN_LN_B, E_QKV_W, E_QKV_B, E_OUT_W, E_OUT_B, E_MLP_LN_W, E_MLP_LN_B,
  E_FC1_W, E_FC1_B, E_FC2_W, E_FC2_B, E_COUNT
};
enum DecSlot {
  D_ATTN_LN_W, D_ATTN_LN_B, D_QKV_W, D_QKV_B, D_OUT_W, D_OUT_B, D_CROSS_LN_W, D_CROSS_LN_B,
  D_CQ_W, D_CQ_B, D_CKV_W, D_CKV_B, D_COUT_W, D_COUT_B, D_MLP_LN_W, D_MLP_LN_B,
  D_FC1_W, D_FC1_B, D_FC2_W, D_FC2_B,
  // LayerNorm folded into the consuming Linear for the fused decoder-layer kernel (dec_layer.cu):
  // *_WF = W (.) gamma in the 16-bit type, *_C1 = fp32 row sums of WF, *_C2 = fp32 W beta + bias
  D_QKV_WF, D_QKV_C1, D_QKV_C2, D_CQ_WF, D_CQ_C1, D_CQ_C2, D_FC1_WF, D_FC1_C1, D_FC1_C2, D_COUNT
};

struct Dims {
  int n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer;
};

struct Model {
  Dims dims;
  int dtype;
  std::vector<const void*> t;   // device pointers owned by the caller
  const void* const* enc_layer(int l) const { return t.data() + G_COUNT + l * E_COUNT; }
  const void* const* dec_layer(int l) const {
    return t.data() + G_COUNT + dims.n_audio_layer * E_COUNT + l * D_COUNT;
  }
  static int num_tensors(const Dims& d) { return G_COUNT + d.n_audio_layer * E_COUNT + d.n_text_layer * D_COUNT; }
};

struct Decoder {
  const Model* m = nullptr;
  wb200_decode_config cfg;
  long long ldv = 0;
  // workspace slices
  void *cross_kv = nullptr, *self_k = nullptr, *self_v = nullptr;
  void *x = nullptr, *ln = nullptr, *qkv = nullptr, *att = nullptr, *q = nullptr, *hid = nullptr, *sel = nullptr;
  float* logits = nullptr;
  float* partial = nullptr;
  int* counters = nullptr;
  int n_counters = 0;
  float* gemm_ws = nullptr;        // split-K slabs for the skinny decode-step GEMMs
  size_t gemm_ws_bytes = 0;
  int* gemm_counters = nullptr;    // per-tile tickets, zero between launches
  int* tokens[2] = {nullptr, nullptr};
  int* indir[2] = {nullptr, nullptr};
  float* sum_lp = nullptr;
  float* no_speech = nullptr;
  float* top_val = nullptr;
  int* top_idx = nullptr;
  int* sources = nullptr;
  int* fin_tokens = nullptr;
  int* fin_len = nullptr;
  float* fin_score = nullptr;
  int* fin_count = nullptr;
  uint32_t* suppress_mask = nullptr;
  uint32_t* blank_mask = nullptr;
  int* init_tokens = nullptr;
  unsigned char* beam_same[2] = {nullptr, nullptr};   // [n_audio][16][16] prefix-equality of beams (ping-pong with tokens)
  int* scalars = nullptr;      // [0] length, [8] done flag, [16] current ping-pong buffer
  int* len_ptr = nullptr;
  int* done_ptr = nullptr;
  int* cur_ptr = nullptr;
  // host mirrors
  int cur = 0;                 // ping-pong buffer the NEXT kernels read (host view)
  int host_len = 0;            // tokens per row as far as the host has issued work
  const float* logits_cur = nullptr;
  int logits_row_div = 1;
  int logits_rows = 0;         // rows currently valid at logits_cur
  bool forward_only = false;   // all_logits sessions cannot step / select
  int* pinned = nullptr;       // pinned host scratch for flag polling
  std::vector<int> align_heads;     // (layer, head) pairs whose cross-attention scores are exported
  float* align_qk = nullptr;        // [n_heads, n_init, n_audio_ctx] fp32 (caller memory)
  // CUDA-graph replay of the decode loop: two iterations (step, select, step, select) per graph so
  // the ping-pong token / parent-table buffers are back where they started after every replay.
  cudaStream_t gstream = nullptr;   // private capture / replay stream (capture is illegal on the legacy stream)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t pair_graph = nullptr;
  int pair_graph_cur = -1;          // value of `cur` the graph was captured at
  int launches_per_pair = 0;        // kernels inside one replay (for wb200_launch_count)
  bool kv_head_major = false;       // kv caches stored per head ([.., head, position, 64]); fixed at create
  bool kv_window = false;           // self caches in the beam-window layout [audio][head][position][slot][64]; fixed at create
  // fused decoder-layer GEMM chain of the step path (dec_layer.cu): per layer [QKV] (layer 0 only; later layers get
  // theirs from the previous layer's tail), [out-proj, cross-query], [cross-out, fc1, fc2, next layer's QKV]
  bool fused = false;
  std::vector<DLLaunch> dl_head, dl_mid, dl_tail;
  // few-rows sessions: the whole decoder stack of an iteration (every Linear chain AND both attentions of every layer;
  // optionally the final LayerNorm and the logits) as ONE launch of dec_rows_kernel driven by a phase table in device memory
  bool stack_ready = false;
  bool stack_has_logits = false;    // the table ends with the final LayerNorm and the logits (mode 2)
  DLLaunch dl_stack;
  std::vector<DLPhase> stack_host;
  DLPhase* stack_table = nullptr;   // device copy of stack_host
  float4* ln_part = nullptr;        // LN partial statistics of the residual stream
  int ln_ld = 0;
  unsigned int* dl_sync = nullptr;  // grid-barrier / exit counters of the fused kernel
  // GreedyDecoder temperature sampling (wb200_decoder_set_sampling); 0 = argmax
  float temperature = 0.f;
  unsigned long long seed = 0;
};

size_t encoder_workspace_bytes(const Model* m, int B);
int encoder_forward(const Model* m, const float* mel, int B, void* out, void* ws, size_t ws_bytes, cudaStream_t s);
size_t decoder_workspace_bytes(const Model* m, const wb200_decode_config* c);
int decoder_create(const Model* m, const wb200_decode_config* c, void* ws, size_t ws_bytes, Decoder** out, cudaStream_t s);
int decoder_set_audio(Decoder* D, const void* features, cudaStream_t s);
int decoder_prefill(Decoder* D, const int32_t* init_tokens_host, cudaStream_t s);
int decoder_step(Decoder* D, cudaStream_t s);
int decoder_select(Decoder* D, cudaStream_t s);
int decoder_set_sampling(Decoder* D, float temperature, unsigned long long seed);
int decoder_append(Decoder* D, const int32_t* next_host, cudaStream_t s);
int decoder_run(Decoder* D, int max_steps, int* steps_issued, cudaStream_t s);
int decoder_state_ptr(Decoder* D, int what, void** ptr, size_t* bytes, cudaStream_t s);

}  // namespace wb
