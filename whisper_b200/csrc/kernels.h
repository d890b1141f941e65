// Internal C++ launch interface shared by the .cu translation units and api.cu.
// Nothing here is exported; the C-ABI lives in include/whisper_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb {

enum { DT_BF16 = 0, DT_F16 = 1 };

// kernels launched by this library since load (reported by bench.py as gpu_launches)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += static_cast<unsigned long long>(n); }

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
  void* C = nullptr;
  long long ldc = 0;
  int gelu = 0;
  int out_f32 = 0;
  int block_n = 0;  // 0 = choose (256 for large M, 64 for skinny M)
};
int launch_linear(const LinearArgs& a, cudaStream_t s);

// Row LayerNorm in fp32 (reference model.py:39-41): y = (x - mean) / sqrt(var + 1e-5) * g + b
int launch_layernorm(int dtype, const void* x, long long ldx, void* y, long long ldy, const float* g,
                     const float* b, int rows, int d, cudaStream_t s);

// (B, C, T) fp32 -> (B, T, C) 16-bit.  Used to put the mel spectrogram time-major for the conv GEMM.
int launch_transpose_to16(int dtype, const float* x, void* y, int B, int C, int T, cudaStream_t s);

// Encoder (non-causal) multi-head attention over packed qkv [B*T, 3*d] -> out [B*T, d]
int launch_enc_attention(int dtype, const void* qkv, void* out, int B, int T, int n_head,
                         cudaStream_t s);

}  // namespace wb
