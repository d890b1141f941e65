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

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
