// C-ABI entry points (include/whisper_b200.h).  Thin argument checking + dispatch to the
// kernels; no torch types, no exceptions, status codes only.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/whisper_b200.h"
#include "kernels.h"

namespace wb {
unsigned long long g_launch_count = 0;
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
uint64_t wb200_launch_count(void) { return g_launch_count; }

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

}  // extern "C"
