// Word-timing math on the device (reference whisper/timing.py:19-151 and the two Triton kernels in
// whisper/triton_ops.py that back it on CUDA): sliding-window median filter and dynamic time warping
// with the backtrace kept on the GPU (the reference copies the trace matrix to the host and walks
// it with numba, timing.py:138).
#include "kernels.h"
#include "ptx.cuh"

namespace wb {

// -------------------------------------------------------------------------------------------------
// median filter along the last dim with reflect padding (timing.py:19-54, triton_ops.py:43-117).
// One thread per output element; the window lives in registers and is sorted by a compare-exchange
// network (comparisons only -> bit-exact with any correct median).
// -------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) median_filter_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            long long rows, int T) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * T) return;
  const long long r = idx / T;
  const int t = static_cast<int>(idx - r * T);
  const float* xr = x + r * T;
  constexpr int P = W / 2;
  float v[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    int j = t - P + i;
    if (j < 0) j = -j;                       // reflect (F.pad mode="reflect", timing.py:35)
    if (j >= T) j = 2 * (T - 1) - j;
    v[i] = xr[j];
  }
  // partial selection: after P+1 passes of bubbling the minimum forward, v[P] is the median
#pragma unroll
  for (int i = 0; i <= P; ++i) {
#pragma unroll
    for (int j = W - 1; j > i; --j) {
      const float a = v[j - 1], b = v[j];
      v[j - 1] = fminf(a, b);
      v[j] = fmaxf(a, b);
    }
  }
  y[idx] = v[P];
}

int launch_median_filter(const float* x, float* y, long long rows, int T, int width, cudaStream_t s) {
  if (rows <= 0 || T <= 0) return 0;
  if (width < 1 || (width & 1) == 0) return 70;
  if (T <= width / 2) return 71;              // caller must pass such inputs through unchanged
  const long long n = rows * T;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  switch (width) {
#define WB_MED(Wv) case Wv: median_filter_kernel<Wv><<<blocks, 256, 0, s>>>(x, y, rows, T); break;
    WB_MED(1) WB_MED(3) WB_MED(5) WB_MED(7) WB_MED(9) WB_MED(11) WB_MED(13) WB_MED(15) WB_MED(17)
    WB_MED(19) WB_MED(21)
#undef WB_MED
    default: return 72;
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 73;
}

// -------------------------------------------------------------------------------------------------
// DTW (timing.py:82-138, triton_ops.py:13-40): anti-diagonal wavefront in one CTA; the three live
// diagonals stay in shared memory, only the 1-byte trace goes to global memory.
// tie_mode 0: the CUDA/Triton rule (diag, then up, then left win ties, `<=`), which is what the
//             reference executes for CUDA tensors;  1: the CPU/numba rule (strict `<`).
// -------------------------------------------------------------------------------------------------
constexpr int kDtwThreads = 1024;

__global__ void __launch_bounds__(kDtwThreads) dtw_kernel(const float* __restrict__ x, int N, int M,
                                                          unsigned char* __restrict__ trace, int tie_mode) {
  extern __shared__ float dtw_sm[];           // 3 x (N + 1)
  float* diag[3] = {dtw_sm, dtw_sm + (N + 1), dtw_sm + 2 * (N + 1)};
  const int tid = threadIdx.x;
  // diagonal k holds cost(i, k - i) at index i.  Seed k = 0 and k = 1 (only boundary cells).
  for (int i = tid; i <= N; i += kDtwThreads) {
    diag[0][i] = (i == 0) ? 0.f : INFINITY;   // k = 0: cost(0,0)
    diag[1][i] = INFINITY;                    // k = 1: cost(0,1), cost(1,0)
    diag[2][i] = INFINITY;
  }
  __syncthreads();
  for (int k = 2; k <= N + M; ++k) {
    float* cur = diag[k % 3];
    const float* p1 = diag[(k + 2) % 3];      // k - 1
    const float* p2 = diag[(k + 1) % 3];      // k - 2
    for (int i = tid; i <= N; i += kDtwThreads) {
      const int j = k - i;
      float c = INFINITY;
      if (i >= 1 && j >= 1 && j <= M) {
        const float c0 = p2[i - 1];           // cost(i-1, j-1)
        const float c1 = p1[i - 1];           // cost(i-1, j)
        const float c2 = p1[i];               // cost(i, j-1)
        int t;
        float best;
        if (tie_mode == 0) {
          t = 2;
          if (c1 <= c0 && c1 <= c2) t = 1;
          if (c0 <= c1 && c0 <= c2) t = 0;
          best = fminf(fminf(c0, c1), c2);
        } else {
          if (c0 < c1 && c0 < c2) { best = c0; t = 0; }
          else if (c1 < c0 && c1 < c2) { best = c1; t = 1; }
          else { best = c2; t = 2; }
        }
        c = x[static_cast<long long>(i - 1) * M + (j - 1)] + best;
        trace[static_cast<long long>(i) * (M + 1) + j] = static_cast<unsigned char>(t);
      } else if (i == 0 && j == 0) {
        c = 0.f;
      }
      cur[i] = c;
    }
    __syncthreads();
  }
}

// single-thread walk from (N, M) to (0, 0); writes the path reversed into out[0..], out[cap..]
__global__ void dtw_backtrace_kernel(const unsigned char* __restrict__ trace, int N, int M, int* __restrict__ path,
                                     int* __restrict__ path_len) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int cap = N + M + 1;
  int i = N, j = M, n = 0;
  while (i > 0 || j > 0) {
    path[n] = i - 1;
    path[cap + n] = j - 1;
    ++n;
    int t;
    if (i == 0) t = 2;                        // trace[0, :] = 2   (timing.py:61)
    else if (j == 0) t = 1;                   // trace[:, 0] = 1   (timing.py:62)
    else t = trace[static_cast<long long>(i) * (M + 1) + j];
    if (t == 0) { --i; --j; }
    else if (t == 1) --i;
    else --j;
  }
  // reverse in place (timing.py:78-79)
  for (int a = 0, b = n - 1; a < b; ++a, --b) {
    int t0 = path[a]; path[a] = path[b]; path[b] = t0;
    int t1 = path[cap + a]; path[cap + a] = path[cap + b]; path[cap + b] = t1;
  }
  *path_len = n;
}

size_t dtw_workspace_bytes(int N, int M) { return static_cast<size_t>(N + 1) * (M + 1) + 64; }

int launch_dtw(const float* x, int N, int M, int* path, int* path_len, void* workspace, int tie_mode,
               cudaStream_t s) {
  if (N <= 0 || M <= 0) return 74;
  const size_t smem = 3 * static_cast<size_t>(N + 1) * sizeof(float);
  if (smem > 200 * 1024) return 75;
  static SmemOptIn optin;
  if (!optin.ensure(dtw_kernel, 200 * 1024)) return 76;
  unsigned char* trace = reinterpret_cast<unsigned char*>(workspace);
  dtw_kernel<<<1, kDtwThreads, smem, s>>>(x, N, M, trace, tie_mode);
  dtw_backtrace_kernel<<<1, 32, 0, s>>>(trace, N, M, path, path_len);
  count_launch(2);
  return cudaGetLastError() == cudaSuccess ? 0 : 77;
}

// -------------------------------------------------------------------------------------------------
// find_alignment tensor part (timing.py:185-216)
// -------------------------------------------------------------------------------------------------
// Pre-softmax cross-attention scores of ONE head: out[i][t] = (q_i . k_t) / sqrt(64).  The reference
// gets them by re-running the whole model with SDPA disabled (model.py:129-137); here the decoder
// prefill exports them for the alignment heads only.  One thread per key, the query row in smem.
template <typename T>
__global__ void __launch_bounds__(256) qk_export_kernel(const T* __restrict__ q, long long ldq,
                                                        const T* __restrict__ k, long long ldk,
                                                        float* __restrict__ out, int T_keys) {
  __shared__ float sq[64];
  const int i = blockIdx.y;
  if (threadIdx.x < 64) sq[threadIdx.x] = Cvt<T>::to_f(q[i * ldq + threadIdx.x]);
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T_keys) return;
  const uint4* kr = reinterpret_cast<const uint4*>(k + static_cast<long long>(t) * ldk);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = kr[c];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = Cvt<T>::unpack2(w[e]);
      acc = fmaf(sq[c * 8 + 2 * e], f.x, acc);
      acc = fmaf(sq[c * 8 + 2 * e + 1], f.y, acc);
    }
  }
  out[static_cast<long long>(i) * T_keys + t] = acc * 0.125f;
}

int launch_qk_export(int dtype, const void* q, long long ldq, const void* k, long long ldk, float* out,
                     int n_q, int T, cudaStream_t s) {
  if (n_q <= 0 || T <= 0) return 0;
  dim3 grid((T + 255) / 256, n_q);
  if (dtype == DT_BF16)
    qk_export_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(q), ldq,
                                                         static_cast<const __nv_bfloat16*>(k), ldk, out, T);
  else
    qk_export_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(q), ldq,
                                                  static_cast<const __half*>(k), ldk, out, T);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 78;
}

// softmax over the first n_frames frames of one (head, token) row (timing.py:209)
__global__ void __launch_bounds__(256) align_softmax_kernel(const float* __restrict__ qk, float* __restrict__ w,
                                                            int t_stride, int n_frames, float qk_scale) {
  __shared__ float red[8];
  const float* x = qk + static_cast<long long>(blockIdx.x) * t_stride;
  float* y = w + static_cast<long long>(blockIdx.x) * n_frames;
  float m = -INFINITY;
  for (int t = threadIdx.x; t < n_frames; t += 256) m = fmaxf(m, x[t] * qk_scale);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < n_frames; t += 256) {
    const float e = expf(x[t] * qk_scale - m);
    y[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int t = threadIdx.x; t < n_frames; t += 256) y[t] *= inv;
}

// z-score over the token axis for every (head, frame): std_mean(dim=-2, unbiased=False) (timing.py:210-211)
__global__ void __launch_bounds__(256) align_zscore_kernel(float* __restrict__ w, int n_tokens, int n_frames) {
  const int h = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames) return;
  float* base = w + static_cast<long long>(h) * n_tokens * n_frames + t;
  float mean = 0.f;
  for (int i = 0; i < n_tokens; ++i) mean += base[static_cast<long long>(i) * n_frames];
  mean /= static_cast<float>(n_tokens);
  float var = 0.f;
  for (int i = 0; i < n_tokens; ++i) {
    const float d = base[static_cast<long long>(i) * n_frames] - mean;
    var += d * d;
  }
  const float stdv = sqrtf(var / static_cast<float>(n_tokens));
  for (int i = 0; i < n_tokens; ++i) {
    float* p = base + static_cast<long long>(i) * n_frames;
    *p = (*p - mean) / stdv;
  }
}

// mean over heads (timing.py:214), optionally negated for dtw(-matrix) (timing.py:216)
__global__ void __launch_bounds__(256) align_head_mean_kernel(const float* __restrict__ w, float* __restrict__ out,
                                                              int n_heads, long long per_head, float scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per_head) return;
  float acc = 0.f;
  for (int h = 0; h < n_heads; ++h) acc += w[h * per_head + i];
  out[i] = acc * scale;
}

int launch_alignment_weights(const float* qk, int n_heads, int n_tokens, int t_stride, int n_frames,
                             float qk_scale, int medfilt_width, int negate, float* out, float* scratch,
                             cudaStream_t s) {
  if (n_heads <= 0 || n_tokens <= 0 || n_frames <= 0 || n_frames > t_stride) return 79;
  const long long per_head = static_cast<long long>(n_tokens) * n_frames;
  float* w0 = scratch;
  float* w1 = scratch + n_heads * per_head;
  align_softmax_kernel<<<n_heads * n_tokens, 256, 0, s>>>(qk, w0, t_stride, n_frames, qk_scale);
  dim3 gz((n_frames + 255) / 256, n_heads);
  align_zscore_kernel<<<gz, 256, 0, s>>>(w0, n_tokens, n_frames);
  count_launch(2);
  const float* filtered = w0;
  if (n_frames > medfilt_width / 2) {                       // timing.py:22-24: shorter rows pass through
    int r = launch_median_filter(w0, w1, static_cast<long long>(n_heads) * n_tokens, n_frames, medfilt_width, s);
    if (r) return r;
    filtered = w1;
  }
  const float scale = (negate ? -1.0f : 1.0f) / static_cast<float>(n_heads);
  align_head_mean_kernel<<<static_cast<unsigned>((per_head + 255) / 256), 256, 0, s>>>(filtered, out, n_heads, per_head, scale);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 80;
}

}  // namespace wb
