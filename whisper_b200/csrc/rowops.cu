// Row-wise HBM-bound kernels: LayerNorm (reference model.py:39-41) and the fp32 (B,C,T) ->
// 16-bit (B,T,C) transpose that puts the mel spectrogram time-major for the conv-as-GEMM path.
#include "kernels.h"
#include "ptx.cuh"

namespace wb {

// One warp per row.  The row (d <= 2048) is held in registers between the statistics pass and the
// normalise pass, so each element is read from HBM exactly once: 16-byte loads, 8 values per lane
// per step, lanes interleaved so a warp reads 512 contiguous bytes per step.
template <typename T, int MAX_STEPS>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, long long ldx,
                                                        T* __restrict__ y, long long ldy,
                                                        const float* __restrict__ g,
                                                        const float* __restrict__ b, int rows, int d,
                                                        const int* skip_flag) {
  pdl_launch_dependents();
  pdl_wait();
  if (skip_flag && *skip_flag) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const T* xr = x + static_cast<long long>(warp) * ldx;
  T* yr = y + static_cast<long long>(warp) * ldy;
  const int steps = (d + 255) / 256;
  float v[MAX_STEPS][8];
  float sum = 0.f;
#pragma unroll
  for (int s = 0; s < MAX_STEPS; ++s) {
    const int c = s * 256 + lane * 8;
    if (s < steps && c < d) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = Cvt<T>::unpack2(w[e]);
        v[s][2 * e] = f.x;
        v[s][2 * e + 1] = f.y;
        sum += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[s][e] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(d);
  float sq = 0.f;
#pragma unroll
  for (int s = 0; s < MAX_STEPS; ++s) {
    const int c = s * 256 + lane * 8;
    if (s < steps && c < d) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = v[s][e] - mean;
        sq += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(d) + 1e-5f);
#pragma unroll
  for (int s = 0; s < MAX_STEPS; ++s) {
    const int c = s * 256 + lane * 8;
    if (s < steps && c < d) {
      const float4 g0 = *reinterpret_cast<const float4*>(g + c);
      const float4 g1 = *reinterpret_cast<const float4*>(g + c + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(b + c);
      const float4 b1 = *reinterpret_cast<const float4*>(b + c + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[s][e] - mean) * rstd * gg[e] + bb[e];
      uint4 u;
      u.x = Cvt<T>::pack2(o[0], o[1]);
      u.y = Cvt<T>::pack2(o[2], o[3]);
      u.z = Cvt<T>::pack2(o[4], o[5]);
      u.w = Cvt<T>::pack2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + c) = u;
    }
  }
}

int launch_layernorm(int dtype, const void* x, long long ldx, void* y, long long ldy, const float* g,
                     const float* b, int rows, int d, cudaStream_t s, const int* skip_flag) {
  if (rows <= 0) return 0;
  if (d % 8 || d > 2048 || ldx % 8 || ldy % 8) return 20;
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  ProfileScope prof(PROF_LAYERNORM, s);
  cudaError_t le;
  if (dtype == DT_BF16)
    le = launch_pdl(layernorm_kernel<__nv_bfloat16, 8>, dim3(blocks), dim3(threads), 0, s,
                    static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y), ldy, g, b, rows, d, skip_flag);
  else
    le = launch_pdl(layernorm_kernel<__half, 8>, dim3(blocks), dim3(threads), 0, s, static_cast<const __half*>(x), ldx,
                    static_cast<__half*>(y), ldy, g, b, rows, d, skip_flag);
  count_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 21;
}

// 32x32 tile transpose through shared memory; reads coalesced along T, writes coalesced along C.
template <typename T>
__global__ void __launch_bounds__(256) transpose_to16_kernel(const float* __restrict__ x,
                                                             T* __restrict__ y, int C, int Tn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* xb = x + static_cast<long long>(b) * C * Tn;
  T* yb = y + static_cast<long long>(b) * C * Tn;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, t = t0 + tx;
    tile[ty + i * 8][tx] = (c < C && t < Tn) ? xb[static_cast<long long>(c) * Tn + t] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + i * 8, c = c0 + tx;
    if (t < Tn && c < C) yb[static_cast<long long>(t) * C + c] = Cvt<T>::from_f(tile[tx][ty + i * 8]);
  }
}

int launch_transpose_to16(int dtype, const float* x, void* y, int B, int C, int T, cudaStream_t s) {
  if (B <= 0) return 0;
  dim3 grid((T + 31) / 32, (C + 31) / 32, B);
  if (dtype == DT_BF16)
    transpose_to16_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, static_cast<__nv_bfloat16*>(y), C, T);
  else
    transpose_to16_kernel<__half><<<grid, 256, 0, s>>>(x, static_cast<__half*>(y), C, T);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 22;
}

}  // namespace wb
