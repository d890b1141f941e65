// Fused log-mel front-end (reference whisper/audio.py:110-157):
//   reflect pad -> periodic Hann(400) -> 400-point real FFT every 160 samples -> |.|^2 ->
//   mel filterbank -> log10(max(., 1e-10))            [kernel 1, also reduces the global max]
//   max(., gmax - 8), (. + 4) / 4                      [kernel 2, elementwise]
// in place of torch.stft (cuFFT) + abs()**2 + a dense sgemm + four elementwise passes.
//
// Kernel 1: one CTA = 32 consecutive frames of one waveform.  The 5360-sample span the frames
// cover is staged in shared memory once (a single cp.async.bulk / TMA 1-D copy for interior CTAs,
// a reflect-indexed gather at the edges), so each audio sample is read from HBM once although
// every sample belongs to 2.5 frames.  Each warp then transforms frames one at a time with a
// mixed-radix FFT, 400 = 16 x 25:
//   lane n2 < 25 : 16-point real DFT over x[25*n1 + n2] (registers), twiddle by W400^(n2*k1)
//   all lanes    : 25-point DFTs across n2 (through shared memory) for the 201 needed bins; bins
//                  with k mod 16 > 8 come from the conjugate-symmetric partner 400 - k.
// The filterbank is applied in its sparse form (<= 16 contiguous non-zero taps per mel row,
// 394 non-zeros in total) and the log-compressed tile is written back frame-contiguous.
#include <math.h>

#include "kernels.h"
#include "ptx.cuh"

namespace wb {

constexpr int kNFFT = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFramesPerCta = 32;
constexpr int kSpan = (kFramesPerCta - 1) * kHop + kNFFT;  // 5360 samples
constexpr int kMelThreads = 256;
constexpr int kMaxTaps = 16;

struct MelTables {
  float hann[kNFFT];
  float2 tw400[25 * 9];  // W400^(n2*k1), index n2*9 + k1
  float2 w25[25];        // W25^m
};
__device__ MelTables g_mel_tables;

struct MelSparse {       // built on the device from the dense (n_mels x 201) matrix
  int start[128];
  int len[128];
  float w[128][kMaxTaps];
  int overflow;
};

__global__ void mel_sparsify_kernel(const float* __restrict__ filters, int n_mels, MelSparse* sp) {
  const int m = threadIdx.x;
  if (m == 0) sp->overflow = 0;
  __syncthreads();
  if (m >= n_mels) return;
  const float* row = filters + m * kBins;
  int lo = kBins, hi = -1;
  for (int k = 0; k < kBins; ++k)
    if (row[k] != 0.f) {
      lo = min(lo, k);
      hi = max(hi, k);
    }
  const int len = hi >= lo ? hi - lo + 1 : 0;
  if (len > kMaxTaps) atomicExch(&sp->overflow, 1);
  sp->start[m] = len ? lo : 0;
  sp->len[m] = min(len, kMaxTaps);
  for (int i = 0; i < kMaxTaps; ++i) sp->w[m][i] = (i < len) ? row[lo + i] : 0.f;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// cos / sin of 2*pi*m/16
__device__ __forceinline__ constexpr float c16(int m) {
  constexpr float t[16] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                           0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                           -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
                           0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
  return t[m & 15];
}
__device__ __forceinline__ constexpr float s16(int m) { return c16(m - 4); }  // sin(x) = cos(x - pi/2)

template <int N_MELS_MAX>
__global__ void __launch_bounds__(kMelThreads)
log_mel_kernel(const float* __restrict__ audio, long long n_samples, int n_frames, int n_mels,
               const MelSparse* __restrict__ sp, float* __restrict__ out, float* __restrict__ gmax,
               int per_row_max) {
  extern __shared__ __align__(128) uint8_t mel_smem[];
  float* s_audio = reinterpret_cast<float*>(mel_smem);                       // kSpan
  float* s_hann = s_audio + kSpan + 16;                                      // 400
  float2* s_tw = reinterpret_cast<float2*>(s_hann + kNFFT);                  // 225
  float2* s_w25 = s_tw + 25 * 9 + 1;                                         // 25
  float2* s_z = s_w25 + 25 + 1;                                              // 8 warps x 9 x 25
  float* s_pow = reinterpret_cast<float*>(s_z + 8 * 9 * 25);                 // 8 warps x 208
  float* s_out = s_pow + 8 * 208;                                            // n_mels x 33
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ float s_wmax[kMelThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerCta;
  const float* x = audio + static_cast<long long>(a) * n_samples;
  const long long span0 = static_cast<long long>(f0) * kHop - kNFFT / 2;     // first sample of the span

  // ---- stage the audio span
  const bool interior = span0 >= 0 && span0 + kSpan <= n_samples &&
                        ((reinterpret_cast<uintptr_t>(x + span0) & 15) == 0);
  if (interior) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_fence_init();
      mbar_expect_tx(&s_bar, kSpan * 4);
      bulk_load_1d(s_audio, x + span0, kSpan * 4, &s_bar);
    }
  } else {
    for (int i = tid; i < kSpan; i += kMelThreads) {
      long long j = span0 + i;
      if (j < 0) j = -j;                                   // reflect (torch.stft center=True)
      if (j >= n_samples) j = 2 * (n_samples - 1) - j;
      s_audio[i] = (j >= 0 && j < n_samples) ? x[j] : 0.f;
    }
  }
  for (int i = tid; i < kNFFT; i += kMelThreads) s_hann[i] = g_mel_tables.hann[i];
  for (int i = tid; i < 25 * 9; i += kMelThreads) s_tw[i] = g_mel_tables.tw400[i];
  if (tid < 25) s_w25[tid] = g_mel_tables.w25[tid];
  __syncthreads();
  if (interior) mbar_wait(&s_bar, 0);

  float2* z = s_z + warp * 9 * 25;
  float* pw = s_pow + warp * 208;
  for (int fi = warp; fi < kFramesPerCta; fi += kMelThreads / 32) {
    const int f = f0 + fi;
    if (f >= n_frames) break;                              // warp-uniform
    const float* fr = s_audio + fi * kHop;
    // ---- step 1+2: lane n2 computes the 16-point real DFT of its decimated sequence
    if (lane < 25) {
      float xs[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) xs[n1] = fr[25 * n1 + lane] * s_hann[25 * n1 + lane];
#pragma unroll
      for (int k1 = 0; k1 < 9; ++k1) {
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          re += xs[n1] * c16(n1 * k1);
          im -= xs[n1] * s16(n1 * k1);
        }
        const float2 t = s_tw[lane * 9 + k1];
        z[k1 * 25 + lane] = make_float2(re * t.x - im * t.y, re * t.y + im * t.x);
      }
    }
    __syncwarp();
    // ---- step 3: 25-point DFT across n2 for each needed bin; power spectrum
    for (int b = lane; b < kBins; b += 32) {
      int k = b;
      if ((k & 15) > 8) k = kNFFT - k;                     // conjugate partner has k1 <= 8
      const int k1 = k & 15, k2 = k >> 4;
      const float2* zr = z + k1 * 25;
      float re = 0.f, im = 0.f;
      int idx = 0;
#pragma unroll 5
      for (int n2 = 0; n2 < 25; ++n2) {
        const float2 w = s_w25[idx];
        const float2 v = zr[n2];
        re += v.x * w.x - v.y * w.y;
        im += v.x * w.y + v.y * w.x;
        idx += k2;
        if (idx >= 25) idx -= 25;
      }
      pw[b] = re * re + im * im;
    }
    __syncwarp();
    // ---- mel projection (sparse rows) + log10
    for (int m = lane; m < n_mels; m += 32) {
      const int st = sp->start[m], ln = sp->len[m];
      float acc = 0.f;
      for (int i = 0; i < ln; ++i) acc += sp->w[m][i] * pw[st + i];
      s_out[m * 33 + fi] = log10f(fmaxf(acc, 1e-10f));
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- write the tile (frames contiguous) and reduce the max
  const int n_valid = min(kFramesPerCta, n_frames - f0);
  float mx = -INFINITY;
  float* ob = out + static_cast<long long>(a) * n_mels * n_frames;
  for (int i = tid; i < n_mels * kFramesPerCta; i += kMelThreads) {
    const int m = i >> 5, fi = i & 31;
    if (fi < n_valid) {
      const float v = s_out[m * 33 + fi];
      ob[static_cast<long long>(m) * n_frames + f0 + fi] = v;
      mx = fmaxf(mx, v);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_wmax[warp] = mx;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kMelThreads / 32; ++w) mx = fmaxf(mx, s_wmax[w]);
    atomic_max_float(gmax + (per_row_max ? a : 0), mx);
  }
}

__global__ void mel_init_max_kernel(float* gmax, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gmax[i] = -INFINITY;
}

__global__ void __launch_bounds__(256) mel_finalize_kernel(float* __restrict__ out, long long per_row,
                                                           long long total, const float* __restrict__ gmax,
                                                           int per_row_max) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  if (i + 4 <= total && (per_row % 4) == 0) {
    const float floor_v = gmax[per_row_max ? i / per_row : 0] - 8.0f;
    float4 v = *reinterpret_cast<float4*>(out + i);
    v.x = (fmaxf(v.x, floor_v) + 4.0f) / 4.0f;
    v.y = (fmaxf(v.y, floor_v) + 4.0f) / 4.0f;
    v.z = (fmaxf(v.z, floor_v) + 4.0f) / 4.0f;
    v.w = (fmaxf(v.w, floor_v) + 4.0f) / 4.0f;
    *reinterpret_cast<float4*>(out + i) = v;
  } else {
    for (long long j = i; j < min(i + 4, total); ++j) {
      const float floor_v = gmax[per_row_max ? j / per_row : 0] - 8.0f;
      out[j] = (fmaxf(out[j], floor_v) + 4.0f) / 4.0f;
    }
  }
}

size_t log_mel_workspace_bytes(int n_audio) {
  return sizeof(MelSparse) + sizeof(float) * static_cast<size_t>(n_audio > 0 ? n_audio : 1) + 64;
}

int launch_log_mel(const float* audio, int n_audio, long long n_samples, int n_mels, const float* filters,
                   float* out, void* workspace, int per_row_max, cudaStream_t s) {
  if (n_audio <= 0) return 0;
  if (n_mels < 1 || n_mels > 128) return 60;
  if (n_samples < kNFFT / 2 + 1) return 61;                // reflect padding needs > 200 samples
  static bool tables_ready = false;
  if (!tables_ready) {
    static MelTables h;
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < kNFFT; ++n) h.hann[n] = static_cast<float>(0.5 - 0.5 * cos(two_pi * n / kNFFT));
    for (int n2 = 0; n2 < 25; ++n2)
      for (int k1 = 0; k1 < 9; ++k1) {
        const double ang = -two_pi * (n2 * k1) / 400.0;
        h.tw400[n2 * 9 + k1] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
      }
    for (int m = 0; m < 25; ++m) {
      const double ang = -two_pi * m / 25.0;
      h.w25[m] = make_float2(static_cast<float>(cos(ang)), static_cast<float>(sin(ang)));
    }
    if (cudaMemcpyToSymbolAsync(g_mel_tables, &h, sizeof(h), 0, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return 62;
    if (cudaStreamSynchronize(s) != cudaSuccess) return 62;   // one-time; h is static host memory
    tables_ready = true;
  }
  MelSparse* sp = reinterpret_cast<MelSparse*>(workspace);
  float* gmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((sizeof(MelSparse) + 63) / 64) * 64);
  const int n_frames = static_cast<int>(n_samples / kHop);
  const int n_max = per_row_max ? n_audio : 1;
  mel_sparsify_kernel<<<1, 128, 0, s>>>(filters, n_mels, sp);
  mel_init_max_kernel<<<(n_max + 255) / 256, 256, 0, s>>>(gmax, n_max);
  const size_t smem = (kSpan + 16 + kNFFT) * 4 + (25 * 9 + 1 + 25 + 1 + 8 * 9 * 25) * 8 + (8 * 208) * 4 +
                      static_cast<size_t>(n_mels) * 33 * 4 + 128;
  auto kern = log_mel_kernel<128>;
  static SmemOptIn optin;
  if (!optin.ensure(kern, 96 * 1024)) return 63;
  dim3 grid((n_frames + kFramesPerCta - 1) / kFramesPerCta, n_audio);
  ProfileScope prof(PROF_MEL, s);
  kern<<<grid, kMelThreads, smem, s>>>(audio, n_samples, n_frames, n_mels, sp, out, gmax, per_row_max);
  const long long per_row = static_cast<long long>(n_mels) * n_frames;
  const long long total = per_row * n_audio;
  mel_finalize_kernel<<<static_cast<unsigned>((total / 4 + 255) / 256 + 1), 256, 0, s>>>(out, per_row, total, gmax,
                                                                                       per_row_max);
  count_launch(4);
  return cudaGetLastError() == cudaSuccess ? 0 : 64;
}

}  // namespace wb
