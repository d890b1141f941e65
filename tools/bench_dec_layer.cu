// Stand-alone timing / tracing harness of the fused decoder-layer kernel (whisper_b200/csrc/dec_layer.cu) at the headline
// shape (R = 320 rows, d = 1280): the three launches of a layer - {QKV}, {out-proj, cross-query}, {cross-out, fc1, fc2, QKV}
// - on random data, timed with CUDA events, plus the kernel's per-phase clock stamps reduced over CTAs.
// Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
//        -o tools/bin/bench_dec_layer tools/bench_dec_layer.cu whisper_b200/csrc/dec_layer.cu
// Usage: bench_dec_layer [R=320] [d=1280] [iters=200]
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../whisper_b200/csrc/dec_layer.h"
#include "../whisper_b200/csrc/kernels.h"

namespace wb {   // the pieces of the library dec_layer.cu links against
unsigned long long g_launch_count = 0;
thread_local unsigned long long t_launch_count = 0;
int g_profile_kernel = 0;
int g_pdl_on = 1;
void profile_mark(cudaStream_t, bool) {}
}  // namespace wb
using namespace wb;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__global__ void fill16(unsigned short* p, size_t n, unsigned seed, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)(i * 2654435761u) ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float v = ((h & 0xffff) / 32768.0f - 1.0f) * scale;
    unsigned u = __float_as_uint(v);
    p[i] = (unsigned short)(u >> 16);   // bf16 truncation
  }
}
__global__ void fill32(float* p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_stats(float4* p, int n, float d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = make_float4(d, 0.f, d * 0.01f, 0.f);
}

template <typename Tp>
static Tp* dalloc(size_t n) {
  Tp* p;
  CK(cudaMalloc(&p, n * sizeof(Tp)));
  return p;
}

int main(int argc, char** argv) {
  const int R = argc > 1 ? atoi(argv[1]) : 320;
  const int d = argc > 2 ? atoi(argv[2]) : 1280;
  const int iters = argc > 3 ? atoi(argv[3]) : 200;
  const int grid = dl_grid_size();
  printf("R=%d d=%d grid=%d supported=%d\n", R, d, grid, (int)dl_supported(R, d, grid));
  if (!dl_supported(R, d, grid)) return 1;
  auto w16 = [&](size_t n, unsigned seed, float scale) {
    unsigned short* p = dalloc<unsigned short>(n);
    fill16<<<1024, 256>>>(p, n, seed, scale);
    return p;
  };
  auto f32 = [&](size_t n, float v) {
    float* p = dalloc<float>(n);
    fill32<<<256, 256>>>(p, n, v);
    return p;
  };
  const float ws = 0.02f;
  unsigned short *x = w16((size_t)R * d, 1, 1.f), *att = w16((size_t)R * d, 2, 1.f), *qkv = w16((size_t)R * 3 * d, 3, 1.f),
                 *q = w16((size_t)R * d, 4, 1.f), *hid = w16((size_t)R * 4 * d, 5, 1.f);
  unsigned short *Wqkv = w16((size_t)3 * d * d, 11, ws), *Wo = w16((size_t)d * d, 12, ws), *Wcq = w16((size_t)d * d, 13, ws),
                 *Wco = w16((size_t)d * d, 14, ws), *W1 = w16((size_t)4 * d * d, 15, ws), *W2 = w16((size_t)4 * d * d, 16, ws);
  unsigned short* bias = w16((size_t)4 * d, 17, 0.01f);
  float *c1 = f32((size_t)4 * d, 0.f), *c2 = f32((size_t)4 * d, 0.f);
  const int ln_ld = (R + 63) / 64 * 64;
  float4* ln_part = dalloc<float4>((size_t)256 * ln_ld);
  fill_stats<<<256, 256>>>(ln_part, 256 * ln_ld, 0.f);
  fill_stats<<<16, 256>>>(ln_part, ln_ld, (float)d);       // slot 0: a full-row partial (as the embedding kernel leaves it)
  unsigned int* sync = dalloc<unsigned int>(64);
  CK(cudaMemset(sync, 0, 256));
  unsigned long long* trace = dalloc<unsigned long long>((size_t)grid * kDLMaxPhases * 8);
  CK(cudaDeviceSynchronize());

  DLLaunch head, mid, tail;
  const int dt = DT_BF16;
  dl_init_launch(head, dt, R, grid, ln_part, ln_ld, sync, nullptr, 0);
  int rc = dl_fill_phase(head, 0, dt, R, grid, x, d, Wqkv, 3 * d, d, nullptr, c1, c2, DL_FOLD, qkv, 3LL * d);
  head.p.n_phases = 1;
  dl_init_launch(mid, dt, R, grid, ln_part, ln_ld, sync, nullptr, 0);
  rc |= dl_fill_phase(mid, 0, dt, R, grid, att, d, Wo, d, d, bias, nullptr, nullptr, DL_RESID | DL_STATS, x, d);
  rc |= dl_fill_phase(mid, 1, dt, R, grid, x, d, Wcq, d, d, nullptr, c1, c2, DL_FOLD, q, d);
  mid.p.n_phases = 2;
  dl_init_launch(tail, dt, R, grid, ln_part, ln_ld, sync, nullptr, 0);
  rc |= dl_fill_phase(tail, 0, dt, R, grid, att, d, Wco, d, d, bias, nullptr, nullptr, DL_RESID | DL_STATS, x, d);
  rc |= dl_fill_phase(tail, 1, dt, R, grid, x, d, W1, 4 * d, d, nullptr, c1, c2, DL_FOLD | DL_GELU, hid, 4LL * d);
  rc |= dl_fill_phase(tail, 2, dt, R, grid, hid, 4LL * d, W2, d, 4 * d, bias, nullptr, nullptr, DL_RESID | DL_STATS, x, d);
  rc |= dl_fill_phase(tail, 3, dt, R, grid, x, d, Wqkv, 3 * d, d, nullptr, c1, c2, DL_FOLD, qkv, 3LL * d);
  tail.p.n_phases = 4;
  if (rc) {
    printf("plan failed %d\n", rc);
    return 1;
  }
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct Item { const char* name; DLLaunch* L; };
  Item items[] = {{"head {qkv}", &head}, {"mid  {out, cq}", &mid}, {"tail {cout, fc1, fc2, qkv}", &tail}};
  // variants of the tail with fewer phases, to see the cost of each phase and of the barriers
  DLLaunch tail1 = tail, tail2 = tail, tail3 = tail;
  tail1.p.n_phases = 1;
  tail2.p.n_phases = 2;
  tail3.p.n_phases = 3;
  Item extra[] = {{"tail[:1] {cout}", &tail1}, {"tail[:2] {cout, fc1}", &tail2}, {"tail[:3] {cout, fc1, fc2}", &tail3}};
  auto time_it = [&](const Item& it) {
    for (int i = 0; i < 10; ++i)
      if (dl_launch(*it.L, s)) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) dl_launch(*it.L, s);
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-32s %8.2f us per launch (back-to-back, PDL on)\n", it.name, ms * 1e3 / iters);
  };
  for (auto& it : items) time_it(it);
  for (auto& it : extra) time_it(it);
  // whole-layer chain the way engine.cu issues it (without the attention kernels)
  {
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) {
      dl_launch(mid, s);
      dl_launch(tail, s);
    }
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-32s %8.2f us per layer\n", "mid + tail", ms * 1e3 / iters);
  }
  // ---- trace of one mid and one tail launch
  for (auto* L : {&mid, &tail}) {
    CK(cudaMemset(trace, 0, (size_t)grid * kDLMaxPhases * 8 * 8));
    L->p.trace = trace;
    dl_launch(*L, s);
    CK(cudaStreamSynchronize(s));
    L->p.trace = nullptr;
    std::vector<unsigned long long> h((size_t)grid * kDLMaxPhases * 8);
    CK(cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost));
    printf("trace of %s (SM clocks, median over CTAs with work; phase-relative to stamp 0):\n", L == &mid ? "mid" : "tail");
    for (int p = 0; p < L->p.n_phases; ++p) {
      const char* names[8] = {"start", "first stage landed", "last MMA committed", "acc seen by epilogue", "stores done",
                              "grid arrive", "next phase released", "LN stats gathered"};
      printf("  phase %d (N=%d K=%d):", p, L->p.ph[p].N, L->p.ph[p].K);
      for (int k = 1; k < 8; ++k) {
        std::vector<long long> v;
        for (int c = 0; c < grid; ++c) {
          const unsigned long long t0 = h[((size_t)c * kDLMaxPhases + p) * 8 + 0], t = h[((size_t)c * kDLMaxPhases + p) * 8 + k];
          if (t0 && t) v.push_back((long long)t - (long long)t0);
        }
        if (v.empty()) continue;
        std::sort(v.begin(), v.end());
        printf("  %s %lld (max %lld)", names[k], v[v.size() / 2], v.back());
      }
      printf("\n");
      if (p + 1 < L->p.n_phases) {   // producer start of the next phase relative to this phase's start
        std::vector<long long> v;
        for (int c = 0; c < grid; ++c) {
          const unsigned long long a = h[((size_t)c * kDLMaxPhases + p) * 8 + 0], b = h[((size_t)c * kDLMaxPhases + p + 1) * 8 + 0];
          if (a && b) v.push_back((long long)b - (long long)a);
        }
        std::sort(v.begin(), v.end());
        if (!v.empty()) printf("    -> next phase starts %lld clk after this one (median), %lld (max)\n", v[v.size() / 2], v.back());
      }
    }
  }
  printf("done\n");
  return 0;
}
