// Hardware constants the fused decoder-layer kernel is designed around, measured on the box it will run on:
//   1. cost of a grid-wide barrier between the phases of a persistent kernel (one CTA per SM)
//   2. L2 -> SM and HBM -> SM pull rate of TMA bulk copies, chip-wide and for a single SM
//   3. the same when many CTAs pull the SAME bytes (the 320-row activation tile every CTA of a GEMM phase reads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
// Every spin is bounded (trap instead of hang).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);         \
      exit(1);                                                                                \
    }                                                                                         \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (long long i = 0; i < (1ll << 26); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// 1. grid barrier: monotonically increasing counter, one arrival per CTA, acquire-polling by thread 0
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) grid_barrier_kernel(unsigned int* counter, int iters, int mode) {
  for (int it = 0; it < iters; ++it) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int target = static_cast<unsigned int>(it + 1) * gridDim.x;
      if (mode == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        long long spins = 0;
        while (true) {
          unsigned int v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (v >= target) break;
          if (++spins > (1ll << 26)) __trap();
        }
      } else {
        // release-reduction + acquire poll (no separate fence)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        long long spins = 0;
        while (true) {
          unsigned int v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (v >= target) break;
          if (++spins > (1ll << 26)) __trap();
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2./3. TMA bulk pull: every CTA streams `chunks_per_cta` chunks of CHUNK bytes through a STAGES-deep smem ring
//   pattern 0: CTA c reads chunks c, c + grid, c + 2 grid, ... of the region (disjoint, whole region covered)
//   pattern 1: every CTA reads the SAME chunks 0, 1, 2, ... (hot rows shared by all CTAs)
//   pattern 2: groups of 30 CTAs share a stream (the activation tile of one 64-row block), groups differ
// ---------------------------------------------------------------------------------------------------------------
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(128, 1) pull_kernel(const uint8_t* base, long long region_chunks, int chunks_per_cta,
                                                      int pattern, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long acc = 0;
  if (threadIdx.x == 0) {
    auto chunk_of = [&](int i) -> long long {
      long long c;
      if (pattern == 0) c = static_cast<long long>(i) * gridDim.x + blockIdx.x;
      else if (pattern == 1) c = i;
      else c = static_cast<long long>(blockIdx.x / 30) * chunks_per_cta + i;
      return c % region_chunks;
    };
    for (int i = 0; i < STAGES && i < chunks_per_cta; ++i) {
      mbar_expect_tx(&full[i], CHUNK);
      bulk_load_1d(smem + i * CHUNK, base + chunk_of(i) * CHUNK, CHUNK, &full[i]);
    }
    uint32_t phase = 0;
    int st = 0;
    for (int i = 0; i < chunks_per_cta; ++i) {
      mbar_wait(&full[st], phase);
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + st * CHUNK);   // touch the data
      const int nx = i + STAGES;
      if (nx < chunks_per_cta) {
        mbar_expect_tx(&full[st], CHUNK);
        bulk_load_1d(smem + st * CHUNK, base + chunk_of(nx) * CHUNK, CHUNK, &full[st]);
      }
      if (++st == STAGES) {
        st = 0;
        phase ^= 1;
      }
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

// plain vectorised LDG pull for comparison (all threads, 16 B per thread per iteration)
__global__ void __launch_bounds__(512, 1) ldg_pull_kernel(const uint4* base, long long n_vec, int iters, unsigned long long* sink) {
  unsigned long long acc = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 8
    for (long long j = i; j < n_vec; j += stride) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + j));
      acc += v.x ^ v.w;
    }
  }
  if (acc == 0x1234567) *sink = acc;
}

template <int CHUNK, int STAGES>
static void run_pull(const char* label, const uint8_t* buf, size_t region_bytes, int grid, int pattern, double total_gb,
                     unsigned long long* sink) {
  auto kern = pull_kernel<CHUNK, STAGES>;
  const int smem = CHUNK * STAGES;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long region_chunks = region_bytes / CHUNK;
  const int chunks_per_cta = static_cast<int>(total_gb * 1e9 / CHUNK / grid);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) kern<<<grid, 128, smem>>>(buf, region_chunks, chunks_per_cta, pattern, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  const int reps = 3;
  for (int r = 0; r < reps; ++r) kern<<<grid, 128, smem>>>(buf, region_chunks, chunks_per_cta, pattern, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double bytes = static_cast<double>(chunks_per_cta) * CHUNK * grid * reps;
  printf("pull %-34s region %7.1f MB grid %3d chunk %3d KB x%d stages pattern %d: %8.1f GB/s  (%.1f B/clk/SM at 1.9 GHz)\n",
         label, region_bytes / 1e6, grid, CHUNK / 1024, STAGES, pattern, bytes / (ms * 1e-3) / 1e9,
         bytes / (ms * 1e-3) / grid / 1.9e9);
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %.1f MB\n", prop.name, sms, prop.l2CacheSize / 1e6);
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));

  // ---- 1. grid barrier
  {
    unsigned int* counter;
    CK(cudaMalloc(&counter, 4));
    for (int mode = 0; mode < 2; ++mode) {
      for (int grid : {sms, sms / 2}) {
        CK(cudaMemset(counter, 0, 4));
        int iters = 200;
        void* args[] = {&counter, &iters, &mode};
        CK(cudaLaunchCooperativeKernel((void*)grid_barrier_kernel, dim3(grid), dim3(256), args, 0, 0));   // warm
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(counter, 0, 4));
        iters = 2000;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((void*)grid_barrier_kernel, dim3(grid), dim3(256), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("grid barrier mode %d (%s), %3d CTAs x 256 thr: %.3f us per barrier\n", mode,
               mode == 0 ? "fence+atomicAdd+ld.acquire" : "red.release+ld.acquire", grid, ms * 1e3 / iters);
      }
    }
  }

  // ---- 2. TMA bulk pulls
  const size_t big = 4ull << 30;
  uint8_t* buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  // HBM-resident region (4 GB, each byte read once per launch)
  run_pull<16384, 8>("HBM disjoint", buf, big, sms, 0, 3.9, sink);
  run_pull<32768, 6>("HBM disjoint", buf, big, sms, 0, 3.9, sink);
  run_pull<8192, 16>("HBM disjoint", buf, big, sms, 0, 3.9, sink);
  run_pull<16384, 12>("HBM disjoint", buf, big, sms, 0, 3.9, sink);
  // L2-resident regions
  for (size_t mb : {8, 24, 48, 96}) {
    run_pull<16384, 8>("L2 disjoint", buf, mb << 20, sms, 0, 4.0, sink);
  }
  run_pull<32768, 6>("L2 disjoint", buf, 24u << 20, sms, 0, 4.0, sink);
  run_pull<8192, 16>("L2 disjoint", buf, 24u << 20, sms, 0, 4.0, sink);
  run_pull<16384, 12>("L2 disjoint", buf, 24u << 20, sms, 0, 4.0, sink);
  // everyone reads the same bytes / groups of 30 share a stream
  run_pull<16384, 8>("L2 same bytes (all CTAs)", buf, 1u << 20, sms, 1, 4.0, sink);
  run_pull<16384, 8>("L2 shared by groups of 30", buf, 8u << 20, sms, 2, 4.0, sink);
  // fewer SMs / a single SM
  run_pull<16384, 8>("L2 disjoint, 100 CTAs", buf, 24u << 20, 100, 0, 3.0, sink);
  run_pull<16384, 8>("L2 disjoint, 74 CTAs", buf, 24u << 20, 74, 0, 2.0, sink);
  run_pull<16384, 8>("L2 disjoint, 16 CTAs", buf, 24u << 20, 16, 0, 0.5, sink);
  run_pull<16384, 8>("L2 disjoint, 1 CTA", buf, 24u << 20, 1, 0, 0.05, sink);
  run_pull<16384, 12>("L2 disjoint, 1 CTA", buf, 24u << 20, 1, 0, 0.05, sink);
  run_pull<16384, 8>("HBM disjoint, 1 CTA", buf, big, 1, 0, 0.05, sink);

  // ---- plain LDG for comparison
  for (size_t mb : {24, 4096}) {
    const long long n_vec = (static_cast<long long>(mb) << 20) / 16;
    const int iters = mb == 24 ? 100 : 1;
    ldg_pull_kernel<<<sms * 2, 512>>>(reinterpret_cast<const uint4*>(buf), n_vec, iters, sink);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    ldg_pull_kernel<<<sms * 2, 512>>>(reinterpret_cast<const uint4*>(buf), n_vec, iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("ldg pull region %6zu MB: %8.1f GB/s\n", mb, static_cast<double>(n_vec) * 16 * iters / (ms * 1e-3) / 1e9);
  }
  printf("done\n");
  return 0;
}
