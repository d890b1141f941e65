// How fast can one SM / the whole chip pull GEMM operand tiles with TENSOR TMA (cp.async.bulk.tensor.2d, 128B swizzle,
// 64 x 16-bit inner box) out of L2, as a function of box height, requests in flight and issuing threads?
// (tools/microbench.cu showed that 1-D bulk copies cost ~545 clk per request regardless of size; the fused decoder-layer
// kernel is sized from the numbers printed here.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench_tma tools/microbench_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static CUtensorMap make_map(const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
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
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// Each CTA streams a slab of `rows_per_cta` rows x K columns (bf16) in k-blocks of 64: per stage REQ requests, each a box of
// BOX_ROWS x 64 (consecutive row groups), through a STAGES-deep ring.  NPROD threads (one per warp) own interleaved stages.
// share_div > 1: CTA c reads the slab of CTA c / share_div (several CTAs pull the same bytes).
template <int BOX_ROWS, int REQ, int STAGES, int NPROD>
__global__ void __launch_bounds__(128, 1) tma_pull_kernel(const __grid_constant__ CUtensorMap map, int K, int slabs, int passes,
                                                          int share_div, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t full[STAGES];
  constexpr int kReqBytes = BOX_ROWS * 128;
  constexpr int kStageBytes = kReqBytes * REQ;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0 || warp >= NPROD) return;
  const int kblocks = K / 64;
  const int total = passes * kblocks;                 // stages this CTA streams
  const int slab = (blockIdx.x / share_div) % slabs;
  const int row0 = slab * BOX_ROWS * REQ;
  unsigned long long acc = 0;
  auto issue = [&](int i, int st) {
    const int kb = i % kblocks;
    mbar_expect_tx(&full[st], kStageBytes);
#pragma unroll
    for (int r = 0; r < REQ; ++r) tma_load_2d(smem + st * kStageBytes + r * kReqBytes, &map, &full[st], kb * 64, row0 + r * BOX_ROWS);
  };
  // this thread owns stages st = warp, warp + NPROD, ...
  for (int st = warp; st < STAGES; st += NPROD)
    if (st < total) issue(st, st);
  uint32_t phase = 0;
  for (int base = 0; base < total; base += STAGES) {
    for (int st = warp; st < STAGES; st += NPROD) {
      const int i = base + st;
      if (i >= total) break;
      mbar_wait(&full[st], phase);
      acc += *reinterpret_cast<volatile unsigned long long*>(smem + st * kStageBytes);
      if (i + STAGES < total) issue(i + STAGES, st);
    }
    phase ^= 1;
  }
  if (acc == 0x1234567) *sink = acc;
}

template <int BOX_ROWS, int REQ, int STAGES, int NPROD>
static void run(const char* label, const void* buf, int K, int grid, int share_div, double total_gb, unsigned long long* sink) {
  auto kern = tma_pull_kernel<BOX_ROWS, REQ, STAGES, NPROD>;
  const int smem = BOX_ROWS * 128 * REQ * STAGES + 1024;
  if (smem > 227 * 1024) {
    printf("skip %s (smem %d)\n", label, smem);
    return;
  }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int slab_rows = BOX_ROWS * REQ;
  const int slabs = (grid + share_div - 1) / share_div;
  const uint64_t rows = static_cast<uint64_t>(slabs) * slab_rows;
  CUtensorMap map = make_map(buf, K, rows, BOX_ROWS);
  const double per_pass = static_cast<double>(slab_rows) * K * 2;
  int passes = static_cast<int>(total_gb * 1e9 / grid / per_pass);
  if (passes < 1) passes = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kern<<<grid, 128, smem>>>(map, K, slabs, passes, share_div, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  kern<<<grid, 128, smem>>>(map, K, slabs, passes, share_div, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double bytes = per_pass * passes * grid;
  printf("tma %-28s box %3dx64 x%d req/stage, %2d stages (%3d KB in flight), %d issuer(s), grid %3d, footprint %6.1f MB: %8.1f GB/s = %5.1f B/clk/SM\n",
         label, BOX_ROWS, REQ, STAGES, BOX_ROWS * 128 * REQ * STAGES / 1024, NPROD, grid, rows * K * 2 / 1e6,
         bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / grid / 1.9e9);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  const size_t big = 1ull << 30;
  void* buf;
  CK(cudaMalloc(&buf, big));
  CK(cudaMemset(buf, 1, big));
  const int K = 1280;
  // --- box height, one request per stage, ~64-128 KB in flight
  run<16, 1, 32, 1>("L2", buf, K, sms, 1, 2.0, sink);
  run<64, 1, 12, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<128, 1, 8, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<256, 1, 4, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<256, 1, 6, 1>("L2", buf, K, sms, 1, 4.0, sink);
  // --- several requests per stage (what a GEMM stage with A + B sub-tiles issues)
  run<64, 2, 8, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<64, 4, 6, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<64, 8, 3, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<128, 2, 6, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<128, 4, 3, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<32, 8, 6, 1>("L2", buf, K, sms, 1, 4.0, sink);
  // --- depth of the ring
  run<64, 1, 2, 1>("L2", buf, K, sms, 1, 2.0, sink);
  run<64, 1, 4, 1>("L2", buf, K, sms, 1, 2.0, sink);
  run<64, 1, 24, 1>("L2", buf, K, sms, 1, 4.0, sink);
  run<128, 1, 2, 1>("L2", buf, K, sms, 1, 2.0, sink);
  run<128, 1, 12, 1>("L2", buf, K, sms, 1, 4.0, sink);
  // --- two / four issuing threads
  run<64, 1, 12, 2>("L2 2 issuers", buf, K, sms, 1, 4.0, sink);
  run<64, 1, 12, 4>("L2 4 issuers", buf, K, sms, 1, 4.0, sink);
  run<128, 1, 8, 2>("L2 2 issuers", buf, K, sms, 1, 4.0, sink);
  run<64, 4, 6, 2>("L2 2 issuers", buf, K, sms, 1, 4.0, sink);
  // --- shared slabs (30 CTAs read the same activation rows), fewer CTAs, a single CTA
  run<64, 1, 12, 1>("L2 slab shared by 30", buf, K, sms, 30, 4.0, sink);
  run<128, 1, 8, 1>("L2 slab shared by 30", buf, K, sms, 30, 4.0, sink);
  run<128, 1, 8, 1>("L2 74 CTAs", buf, K, 74, 1, 2.0, sink);
  run<128, 1, 8, 1>("L2 1 CTA", buf, K, 1, 1, 0.05, sink);
  run<64, 4, 6, 1>("L2 1 CTA", buf, K, 1, 1, 0.05, sink);
  run<256, 1, 4, 1>("L2 1 CTA", buf, K, 1, 1, 0.05, sink);
  // --- long rows (K = 5120) and an HBM-sized footprint (K = 1280 x many rows does not fit: use big K)
  run<128, 1, 8, 1>("K=5120 (194 MB, HBM)", buf, 5120, sms, 1, 4.0, sink);
  run<64, 4, 6, 1>("K=5120 (194 MB, HBM)", buf, 5120, sms, 1, 4.0, sink);
  printf("done\n");
  return 0;
}
