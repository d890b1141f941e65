// tcgen05 GEMM for sm_100a: every Linear / Conv1d on the Whisper hot path.
//
//   C[M, N] = epilogue( A[M, K] * W[N, K]^T )      A, W 16-bit K-major; fp32 accumulate in TMEM
//
// Replaces the cuBLAS / cuDNN call sites of the reference: Linear (model.py:44-50), Conv1d
// (model.py:53-59, used at :193-194), the MLP (model.py:155-157) and the tied-embedding logits
// product (model.py:245-247).  Conv1d(k=3) is run as three accumulated GEMM "taps" whose A tiles
// are the same activation matrix shifted by one row; the zero padding comes from TMA's
// out-of-bounds fill, the stride-2 of conv2 from a tensor map with a doubled row stride.
//
// Structure (persistent, one CTA per SM, 384 threads = 4 control warps + 8 epilogue warps):
//   warp 0    : TMA producer (one elected thread) - A tile BMx64, W tile BNx64 (x KS sub-blocks), 128B swizzle
//   warp 1    : tcgen05.mma issuer (one thread)   - UMMA BM x BN x 16, accumulators in TMEM
//   warp 2    : TMEM allocator / deallocator
//   warps 4-11: epilogue - tcgen05.ld -> bias / GELU / positional add / residual -> global
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA; two TMEM accumulator
// stages (tmem_full/tmem_empty) between MMA and epilogue so tile i+1's main loop overlaps
// tile i's epilogue.
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tmap.cuh"

namespace wb {

int g_pdl_on = -1;      // -1: read WB200_PDL on first use (default on); wb200_set_pdl() overrides
int g_kv_head_major = -1;   // -1: read WB200_KV_HEAD_MAJOR on first decoder_create (default on)
int g_sattn_tma = -1;       // -1: read WB200_SATTN_TMA on first decoder_create (default off)
int g_xattn_tma = -1;       // -1: read WB200_XATTN_TMA on first decoder_create (default on)
int g_bm64_on = 1;      // wb200_set_option("bm64", 0/1): 64-row tiles for skinny problems
int g_splitk_on = -1;   // -1: read WB200_SPLITK on first use; wb200_set_splitk() overrides

constexpr int kBM = 128;   // default tile height; BM = 64 (UMMA M = 64) is used for skinny problems
constexpr int kBK = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int kGemmThreads = 384;  // 4 control warps + 8 epilogue warps

struct GemmParams {
  int batch, rows_per_batch, m_tiles_per_batch;
  int N, n_tiles, k_blocks_per_tap, taps, K_tap;
  int a_row_off[3];
  int a_map_sel[3];
  void* C;
  long long ldc;
  const void* bias;
  const void* residual;
  long long ldr;
  const float* pos;
  int gelu;
  const int* skip_flag;
  // split-K (skinny decode-step GEMMs): work item = (tile, split); partial fp32 tiles go to
  // `partial[split][row][n]`, the last CTA to finish a tile sums them in split order and runs the epilogue
  int splits, k_per_split;
  float* partial;
  long long partial_stride;   // floats per split slab
  int* tile_counters;
  int hm_T;                   // > 0: head-major 16-bit output [rows / hm_T][N / 64][hm_T][64] (LinearArgs::head_major_T)
};

// KS = 64-wide K sub-blocks per pipeline stage.  The single MMA-issuing thread pays ~350 cycles of
// mbarrier wait / fence / commit per stage; with 64x64 tiles a 64-wide stage holds only 128 cycles of
// tensor work, so the skinny variant moves 256 of K per stage (KS = 4) to amortise it.
template <int BN, int BM = kBM, int KS = 1>
struct GemmCfg {
  static constexpr int kASub = BM * kBK * 2;
  static constexpr int kBSub = BN * kBK * 2;
  static constexpr int kABytes = kASub * KS;
  static constexpr int kStages = KS > 1 ? 3 : (BN == 256 ? 4 : (BN == 128 ? 6 : (BM == 64 ? 12 : 8)));
  static constexpr int kBBytes = kBSub * KS;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // BN in {64,128,256} -> pow2
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// bias / GELU / positional add / residual, then the store of 32 consecutive columns of one row
template <typename T, bool OUT_F32>
__device__ __forceinline__ void epilogue_store(float (&v)[32], const GemmParams& p, int t, long long grow, int nb) {
  const T* bias = reinterpret_cast<const T*>(p.bias);
  const T* resid = reinterpret_cast<const T*>(p.residual);
  const bool full = nb + 32 <= p.N;
  if (bias) {
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(bias + nb) + q);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = Cvt<T>::unpack2(w[e]);
          v[q * 8 + e * 2] += f.x;
          v[q * 8 + e * 2 + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb + j < p.N) v[j] += Cvt<T>::to_f(bias[nb + j]);
    }
  }
  if (p.gelu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(round_to<T>(v[j]));
  }
  if (p.pos) {
    const float* pr = p.pos + static_cast<long long>(t) * p.N + nb;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || nb + j < p.N) v[j] = round_to<T>(v[j]) + pr[j];
  }
  if (resid) {
    const T* rr = resid + grow * p.ldr + nb;
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u = *reinterpret_cast<const uint4*>(rr + q * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = Cvt<T>::unpack2(w[e]);
          v[q * 8 + e * 2] = round_to<T>(v[q * 8 + e * 2]) + f.x;
          v[q * 8 + e * 2 + 1] = round_to<T>(v[q * 8 + e * 2 + 1]) + f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb + j < p.N) v[j] = round_to<T>(v[j]) + Cvt<T>::to_f(rr[j]);
    }
  }
  if constexpr (OUT_F32) {
    float* out = reinterpret_cast<float*>(p.C) + grow * p.ldc + nb;
    if (full) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(out + q * 4) =
            make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb + j < p.N) out[j] = v[j];
    }
  } else {
    T* out = reinterpret_cast<T*>(p.C) + grow * p.ldc + nb;
    if (p.hm_T > 0) {          // [batch][head][t][64]: the 32 columns handled here lie inside one head
      const long long bt = grow / p.hm_T, tt = grow % p.hm_T;
      out = reinterpret_cast<T*>(p.C) + ((bt * (p.N >> 6) + (nb >> 6)) * p.hm_T + tt) * 64 + (nb & 63);
    }
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = Cvt<T>::pack2(v[q * 8 + 0], v[q * 8 + 1]);
        u.y = Cvt<T>::pack2(v[q * 8 + 2], v[q * 8 + 3]);
        u.z = Cvt<T>::pack2(v[q * 8 + 4], v[q * 8 + 5]);
        u.w = Cvt<T>::pack2(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(out + q * 8) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb + j < p.N) out[j] = Cvt<T>::from_f(v[j]);
    }
  }
}

template <typename T, int BN, bool OUT_F32, int BM = kBM, int KS = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const GemmParams p, const __grid_constant__ CUtensorMap mapA0,
                    const __grid_constant__ CUtensorMap mapA1,
                    const __grid_constant__ CUtensorMap mapB) {
  using Cfg = GemmCfg<BN, BM, KS>;
  pdl_launch_dependents();   // the next kernel may be scheduled as soon as every CTA of this one is resident
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.batch * p.m_tiles_per_batch * p.n_tiles;
  const int groups_per_tap = (p.k_blocks_per_tap + KS - 1) / KS;
  const int k_blocks = p.taps * groups_per_tap;          // pipeline stages per full K sweep
  __shared__ int s_ticket;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) touches no global data and
  // overlaps the tail of the previous kernel; from here on its results are needed.
  pdl_wait();
  const bool skip = p.skip_flag && *p.skip_flag;
  const int total_items = skip ? 0 : total_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / p.splits, split = item % p.splits;
      const int m_tile = tile / p.n_tiles;
      const int n0 = (tile % p.n_tiles) * BN;
      const int b = m_tile / p.m_tiles_per_batch;
      const int t0 = (m_tile % p.m_tiles_per_batch) * BM;
      const int kb_lo = split * p.k_per_split, kb_hi = min(k_blocks, kb_lo + p.k_per_split);
      for (int tap = 0; tap < p.taps; ++tap) {
        const CUtensorMap* ma = p.a_map_sel[tap] ? &mapA1 : &mapA0;
        const int row0 = t0 + p.a_row_off[tap];
        for (int kg = 0; kg < groups_per_tap; ++kg) {
          const int kidx = tap * groups_per_tap + kg;
          if (kidx < kb_lo || kidx >= kb_hi) continue;
          uint8_t* sa = tiles + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
          for (int sub = 0; sub < KS; ++sub) {     // sub-blocks past K are zero-filled by TMA
            const int kb = kg * KS + sub;
            tma_load_3d(sa + sub * Cfg::kASub, ma, &full_bar[stage], kb * kBK, row0, b);
            tma_load_2d(sb + sub * Cfg::kBSub, &mapB, &full_bar[stage], tap * p.K_tap + kb * kBK, n0);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc(Cvt<T>::kUmmaFmt, BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int split = item % p.splits;
      const int kb_lo = split * p.k_per_split, kb_hi = min(k_blocks, kb_lo + p.k_per_split);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int sub = 0; sub < KS; ++sub) {
          const uint64_t adesc = umma_desc_sw128(sa + sub * Cfg::kASub, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sb + sub * Cfg::kBSub, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in 16 B units
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != kb_lo) || (sub != 0) || (k != 0));
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(&tmem_full[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // Eight epilogue warps: two per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31),
    // each owning half of the tile's columns.  With K = 1280 the main loop of a 128x256 tile lasts
    // ~10k cycles; four warps needed longer than that for the bias/GELU/residual epilogue and stalled it.
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kColsPerWarp = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / p.splits, split = item % p.splits;
      const int m_tile = tile / p.n_tiles;
      const int n0 = (tile % p.n_tiles) * BN;
      const int b = m_tile / p.m_tiles_per_batch;
      // UMMA M = 128: accumulator row i sits in TMEM lane i.  M = 64: the 64 rows use lanes 0-15 of each
      // 32-lane quadrant (row = 16 * quadrant + lane), the upper 16 lanes of every quadrant are unused.
      const int t = (m_tile % p.m_tiles_per_batch) * BM + (BM == 128 ? quad * 32 + lane : quad * 16 + lane);
      const bool row_ok = t < p.rows_per_batch && (BM == 128 || lane < 16);
      const long long grow = static_cast<long long>(b) * p.rows_per_batch + t;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * kColsPerWarp;
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        const int nb = n0 + half * kColsPerWarp + c * 32;
        if (row_ok && nb < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.splits == 1) {
            epilogue_store<T, OUT_F32>(v, p, t, grow, nb);
          } else {
            // raw fp32 partial sums of this K-slice (only in-range columns exist in the slab)
            float* ps = p.partial + split * p.partial_stride + grow * p.N + nb;
            if (nb + 32 <= p.N && (p.N & 3) == 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                __stcg(reinterpret_cast<float4*>(ps) + q, make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) __stcg(ps + j, v[j]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      if (p.splits > 1) {
        // ticket: the last of the tile's `splits` CTAs reduces the slabs (in split order, so the sum does
        // not depend on arrival order) and applies the epilogue
        // one gpu-scope fence by the ticket thread AFTER the CTA barrier publishes all 256 threads' slab
        // stores (fences are cumulative); a fence per thread costs microseconds here (it also drops L1)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) {
          __threadfence();
          s_ticket = atomicAdd(&p.tile_counters[tile], 1);
          __threadfence();
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const bool last = (s_ticket == p.splits - 1);
        if (last) {
          if (threadIdx.x == 128) p.tile_counters[tile] = 0;
#pragma unroll 1
          for (int c = 0; c < kColsPerWarp / 32; ++c) {
            const int nb = n0 + half * kColsPerWarp + c * 32;
            if (row_ok && nb < p.N) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
              const bool vec = nb + 32 <= p.N && (p.N & 3) == 0;
              for (int sp = 0; sp < p.splits; ++sp) {
                const float* ps = p.partial + sp * p.partial_stride + grow * p.N + nb;
                if (vec) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 f = __ldcg(reinterpret_cast<const float4*>(ps) + q);
                    v[q * 4] += f.x;
                    v[q * 4 + 1] += f.y;
                    v[q * 4 + 2] += f.z;
                    v[q * 4 + 3] += f.w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (nb + j < p.N) v[j] += __ldcg(ps + j);
                }
              }
              epilogue_store<T, OUT_F32>(v, p, t, grow, nb);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // s_ticket is reused by the next item
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <typename T, int BN, bool OUT_F32, int BM = kBM, int KS = 1>
static int launch_impl(const GemmParams& p, const CUtensorMap& a0, const CUtensorMap& a1,
                       const CUtensorMap& b, cudaStream_t s) {
  using Cfg = GemmCfg<BN, BM, KS>;
  static SmemOptIn optin;
  auto kern = gemm_tcgen05_kernel<T, BN, OUT_F32, BM, KS>;
  if (!optin.ensure(kern, Cfg::kSmemBytes)) return 10;
  const int num_sms = sm_count();
  const int total = p.batch * p.m_tiles_per_batch * p.n_tiles * p.splits;
  const int grid = total < num_sms ? total : num_sms;
  ProfileScope prof(PROF_GEMM, s);
  const cudaError_t le = launch_pdl(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, s, p, a0, a1, b);
  count_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 11;
}

int launch_linear(const LinearArgs& a, cudaStream_t s) {
  if (a.rows_per_batch <= 0 || a.batch <= 0 || a.N <= 0 || a.K_tap <= 0) return 0;
  if (a.taps < 1 || a.taps > 3) return 3;
  // TMA / vector-store alignment requirements
  const long long esz = 2;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15) ||
      (a.lda * esz) % 16 || (a.ldw * esz) % 16 || (a.a_batch_stride * esz) % 16)
    return 4;
  const long long osz = a.out_f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(a.C) & 15) || (a.ldc * osz) % 16) return 5;
  if (a.residual && ((reinterpret_cast<uintptr_t>(a.residual) & 15) || (a.ldr * esz) % 16)) return 6;

  int bn = a.block_n;
  int bm = a.block_m ? a.block_m : kBM;
  const long long rows = static_cast<long long>(a.batch) * a.rows_per_batch;
  // wide tiles when M is large (encoder) or N is huge (logits: fewer passes over the A tile);
  // 64-wide tiles for the skinny decode-step GEMMs so that more CTAs pull weights concurrently
  if (bn == 0) {
    if (rows >= 4096) {
      bn = 256;
    } else {
      // Skinny (decode-step) problems are bound by the ~40 B/clk a single SM's TMA unit can pull
      // (profiles/r1_summary.md): pick the tile width that minimises rounds x bytes per k-block per CTA.
      long long best = -1;
      const bool allow_bm64 = a.block_m == 0 && a.batch == 1 && a.taps == 1 && g_bm64_on;
      for (int cbm = allow_bm64 ? 64 : kBM; cbm <= kBM; cbm *= 2) {
        const int m_tiles = static_cast<int>((rows + cbm - 1) / cbm);
        for (int cand = 64; cand <= (cbm == 64 ? 64 : 256); cand *= 2) {
          const long long tiles = static_cast<long long>(m_tiles) * ((a.N + cand - 1) / cand);
          const long long rounds = (tiles + 147) / 148;
          const long long cost = rounds * (cbm / 8 + cand / 8);      // KB per k-block per CTA
          if (best < 0 || cost < best) {
            best = cost;
            bn = cand;
            bm = cbm;
          }
        }
      }
    }
  }
  if (bn == 256 && a.N < 256) bn = a.N >= 128 ? 128 : 64;

  GemmParams p;
  p.batch = a.batch;
  p.rows_per_batch = a.rows_per_batch;
  if (bm == 64 && bn != 64) bm = kBM;        // only the 64 x 64 variant is instantiated
  p.m_tiles_per_batch = (a.rows_per_batch + bm - 1) / bm;
  p.N = a.N;
  p.n_tiles = (a.N + bn - 1) / bn;
  p.k_blocks_per_tap = (a.K_tap + kBK - 1) / kBK;
  p.taps = a.taps;
  p.K_tap = a.K_tap;
  p.C = a.C;
  p.ldc = a.ldc;
  p.bias = a.bias;
  p.residual = a.residual;
  p.ldr = a.ldr;
  p.pos = a.pos;
  p.gelu = a.gelu;
  p.skip_flag = a.skip_flag;
  // split-K only for skinny problems that would otherwise leave most SMs idle
  p.splits = 1;
  const int ks_host = (bm == 64) ? 4 : (bn == 128 ? 2 : 1);   // must match the kernel variant dispatched below
  p.k_per_split = p.taps * ((p.k_blocks_per_tap + ks_host - 1) / ks_host);
  p.partial = nullptr;
  p.partial_stride = rows * static_cast<long long>(a.N);
  p.tile_counters = a.splitk_counters;
  p.hm_T = 0;
  if (a.head_major_T > 0) {
    if (a.out_f32 || a.N % 64 || a.batch != 1 || a.residual || a.rows_per_batch % a.head_major_T) return 14;
    p.hm_T = a.head_major_T;
  }
  {
    const int tiles = p.batch * p.m_tiles_per_batch * p.n_tiles;
    const int kblocks = p.taps * p.k_blocks_per_tap;
    int& splitk_on = g_splitk_on;
    if (splitk_on < 0) {
      // Off by default: as measured on B200 (profiles/r1_summary.md) the slab write + ticket + reduce costs
      // more than the shorter K chain saves at M = 320; kept behind WB200_SPLITK=1 for the next round.
      const char* e = getenv("WB200_SPLITK");
      splitk_on = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    if (splitk_on && !a.head_major_T && ks_host == 1 && a.splitk_ws && a.splitk_counters && a.taps == 1 && tiles < 148 && kblocks >= 8 &&
        tiles <= a.splitk_max_tiles) {
      int sp = (2 * 148 + tiles - 1) / tiles;          // aim at ~2 work items per SM
      if (sp > kblocks / 4) sp = kblocks / 4;          // keep >= 4 k-blocks (256 of K) per item
      static int sp_cap = 0;
      if (!sp_cap) {
        const char* e = getenv("WB200_SPLITK_MAX");
        sp_cap = (e && atoi(e) > 0) ? atoi(e) : 8;
      }
      if (sp > sp_cap) sp = sp_cap;
      while (sp > 1 && static_cast<size_t>(sp) * p.partial_stride * 4 > a.splitk_ws_bytes) --sp;
      if (sp > 1) {
        p.splits = sp;
        p.k_per_split = (kblocks + sp - 1) / sp;
        p.splits = (kblocks + p.k_per_split - 1) / p.k_per_split;   // no empty slices
        p.partial = a.splitk_ws;
      }
    }
  }

  // A maps: distinct base offsets -> at most two tensor maps
  CUtensorMap mapA[2];
  long long base_of[2] = {a.a_base_off[0], a.a_base_off[0]};
  int n_maps = 1;
  for (int t = 0; t < 3; ++t) {
    p.a_row_off[t] = t < a.taps ? a.a_row_off[t] : 0;
    p.a_map_sel[t] = 0;
    if (t >= a.taps) continue;
    if (a.a_base_off[t] == base_of[0]) {
      p.a_map_sel[t] = 0;
    } else if (n_maps == 2 && a.a_base_off[t] == base_of[1]) {
      p.a_map_sel[t] = 1;
    } else if (n_maps == 1) {
      base_of[1] = a.a_base_off[t];
      n_maps = 2;
      p.a_map_sel[t] = 1;
    } else {
      return 7;
    }
  }
  for (int i = 0; i < 2; ++i) {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a.A) + base_of[i] * esz;
    if (reinterpret_cast<uintptr_t>(base) & 15) return 4;
    uint64_t dims[3] = {static_cast<uint64_t>(a.K_tap), static_cast<uint64_t>(a.a_rows_per_batch),
                        static_cast<uint64_t>(a.batch)};
    uint64_t strides[2] = {static_cast<uint64_t>(a.lda * esz),
                           static_cast<uint64_t>((a.batch > 1 ? a.a_batch_stride : a.lda * a.a_rows_per_batch) * esz)};
    uint32_t box[3] = {kBK, static_cast<uint32_t>(bm), 1};
    if (make_tmap_16bit(&mapA[i], a.dtype, base, 3, dims, strides, box)) return 8;
  }
  CUtensorMap mapB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(a.taps) * a.K_tap, static_cast<uint64_t>(a.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(a.ldw * esz)};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    if (make_tmap_16bit(&mapB, a.dtype, a.W, 2, dims, strides, box)) return 9;
  }

#define WB_DISPATCH(TT, BNN)                                                              \
  return a.out_f32 ? launch_impl<TT, BNN, true>(p, mapA[0], mapA[1], mapB, s)             \
                   : launch_impl<TT, BNN, false>(p, mapA[0], mapA[1], mapB, s)
  if (bm == 64) {
    if (a.dtype == DT_BF16)
      return a.out_f32 ? launch_impl<__nv_bfloat16, 64, true, 64, 4>(p, mapA[0], mapA[1], mapB, s)
                       : launch_impl<__nv_bfloat16, 64, false, 64, 4>(p, mapA[0], mapA[1], mapB, s);
    return a.out_f32 ? launch_impl<__half, 64, true, 64, 4>(p, mapA[0], mapA[1], mapB, s)
                     : launch_impl<__half, 64, false, 64, 4>(p, mapA[0], mapA[1], mapB, s);
  }
  if (a.dtype == DT_BF16) {
    if (bn == 256) { WB_DISPATCH(__nv_bfloat16, 256); }
    if (bn == 128)
      return a.out_f32 ? launch_impl<__nv_bfloat16, 128, true, 128, 2>(p, mapA[0], mapA[1], mapB, s)
                       : launch_impl<__nv_bfloat16, 128, false, 128, 2>(p, mapA[0], mapA[1], mapB, s);
    WB_DISPATCH(__nv_bfloat16, 64);
  } else {
    if (bn == 256) { WB_DISPATCH(__half, 256); }
    if (bn == 128)
      return a.out_f32 ? launch_impl<__half, 128, true, 128, 2>(p, mapA[0], mapA[1], mapB, s)
                       : launch_impl<__half, 128, false, 128, 2>(p, mapA[0], mapA[1], mapB, s);
    WB_DISPATCH(__half, 64);
  }
#undef WB_DISPATCH
}

}  // namespace wb
