// Fused decoder-layer GEMM chain for the autoregressive step (reference whisper/model.py:142-171,
// ResidualAttentionBlock.forward, as driven once per generated token by decoding.py:680-710).
//
// One step of one layer is 3 LayerNorms and 6 Linears over R = n_audio x beams rows (320 at the headline
// configuration): ~46 MB of weights against 0.8 MB of activations.  Launched one by one (round 1) every Linear
// paid a launch, a prologue (barrier init, TMEM allocation, descriptor fetch), a cold pipeline and a tail: ~10-18 us
// each for 0.5-2 us worth of HBM traffic, plus ~6 us per LayerNorm.  Here a run of consecutive Linears is ONE
// persistent kernel (one CTA per SM) that walks through them as PHASES separated by a grid-wide barrier
// (1.3 us measured on B200, tools/microbench.cu):
//
//   * LayerNorm never runs as a pass of its own.  For y = LN(x) W^T + b the kernel multiplies the RAW residual rows by
//     W' = W (.) gamma (folded once at load time) and finishes in the epilogue:
//         y[r, n] = rstd_r * (acc[r, n] - mean_r * c1[n]) + c2[n],   c1 = sum_k W'[n, k],  c2 = W beta + b
//     The row statistics come from the PRODUCER of x: the epilogue that writes the residual stream also writes, per
//     row and per column slice, a (count, mean, M2) partial, and the consumer merges the ~30 partials of a row with
//     Chan's formula - no re-read of the activations, no extra launch, fp32 throughout like nn.LayerNorm
//     (model.py:39-41).
//   * Work split of a phase: the R rows are cut into 64-row blocks (UMMA M = 64); the CTAs assigned to a row block
//     split the N output features evenly in units of 16 columns, so every CTA streams its own weight slab exactly
//     once plus one 64-row activation block: bytes per CTA = 2K (64 + N / ctas_per_block), the minimum over tile
//     shapes for 148 CTAs, and all SMs pull weights concurrently.
//   * Roles per CTA (416 threads): warps 0, 2, 3 = TMA producers, ONE PER RING STAGE (activation block + weight slab
//     through a 3 x 64 KB smem ring that lives across phases), warp 1 = tcgen05.mma issuer (accumulator 64 x <=192 fp32
//     in TMEM, two buffers), warp 2 also allocates TMEM, warps 4-11 = epilogue (tcgen05.ld -> LN fold / bias / erf-GELU /
//     residual add / LN partials -> global), warp 12 = grid-barrier poller.
//     Why a producer per stage: measured on B200 (tools/microbench_tma.cu, profiles/r2_microbench_tma.txt) one thread
//     that waits on an mbarrier, re-arms it and issues a TMA tile sustains one such round every ~500 clk whatever the
//     box size (8 KB boxes: 16 B/clk/SM, 32 KB boxes: 62 B/clk/SM) while every further back-to-back TMA costs ~60 clk,
//     and N issuing threads scale N-fold (4 threads x 8 KB: 62 B/clk/SM; L2 delivers 17-20 TB/s chip-wide).  So each
//     ring stage has its own issuing thread, weight slabs travel as ONE box of up to 192 rows, and a stage carries as many
//     64-wide K sub-blocks as fit in 64 KB (4 for the 1280-wide projections, 2 for QKV / fc1).
//
// Phases of a decoder layer (engine.cu strings them together; the two attention kernels stay separate launches):
//     [QKV] | self-attention | [out-proj + residual, cross-query] | cross-attention |
//     [cross-out + residual, fc1 + GELU, fc2 + residual, QKV of the NEXT layer]
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "dec_layer.h"
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.cuh"

namespace wb {

int g_fused_layer = -1;   // -1: read WB200_FUSED_LAYER on first use; wb200_set_fused_decoder_layer() overrides
int g_fused_rows = -1;    // -1: read WB200_FUSED_ROWS on first use; wb200_set_fused_decoder_rows() overrides
int g_fused_stack = -1;   // -1: read WB200_FUSED_STACK on first use; wb200_set_fused_decoder_stack() overrides

constexpr int kDLThreads = 416;                      // 13 warps
constexpr int kDLRowBytes = 128;                     // one row of a 64-wide 16-bit k-block
constexpr int kDLUnitBytes = kDLUnit * 128;          // 16 weight rows x 64 x 16-bit
constexpr int kDLSlotBytes = 64 * 1024;              // one ring stage
constexpr int kDLMaxKS = 4;
constexpr int kDLTmemCols = 512;                     // two accumulator buffers of 256 columns
static_assert(2 * (64 * kDLRowBytes + kDLMaxUnits * kDLUnitBytes) <= kDLSlotBytes, "a stage must hold two K sub-blocks of the widest tile");

// 64-wide K sub-blocks per ring stage for a row block of bm rows and a weight box of `box` units (per phase)
__host__ __device__ __forceinline__ int dl_ks(int bm, int box) {
  const int k = kDLSlotBytes / (bm * kDLRowBytes + box * kDLUnitBytes);
  return k > kDLMaxKS ? kDLMaxKS : (k < 1 ? 1 : k);
}

// Measured and rejected (profiles/r2_dec_layer_trace_v4_chains_rejected.txt): spreading the K steps of a narrow tile
// round-robin over up to four accumulators, on the theory that back-to-back tcgen05.mma into ONE accumulator are a
// latency-bound dependent chain.  They are not: alternating accumulators doubled the main-loop time (7400 -> 14900 clk
// for the K = 1280 phases).  What the traces do show is the cost of UMMA M = 64 with both operands in shared memory:
// ~40 + 1.1 x N clk per 16-deep MMA (N = 48: 92 clk, N = 144: 180, N = 176: 207) against the 0.5 x N of the
// M = 128 form - the tensor pipe is ~2.3x slower per column at M = 64, which bounds the wide phases (QKV, fc1).
template <int STAGES>
struct DLCfg {
  static constexpr int kStageBytes = kDLSlotBytes;
  static constexpr int kTileBytes = STAGES * kStageBytes;
  // tail: barriers (full, empty, tmem_full[2], tmem_empty[2], ready[kDLMaxPhases]) + tmem pointer + LN scratch +
  // the epilogue's per-column vectors (c1 | c2 or bias) of this CTA's columns
  static constexpr int kTailBytes = 8 * (2 * STAGES + 4 + kDLMaxPhases) + 16 + 2 * 64 * 16 + 2 * kDLMaxUnits * kDLUnit * 4 + 64;
  static constexpr int kSmemBytes = kTileBytes + kTailBytes + 1024;
};

// bounded waits: a protocol bug must end in a trap (an error the host sees), never in a hung GPU
__device__ __forceinline__ void dl_mbar_wait(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// (count, mean, M2) merge of two disjoint samples (Chan et al.)
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
  if (nb == 0.f) return;
  const float nt = n + nb;
  const float delta = meanb - mean;
  const float f = nb / nt;
  mean = fmaf(delta, f, mean);
  m2 = m2 + m2b + delta * delta * n * f;
  n = nt;
}

template <typename T, int STAGES>
__global__ void __launch_bounds__(kDLThreads, 1)
dec_layer_kernel(const DLParams P, const __grid_constant__ DLMaps M) {
  using Cfg = DLCfg<STAGES>;
  static_assert(STAGES == 3, "one producer warp per stage: warps 0, 2, 3");
  pdl_launch_dependents();
  extern __shared__ uint8_t dl_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dl_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kTileBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* ready_bar = tmem_empty + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ready_bar + kDLMaxPhases);
  float4* s_stats = reinterpret_cast<float4*>(tmem_ptr_smem + 4);      // [2][64] (count, mean, M2, -)
  float* s_vec = reinterpret_cast<float*>(s_stats + 128);              // [2][kDLMaxUnits * 16]: c1 | c2, or bias

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grid = gridDim.x;
  const int cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.n_phases; ++p) {
      tma_prefetch_desc(&M.a[p]);
      tma_prefetch_desc(&M.b[p]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    for (int i = 0; i < kDLMaxPhases; ++i) mbar_init(&ready_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kDLTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();       // everything above overlaps the tail of the previous kernel
  const bool skip = P.skip_flag && *P.skip_flag;
  const int n_phases = skip ? 0 : P.n_phases;
  auto stamp = [&](int p, int slot_id) {
    if (P.trace) P.trace[(static_cast<long long>(cta) * kDLMaxPhases + p) * 8 + slot_id] = clock64();
  };

  // this CTA's share of a phase: row block m_blk (bm rows), columns [u0, u0 + nu) in units of 16; the CTAs of a row
  // block are numbered by `slot`
  struct Share {
    int m_blk, slot, n_slots, u0, nu;
  };
  auto share = [&](int p) {
    Share sh;
    const int m_tiles = (P.R + P.ph[p].bm - 1) / P.ph[p].bm;
    sh.m_blk = cta % m_tiles;
    sh.slot = cta / m_tiles;
    sh.n_slots = (grid - sh.m_blk + m_tiles - 1) / m_tiles;
    const int total = P.ph[p].N / kDLUnit;
    sh.u0 = static_cast<int>(static_cast<long long>(sh.slot) * total / sh.n_slots);
    sh.nu = static_cast<int>(static_cast<long long>(sh.slot + 1) * total / sh.n_slots) - sh.u0;
    return sh;
  };

  const int prod = warp == 0 ? 0 : (warp == 2 ? 1 : (warp == 3 ? 2 : -1));
  if (prod >= 0 && lane == 0) {
    // ===================== TMA producers: producer j owns ring stage j =====================
    int q = 0;                                 // running stage number across phases (the ring never drains)
    for (int p = 0; p < n_phases; ++p) {
      const Share sh = share(p);
      const int u0 = sh.u0, nu = sh.nu, m_blk = sh.m_blk;
      if (nu == 0) continue;
      const DLPhase& ph = P.ph[p];
      const int asub = ph.bm * kDLRowBytes;
      // ONE weight box per K sub-block, as tall as the widest share of the phase (the TMA unit costs ~120 clk per request
      // whatever its size: a second 2 KB box for the odd unit cost as much as 8 KB of payload); rows past this CTA's own
      // share are loaded and ignored, rows past N are zero-filled
      const int box = ph.units_box;
      const int ks = dl_ks(ph.bm, box);
      const int kblocks = (ph.K + 63) / 64;
      const int groups = (kblocks + ks - 1) / ks;
      const int bsub = box * kDLUnitBytes;
      const uint32_t bytes = static_cast<uint32_t>(ks) * (asub + bsub);
      bool released = (p == 0);
      for (int g = 0; g < groups; ++g, ++q) {
        if (q % STAGES != prod) continue;
        dl_mbar_wait(&empty_bar[prod], ((q / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full_bar[prod], bytes);
        uint8_t* sa = tiles + prod * Cfg::kStageBytes;
        uint8_t* sb = sa + ks * asub;
        // ONE request per operand per stage: the tensor maps view [rows, K] as {64 k, rows, K / 64} so a box of `ks` k-blocks
        // lands as `ks` consecutive 128B-swizzled [rows x 64] sub-tiles - the TMA unit serves one request per ~130-190 clk
        // whatever its size (8 KB boxes: 37 B/clk/SM measured in this kernel; 32 KB boxes: 62 B/clk/SM)
        // weights first: they are constants, so the first stage of a phase streams them in while the grid barrier of the
        // previous phase is still closing; the activation rows follow once that phase is complete grid-wide
        if (ph.kouter) {
          tma_load_3d(sb, &M.b[p], &full_bar[prod], 0, u0 * kDLUnit, g * ks);
        } else {
          for (int sub = 0; sub < ks; ++sub)       // sub-blocks past K are zero-filled by TMA (full byte count)
            tma_load_2d(sb + sub * bsub, &M.b[p], &full_bar[prod], (g * ks + sub) * 64, u0 * kDLUnit);
        }
        if (!released) {
          dl_mbar_wait(&ready_bar[p], 0);
          fence_proxy_async_global();              // generic-proxy writes of other CTAs -> this thread's async-proxy (TMA) reads
          released = true;
        }
        if (g == 0) stamp(p, 0);
        if (ph.kouter) {
          tma_load_3d(sa, &M.a[p], &full_bar[prod], 0, m_blk * ph.bm, g * ks);
        } else {
          for (int sub = 0; sub < ks; ++sub)
            tma_load_2d(sa + sub * asub, &M.a[p], &full_bar[prod], (g * ks + sub) * 64, m_blk * ph.bm);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    int q = 0;
    int acc = 0;
    uint32_t acc_par = 0;
    for (int p = 0; p < n_phases; ++p) {
      const Share sh = share(p);
      const int nu = sh.nu;
      if (nu == 0) continue;
      const DLPhase& ph = P.ph[p];
      const int asub = ph.bm * kDLRowBytes;
      const int ks = dl_ks(ph.bm, ph.units_box);
      const int kblocks = (ph.K + 63) / 64;
      const int groups = (kblocks + ks - 1) / ks;
      const int bsub = ph.units_box * kDLUnitBytes;
      const uint32_t idesc = umma_idesc(Cvt<T>::kUmmaFmt, static_cast<uint32_t>(ph.bm), static_cast<uint32_t>(nu * kDLUnit), 0, 0);
      dl_mbar_wait(&tmem_empty[acc], acc_par ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int g = 0; g < groups; ++g, ++q) {
        const int stage = q % STAGES;
        dl_mbar_wait(&full_bar[stage], (q / STAGES) & 1);
        if (g == 0) stamp(p, 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + ks * asub;
        for (int sub = 0; sub < ks; ++sub) {
          const uint64_t adesc = umma_desc_sw128(sa + sub * asub, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sb + sub * bsub, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (g != 0) || (sub != 0) || (k != 0));
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(&tmem_full[acc]);
      stamp(p, 2);
      acc ^= 1;
      if (acc == 0) acc_par ^= 1;
    }
  } else if (warp == 12 && lane == 0) {
    // ===================== grid-barrier poller =====================
    for (int p = 1; p < n_phases; ++p) {
      const unsigned int target = static_cast<unsigned int>(p) * static_cast<unsigned int>(grid);
      long long spins = 0;
      while (true) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.sync) : "memory");
        if (v >= target) break;
        if (++spins > (1ll << 23)) __trap();
      }
      mbar_arrive(&ready_bar[p]);
      stamp(p - 1, 6);
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue =====================
    // UMMA M = 64: accumulator row i sits in lane (i % 16) of TMEM quadrant i / 16, so lanes 16-31 of every warp idle;
    // M = 128: row i = lane i.
    const int ct = threadIdx.x - 128;
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_par = 0;
    for (int p = 0; p < n_phases; ++p) {
      const Share sh = share(p);
      const int u0 = sh.u0, nu = sh.nu, m_blk = sh.m_blk, slot = sh.slot;
      const DLPhase& ph = P.ph[p];
      const int trow = ph.bm == 128 ? quad * 32 + lane : quad * 16 + lane;     // row inside the row block
      const long long grow = static_cast<long long>(m_blk) * ph.bm + trow;
      const bool row_ok = (ph.bm == 128 || lane < 16) && grow < P.R;
      const bool fold = (ph.flags & DL_FOLD) != 0;
      // ---- per-column vectors of this CTA's columns into shared memory while the main loop runs: c1 | c2 (LN fold) or
      //      the bias; constants of the model, so no need to wait for the previous phase
      for (int i = ct; i < nu * kDLUnit; i += 256) {
        const int n = u0 * kDLUnit + i;
        if (fold) {
          s_vec[i] = __ldg(ph.c1 + n);
          s_vec[kDLMaxUnits * kDLUnit + i] = __ldg(ph.c2 + n);
        } else {
          s_vec[i] = Cvt<T>::to_f(__ldg(reinterpret_cast<const T*>(ph.bias) + n));
        }
      }
      if (p > 0) dl_mbar_wait(&ready_bar[p], 0);
      // ---- LayerNorm statistics of this row from the partials its producer left (model.py:39-41, eps 1e-5)
      float mean = 0.f, rstd = 0.f;
      if (fold && row_ok) {
        // partials were left by the writer's 64-row split: the CTAs of row block grow / 64
        const int wt = (P.R + 63) / 64, wm = static_cast<int>(grow >> 6);
        const int slots = (p == 0 && P.ln_slots_in > 0) ? P.ln_slots_in : (grid - wm + wt - 1) / wt;
        float n = 0.f, m2 = 0.f;
        for (int s0 = 0; s0 < slots; s0 += 16) {
          float4 part[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            part[j] = (s0 + j < slots) ? __ldcg(P.ln_part + static_cast<long long>(s0 + j) * P.ln_ld + grow)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 16; ++j) chan_merge(n, mean, m2, part[j].x, part[j].y, part[j].z);
        }
        rstd = rsqrtf(m2 / n + 1e-5f);
      }
      // ---- residual rows of the first two chunks of this thread, fetched before the accumulator is ready
      uint4 xo[2][2];
      if ((ph.flags & DL_RESID) && row_ok) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int e = half + 2 * j;
          if (e < nu) {
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(ph.out) + grow * ph.ldo + (u0 + e) * kDLUnit);
            xo[j][0] = __ldcg(src);
            xo[j][1] = __ldcg(src + 1);
          }
        }
      }
      if (ct == 0) stamp(p, 7);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // s_vec is complete
      float sn = 0.f, smean = 0.f, sm2 = 0.f;          // LN partial of the values this thread writes
      if (nu > 0) {
        dl_mbar_wait(&tmem_full[acc], acc_par);
        if (ct == 0) stamp(p, 3);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
        for (int e = half, j = 0; e < nu; e += 2, ++j) {
          uint32_t r[16];
          tmem_ld16(taddr + e * kDLUnit, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          if (row_ok) {
            const int nb = (u0 + e) * kDLUnit;
            const float* v0 = s_vec + e * kDLUnit;
            if (fold) {
              const float* v1 = v0 + kDLMaxUnits * kDLUnit;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaf(rstd, v[i] - mean * v0[i], v1[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += v0[i];
            }
            if (ph.flags & DL_GELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = gelu_erf(round_to<T>(v[i]));
            }
            T* out = reinterpret_cast<T*>(ph.out) + grow * ph.ldo + nb;
            if (ph.flags & DL_RESID) {
              uint4 u0v, u1v;
              if (j == 0) {
                u0v = xo[0][0];
                u1v = xo[0][1];
              } else if (j == 1) {
                u0v = xo[1][0];
                u1v = xo[1][1];
              } else {
                u0v = __ldcg(reinterpret_cast<const uint4*>(out));
                u1v = __ldcg(reinterpret_cast<const uint4*>(out) + 1);
              }
              const uint32_t w[8] = {u0v.x, u0v.y, u0v.z, u0v.w, u1v.x, u1v.y, u1v.z, u1v.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 f = Cvt<T>::unpack2(w[i]);
                v[2 * i] = round_to<T>(v[2 * i]) + f.x;
                v[2 * i + 1] = round_to<T>(v[2 * i + 1]) + f.y;
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = Cvt<T>::pack2(v[2 * i], v[2 * i + 1]);
            *reinterpret_cast<uint4*>(out) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *(reinterpret_cast<uint4*>(out) + 1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (ph.flags & DL_STATS) {
              // statistics of the STORED (16-bit rounded) values: what the next LayerNorm would read
              float w16[16], cs = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 f = Cvt<T>::unpack2(pk[i]);
                w16[2 * i] = f.x;
                w16[2 * i + 1] = f.y;
                cs += f.x + f.y;
              }
              const float cm = cs * (1.0f / 16.0f);
              float c2 = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float t = w16[i] - cm;
                c2 = fmaf(t, t, c2);
              }
              chan_merge(sn, smean, sm2, 16.f, cm, c2);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        acc ^= 1;
        if (acc == 0) acc_par ^= 1;
      }
      if (ct == 0) stamp(p, 4);
      if (ph.flags & DL_STATS) {
        if (lane < 16) s_stats[half * 64 + trow] = make_float4(sn, smean, sm2, 0.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct < 64) {
          const float4 a = s_stats[ct], b = s_stats[64 + ct];
          float n = a.x, mu = a.y, m2 = a.z;
          chan_merge(n, mu, m2, b.x, b.y, b.z);
          __stcg(P.ln_part + static_cast<long long>(slot) * P.ln_ld + m_blk * 64 + ct, make_float4(n, mu, m2, 0.f));
        }
      }
      if (p + 1 < n_phases) {
        // publish this CTA's outputs: the CTA barrier orders every epilogue thread's stores before thread 0, whose
        // gpu-scope release (cumulative) publishes them with the arrival; the consumers' producer threads add the
        // generic -> async proxy fence on their side before any TMA read
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct == 0) {
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(P.sync), "r"(1u) : "memory");
          stamp(p, 5);
        }
      }
    }
    // the last CTA out re-arms the counters for the next launch (which cannot touch them before this grid is complete)
    if (ct == 0 && n_phases > 1) {
      const unsigned int prev = atomicAdd(P.sync + 1, 1u);
      if (prev == static_cast<unsigned int>(grid) - 1) {
        P.sync[0] = 0;
        P.sync[1] = 0;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kDLTmemCols);
}

// =================================================================================================
// few-rows form: R <= 32 where it fits (one audio greedy, one audio x 5 beams, 32 streams of a small model)
// =================================================================================================
// With a handful of rows the tile form above is all fixed cost: a 64-row UMMA tile holds <= 8 real rows, every K = 16
// step still costs ~58 clk of tensor-pipe issue at N = 16, and a phase takes 6-18 us for 20-90 KB of weights per SM
// (measured on the turbo batch-1 run, profiles/r2_launches_turbo_b1_tileform.csv: 46 us per launch, 9 launches per token).
// This form keeps the phase structure, the LayerNorm folding and the grid barrier, and replaces the main loop by a
// weight-stationary matrix-vector product on mma.sync with the operands swapped:
//   * every CTA owns N / grid consecutive output features (8-35 weight rows); ONE bulk copy per weight row brings its
//     whole slab (<= 92 KB) into shared memory, issued by the 32 lanes of a control warp as soon as the previous
//     phase's main loop has released the buffer - so the slab of phase p + 1 streams in from HBM while phase p runs its
//     epilogue and the grid barrier closes (the slab of the first phase: before griddepcontrol.wait);
//   * the input rows (R x K 16-bit; R <= 8 for the large models, up to 32 where slab + rows fit in 227 KB) follow by bulk
//     copy once the barrier has opened;
//   * D[16 features x 8 rows] += W[16 x 16] . X^T[16 x 8] per 8-row operand tile: the weight rows are the M side of
//     m16n8k16, the input rows the N side; the 8 compute warps split K eight ways, partial sums meet in shared memory.  Rows are
//     padded by 16 bytes so that ldmatrix is conflict-free; features / input rows that do not exist read whatever the
//     buffer holds and only ever reach accumulator entries nobody looks at.
//   * epilogue per output element as in the tile form (fold / bias / erf-GELU / residual), LN partials per
//     (row, CTA): count = this CTA's features, merged by the consumer with Chan's formula.
// Whole stack in one launch: the phases may also come from a table in global memory that strings every layer's chains
// together with its self-attention (kv append + attention over the row's lineage through the parent table), its
// cross-attention ((row, head, key slice) items over all SMs + a merge phase), the final LayerNorm and the logits (a
// Linear whose share of 351 features per CTA passes through the slab buffer in several slabs).  The control warp keeps
// prefetching the slab of the next Linear through the attention phases, so the weights stream continuously.
constexpr int kDRComputeWarps = 8;
constexpr int kDRThreads = (kDRComputeWarps + 1) * 32;
constexpr int kDRMaxTiles = 3;                                  // 16-feature tiles per CTA and phase (N / grid <= 48)
constexpr int kDRCols = kDRMaxTiles * 16;
// scratch tail for NT 8-row operand tiles (R <= 8 NT): fp32 partial sums [warp][tile][16][8 NT], stored values
// [8 NT][48], mean | rstd per row, six mbarriers
__host__ __device__ constexpr int dr_red_floats(int nt) { return kDRComputeWarps * kDRMaxTiles * 16 * 8 * nt; }
__host__ __device__ constexpr int dr_tail_bytes(int nt) { return (dr_red_floats(nt) + 8 * nt * kDRCols + 2 * 8 * nt) * 4 + 6 * 8 + 64; }

// ---- attention inside the few-rows kernel: one query row, keys in blocks of four per warp.  Lane = (key of the block
// g = lane / 8, 16-byte chunk c = lane % 8 of the 64-wide head): every load instruction fetches four complete 128-byte K
// (or V) rows, the dot product finishes with three xor-shuffles inside the 8-lane group, every group keeps an online
// softmax state that is merged across the groups, then across the warps through shared memory.
constexpr float kDRScaleLog2 = 0.125f * 1.4426950408889634f;      // (1 / sqrt(64)) * log2(e)
struct AttAcc {
  float m, l, o[8];
};
__device__ __forceinline__ void att_init(AttAcc& a) {
  a.m = -INFINITY;
  a.l = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) a.o[e] = 0.f;
}
template <typename T>
__device__ __forceinline__ void att_load_q(float (&q)[8], const void* src) {
  const uint4 u = __ldcg(reinterpret_cast<const uint4*>(src));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = Cvt<T>::unpack2(w[e]);
    q[2 * e] = f.x * kDRScaleLog2;
    q[2 * e + 1] = f.y * kDRScaleLog2;
  }
}
// one key per 8-lane group (all 32 lanes call this together)
template <typename T>
__device__ __forceinline__ void att_update(AttAcc& a, const float (&q)[8], uint4 kk, uint4 vv, bool valid) {
  const uint32_t wk[4] = {kk.x, kk.y, kk.z, kk.w};
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = Cvt<T>::unpack2(wk[e]);
    s = fmaf(q[2 * e], f.x, s);
    s = fmaf(q[2 * e + 1], f.y, s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (!valid) return;
  const float mn = fmaxf(a.m, s);
  const float al = fast_exp2(a.m - mn);           // first key: exp2(-inf) = 0
  const float pr = fast_exp2(s - mn);
  a.m = mn;
  a.l = fmaf(a.l, al, pr);
  const uint32_t wv[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = Cvt<T>::unpack2(wv[e]);
    a.o[2 * e] = fmaf(a.o[2 * e], al, pr * f.x);
    a.o[2 * e + 1] = fmaf(a.o[2 * e + 1], al, pr * f.y);
  }
}
__device__ __forceinline__ void att_merge(float& m, float& l, float (&o)[8], float m2, float l2, const float (&o2)[8]) {
  const float mn = fmaxf(m, m2);
  const float a1 = m == -INFINITY ? 0.f : fast_exp2(m - mn);
  const float a2 = m2 == -INFINITY ? 0.f : fast_exp2(m2 - mn);
  l = l * a1 + l2 * a2;
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = o[e] * a1 + o2[e] * a2;
  m = mn;
}
// the four key groups of a warp -> every lane holds the warp's state for its chunk; lanes 0-7 park it in shared memory
__device__ __forceinline__ void att_park(AttAcc& a, float* slot /* [68] of this warp */) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int sh = 8; sh <= 16; sh <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, a.m, sh), l2 = __shfl_xor_sync(0xffffffffu, a.l, sh);
    float o2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o2[e] = __shfl_xor_sync(0xffffffffu, a.o[e], sh);
    att_merge(a.m, a.l, a.o, m2, l2, o2);
  }
  if (lane < 8) {
#pragma unroll
    for (int e = 0; e < 8; ++e) slot[lane * 8 + e] = a.o[e];
  }
  if (lane == 0) {
    slot[64] = a.m;
    slot[65] = a.l;
  }
}
// lanes 0-7 of one warp: merge the parked states of the compute warps for chunk c = lane
__device__ __forceinline__ void att_gather(const float* slots, int n_warps, int c, float& m, float& l, float (&o)[8]) {
  m = -INFINITY;
  l = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  for (int w = 0; w < n_warps; ++w) {
    const float* sl = slots + w * 68;
    float o2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o2[e] = sl[c * 8 + e];
    att_merge(m, l, o, sl[64], sl[65], o2);
  }
}
template <typename T>
__device__ __forceinline__ uint4 att_pack(const float (&o)[8], float inv) {
  uint4 u;
  u.x = Cvt<T>::pack2(o[0] * inv, o[1] * inv);
  u.y = Cvt<T>::pack2(o[2] * inv, o[3] * inv);
  u.z = Cvt<T>::pack2(o[4] * inv, o[5] * inv);
  u.w = Cvt<T>::pack2(o[6] * inv, o[7] * inv);
  return u;
}

// features of a CTA's share that fit the slab buffer (the region in front of the input rows) at once
__host__ __device__ __forceinline__ int dr_slab_cols(int w_region_bytes, int K) {
  const int c = w_region_bytes / (K * 2 + 16);
  return c > kDRCols ? kDRCols : (c < 1 ? 1 : c);
}
// A Linear whose per-CTA share exceeds 48 features (the logits) goes through the buffer in SEVERAL slabs; those use the
// two halves of the region alternately, so that slab j + 1 streams in while slab j is multiplied.
struct DRSlabs {
  bool multi;
  int half;      // byte offset of the second half (multi only)
  int cap;       // features per slab
};
__host__ __device__ __forceinline__ DRSlabs dr_slabs(int w_region_bytes, int N, int K, int grid) {
  DRSlabs g;
  g.multi = (N + grid - 1) / grid > kDRCols;
  g.half = (w_region_bytes / 2) & ~127;
  g.cap = dr_slab_cols(g.multi ? g.half : w_region_bytes, K);
  return g;
}

template <typename T, int NT>
__global__ void __launch_bounds__(kDRThreads, 1) dec_rows_kernel(const DLParams P) {
  constexpr int RMAX = 8 * NT;
  constexpr int J = (kDRCols * RMAX + 255) / 256;             // output elements per thread
  pdl_launch_dependents();
  extern __shared__ uint8_t dr_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dr_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* sW = smem;
  uint8_t* sA = smem + P.dr_a_off;
  float* s_red = reinterpret_cast<float*>(smem + P.dr_tail_off);   // [warp][tile][16 features][RMAX rows]; attention: [warp][68]
  float* s_out = s_red + dr_red_floats(NT);                         // [row][feature]: the values as stored
  float* s_stat = s_out + RMAX * kDRCols;                           // mean[RMAX] | rstd[RMAX]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(s_stat + 2 * RMAX);   // [2]: a weight slab has landed (buffer 0 / second half)
  uint64_t* w_empty = w_full + 2;     // [2]: ... and has been multiplied
  uint64_t* a_full = w_full + 4;      // the input rows of a Linear phase have landed
  uint64_t* go = w_full + 5;          // phase p may start: the grid barrier behind phase p - 1 has opened
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int R = P.R;
  if (tid == 0) {
    mbar_init(&w_full[0], 1);
    mbar_init(&w_full[1], 1);
    mbar_init(&w_empty[0], 1);
    mbar_init(&w_empty[1], 1);
    mbar_init(a_full, 1);
    mbar_init(go, 1);
    mbar_fence_init();
  }
  __syncthreads();
  // phase p: from the kernel parameters (a chain of Linears) or from the table in global memory (a whole decoder stack)
  auto phase = [&](int p) -> DLPhase {
    if (P.table) return P.table[p];
    return P.ph[p];
  };
  auto phase_type = [&](int p) -> int { return P.table ? P.table[p].type : DS_LINEAR; };
  auto next_linear = [&](int p) {        // first Linear phase at or after p (n_phases if none)
    while (p < P.n_phases && phase_type(p) != DS_LINEAR) ++p;
    return p;
  };

  if (warp == kDRComputeWarps) {
    // ===================== control warp: bulk copies and the grid barrier =====================
    // slab j of Linear phase p: the weight rows of this CTA's features [j * cap, (j + 1) * cap) - one slab per phase (buffer
    // 0, the whole region) unless the CTA owns more features than fit (the logits: 351 features of K = 1280 -> 20 slabs
    // alternating between the two halves of the region).  A buffer is re-used once its previous slab has been multiplied.
    int issued0 = 0, issued1 = 0, waited0 = 0, waited1 = 0;      // per buffer (scalars: no local-memory arrays)
    auto ensure_free = [&](int b) {
      if (b == 0) {
        if (waited0 < issued0) {
          dl_mbar_wait(&w_empty[0], waited0 & 1);
          ++waited0;
        }
      } else if (waited1 < issued1) {
        dl_mbar_wait(&w_empty[1], waited1 & 1);
        ++waited1;
      }
    };
    auto issue_w = [&](int p, int j) {
      const DLPhase ph = phase(p);
      const int n0 = static_cast<int>(static_cast<long long>(cta) * ph.N / grid);
      const int nc = static_cast<int>(static_cast<long long>(cta + 1) * ph.N / grid) - n0;
      const uint32_t row_bytes = static_cast<uint32_t>(ph.K) * 2;
      const DRSlabs g = dr_slabs(P.dr_a_off, ph.N, ph.K, grid);
      const int b = g.multi ? (j & 1) : 0;
      ensure_free(b);
      if (!g.multi) ensure_free(1);      // a whole-region slab also covers the second half
      const int c0 = j * g.cap, ncc = max(0, min(g.cap, nc - c0));
      uint8_t* dst = sW + (b ? g.half : 0);
      if (lane == 0) mbar_expect_tx(&w_full[b], static_cast<uint32_t>(ncc) * row_bytes);
      __syncwarp();
      for (int i = lane; i < ncc; i += 32)
        bulk_load_1d(dst + static_cast<size_t>(i) * (row_bytes + 16),
                     static_cast<const uint8_t*>(ph.w) + static_cast<size_t>(n0 + c0 + i) * row_bytes, row_bytes, &w_full[b]);
      if (b == 0) ++issued0; else ++issued1;
    };
    int lin = next_linear(0);            // the Linear phase whose slab is in flight / resident
    if (lin < P.n_phases) issue_w(lin, 0);  // weights are constants: the first slab streams in under the tail of the previous kernel
    pdl_wait();
    if (P.skip_flag && *P.skip_flag) {
      if (lin < P.n_phases) dl_mbar_wait(&w_full[0], 0);  // never leave with a copy into this CTA's shared memory in flight
      return;
    }
    for (int p = 0; p < P.n_phases; ++p) {
      if (p > 0) {
        const unsigned int target = static_cast<unsigned int>(p) * static_cast<unsigned int>(grid);
        long long spins = 0;
        while (true) {
          unsigned int v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.sync) : "memory");
          if (v >= target) break;
          if (++spins > (1ll << 23)) __trap();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(go);    // the compute warps start on what needs no staging (LN statistics, residual rows, attention)
      if (p != lin) continue;            // an attention / LayerNorm phase: nothing to stage
      const DLPhase ph = phase(p);
      fence_proxy_async_global();        // other CTAs' generic-proxy stores -> this lane's bulk (async-proxy) reads
      const uint32_t row_bytes = static_cast<uint32_t>(ph.K) * 2;
      if (lane == 0) mbar_expect_tx(a_full, static_cast<uint32_t>(R) * row_bytes);
      __syncwarp();
      if (lane < R)                      // R <= 32: one input row per lane
        bulk_load_1d(sA + static_cast<size_t>(lane) * (row_bytes + 16), static_cast<const uint8_t*>(ph.a) + static_cast<size_t>(lane) * ph.lda * 2,
                     row_bytes, a_full);
      {
        const int nc = static_cast<int>(static_cast<long long>(cta + 1) * ph.N / grid) - static_cast<int>(static_cast<long long>(cta) * ph.N / grid);
        const int cap = dr_slabs(P.dr_a_off, ph.N, ph.K, grid).cap;
        for (int j = 1; j * cap < nc; ++j) issue_w(p, j);   // further slabs of this phase, each into the half that is free
      }
      lin = next_linear(p + 1);
      if (lin < P.n_phases) issue_w(lin, 0);   // streams in while phase p finishes and the attention phases in between run
    }
    return;
  }

  // ===================== compute warps =====================
  pdl_wait();
  if (P.skip_flag && *P.skip_flag) return;
  const int g = lane >> 2, t4 = lane & 3;
  const int g4 = lane >> 3, c8 = lane & 7;                    // attention: key of the block, 16-byte chunk
  int used0 = 0, used1 = 0, n_lin = 0;     // slabs consumed per buffer, Linear phases done
  for (int p = 0; p < P.n_phases; ++p) {
    const DLPhase ph = phase(p);
    if (ph.type == DS_LINEAR) {
      const int n0 = static_cast<int>(static_cast<long long>(cta) * ph.N / grid);
      const int nc = static_cast<int>(static_cast<long long>(cta + 1) * ph.N / grid) - n0;
      const DRSlabs sl = dr_slabs(P.dr_a_off, ph.N, ph.K, grid);
      const int cap = sl.cap;            // features per slab (all of them except for the logits)
      const int stride = ph.K * 2 + 16;
      const bool fold = (ph.flags & DL_FOLD) != 0;
      dl_mbar_wait(go, p & 1);           // the previous phase is complete grid-wide (the control warp saw the barrier open)
      // ---- LayerNorm statistics of rows warp, warp + 8, .. from the partials their producer left (model.py:39-41, eps 1e-5)
      if (fold) {
        const int slots = (p == 0 && P.ln_slots_in > 0) ? P.ln_slots_in : grid;
        for (int r = warp; r < R; r += kDRComputeWarps) {
          float n = 0.f, mean = 0.f, m2 = 0.f;
          for (int s0 = lane; s0 < slots; s0 += 32) {
            const float4 part = __ldcg(P.ln_part + static_cast<long long>(s0) * P.ln_ld + r);
            chan_merge(n, mean, m2, part.x, part.y, part.z);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                        qb = __shfl_xor_sync(0xffffffffu, m2, o);
            chan_merge(n, mean, m2, nb, mb, qb);
          }
          if (lane == 0) {
            s_stat[r] = mean;
            s_stat[RMAX + r] = rsqrtf(m2 / n + 1e-5f);
          }
        }
      }
      for (int c0 = 0, js = 0; c0 == 0 || c0 < nc; c0 += cap, ++js) {   // one slab per pass (an empty share still consumes its slab)
        const int wb = sl.multi ? (js & 1) : 0;
        const uint8_t* sWs = sW + (wb ? sl.half : 0);
        const int ncc = max(0, min(cap, nc - c0));
        const int n_tiles = (ncc + 15) >> 4;
        // ---- this thread's <= J output elements (row, feature) of the slab and their constants
        const int n_out = ncc * R;
        int o_r[J], o_n[J];
        float o_c1[J], o_c2[J], o_x[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const int o = tid + j * 256;
          o_r[j] = o < n_out ? o / ncc : -1;
          o_n[j] = o < n_out ? o - o_r[j] * ncc : 0;
          o_c1[j] = o_c2[j] = o_x[j] = 0.f;
          if (o_r[j] >= 0) {
            const int n = n0 + c0 + o_n[j];
            if (fold) {
              o_c1[j] = __ldg(ph.c1 + n);
              o_c2[j] = __ldg(ph.c2 + n);
            } else if (ph.bias) {
              o_c1[j] = Cvt<T>::to_f(__ldg(reinterpret_cast<const T*>(ph.bias) + n));
            }
            if (ph.flags & DL_RESID) o_x[j] = Cvt<T>::to_f(__ldcg(reinterpret_cast<const T*>(ph.out) + o_r[j] * ph.ldo + n));
          }
        }
        // ---- main loop: this warp's eighth of K for every 16-feature tile of the slab
        float acc[kDRMaxTiles][NT][4];
#pragma unroll
        for (int t = 0; t < kDRMaxTiles; ++t)
#pragma unroll
          for (int u = 0; u < NT; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[t][u][i] = 0.f;
        if (c0 == 0) dl_mbar_wait(a_full, n_lin & 1);
        dl_mbar_wait(&w_full[wb], (wb ? used1 : used0) & 1);
        {
          const int kw = ph.K / kDRComputeWarps;
          // ldmatrix x4 on the slab: matrices (features 0-7, k 0-7), (8-15, k 0-7), (0-7, k 8-15), (8-15, k 8-15) = a0..a3
          const uint8_t* wrow = sWs + static_cast<size_t>((lane & 7) + ((lane >> 3) & 1) * 8) * stride + (lane >> 4) * 16;
          // ldmatrix x2 on 8 input rows: (rows 0-7, k 0-7), (rows 0-7, k 8-15) = b0, b1
          const uint8_t* arow = sA + static_cast<size_t>(lane & 7) * stride + ((lane >> 3) & 1) * 16;
          for (int k0 = warp * kw; k0 < (warp + 1) * kw; k0 += 16) {
            uint32_t b[NT][2];
#pragma unroll
            for (int u = 0; u < NT; ++u) ldmatrix_x2(b[u], arow + static_cast<size_t>(u) * 8 * stride + k0 * 2);
#pragma unroll
            for (int t = 0; t < kDRMaxTiles; ++t)
              if (t < n_tiles) {
                uint32_t a[4];
                ldmatrix_x4(a, wrow + static_cast<size_t>(t) * 16 * stride + k0 * 2);
#pragma unroll
                for (int u = 0; u < NT; ++u) mma16816<T>(acc[t][u], a, b[u][0], b[u][1]);
              }
          }
        }
#pragma unroll
        for (int t = 0; t < kDRMaxTiles; ++t)
          if (t < n_tiles) {
            float* dst = s_red + ((warp * kDRMaxTiles + t) * 16) * RMAX;
#pragma unroll
            for (int u = 0; u < NT; ++u) {
              *reinterpret_cast<float2*>(dst + g * RMAX + u * 8 + 2 * t4) = make_float2(acc[t][u][0], acc[t][u][1]);
              *reinterpret_cast<float2*>(dst + (g + 8) * RMAX + u * 8 + 2 * t4) = make_float2(acc[t][u][2], acc[t][u][3]);
            }
          }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) mbar_arrive(&w_empty[wb]);     // the buffer may be refilled while the epilogue runs
        if (wb) ++used1; else ++used0;
        // ---- epilogue
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (o_r[j] >= 0) {
            const int r = o_r[j], nl = o_n[j];
            const float* src = s_red + ((nl >> 4) * 16 + (nl & 15)) * RMAX + r;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kDRComputeWarps; ++w) v += src[w * kDRMaxTiles * 16 * RMAX];
            if (fold)
              v = fmaf(s_stat[RMAX + r], v - s_stat[r] * o_c1[j], o_c2[j]);
            else
              v += o_c1[j];
            if (ph.flags & DL_GELU) v = gelu_erf(round_to<T>(v));
            if (ph.flags & DL_RESID) v = round_to<T>(v) + o_x[j];
            if (ph.flags & DL_OUTF32) {
              reinterpret_cast<float*>(ph.out)[r * ph.ldo + n0 + c0 + nl] = v;
            } else {
              const T tv = Cvt<T>::from_f(v);
              reinterpret_cast<T*>(ph.out)[r * ph.ldo + n0 + c0 + nl] = tv;
              s_out[r * kDRCols + nl] = Cvt<T>::to_f(tv);
            }
          }
        if (c0 + cap < nc) asm volatile("bar.sync 1, 256;" ::: "memory");   // the partial sums are re-used by the next slab
      }
      ++n_lin;
      if (ph.flags & DL_STATS) {         // (single-slab phases only)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < R) {
          // statistics of the STORED (16-bit rounded) values: what the next LayerNorm would read
          const float* row = s_out + tid * kDRCols;
          float sum = 0.f;
          for (int i = 0; i < nc; ++i) sum += row[i];
          const float mean = nc > 0 ? sum / static_cast<float>(nc) : 0.f;
          float m2 = 0.f;
          for (int i = 0; i < nc; ++i) {
            const float d = row[i] - mean;
            m2 = fmaf(d, d, m2);
          }
          __stcg(P.ln_part + static_cast<long long>(cta) * P.ln_ld + tid, make_float4(static_cast<float>(nc), mean, m2, 0.f));
        }
      }
    } else if (ph.type == DS_LN) {
      // ---- LayerNorm of the rows (the decoder's final ln, model.py:243-245): a warp per row, fp32 statistics in two
      //      passes like layernorm_kernel
      dl_mbar_wait(go, p & 1);
      for (int row = cta; row < R; row += grid) {
        if (warp != 0) continue;
        const T* xr = reinterpret_cast<const T*>(ph.a) + static_cast<long long>(row) * ph.lda;
        T* yr = reinterpret_cast<T*>(ph.out) + static_cast<long long>(row) * ph.ldo;
        const int d = ph.N;
        float sum = 0.f;
        for (int c = lane * 8; c < d; c += 256) {
          const uint4 u = __ldcg(reinterpret_cast<const uint4*>(xr + c));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = Cvt<T>::unpack2(w[e]);
            sum += f.x + f.y;
          }
        }
        const float mean = warp_sum(sum) / static_cast<float>(d);
        float sq = 0.f;
        for (int c = lane * 8; c < d; c += 256) {
          const uint4 u = __ldcg(reinterpret_cast<const uint4*>(xr + c));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = Cvt<T>::unpack2(w[e]);
            sq += (f.x - mean) * (f.x - mean);
            sq += (f.y - mean) * (f.y - mean);
          }
        }
        const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(d) + 1e-5f);
        for (int c = lane * 8; c < d; c += 256) {
          const uint4 u = __ldcg(reinterpret_cast<const uint4*>(xr + c));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          float o[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = Cvt<T>::unpack2(w[e]);
            o[2 * e] = (f.x - mean) * rstd * __ldg(ph.c1 + c + 2 * e) + __ldg(ph.c2 + c + 2 * e);
            o[2 * e + 1] = (f.y - mean) * rstd * __ldg(ph.c1 + c + 2 * e + 1) + __ldg(ph.c2 + c + 2 * e + 1);
          }
          uint4 y;
          y.x = Cvt<T>::pack2(o[0], o[1]);
          y.y = Cvt<T>::pack2(o[2], o[3]);
          y.z = Cvt<T>::pack2(o[4], o[5]);
          y.w = Cvt<T>::pack2(o[6], o[7]);
          *reinterpret_cast<uint4*>(yr + c) = y;
        }
      }
    } else if (ph.type == DS_SELF) {
      // ---- self-attention of the new position of every (row, head) over the row's lineage + kv append: model.py:124-127
      //      with the cache of model.py:327-333 read through the parent table (decoding.py:172-176)
      dl_mbar_wait(go, p & 1);
      const int H = P.n_head, ctx = P.ctx, d = P.d;
      const int L = *P.len_ptr, pos_new = L - 1;
      const long long row_stride = static_cast<long long>(H) * ctx * 128;
      for (int pair = cta; pair < R * H; pair += grid) {
        const int row = pair / H, h = pair - row * H;
        const uint8_t* qrow = static_cast<const uint8_t*>(P.qkv) + (static_cast<long long>(row) * 3 * d + h * 64) * 2 + c8 * 16;
        float q[8];
        att_load_q<T>(q, qrow);
        const int* ind = P.indir + static_cast<long long>(row) * ctx;
        uint8_t* kc = static_cast<uint8_t*>(ph.kc) + static_cast<long long>(h) * ctx * 128 + c8 * 16;
        uint8_t* vc = static_cast<uint8_t*>(ph.vc) + static_cast<long long>(h) * ctx * 128 + c8 * 16;
        AttAcc a;
        att_init(a);
        for (int b = warp; b * 4 < L; b += 2 * kDRComputeWarps) {          // two blocks of four keys in flight
          uint4 kk[2], vv[2];
          bool ok[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int key = (b + u * kDRComputeWarps) * 4 + g4;
            ok[u] = key < L;
            kk[u] = vv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (key < pos_new) {
              const long long off = static_cast<long long>(__ldg(ind + key)) * row_stride + static_cast<long long>(key) * 128;
              kk[u] = __ldcg(reinterpret_cast<const uint4*>(kc + off));
              vv[u] = __ldcg(reinterpret_cast<const uint4*>(vc + off));
            } else if (key == pos_new) {
              kk[u] = __ldcg(reinterpret_cast<const uint4*>(qrow + static_cast<long long>(d) * 2));
              vv[u] = __ldcg(reinterpret_cast<const uint4*>(qrow + static_cast<long long>(d) * 4));
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) att_update<T>(a, q, kk[u], vv[u], ok[u]);
        }
        att_park(a, s_red + warp * 68);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 0 && lane < 8) {
          float m, l, o[8];
          att_gather(s_red, kDRComputeWarps, lane, m, l, o);
          *reinterpret_cast<uint4*>(static_cast<uint8_t*>(P.att) + (static_cast<long long>(row) * d + h * 64) * 2 + lane * 16) =
              att_pack<T>(o, 1.0f / l);
        } else if (warp == 1 && lane < 16) {
          // the new position joins the cache of its own row (torch.cat of model.py:327-333)
          const long long off = static_cast<long long>(row) * row_stride + static_cast<long long>(pos_new) * 128;
          const uint4 v = __ldcg(reinterpret_cast<const uint4*>(qrow + static_cast<long long>(d) * (lane < 8 ? 2 : 4)));
          *reinterpret_cast<uint4*>((lane < 8 ? kc : vc) + off) = v;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    } else if (ph.type == DS_CROSS) {
      // ---- cross-attention (model.py:101-109 + SDPA) of every (row, head) over a slice of the audio's keys; with one slice
      //      the output is final, otherwise (m, l, o) partials for the DS_COMBINE phase
      dl_mbar_wait(go, p & 1);
      const int H = P.n_head, d = P.d, S = P.splits, Tn = P.T;
      for (int item = cta; item < R * H * S; item += grid) {
        const int pair = item / S, sp = item - pair * S;
        const int row = pair / H, h = pair - row * H, audio = row / P.G;
        const int k0 = static_cast<int>(static_cast<long long>(sp) * Tn / S), k1 = static_cast<int>(static_cast<long long>(sp + 1) * Tn / S);
        float q[8];
        att_load_q<T>(q, static_cast<const uint8_t*>(P.q) + (static_cast<long long>(row) * d + h * 64) * 2 + c8 * 16);
        const uint8_t* kb = static_cast<const uint8_t*>(ph.kc) + (static_cast<long long>(audio) * 2 * H + h) * Tn * 128 + c8 * 16;
        const uint8_t* vb = kb + static_cast<long long>(H) * Tn * 128;
        AttAcc a;
        att_init(a);
        // eight blocks of four keys in flight per warp: 64 KB per CTA, what it takes to pull ~40 GB/s per SM out of HBM
        for (int b = warp; k0 + b * 4 < k1; b += 8 * kDRComputeWarps) {
          uint4 kk[8], vv[8];
          bool ok[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int key = k0 + (b + u * kDRComputeWarps) * 4 + g4;
            ok[u] = key < k1;
            kk[u] = vv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (ok[u]) {
              kk[u] = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(key) * 128));
              vv[u] = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(key) * 128));
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) att_update<T>(a, q, kk[u], vv[u], ok[u]);
        }
        att_park(a, s_red + warp * 68);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 0 && lane < 8) {
          float m, l, o[8];
          att_gather(s_red, kDRComputeWarps, lane, m, l, o);
          if (S == 1) {
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(P.att) + (static_cast<long long>(row) * d + h * 64) * 2 + lane * 16) =
                att_pack<T>(o, 1.0f / l);
          } else {
            float* dst = P.xpart + static_cast<long long>(item) * 66;
#pragma unroll
            for (int e = 0; e < 8; ++e) __stcg(dst + lane * 8 + e, o[e]);
            if (lane == 0) {
              __stcg(dst + 64, m);
              __stcg(dst + 65, l);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    } else {
      // ---- DS_COMBINE: merge the key slices of every (row, head); a warp per pair, two output features per lane
      dl_mbar_wait(go, p & 1);
      const int H = P.n_head, d = P.d, S = P.splits;
      for (int pair = cta + warp * grid; pair < R * H; pair += grid * kDRComputeWarps) {
        const int row = pair / H, h = pair - row * H;
        const float* src = P.xpart + static_cast<long long>(pair) * S * 66;
        float m = -INFINITY;
        for (int sp = 0; sp < S; ++sp) m = fmaxf(m, __ldcg(src + sp * 66 + 64));
        float l = 0.f, o0 = 0.f, o1 = 0.f;
        for (int sp = 0; sp < S; ++sp) {
          const float ms = __ldcg(src + sp * 66 + 64);
          const float f = ms == -INFINITY ? 0.f : fast_exp2(ms - m);
          l = fmaf(__ldcg(src + sp * 66 + 65), f, l);
          const float2 v = __ldcg(reinterpret_cast<const float2*>(src + sp * 66 + 2 * lane));
          o0 = fmaf(v.x, f, o0);
          o1 = fmaf(v.y, f, o1);
        }
        const float inv = 1.0f / l;
        *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(P.att) + (static_cast<long long>(row) * d + h * 64 + 2 * lane) * 2) =
            Cvt<T>::pack2(o0 * inv, o1 * inv);
      }
    }
    if (p + 1 < P.n_phases) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(P.sync), "r"(1u) : "memory");
    }
  }
  if (tid == 0 && P.n_phases > 1) {
    const unsigned int prev = atomicAdd(P.sync + 1, 1u);
    if (prev == static_cast<unsigned int>(grid) - 1) {
      P.sync[0] = 0;
      P.sync[1] = 0;
      __threadfence();
    }
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
int dl_grid_size() { return sm_count(); }

// rows per row block of a phase: the wide phases without LN partials (QKV, fc1) use UMMA M = 128 - at M = 64 the tensor
// pipe runs at less than half its per-column rate and those phases were MMA-bound (profiles/r2_dec_layer_trace_v3.txt)
static int dl_auto_bm(int R, int N, int K, int flags) {
  return (!(flags & DL_STATS) && R > 64 && N >= 2 * K) ? 128 : 64;
}

static int dl_box_units(int R, int grid, int N, int bm) {
  const int m_tiles = (R + bm - 1) / bm;
  if (m_tiles > grid) return 1 << 20;
  const int min_slots = grid / m_tiles;                         // fewest CTAs a row block gets
  const int box = ((N / kDLUnit) + min_slots - 1) / min_slots;  // widest share: the weight box every CTA loads
  return box < 1 ? 1 : box;
}

bool dl_supported(int R, int d, int grid) {
  if (R <= 0 || d % 64 != 0 || grid <= 0) return false;
  const int shapes[4][3] = {{3 * d, d, DL_FOLD}, {d, d, DL_RESID | DL_STATS}, {4 * d, d, DL_FOLD | DL_GELU}, {d, 4 * d, DL_RESID | DL_STATS}};
  for (auto& sh : shapes) {
    const int bm = dl_auto_bm(R, sh[0], sh[1], sh[2]);
    if (dl_box_units(R, grid, sh[0], bm) > kDLMaxUnits) return false;
  }
  return true;
}

int dl_fill_phase(DLLaunch& L, int idx, int dtype, int R, int grid, const void* A, long long lda, const void* W, int N, int K,
                  const void* bias, const float* c1, const float* c2, int flags, void* out, long long ldo, int bm) {
  if (idx >= kDLMaxPhases || N % kDLUnit || K % 64) return 1;
  DLPhase& ph = L.p.ph[idx];
  ph.N = N;
  ph.K = K;
  ph.flags = flags;
  ph.bias = bias;
  ph.c1 = c1;
  ph.c2 = c2;
  ph.out = out;
  ph.ldo = ldo;
  ph.a = A;
  ph.lda = lda;
  ph.w = W;
  ph.type = DS_LINEAR;
  ph.kc = ph.vc = nullptr;
  if (bm == 0) bm = dl_auto_bm(R, N, K, flags);
  if ((bm != 64 && bm != 128) || (bm == 128 && (flags & DL_STATS))) return 7;
  ph.bm = bm;
  const int box = dl_box_units(R, grid, N, bm);
  if (box > kDLMaxUnits) return 5;
  ph.units_box = box;
  const int ks = dl_ks(bm, box);
  static int kouter_ok = -1;          // the driver may refuse a k-block stride (128 B) below the row stride: fall back to 2-D
  if (kouter_ok < 0) {
    const char* e = getenv("WB200_DL_KOUTER");
    kouter_ok = (e && e[0] == '0') ? 0 : 1;
  }
  ph.kouter = 0;
  if (kouter_ok) {
    uint64_t da[3] = {64, static_cast<uint64_t>(R), static_cast<uint64_t>(K / 64)};
    uint64_t sa[2] = {static_cast<uint64_t>(lda * 2), 128};
    uint32_t ba[3] = {64, static_cast<uint32_t>(bm), static_cast<uint32_t>(ks)};
    uint64_t db[3] = {64, static_cast<uint64_t>(N), static_cast<uint64_t>(K / 64)};
    uint64_t sb[2] = {static_cast<uint64_t>(K) * 2, 128};
    uint32_t bb[3] = {64, static_cast<uint32_t>(box * kDLUnit), static_cast<uint32_t>(ks)};
    if (make_tmap_16bit(&L.maps.a[idx], dtype, A, 3, da, sa, ba) == 0 && make_tmap_16bit(&L.maps.b[idx], dtype, W, 3, db, sb, bb) == 0)
      ph.kouter = 1;
    else
      kouter_ok = 0;
  }
  if (!ph.kouter) {
    uint64_t da[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(R)};
    uint64_t sa[1] = {static_cast<uint64_t>(lda * 2)};
    uint32_t ba[2] = {64, static_cast<uint32_t>(bm)};
    if (make_tmap_16bit(&L.maps.a[idx], dtype, A, 2, da, sa, ba)) return 2;
    uint64_t db[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t sb[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t bb[2] = {64, static_cast<uint32_t>(box * kDLUnit)};
    if (make_tmap_16bit(&L.maps.b[idx], dtype, W, 2, db, sb, bb)) return 3;
  }
  return 0;
}

void dl_init_launch(DLLaunch& L, int dtype, int R, int grid, float4* ln_part, int ln_ld, unsigned int* sync,
                    const int* skip_flag, int ln_slots_in) {
  L.dtype = dtype;
  L.grid = grid;
  L.p.n_phases = 0;
  L.p.R = R;
  L.p.m_tiles = (R + 63) / 64;
  L.p.ln_slots_in = ln_slots_in;
  L.p.ln_part = ln_part;
  L.p.ln_ld = ln_ld;
  L.p.sync = sync;
  L.p.skip_flag = skip_flag;
  L.p.trace = nullptr;
  L.p.dr_a_off = L.p.dr_tail_off = 0;
  L.p.table = nullptr;
  L.p.qkv = L.p.q = nullptr;
  L.p.att = nullptr;
  L.p.indir = L.p.len_ptr = nullptr;
  L.p.xpart = nullptr;
  L.p.n_head = L.p.ctx = L.p.T = L.p.G = L.p.splits = L.p.d = 0;
  L.rows_smem = 0;
}

// shared-memory layout of the few-rows form over a set of Linear phases; 0 if it does not fit
static int dr_layout(const DLPhase* ph, int n, int R, int grid, int* a_off_out, int* tail_off_out) {
  const int nt = R <= 8 ? 1 : (R <= 16 ? 2 : 4);       // 8-row operand tiles of the input rows
  long long w_bytes = 0, a_bytes = 0;
  int n_linear = 0;
  for (int p = 0; p < n; ++p) {
    if (ph[p].type != DS_LINEAR) continue;
    ++n_linear;
    if (ph[p].K % (16 * kDRComputeWarps) || ph[p].N < 1) return 0;
    const long long stride = ph[p].K * 2LL + 16;
    const int nc_max = (ph[p].N + grid - 1) / grid;
    // a share of more than 48 features (the logits) goes through the buffer in several slabs: it does not size it
    if (nc_max <= kDRCols) w_bytes = std::max(w_bytes, nc_max * stride);
    else if (ph[p].flags & DL_STATS) return 0;
    a_bytes = std::max(a_bytes, 8LL * nt * stride);
  }
  if (!n_linear) return 0;
  if (w_bytes == 0) w_bytes = 64 * 1024;
  const long long a_off = (w_bytes + 127) / 128 * 128;
  long long extent = 0;
  for (int p = 0; p < n; ++p) {
    if (ph[p].type != DS_LINEAR) continue;
    const long long stride = ph[p].K * 2LL + 16;
    const int nc_max = (ph[p].N + grid - 1) / grid;
    const DRSlabs g = dr_slabs(static_cast<int>(a_off), ph[p].N, ph[p].K, grid);
    const int cols = std::min(nc_max, g.cap);
    if (g.multi && cols < 16) return 0;                                // slabs too thin to be worth it
    extent = std::max(extent, (g.multi ? g.half : 0) + ((cols + 15) / 16 * 16) * stride);     // ldmatrix reads whole 16-row tiles
  }
  const long long tail_off = (std::max(a_off + a_bytes, extent) + 127) / 128 * 128;
  const long long total = tail_off + dr_tail_bytes(nt) + 128;
  if (total > 227 * 1024) return 0;
  *a_off_out = static_cast<int>(a_off);
  *tail_off_out = static_cast<int>(tail_off);
  return static_cast<int>(total);
}

static bool rows_form_enabled() {
  if (g_fused_rows < 0) {
    const char* e = getenv("WB200_FUSED_ROWS");
    g_fused_rows = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fused_rows != 0;
}

bool dl_use_rows_form(DLLaunch& L) {
  L.rows_smem = 0;
  L.p.table = nullptr;
  if (!rows_form_enabled() || L.p.R > kDRMaxRows || L.p.n_phases <= 0) return false;
  L.rows_smem = dr_layout(L.p.ph, L.p.n_phases, L.p.R, L.grid, &L.p.dr_a_off, &L.p.dr_tail_off);
  return L.rows_smem > 0;
}

bool dl_plan_stack(DLLaunch& L, const DLPhase* host_table, int n, const DLPhase* device_table) {
  L.rows_smem = 0;
  if (!rows_form_enabled() || L.p.R > kDRMaxRows || n <= 0) return false;
  L.rows_smem = dr_layout(host_table, n, L.p.R, L.grid, &L.p.dr_a_off, &L.p.dr_tail_off);
  if (L.rows_smem <= 0) return false;
  L.p.table = device_table;
  L.p.n_phases = n;
  return true;
}

template <typename T>
static int dl_launch_t(const DLLaunch& L, cudaStream_t s) {
  constexpr int STAGES = 3;
  using Cfg = DLCfg<STAGES>;
  auto kern = dec_layer_kernel<T, STAGES>;
  static SmemOptIn optin;
  if (!optin.ensure(kern, Cfg::kSmemBytes)) return 60;
  ProfileScope prof(PROF_DEC_LAYER, s);
  const cudaError_t le = launch_pdl(kern, dim3(L.grid), dim3(kDLThreads), Cfg::kSmemBytes, s, L.p, L.maps);
  count_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 61;
}

template <typename T, int NT>
static int dr_launch_t(const DLLaunch& L, cudaStream_t s) {
  auto kern = dec_rows_kernel<T, NT>;
  static SmemOptIn optin;
  if (!optin.ensure(kern, 227 * 1024)) return 62;
  ProfileScope prof(PROF_DEC_LAYER, s);
  const cudaError_t le = launch_pdl(kern, dim3(L.grid), dim3(kDRThreads), L.rows_smem, s, L.p);
  count_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 63;
}

int dl_launch(const DLLaunch& L, cudaStream_t s) {
  if (L.p.n_phases <= 0) return 0;
  if (L.rows_smem > 0) {
    const bool bf = L.dtype == DT_BF16;
    if (L.p.R <= 8) return bf ? dr_launch_t<__nv_bfloat16, 1>(L, s) : dr_launch_t<__half, 1>(L, s);
    if (L.p.R <= 16) return bf ? dr_launch_t<__nv_bfloat16, 2>(L, s) : dr_launch_t<__half, 2>(L, s);
    return bf ? dr_launch_t<__nv_bfloat16, 4>(L, s) : dr_launch_t<__half, 4>(L, s);
  }
  return L.dtype == DT_BF16 ? dl_launch_t<__nv_bfloat16>(L, s) : dl_launch_t<__half>(L, s);
}

}  // namespace wb
