// Decoder attention for the autoregressive step (and its n_init-token prefill): the HBM-bound
// part of TextDecoder.forward (reference whisper/model.py:81-139 as driven by the kv-cache hooks
// of model.py:310-341 and decoding.py:144-176).
//
// Both kernels stream K/V rows (64 x 16-bit = 128 B per head per position) through a 4-stage
// cp.async ring in shared memory and do the (tiny) math on mma.sync m16n8k16 with fp32
// accumulation and an online softmax: the "query" side of the tile is the <= 16 queries that
// share one K/V stream, so every K/V byte is read from HBM once per step.
//
//   cross_attention_kernel : queries = the G beams (step) or n_init prompt positions (prefill) of
//                            ONE audio; K/V = that audio's 1500 encoder positions, shared by all
//                            of them ("each audio's K/V counted once").  The 1500
//                            keys are split across CTAs (flash-decoding); the last CTA to finish a
//                            (audio, head, q-tile) combines the partials - no second launch.
//   self_attention_kernel  : one WARP per (row, head), plain streaming reduction through a per-lane
//                            cp.async ring (no tensor-core tile: a single query has nothing to
//                            share); keys are gathered through the beam
//                            indirection table (position p of row r lives in physical row
//                            indir[r][p]), so a beam reorder is a table update, not the physical
//                            gather of every cache tensor that decoding.py:172-176 performs.  In
//                            step mode the kernel also APPENDS the new token's K/V to the cache
//                            (the torch.cat of model.py:327-333).
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tmap.cuh"

namespace wb {

constexpr int kDaThreads = 128;  // 4 warps, each owns 16 keys of every 64-key tile
constexpr int kDaTileKeys = 64;
constexpr int kDaTileBytes = kDaTileKeys * 128;                 // K or V tile

constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;      // (1/sqrt(64)) * log2(e)

struct RowSrc {
  const uint8_t* k;
  const uint8_t* v;
};

// Per-warp online-softmax state for a 16-query tile: rows g and g+8 of the mma fragment.
struct WarpAcc {
  float o[8][4];
  float m[2];
  float l[2];
};

__device__ __forceinline__ void acc_init(WarpAcc& a) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a.o[i][j] = 0.f;
  a.m[0] = a.m[1] = -INFINITY;
  a.l[0] = a.l[1] = 0.f;
}

// One warp consumes its 16 keys (rows wrow0..wrow0+15) of the staged K/V tile.
// n_valid: number of valid keys among those 16 (<= 0 means none).
template <typename T>
__device__ __forceinline__ void warp_tile(const uint8_t* sK, const uint8_t* sV, int wrow0,
                                          const uint32_t (&qa)[4][4], int n_valid, WarpAcc& acc) {
  const int lane = threadIdx.x & 31;
  const int t = lane & 3;
  if (n_valid <= 0) return;
  // ---- S = Q K^T for 16 keys: two n-tiles of 8 keys, four k-steps over dh = 64
  float s[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
  {
    const int m = lane >> 3;                       // matrix id for ldmatrix.x4
    const int krow = wrow0 + (m >> 1) * 8 + (lane & 7);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int chunk = ks * 2 + (m & 1);          // 16-byte chunk index along dh
      uint32_t kb[4];
      ldmatrix_x4(kb, sK + krow * 128 + ((chunk ^ (krow & 7)) << 4));
      mma16816<T>(s[0], qa[ks], kb[0], kb[1]);
      mma16816<T>(s[1], qa[ks], kb[2], kb[3]);
    }
  }
  // ---- mask keys beyond n_valid, online softmax (rows g -> idx 0,1 ; g+8 -> idx 2,3)
  if (n_valid < 16) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = nt * 8 + 2 * t + (j & 1);
        if (key >= n_valid) s[nt][j] = -INFINITY;
      }
  }
  float mx0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
  float mx1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float mn0 = fmaxf(acc.m[0], mx0 * kScaleLog2);
  const float mn1 = fmaxf(acc.m[1], mx1 * kScaleLog2);
  const float al0 = fast_exp2(acc.m[0] - mn0);     // mn finite here (>=1 valid key)
  const float al1 = fast_exp2(acc.m[1] - mn1);
  acc.m[0] = mn0;
  acc.m[1] = mn1;
  float p[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    p[nt][0] = fast_exp2(s[nt][0] * kScaleLog2 - mn0);
    p[nt][1] = fast_exp2(s[nt][1] * kScaleLog2 - mn0);
    p[nt][2] = fast_exp2(s[nt][2] * kScaleLog2 - mn1);
    p[nt][3] = fast_exp2(s[nt][3] * kScaleLog2 - mn1);
  }
  acc.l[0] = acc.l[0] * al0 + (p[0][0] + p[0][1] + p[1][0] + p[1][1]);
  acc.l[1] = acc.l[1] * al1 + (p[0][2] + p[0][3] + p[1][2] + p[1][3]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc.o[i][0] *= al0;
    acc.o[i][1] *= al0;
    acc.o[i][2] *= al1;
    acc.o[i][3] *= al1;
  }
  // ---- O += P V : A = P (16 q x 16 keys) from the S fragments, B = V via ldmatrix.trans
  uint32_t pa[4];
  pa[0] = Cvt<T>::pack2(p[0][0], p[0][1]);
  pa[1] = Cvt<T>::pack2(p[0][2], p[0][3]);
  pa[2] = Cvt<T>::pack2(p[1][0], p[1][1]);
  pa[3] = Cvt<T>::pack2(p[1][2], p[1][3]);
  {
    const int m = lane >> 3;
    const int vrow = wrow0 + (m & 1) * 8 + (lane & 7);
#pragma unroll
    for (int np = 0; np < 4; ++np) {               // pairs of dh n-tiles
      const int chunk = np * 2 + (m >> 1);
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, sV + vrow * 128 + ((chunk ^ (vrow & 7)) << 4));
      mma16816<T>(acc.o[np * 2], pa, vb[0], vb[1]);
      mma16816<T>(acc.o[np * 2 + 1], pa, vb[2], vb[3]);
    }
  }
}

// Q fragments (A operand, 16 queries x 64 dh) straight from global memory; rows >= n_q are zero.
template <typename T>
__device__ __forceinline__ void load_q_frags(uint32_t (&qa)[4][4], const T* q0, long long ldq, int n_q) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int c = ks * 16 + 2 * t;
    qa[ks][0] = g < n_q ? *reinterpret_cast<const uint32_t*>(q0 + g * ldq + c) : 0u;
    qa[ks][1] = g + 8 < n_q ? *reinterpret_cast<const uint32_t*>(q0 + (g + 8) * ldq + c) : 0u;
    qa[ks][2] = g < n_q ? *reinterpret_cast<const uint32_t*>(q0 + g * ldq + c + 8) : 0u;
    qa[ks][3] = g + 8 < n_q ? *reinterpret_cast<const uint32_t*>(q0 + (g + 8) * ldq + c + 8) : 0u;
  }
}

// Merge the four warps' (m, l, O) through shared memory into warp 0's registers.
// red: float[4][32][36+]; uses 4*32*40 floats.
__device__ __forceinline__ void cta_merge(WarpAcc& acc, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // finish the row sums inside each quad first
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    acc.l[r] += __shfl_xor_sync(0xffffffffu, acc.l[r], 1);
    acc.l[r] += __shfl_xor_sync(0xffffffffu, acc.l[r], 2);
  }
  float* mine = red + (warp * 32 + lane) * 40;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) mine[i * 4 + j] = acc.o[i][j];
  mine[32] = acc.m[0];
  mine[33] = acc.m[1];
  mine[34] = acc.l[0];
  mine[35] = acc.l[1];
  __syncthreads();
  if (warp == 0) {
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      m0 = fmaxf(m0, red[(w * 32 + lane) * 40 + 32]);
      m1 = fmaxf(m1, red[(w * 32 + lane) * 40 + 33]);
    }
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc.o[i][j] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float* o = red + (w * 32 + lane) * 40;
      const float f0 = o[32] == -INFINITY ? 0.f : fast_exp2(o[32] - m0);
      const float f1 = o[33] == -INFINITY ? 0.f : fast_exp2(o[33] - m1);
      l0 += o[34] * f0;
      l1 += o[35] * f1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc.o[i][0] += o[i * 4 + 0] * f0;
        acc.o[i][1] += o[i * 4 + 1] * f0;
        acc.o[i][2] += o[i * 4 + 2] * f1;
        acc.o[i][3] += o[i * 4 + 3] * f1;
      }
    }
    acc.m[0] = m0;
    acc.m[1] = m1;
    acc.l[0] = l0;
    acc.l[1] = l1;
  }
}

// Stage one 64-key tile: thread i copies 16-byte chunk (i % 8) of rows (i / 8) + {0,16,32,48}.
template <typename F>
__device__ __forceinline__ void stage_tile(uint8_t* sK, uint8_t* sV, int key0, int kv_len, F src_of) {
  const int tid = threadIdx.x;
  const int chunk = tid & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = (tid >> 3) + i * 16;
    const int key = key0 + row;
    const bool ok = key < kv_len;
    RowSrc s = src_of(ok ? key : 0);
    const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
    cp_async16_zfill(sK + off, s.k + chunk * 16, ok);
    cp_async16_zfill(sV + off, s.v + chunk * 16, ok);
  }
}

// Shared main loop: stream keys [key_begin, key_end) in 64-key tiles through the cp.async ring.
template <typename T, int STAGES, typename F>
__device__ __forceinline__ void stream_keys(uint8_t* smem, int key_begin, int key_end, int kv_len,
                                            const uint32_t (&qa)[4][4], WarpAcc& acc, F src_of) {
  const int warp = threadIdx.x >> 5;
  const int n_tiles = (key_end - key_begin + kDaTileKeys - 1) / kDaTileKeys;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n_tiles)
      stage_tile(smem + s * 2 * kDaTileBytes, smem + s * 2 * kDaTileBytes + kDaTileBytes,
                 key_begin + s * kDaTileKeys, min(kv_len, key_end), src_of);
    cp_async_commit();
  }
  for (int it = 0; it < n_tiles; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nx = it + STAGES - 1;
    if (nx < n_tiles) {
      const int st = nx % STAGES;
      stage_tile(smem + st * 2 * kDaTileBytes, smem + st * 2 * kDaTileBytes + kDaTileBytes,
                 key_begin + nx * kDaTileKeys, min(kv_len, key_end), src_of);
    }
    cp_async_commit();
    const int st = it % STAGES;
    const int k0 = key_begin + it * kDaTileKeys + warp * 16;
    warp_tile<T>(smem + st * 2 * kDaTileBytes, smem + st * 2 * kDaTileBytes + kDaTileBytes, warp * 16,
                 qa, min(kv_len, key_end) - k0, acc);
  }
  cp_async_wait<0>();
  __syncthreads();
}

// =================================================================================================
// cross attention
// =================================================================================================
struct CrossParams {
  const void* q;        // [n_audio * n_q, d] queries (already projected)
  const void* k;        // [n_audio, T, d]
  const void* v;        // [n_audio, T, d]
  void* out;            // [n_audio * n_q, d]
  float* partial;       // [n_audio * q_tiles, H, splits, 16, 66]
  int* counters;        // [n_audio * q_tiles * H], zero on entry, zero on exit
  const int* skip_flag; // optional device flag: non-zero -> kernel does nothing
  int n_q, q_tiles, T, d, splits, keys_per_split, kv_ld;
  int head_major;       // k = one layer's [n_audio][2H][T][64] block (K heads, then V heads); v unused
};

template <typename T, int STAGES>
__global__ void __launch_bounds__(kDaThreads) cross_attention_kernel(const CrossParams p) {
  pdl_launch_dependents();
  pdl_wait();
  if (p.skip_flag && *p.skip_flag) return;
  extern __shared__ __align__(1024) uint8_t da_smem[];
  __shared__ int s_last;
  const int split = blockIdx.x, h = blockIdx.y;
  const int audio = blockIdx.z / p.q_tiles, qt = blockIdx.z % p.q_tiles;
  const int q_first = qt * 16;
  const int n_q = min(16, p.n_q - q_first);
  const T* q0 = reinterpret_cast<const T*>(p.q) + (static_cast<long long>(audio) * p.n_q + q_first) * p.d + h * 64;
  uint32_t qa[4][4];
  load_q_frags<T>(qa, q0, p.d, n_q);

  const uint8_t* kbase = reinterpret_cast<const uint8_t*>(p.k) + (static_cast<long long>(audio) * p.T * p.kv_ld + h * 64) * 2;
  const uint8_t* vbase = reinterpret_cast<const uint8_t*>(p.v) + (static_cast<long long>(audio) * p.T * p.kv_ld + h * 64) * 2;
  long long row_bytes = static_cast<long long>(p.kv_ld) * 2;
  if (p.head_major) {       // each (audio, head) streams one contiguous T x 128-byte block
    const long long H = gridDim.y, hb = static_cast<long long>(p.T) * 128;
    kbase = reinterpret_cast<const uint8_t*>(p.k) + (static_cast<long long>(audio) * 2 * H + h) * hb;
    vbase = reinterpret_cast<const uint8_t*>(p.k) + (static_cast<long long>(audio) * 2 * H + H + h) * hb;
    row_bytes = 128;
  }
  auto src_of = [&](int key) {
    RowSrc s;
    s.k = kbase + key * row_bytes;
    s.v = vbase + key * row_bytes;
    return s;
  };
  WarpAcc acc;
  acc_init(acc);
  const int kb = split * p.keys_per_split;
  const int ke = min(p.T, kb + p.keys_per_split);
  stream_keys<T, STAGES>(da_smem, kb, ke, p.T, qa, acc, src_of);

  float* red = reinterpret_cast<float*>(da_smem);
  cta_merge(acc, red);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long tile_id = (static_cast<long long>(blockIdx.z) * gridDim.y + h);
  float* part = p.partial + (tile_id * p.splits + split) * (16 * 66);
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      part[g * 66 + i * 8 + 2 * t] = acc.o[i][0];
      part[g * 66 + i * 8 + 2 * t + 1] = acc.o[i][1];
      part[(g + 8) * 66 + i * 8 + 2 * t] = acc.o[i][2];
      part[(g + 8) * 66 + i * 8 + 2 * t + 1] = acc.o[i][3];
    }
    if (t == 0) {
      part[g * 66 + 64] = acc.m[0];
      part[g * 66 + 65] = acc.l[0];
      part[(g + 8) * 66 + 64] = acc.m[1];
      part[(g + 8) * 66 + 65] = acc.l[1];
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(&p.counters[tile_id], 1);
    s_last = (prev == p.splits - 1);
    if (s_last) p.counters[tile_id] = 0;          // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- combine the splits: 128 threads = 16 rows x 8 column groups of 8
  const int row = threadIdx.x >> 3, cg = threadIdx.x & 7;
  if (row < n_q) {
    const float* base = p.partial + tile_id * p.splits * (16 * 66) + row * 66;
    float m = -INFINITY;
    for (int s = 0; s < p.splits; ++s) m = fmaxf(m, __ldcg(base + s * 16 * 66 + 64));
    float l = 0.f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float* ps = base + s * 16 * 66;
      const float ms = __ldcg(ps + 64);
      const float f = ms == -INFINITY ? 0.f : fast_exp2(ms - m);
      l += __ldcg(ps + 65) * f;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += __ldcg(ps + cg * 8 + e) * f;
    }
    const float inv = 1.0f / l;
    T* orow = reinterpret_cast<T*>(p.out) + (static_cast<long long>(audio) * p.n_q + q_first + row) * p.d + h * 64 + cg * 8;
    uint4 u;
    u.x = Cvt<T>::pack2(o[0] * inv, o[1] * inv);
    u.y = Cvt<T>::pack2(o[2] * inv, o[3] * inv);
    u.z = Cvt<T>::pack2(o[4] * inv, o[5] * inv);
    u.w = Cvt<T>::pack2(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(orow) = u;
  }
}

// =================================================================================================
// cross attention, step mode, TMA-fed and persistent (head-major K/V only)
// =================================================================================================
// The cp.async kernel above plateaus at ~5.2 TB/s whatever its ring depth or key split (profiles/r1_xattn_sweep.txt)
// while a read-only stream reaches 7.3 TB/s on the same part (tools/microbench.cu): each of its 2560 short-lived
// CTAs pays its own pipeline fill, query load, 4-warp merge and drain for only 192 KB of K/V.  Here ONE CTA per SM
// stays resident and walks through its share of the (audio, head[, key-split]) items with the memory pipeline kept full
// ACROSS items, three decoupled roles talking through mbarriers only:
//   producer warp : streams 128-key K and V tiles (one 16 KB TMA box each, 128-byte swizzle = the layout ldmatrix
//                   wants) and the next item's 16-row query tile, never waiting for anything but free ring slots
//   8 consumer warps : 16 keys of every tile each, the same mma.sync online-softmax tile code as above; at the end
//                   of an item they drop their (m, l, O) fragments into one of two exchange buffers and move on
//   epilogue warp : merges the eight fragments, normalises and stores the output (or, with key splits, writes the
//                   partial and runs the last-arriver combine) while the consumers are already in the next item.
// (A first version that let consumer warp 0 do the merge, fence and ticket while the other seven waited at a CTA barrier
// ran at 171 us against the cp.async kernel's 94: nothing else on the SM overlapped that tail.)
constexpr int kX2Consumers = 8;
constexpr int kX2Threads = (kX2Consumers + 2) * 32;
constexpr int kX2TileKeys = 128;
constexpr int kX2Stages = 4;
constexpr int kX2TileBytes = kX2TileKeys * 128;                  // one K (or V) tile
constexpr int kX2StageBytes = 2 * kX2TileBytes;
constexpr int kX2QBytes = 16 * 128;                              // 16 query rows x 64 dims
constexpr int kX2RedFloats = kX2Consumers * 32 * 36;             // per buffer: 8 warps x 32 lanes x (32 O + m0 m1 l0 l1)
constexpr int kX2SmemBytes = kX2Stages * kX2StageBytes + 2 * kX2QBytes + 2 * kX2RedFloats * 4 + 256 + 1024;

struct Cross2Params {
  void* out;            // [n_audio * n_q, d]
  float* partial;       // [n_audio * H, splits, 16, 66]
  int* counters;        // [n_audio * H], zero on entry, zero on exit
  const int* skip_flag;
  int n_q, n_head, T, d, splits, tiles_per_split, total_tiles, total_items;
};

__device__ __forceinline__ void x2_wait(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();                                                      // protocol bug: fail loudly instead of hanging the GPU
}

template <typename T>
__global__ void __launch_bounds__(kX2Threads, 1)
cross_attention_tma_kernel(const Cross2Params p, const __grid_constant__ CUtensorMap mapQ,
                           const __grid_constant__ CUtensorMap mapKV) {
  pdl_launch_dependents();
  extern __shared__ uint8_t x2_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(x2_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ring = smem;
  uint8_t* sQ = ring + kX2Stages * kX2StageBytes;
  float* red = reinterpret_cast<float*>(sQ + 2 * kX2QBytes);     // [2][kX2RedFloats]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * kX2RedFloats);
  uint64_t* empty_bar = full_bar + kX2Stages;
  uint64_t* qfull_bar = empty_bar + kX2Stages;
  uint64_t* qempty_bar = qfull_bar + 2;
  uint64_t* rfull_bar = qempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kX2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kX2Consumers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qfull_bar[i], 1);
      mbar_init(&qempty_bar[i], kX2Consumers);
      mbar_init(&rfull_bar[i], kX2Consumers);
      mbar_init(&rempty_bar[i], 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapKV);
  }
  __syncthreads();
  pdl_wait();
  if (p.skip_flag && *p.skip_flag) return;
  const int H = p.n_head;

  if (warp == kX2Consumers) {
    // ===================== producer =====================
    if (lane == 0) {
      int q = 0, iq = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iq) {
        const int split = item % p.splits, ah = item / p.splits;
        const int head = ah % H, audio = ah / H;
        const int qs = iq & 1;
        x2_wait(&qempty_bar[qs], ((iq >> 1) & 1) ^ 1);
        mbar_expect_tx(&qfull_bar[qs], kX2QBytes);
        tma_load_2d(sQ + qs * kX2QBytes, &mapQ, &qfull_bar[qs], head * 64, audio * p.n_q);
        const int t0 = split * p.tiles_per_split;
        const int nt = min(p.tiles_per_split, p.total_tiles - t0);
        for (int t = 0; t < nt; ++t, ++q) {
          const int st = q % kX2Stages;
          x2_wait(&empty_bar[st], ((q / kX2Stages) & 1) ^ 1);
          mbar_expect_tx(&full_bar[st], kX2StageBytes);
          uint8_t* dst = ring + st * kX2StageBytes;
          tma_load_3d(dst, &mapKV, &full_bar[st], 0, (t0 + t) * kX2TileKeys, audio * 2 * H + head);
          tma_load_3d(dst + kX2TileBytes, &mapKV, &full_bar[st], 0, (t0 + t) * kX2TileKeys, audio * 2 * H + H + head);
        }
      }
    }
    return;
  }
  if (warp == kX2Consumers + 1) {
    // ===================== epilogue warp =====================
    const int g = lane >> 2, t4 = lane & 3;
    int iq = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iq) {
      const int split = item % p.splits, ah = item / p.splits;
      const int head = ah % H, audio = ah / H;
      const int rb = iq & 1;
      const float* rbuf = red + rb * kX2RedFloats;
      x2_wait(&rfull_bar[rb], (iq >> 1) & 1);
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int w = 0; w < kX2Consumers; ++w) {
        const float4 ml = *reinterpret_cast<const float4*>(rbuf + (w * 32 + lane) * 36 + 32);
        m0 = fmaxf(m0, ml.x);
        m1 = fmaxf(m1, ml.y);
      }
      float l0 = 0.f, l1 = 0.f, o[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
#pragma unroll
      for (int w = 0; w < kX2Consumers; ++w) {
        const float* src = rbuf + (w * 32 + lane) * 36;
        const float4 ml = *reinterpret_cast<const float4*>(src + 32);
        const float f0 = ml.x == -INFINITY ? 0.f : fast_exp2(ml.x - m0);
        const float f1 = ml.y == -INFINITY ? 0.f : fast_exp2(ml.y - m1);
        l0 += ml.z * f0;
        l1 += ml.w * f1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
          o[i][0] += v.x * f0;
          o[i][1] += v.y * f0;
          o[i][2] += v.z * f1;
          o[i][3] += v.w * f1;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rempty_bar[rb]);               // the consumers may refill this buffer
      if (p.splits == 1) {
        // the whole key range was here: normalise and store rows g and g + 8 of this (audio, head)
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        T* o0 = reinterpret_cast<T*>(p.out) + (static_cast<long long>(audio) * p.n_q + g) * p.d + head * 64 + 2 * t4;
        T* o1 = o0 + 8LL * p.d;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (g < p.n_q) *reinterpret_cast<uint32_t*>(o0 + i * 8) = Cvt<T>::pack2(o[i][0] * i0, o[i][1] * i0);
          if (g + 8 < p.n_q) *reinterpret_cast<uint32_t*>(o1 + i * 8) = Cvt<T>::pack2(o[i][2] * i1, o[i][3] * i1);
        }
        continue;
      }
      const long long tile_id = ah;
      float* part = p.partial + (tile_id * p.splits + split) * (16 * 66);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        *reinterpret_cast<float2*>(part + g * 66 + i * 8 + 2 * t4) = make_float2(o[i][0], o[i][1]);
        *reinterpret_cast<float2*>(part + (g + 8) * 66 + i * 8 + 2 * t4) = make_float2(o[i][2], o[i][3]);
      }
      if (t4 == 0) {
        *reinterpret_cast<float2*>(part + g * 66 + 64) = make_float2(m0, l0);
        *reinterpret_cast<float2*>(part + (g + 8) * 66 + 64) = make_float2(m1, l1);
      }
      __threadfence();
      __syncwarp();
      int last = 0;
      if (lane == 0) {
        const int prev = atomicAdd(&p.counters[tile_id], 1);
        last = (prev == p.splits - 1);
        if (last) p.counters[tile_id] = 0;          // ready for the next launch
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) continue;
      __threadfence();
      // ---- combine the splits: lane = (row = lane / 2, column half = lane % 2)
      const int row = lane >> 1, c0 = (lane & 1) * 32;
      if (row < p.n_q) {
        const float* base = p.partial + tile_id * p.splits * (16 * 66) + row * 66;
        float m = -INFINITY;
        for (int sp = 0; sp < p.splits; ++sp) m = fmaxf(m, __ldcg(base + sp * 16 * 66 + 64));
        float l = 0.f, acc[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[e] = 0.f;
        for (int sp = 0; sp < p.splits; ++sp) {
          const float* ps = base + sp * 16 * 66;
          const float ms = __ldcg(ps + 64);
          const float f = ms == -INFINITY ? 0.f : fast_exp2(ms - m);
          l += __ldcg(ps + 65) * f;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(ps + c0 + e));
            acc[e] += v.x * f;
            acc[e + 1] += v.y * f;
          }
        }
        const float inv = 1.0f / l;
        T* orow = reinterpret_cast<T*>(p.out) + (static_cast<long long>(audio) * p.n_q + row) * p.d + head * 64 + c0;
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 u;
          u.x = Cvt<T>::pack2(acc[e] * inv, acc[e + 1] * inv);
          u.y = Cvt<T>::pack2(acc[e + 2] * inv, acc[e + 3] * inv);
          u.z = Cvt<T>::pack2(acc[e + 4] * inv, acc[e + 5] * inv);
          u.w = Cvt<T>::pack2(acc[e + 6] * inv, acc[e + 7] * inv);
          *reinterpret_cast<uint4*>(orow + e) = u;
        }
      }
    }
    return;
  }
  // ===================== consumers =====================
  int q = 0, iq = 0;
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iq) {
    const int split = item % p.splits;
    const int qs = iq & 1;
    uint32_t qa[4][4];
    x2_wait(&qfull_bar[qs], (iq >> 1) & 1);
    {
      const uint8_t* sq = sQ + qs * kX2QBytes;
      const int row = lane & 15;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(qa[ks], sq + row * 128 + ((chunk ^ (row & 7)) << 4));
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&qempty_bar[qs]);
    WarpAcc acc;
    acc_init(acc);
    const int t0 = split * p.tiles_per_split;
    const int nt = min(p.tiles_per_split, p.total_tiles - t0);
    for (int t = 0; t < nt; ++t, ++q) {
      const int st = q % kX2Stages;
      x2_wait(&full_bar[st], (q / kX2Stages) & 1);
      const uint8_t* sK = ring + st * kX2StageBytes;
      const int key0 = (t0 + t) * kX2TileKeys + warp * 16;
      warp_tile<T>(sK, sK + kX2TileBytes, warp * 16, qa, p.T - key0, acc);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
    // ---- hand the fragment to the epilogue warp
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      acc.l[r] += __shfl_xor_sync(0xffffffffu, acc.l[r], 1);
      acc.l[r] += __shfl_xor_sync(0xffffffffu, acc.l[r], 2);
    }
    const int rb = iq & 1;
    x2_wait(&rempty_bar[rb], ((iq >> 1) & 1) ^ 1);
    float* mine = red + rb * kX2RedFloats + (warp * 32 + lane) * 36;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(mine + i * 4) = make_float4(acc.o[i][0], acc.o[i][1], acc.o[i][2], acc.o[i][3]);
    *reinterpret_cast<float4*>(mine + 32) = make_float4(acc.m[0], acc.m[1], acc.l[0], acc.l[1]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&rfull_bar[rb]);
  }
}

// =================================================================================================
// self attention (+ kv-cache append)
// =================================================================================================
struct SelfParams {
  const void* qkv;      // [n_rows, 3*d]: q | k | v of the NEW position(s)
  void* kcache;         // [phys_rows, max_ctx, d]
  void* vcache;
  void* out;            // [n_rows, d]
  const int* indir;     // step mode: [R, max_ctx] position -> physical row; null in prefill mode
  const int* len_ptr;   // step mode: device int, current token count L (new token is position L-1)
  const int* skip_flag;
  int d, max_ctx;
  int n_init;           // prefill mode: tokens per audio (query row = audio * n_init + i)
  int group;            // prefill mode: physical row of audio a is a * group
  int head_major;       // 0: [phys_row][max_ctx][d]; 1: [phys_row][head][max_ctx][64]; 2: beam window [audio][head][max_ctx][slot][64]
};

// One WARP per (row, head): with a single query there is nothing for a tensor-core tile to share, so
// the kernel is a pure streaming reduction.  Lane = (key sub-index g = lane / 8, 16-byte chunk
// c = lane % 8): every copy instruction fetches four complete 128-byte K (or V) rows, fully
// coalesced; the 64-dim dot product is finished with three xor-shuffles inside the 8-lane group, and
// each group keeps its own online-softmax state (m, l, o[8]) - rescaled once per block of U keys -
// that is merged across the four groups once at the end.
//
// The kernel is latency-bound (short sequences, a dependent indirection lookup in front of every
// K/V fetch), so everything is software-pipelined: the in-flight K/V blocks live in a per-lane
// cp.async ring in shared memory (every lane copies and later re-reads only its own 16-byte slots, so
// no barrier of any kind is needed - cp.async.wait_group orders a thread's own copies), ST-1 blocks
// of 4*U keys are outstanding per warp at ~60 registers, and the position -> physical-row lookups
// run one block further ahead.  Measured on the C3 decode (profiles/r1_selfattn_sweep*.txt): the
// first version (register buffers, serial lookup -> load -> math, 16 warps/SM) averaged 81 us per
// launch, register double-buffering 57 us, this ring with (U, ST, warps/CTA) = (2, 4, 4) 41 us;
// deeper rings or larger CTAs lose more in occupancy than they gain in bytes in flight.
template <typename T, int U, int ST>
__global__ void __launch_bounds__(640) self_attention_kernel(const SelfParams p, int n_rows, int n_head) {
  extern __shared__ __align__(16) uint8_t sa_smem[];
  pdl_launch_dependents();
  pdl_wait();
  if (p.skip_flag && *p.skip_flag) return;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int g = lane >> 3, c = lane & 7;
  const int w = blockIdx.x * (blockDim.x >> 5) + wi;
  if (w >= n_rows * n_head) return;
  const int row = w / n_head, h = w % n_head;
  const bool step = p.indir != nullptr;
  int kv_len, phys_fixed = 0;
  if (step) {
    kv_len = *p.len_ptr;
  } else {
    kv_len = row % p.n_init + 1;
    phys_fixed = (row / p.n_init) * p.group;
  }
  const int pos_new = step ? kv_len - 1 : kv_len;       // prefill: every key is already in the cache
  const long long row_bytes = static_cast<long long>(p.d) * 2;
  const uint8_t* qrow = reinterpret_cast<const uint8_t*>(p.qkv) + static_cast<long long>(row) * 3 * row_bytes + h * 128;
  const uint8_t* knew = qrow + row_bytes + c * 16;
  const uint8_t* vnew = qrow + 2 * row_bytes + c * 16;
  // byte offset of (physical row, position) for this head.  Row-major: [row][pos][d].  Head-major: [row][head][pos][64].
  // Beam window: [audio][head][pos][beam slot][64] - the G rows of an audio interleaved per position, so that the whole
  // history of an (audio, head) is ONE contiguous block (what self_attention_tma_kernel streams); physical row =
  // audio * G + slot.
  const bool window = p.head_major == 2;
  const int Gw = window ? p.group : 1;
  const long long pos_bytes = p.head_major ? 128LL * Gw : row_bytes;
  const long long row_stride = static_cast<long long>(p.max_ctx) * row_bytes;      // bytes per physical row, either layout
  const long long head_off = p.head_major ? static_cast<long long>(h) * p.max_ctx * 128 * Gw : static_cast<long long>(h) * 128;
  auto phys_off = [&](int ph) -> long long {       // offset of position 0 of physical row ph (before head_off)
    // (the division sits in the dependent address chain of every key: only the window layout pays for it)
    return window ? static_cast<long long>(ph / Gw) * (row_stride * Gw) + static_cast<long long>(ph % Gw) * 128
                  : static_cast<long long>(ph) * row_stride;
  };
  uint8_t* kc = reinterpret_cast<uint8_t*>(p.kcache) + head_off + c * 16;
  uint8_t* vc = reinterpret_cast<uint8_t*>(p.vcache) + head_off + c * 16;
  if (step && lane < 16) {
    const long long off = phys_off(row) + pos_new * pos_bytes;
    if (lane < 8)
      *reinterpret_cast<uint4*>(kc + off) = *reinterpret_cast<const uint4*>(knew);
    else
      *reinterpret_cast<uint4*>(vc + off) = *reinterpret_cast<const uint4*>(vnew);
  }
  float q[8];
  {
    const uint4 u = *reinterpret_cast<const uint4*>(qrow + c * 16);
    const uint32_t wq[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = Cvt<T>::unpack2(wq[e]);
      q[2 * e] = f.x * kScaleLog2;
      q[2 * e + 1] = f.y * kScaleLog2;
    }
  }
  const int* ind = step ? p.indir + static_cast<long long>(row) * p.max_ctx : nullptr;
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;

  constexpr int KPB = 4 * U;                              // keys per block (= per ring stage)
  // ring layout: [warp][stage][j][k|v][lane] x 16 B
  uint8_t* ring = sa_smem + (static_cast<size_t>(wi) * ST * U * 2 * 32 + lane) * 16;
  auto load_phys = [&](int (&ph)[U], int b) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int key = b + j * 4 + g;
      ph[j] = (step && key < pos_new) ? __ldg(ind + key) : phys_fixed;
    }
  };
  auto issue = [&](int stage, const int (&ph)[U], int b) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int key = b + j * 4 + g;
      const bool valid = key < kv_len;
      const uint8_t *ks = knew, *vs = vnew;               // new token (or dummy address of a masked key)
      if (key < pos_new) {
        const long long off = phys_off(ph[j]) + key * pos_bytes;
        ks = kc + off;
        vs = vc + off;
      }
      uint8_t* dst = ring + ((stage * U + j) * 2) * 512;
      cp_async16_zfill(dst, ks, valid);
      cp_async16_zfill(dst + 512, vs, valid);
    }
    cp_async_commit();
  };
  auto consume = [&](int stage, int b) {
    float sd[U];
    float mx = m;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const uint4 kk = *reinterpret_cast<const uint4*>(ring + ((stage * U + j) * 2) * 512);
      const uint32_t wk[4] = {kk.x, kk.y, kk.z, kk.w};
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = Cvt<T>::unpack2(wk[e]);
        acc = fmaf(q[2 * e], f.x, acc);
        acc = fmaf(q[2 * e + 1], f.y, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      sd[j] = (b + j * 4 + g) < kv_len ? acc : -INFINITY;
      mx = fmaxf(mx, sd[j]);
    }
    if (mx == -INFINITY) return;                          // this lane group has not seen a valid key yet
    const float al = fast_exp2(m - mx);                   // m = -inf -> 0
    m = mx;
    l *= al;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] *= al;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const float pr = fast_exp2(sd[j] - mx);             // masked key: exp2(-inf) = 0, and its V slot is zero-filled
      l += pr;
      const uint4 vv = *reinterpret_cast<const uint4*>(ring + ((stage * U + j) * 2 + 1) * 512);
      const uint32_t wv[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = Cvt<T>::unpack2(wv[e]);
        o[2 * e] = fmaf(pr, f.x, o[2 * e]);
        o[2 * e + 1] = fmaf(pr, f.y, o[2 * e + 1]);
      }
    }
  };

  int ph[U];
#pragma unroll
  for (int st = 0; st < ST - 1; ++st) {
    load_phys(ph, st * KPB);
    issue(st, ph, st * KPB);
  }
  load_phys(ph, (ST - 1) * KPB);
  int cs = 0, is = ST - 1;                                 // consume / issue stage
  for (int b = 0; b < kv_len; b += KPB) {
    cp_async_wait<ST - 2>();                               // block b has landed
    issue(is, ph, b + (ST - 1) * KPB);
    load_phys(ph, b + ST * KPB);
    consume(cs, b);
    cs = cs + 1 == ST ? 0 : cs + 1;
    is = is + 1 == ST ? 0 : is + 1;
  }
  cp_async_wait<0>();
  // merge the four key groups (lanes differing in bits 3 and 4)
#pragma unroll
  for (int sh = 8; sh <= 16; sh <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, sh);
    const float l2 = __shfl_xor_sync(0xffffffffu, l, sh);
    const float mn = fmaxf(m, m2);
    const float a1 = m == -INFINITY ? 0.f : fast_exp2(m - mn);
    const float a2 = m2 == -INFINITY ? 0.f : fast_exp2(m2 - mn);
    l = l * a1 + l2 * a2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[e], sh);
      o[e] = o[e] * a1 + o2 * a2;
    }
    m = mn;
  }
  if (lane < 8) {
    const float inv = 1.0f / l;
    uint4 u;
    u.x = Cvt<T>::pack2(o[0] * inv, o[1] * inv);
    u.y = Cvt<T>::pack2(o[2] * inv, o[3] * inv);
    u.z = Cvt<T>::pack2(o[4] * inv, o[5] * inv);
    u.w = Cvt<T>::pack2(o[6] * inv, o[7] * inv);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + static_cast<long long>(row) * row_bytes + h * 128 + c * 16) = u;
  }
}

template <typename T, int U, int ST>
static int launch_sa(const SelfParams& p, int n_rows, int n_head, int wpc, cudaStream_t s) {
  auto kern = self_attention_kernel<T, U, ST>;
  const int smem = wpc * ST * U * 1024;
  static SmemOptIn optin;
  if (!optin.ensure(kern, smem)) return 44;
  const int n_pairs = n_rows * n_head;
  return launch_pdl(kern, dim3((n_pairs + wpc - 1) / wpc), dim3(wpc * 32), smem, s, p, n_rows, n_head) == cudaSuccess ? 0 : 43;
}

template <typename T>
static int dispatch_sa(const SelfParams& p, int n_rows, int n_head, int u, int st, int wpc, cudaStream_t s) {
  if (u == 2 && st == 4) return launch_sa<T, 2, 4>(p, n_rows, n_head, wpc, s);
  if (u == 2 && st == 3) return launch_sa<T, 2, 3>(p, n_rows, n_head, wpc, s);
  if (u == 2 && st == 5) return launch_sa<T, 2, 5>(p, n_rows, n_head, wpc, s);
  if (u == 1 && st == 6) return launch_sa<T, 1, 6>(p, n_rows, n_head, wpc, s);
  if (u == 1 && st == 8) return launch_sa<T, 1, 8>(p, n_rows, n_head, wpc, s);
  return 45;
}

// =================================================================================================
// self attention, step mode, beams of an audio together, TMA-fed and persistent (beam-window layout only)
// =================================================================================================
// With beam search the G rows of an audio attend over histories that are piecewise the same physical cache rows, and
// after a few reorders the physical row of consecutive positions of one lineage is effectively random: the warp-per-
// (row, head) kernel above then gathers isolated 128-byte pieces (measured at the mean length of the headline run:
// 148 MB of DRAM traffic per launch at 3.85 TB/s, 12 % L2 hits, profiles/r2_self_attn_midL_ncu_full_selected.csv).
// In the beam-window layout [audio][head][pos][slot][64] the union of those histories is ONE contiguous block of
// (length x G) rows per (audio, head), so this kernel does what the cross-attention kernel does - producer warp streams
// 128-row K / V tiles by TMA through a ring that stays full across items, eight consumer warps run mma.sync tiles
// with the G beams as the M rows, an epilogue warp merges and stores - plus a mask: beam b attends row (pos, slot) iff
// its parent table says indir[b][pos] == slot (the reference gathers the caches instead, decoding.py:172-176).  The new
// token's K / V come straight from the QKV projection's output (one 16-row tile per item), are used from shared
// memory and appended to the cache by a consumer warp (the torch.cat of model.py:327-333).
constexpr int kS2Consumers = 8;
constexpr int kS2Threads = (kS2Consumers + 2) * 32;
constexpr int kS2TileRows = 128;
constexpr int kS2Stages = 4;
constexpr int kS2TileBytes = kS2TileRows * 128;
constexpr int kS2StageBytes = 2 * kS2TileBytes;
constexpr int kS2MaxG = 8;
constexpr int kS2IndBytes = 448 * 4;                              // one beam's parent table (n_text_ctx <= 448)
constexpr int kS2MetaBytes = 3 * kX2QBytes + kS2MaxG * kS2IndBytes;   // Q | Knew | Vnew tiles + G parent tables
constexpr int kS2RedFloats = kS2Consumers * 32 * 20;       // per lane: o of its beam row (16) | m | l | pad
constexpr int kS2SmemBytes = kS2Stages * kS2StageBytes + 2 * kS2MetaBytes + 2 * kS2RedFloats * 4 + 256 + 1024;

struct Self2Params {
  void* out;              // [R, d]
  void* kcache;           // this layer's K block [n_audio][H][ctx][G][64]
  void* vcache;
  const int* indir;       // [R, ctx]
  const int* len_ptr;
  const int* skip_flag;
  int n_audio, n_head, G, ctx, d;
  int g_magic;            // ceil(2^16 / G): (r * g_magic) >> 16 == r / G for every cache row index r < 9362
};

// 16 keys of a staged tile for a query tile whose rows 8..15 are padding (G <= 8 beams live in rows 0..7): only the
// row-g half of every fragment carries scores, the other half is fed zero probabilities.  kmask[j] is the set of beams
// (bit b) that attend this lane's j-th key (keys 2t, 2t + 1, 8 + 2t, 9 + 2t of the slice); safe when a beam has no
// valid key in the slice (its running maximum stays -inf and everything stays zero).
template <typename T>
__device__ __forceinline__ void warp_tile_beams(const uint8_t* sK, const uint8_t* sV, int wrow0, const uint32_t (&qa)[4][4],
                                                WarpAcc& acc, const uint32_t (&kmask)[4]) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;
  float s[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
  {
    const int m = lane >> 3;
    const int krow = wrow0 + (m >> 1) * 8 + (lane & 7);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int chunk = ks * 2 + (m & 1);
      uint32_t kb[4];
      ldmatrix_x4(kb, sK + krow * 128 + ((chunk ^ (krow & 7)) << 4));
      mma16816<T>(s[0], qa[ks], kb[0], kb[1]);
      mma16816<T>(s[1], qa[ks], kb[2], kb[3]);
    }
  }
  float sc[4];
  sc[0] = ((kmask[0] >> g) & 1u) ? s[0][0] * kScaleLog2 : -INFINITY;
  sc[1] = ((kmask[1] >> g) & 1u) ? s[0][1] * kScaleLog2 : -INFINITY;
  sc[2] = ((kmask[2] >> g) & 1u) ? s[1][0] * kScaleLog2 : -INFINITY;
  sc[3] = ((kmask[3] >> g) & 1u) ? s[1][1] * kScaleLog2 : -INFINITY;
  float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  const float mn = fmaxf(acc.m[0], mx);
  const float b0 = mn == -INFINITY ? 0.f : mn;         // a row that has seen no key yet: keep everything at zero
  const float al = fast_exp2(acc.m[0] - b0);
  acc.m[0] = mn;
  const float p0 = fast_exp2(sc[0] - b0), p1 = fast_exp2(sc[1] - b0), p2 = fast_exp2(sc[2] - b0), p3 = fast_exp2(sc[3] - b0);
  acc.l[0] = acc.l[0] * al + ((p0 + p1) + (p2 + p3));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc.o[i][0] *= al;
    acc.o[i][1] *= al;
  }
  uint32_t pa[4];
  pa[0] = Cvt<T>::pack2(p0, p1);
  pa[1] = 0u;
  pa[2] = Cvt<T>::pack2(p2, p3);
  pa[3] = 0u;
  {
    const int m = lane >> 3;
    const int vrow = wrow0 + (m & 1) * 8 + (lane & 7);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int chunk = np * 2 + (m >> 1);
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, sV + vrow * 128 + ((chunk ^ (vrow & 7)) << 4));
      mma16816<T>(acc.o[np * 2], pa, vb[0], vb[1]);
      mma16816<T>(acc.o[np * 2 + 1], pa, vb[2], vb[3]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kS2Threads, 1)
self_attention_tma_kernel(const Self2Params p, const __grid_constant__ CUtensorMap mapQKV,
                          const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV) {
  pdl_launch_dependents();
  extern __shared__ uint8_t s2_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s2_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ring = smem;
  uint8_t* meta = ring + kS2Stages * kS2StageBytes;              // [2][Q | Knew | Vnew | G parent tables]
  float* red = reinterpret_cast<float*>(meta + 2 * kS2MetaBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + 2 * kS2RedFloats);
  uint64_t* empty_bar = full_bar + kS2Stages;
  uint64_t* mfull_bar = empty_bar + kS2Stages;
  uint64_t* mempty_bar = mfull_bar + 2;
  uint64_t* rfull_bar = mempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kS2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kS2Consumers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mfull_bar[i], 1);
      mbar_init(&mempty_bar[i], kS2Consumers);
      mbar_init(&rfull_bar[i], kS2Consumers);
      mbar_init(&rempty_bar[i], 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&mapQKV);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
  }
  __syncthreads();
  pdl_wait();
  if (p.skip_flag && *p.skip_flag) return;
  const int H = p.n_head, G = p.G;
  const int L = *p.len_ptr;                      // tokens per row including the new one
  const int pos_new = L - 1;
  const int cache_rows = pos_new * G;            // rows of an (audio, head) block that hold history
  const int n_tiles = (cache_rows + kS2TileRows - 1) / kS2TileRows;
  const int total_items = p.n_audio * H;
  const uint32_t ind_bytes = static_cast<uint32_t>((pos_new * 4 + 15) & ~15);

  if (warp == kS2Consumers) {
    // ===================== producer =====================
    if (lane == 0) {
      int q = 0, iq = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++iq) {
        const int head = item % H, audio = item / H;
        const int ms = iq & 1;
        uint8_t* mb = meta + ms * kS2MetaBytes;
        x2_wait(&mempty_bar[ms], ((iq >> 1) & 1) ^ 1);
        mbar_expect_tx(&mfull_bar[ms], 3 * kX2QBytes + G * ind_bytes);
        tma_load_2d(mb, &mapQKV, &mfull_bar[ms], head * 64, audio * G);
        tma_load_2d(mb + kX2QBytes, &mapQKV, &mfull_bar[ms], p.d + head * 64, audio * G);
        tma_load_2d(mb + 2 * kX2QBytes, &mapQKV, &mfull_bar[ms], 2 * p.d + head * 64, audio * G);
        if (ind_bytes)
          for (int b = 0; b < G; ++b)
            bulk_load_1d(mb + 3 * kX2QBytes + b * kS2IndBytes, p.indir + static_cast<long long>(audio * G + b) * p.ctx, ind_bytes,
                         &mfull_bar[ms]);
        for (int t = 0; t < n_tiles; ++t, ++q) {
          const int st = q % kS2Stages;
          x2_wait(&empty_bar[st], ((q / kS2Stages) & 1) ^ 1);
          mbar_expect_tx(&full_bar[st], kS2StageBytes);
          uint8_t* dst = ring + st * kS2StageBytes;
          tma_load_3d(dst, &mapK, &full_bar[st], 0, t * kS2TileRows, audio * H + head);
          tma_load_3d(dst + kS2TileBytes, &mapV, &full_bar[st], 0, t * kS2TileRows, audio * H + head);
        }
      }
    }
    return;
  }
  if (warp == kS2Consumers + 1) {
    // ===================== epilogue warp: merge the eight fragments, normalise, store =====================
    const int g = lane >> 2, t4 = lane & 3;
    int iq = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++iq) {
      const int head = item % H, audio = item / H;
      const int rb = iq & 1;
      const float* rbuf = red + rb * kS2RedFloats;
      x2_wait(&rfull_bar[rb], (iq >> 1) & 1);
      float m0 = -INFINITY;
#pragma unroll
      for (int w = 0; w < kS2Consumers; ++w) m0 = fmaxf(m0, rbuf[(w * 32 + lane) * 20 + 16]);
      float l0 = 0.f, o[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = 0.f;
#pragma unroll
      for (int w = 0; w < kS2Consumers; ++w) {
        const float* src = rbuf + (w * 32 + lane) * 20;
        const float2 ml = *reinterpret_cast<const float2*>(src + 16);
        const float f0 = ml.x == -INFINITY ? 0.f : fast_exp2(ml.x - m0);
        l0 += ml.y * f0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
          o[2 * i][0] += v.x * f0;
          o[2 * i][1] += v.y * f0;
          o[2 * i + 1][0] += v.z * f0;
          o[2 * i + 1][1] += v.w * f0;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rempty_bar[rb]);
      if (g < G) {                                   // G <= 8: only the rows g of the fragment are beams
        const float i0 = 1.0f / l0;
        T* o0 = reinterpret_cast<T*>(p.out) + (static_cast<long long>(audio) * G + g) * p.d + head * 64 + 2 * t4;
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint32_t*>(o0 + i * 8) = Cvt<T>::pack2(o[i][0] * i0, o[i][1] * i0);
      }
    }
    return;
  }
  // ===================== consumers =====================
  int q = 0, iq = 0;
  for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++iq) {
    const int head = item % H, audio = item / H;
    const int ms = iq & 1;
    const uint8_t* mb = meta + ms * kS2MetaBytes;
    const int* s_ind = reinterpret_cast<const int*>(mb + 3 * kX2QBytes);      // [G][448]
    uint32_t qa[4][4];
    x2_wait(&mfull_bar[ms], (iq >> 1) & 1);
    {
      const int row = lane & 15;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(qa[ks], mb + row * 128 + ((chunk ^ (row & 7)) << 4));
      }
    }
    WarpAcc acc;
    acc_init(acc);
    const int base_phys = audio * G;
    const int t2 = (lane & 3) * 2;
    for (int t = 0; t < n_tiles; ++t, ++q) {
      const int st = q % kS2Stages;
      x2_wait(&full_bar[st], (q / kS2Stages) & 1);
      const uint8_t* sK = ring + st * kS2StageBytes;
      const int row0 = t * kS2TileRows + warp * 16;          // first cache row of this warp's slice
      if (row0 < cache_rows) {
        // lanes i and i + 16: the set of beams whose parent table points at cache row row0 + i = (position, slot)
        uint32_t bits = 0;
        {
          const int kr = row0 + (lane & 15);
          const int pos = (kr * p.g_magic) >> 16, slot = kr - pos * G;
          if (pos < pos_new) {
            const int want = base_phys + slot;
#pragma unroll
            for (int b = 0; b < kS2MaxG; ++b)
              if (b < G) bits |= static_cast<uint32_t>(s_ind[b * (kS2IndBytes / 4) + pos] == want) << b;
          }
        }
        uint32_t kmask[4];
        kmask[0] = __shfl_sync(0xffffffffu, bits, t2);
        kmask[1] = __shfl_sync(0xffffffffu, bits, t2 + 1);
        kmask[2] = __shfl_sync(0xffffffffu, bits, t2 + 8);
        kmask[3] = __shfl_sync(0xffffffffu, bits, t2 + 9);
        warp_tile_beams<T>(sK, sK + kS2TileBytes, warp * 16, qa, acc, kmask);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);
    }
    // the two warps with the emptiest slice of a ragged last tile take the new position
    if (warp == kS2Consumers - 1) {
      // beam b attends its own new key (row b of the Knew / Vnew tiles)
      uint32_t kmask[4];
      kmask[0] = t2 < G ? 1u << t2 : 0u;
      kmask[1] = t2 + 1 < G ? 2u << t2 : 0u;
      kmask[2] = kmask[3] = 0u;
      warp_tile_beams<T>(mb + kX2QBytes, mb + 2 * kX2QBytes, 0, qa, acc, kmask);
    } else if (warp == kS2Consumers - 2) {
      // append the new K / V rows to the cache: [audio][head][pos_new][slot b][64]
      const long long blk = (static_cast<long long>(audio) * H + head) * p.ctx + pos_new;
      uint8_t* kdst = reinterpret_cast<uint8_t*>(p.kcache) + blk * G * 128;
      uint8_t* vdst = reinterpret_cast<uint8_t*>(p.vcache) + blk * G * 128;
      for (int i = lane; i < G * 8; i += 32) {
        const int b = i >> 3, c = i & 7;
        const int soff = b * 128 + ((c ^ (b & 7)) << 4);
        *reinterpret_cast<uint4*>(kdst + b * 128 + c * 16) = *reinterpret_cast<const uint4*>(mb + kX2QBytes + soff);
        *reinterpret_cast<uint4*>(vdst + b * 128 + c * 16) = *reinterpret_cast<const uint4*>(mb + 2 * kX2QBytes + soff);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&mempty_bar[ms]);
    // ---- hand the fragment (beam rows only) to the epilogue warp
    acc.l[0] += __shfl_xor_sync(0xffffffffu, acc.l[0], 1);
    acc.l[0] += __shfl_xor_sync(0xffffffffu, acc.l[0], 2);
    const int rb = iq & 1;
    x2_wait(&rempty_bar[rb], ((iq >> 1) & 1) ^ 1);
    float* mine = red + rb * kS2RedFloats + (warp * 32 + lane) * 20;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(mine + i * 4) = make_float4(acc.o[2 * i][0], acc.o[2 * i][1], acc.o[2 * i + 1][0], acc.o[2 * i + 1][1]);
    *reinterpret_cast<float4*>(mine + 16) = make_float4(acc.m[0], acc.l[0], 0.f, 0.f);
    __syncwarp();
    if (lane == 0) mbar_arrive(&rfull_bar[rb]);
  }
}

// Prefill-mode append: copy k|v of qkv[(a, i)] into cache[(a*group, i)].  One warp per (row, k/v).
template <typename T>
__global__ void kv_append_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                                 int n_rows, int n_init, int group, int d, int max_ctx, int head_major) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_rows * 2) return;
  const int row = w >> 1, which = w & 1;
  const int a = row / n_init, i = row % n_init;
  const T* src = qkv + static_cast<long long>(row) * 3 * d + (1 + which) * d;
  T* cache = (which ? vcache : kcache) + static_cast<long long>(a) * group * max_ctx * d;   // physical row a * group
  for (int c = lane * 8; c < d; c += 256) {
    // head-major [row][head][pos][64]; beam window [audio][head][pos][slot][64]: the prompt lives in slot 0 of its audio
    T* dst = head_major ? cache + ((static_cast<long long>(c >> 6) * max_ctx + i) * (head_major == 2 ? group : 1)) * 64 + (c & 63)
                        : cache + static_cast<long long>(i) * d + c;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src + c);
  }
}

// -------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------
// Number of key splits per (audio, head, q-tile).  Every CTA pays ~2 us of pipeline fill before its
// first tile is consumed, so splits are kept as few as still give ~6 CTAs per SM: 2 at C3
// (64 audios x 20 heads), up to 8 for a single audio.
int cross_attention_splits(int T, int n_groups) {
  if (T < 256) return 1;
  int s = (888 + n_groups - 1) / n_groups;
  if (s < 2) s = 2;
  if (s > 8) s = 8;
  const int max_by_len = T / 128;              // at least two 64-key tiles per split
  if (s > max_by_len) s = max_by_len;
  return s < 1 ? 1 : s;
}

size_t cross_attention_partial_floats(int n_audio, int n_q, int n_head, int T) {
  const int q_tiles = (n_q + 15) / 16;
  return static_cast<size_t>(n_audio) * q_tiles * n_head * 8 /* max splits */ * 16 * 66;
}

int launch_cross_attention(int dtype, const void* q, const void* k, const void* v, void* out,
                           float* partial, int* counters, const int* skip_flag, int n_audio, int n_q,
                           int T, int n_head, int kv_ld, cudaStream_t s, int head_major) {
  if (n_audio <= 0 || n_q <= 0) return 0;
  CrossParams p;
  p.head_major = head_major;
  p.q = q;
  p.k = k;
  p.v = v;
  p.out = out;
  p.partial = partial;
  p.counters = counters;
  p.skip_flag = skip_flag;
  p.n_q = n_q;
  p.q_tiles = (n_q + 15) / 16;
  p.T = T;
  p.d = n_head * 64;
  p.kv_ld = kv_ld;
  p.splits = cross_attention_splits(T, n_audio * p.q_tiles * n_head);
  p.keys_per_split = ((T + p.splits - 1) / p.splits + 63) / 64 * 64;
  static int stages_opt = 0, splits_opt = -1;
  if (!stages_opt) {
    const char* e = getenv("WB200_XATTN_STAGES");
    stages_opt = (e && (atoi(e) == 3 || atoi(e) == 6)) ? atoi(e) : 4;
    const char* e2 = getenv("WB200_XATTN_SPLITS");
    splits_opt = (e2 && atoi(e2) >= 1 && atoi(e2) <= 8) ? atoi(e2) : 0;
  }
  if (splits_opt) {
    p.splits = splits_opt;
    p.keys_per_split = ((T + p.splits - 1) / p.splits + 63) / 64 * 64;
  }
  dim3 grid(p.splits, n_head, n_audio * p.q_tiles);
  ProfileScope prof(PROF_CROSS_ATTN, s);
  const int smem = stages_opt * 2 * kDaTileBytes;
#define WB_XATTN(TT, ST)                                                                                   \
  {                                                                                                        \
    auto kern = cross_attention_kernel<TT, ST>;                                                            \
    static SmemOptIn optin;                                                                                \
    if (!optin.ensure(kern, smem)) return 40;                                                              \
    if (launch_pdl(kern, grid, dim3(kDaThreads), smem, s, p) != cudaSuccess) return 41;                   \
  }
  if (dtype == DT_BF16) {
    if (stages_opt == 3) WB_XATTN(__nv_bfloat16, 3) else if (stages_opt == 6) WB_XATTN(__nv_bfloat16, 6) else WB_XATTN(__nv_bfloat16, 4)
  } else {
    if (stages_opt == 3) WB_XATTN(__half, 3) else if (stages_opt == 6) WB_XATTN(__half, 6) else WB_XATTN(__half, 4)
  }
#undef WB_XATTN
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 41;
}

int launch_self_attention(int dtype, const void* qkv, void* kcache, void* vcache, void* out,
                          const int* indir, const int* len_ptr, const int* skip_flag, int n_rows,
                          int n_head, int max_ctx, int n_init, int group, cudaStream_t s, int head_major) {
  if (n_rows <= 0) return 0;
  SelfParams p;
  p.head_major = head_major;
  p.qkv = qkv;
  p.kcache = kcache;
  p.vcache = vcache;
  p.out = out;
  p.indir = indir;
  p.len_ptr = len_ptr;
  p.skip_flag = skip_flag;
  p.d = n_head * 64;
  p.max_ctx = max_ctx;
  p.n_init = n_init > 0 ? n_init : 1;
  p.group = group;
  static int cfg_u = 0, cfg_st = 4, cfg_wpc = 4;
  if (!cfg_u) {
    cfg_u = 2;
    const char* c = getenv("WB200_SA_CFG");         // tuning override: "U,ST,WPC" (keys/4 per block, ring depth, warps per CTA)
    if (c) sscanf(c, "%d,%d,%d", &cfg_u, &cfg_st, &cfg_wpc);
    if (cfg_wpc < 1 || cfg_wpc > 20) cfg_wpc = 4;
  }
  if (dtype == DT_BF16) {
    if (!indir) {
      kv_append_kernel<__nv_bfloat16><<<(n_rows * 2 * 32 + 255) / 256, 256, 0, s>>>(
          static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(kcache),
          static_cast<__nv_bfloat16*>(vcache), n_rows, p.n_init, group, p.d, max_ctx, head_major);
      count_launch();
    }
    ProfileScope prof(PROF_SELF_ATTN, s);
    const int rc = dispatch_sa<__nv_bfloat16>(p, n_rows, n_head, cfg_u, cfg_st, cfg_wpc, s);
    if (rc) return rc;
  } else {
    if (!indir) {
      kv_append_kernel<__half><<<(n_rows * 2 * 32 + 255) / 256, 256, 0, s>>>(
          static_cast<const __half*>(qkv), static_cast<__half*>(kcache), static_cast<__half*>(vcache),
          n_rows, p.n_init, group, p.d, max_ctx, head_major);
      count_launch();
    }
    ProfileScope prof(PROF_SELF_ATTN, s);
    const int rc = dispatch_sa<__half>(p, n_rows, n_head, cfg_u, cfg_st, cfg_wpc, s);
    if (rc) return rc;
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 43;
}

// Step-mode cross attention through the TMA kernel: q [n_audio * n_q, d] (n_q <= 16), kv = ONE layer's head-major block
// [n_audio][2H][T][64].  Returns -1 when the shape is not covered (the caller falls back to the cp.async kernel).
int launch_cross_attention_tma(int dtype, const void* q, const void* kv, void* out, float* partial, int* counters,
                               const int* skip_flag, int n_audio, int n_q, int T, int n_head, cudaStream_t s) {
  if (n_q > 16 || n_audio <= 0 || T < kX2TileKeys) return -1;
  const int d = n_head * 64;
  Cross2Params p;
  p.out = out;
  p.partial = partial;
  p.counters = counters;
  p.skip_flag = skip_flag;
  p.n_q = n_q;
  p.n_head = n_head;
  p.T = T;
  p.d = d;
  p.total_tiles = (T + kX2TileKeys - 1) / kX2TileKeys;
  const int sms = sm_count();
  const int pairs = n_audio * n_head;
  // key splits: the count (1..4, at least 2 tiles each) that balances items over the persistent CTAs best
  int best = 1;
  double best_eff = 0.0;
  for (int sp = 1; sp <= 4; ++sp) {
    const int tps = (p.total_tiles + sp - 1) / sp;
    if (sp > 1 && (tps < 2 || (sp - 1) * tps >= p.total_tiles)) continue;
    const long long items = static_cast<long long>(pairs) * sp;
    const long long rounds = (items + sms - 1) / sms;
    const double eff = static_cast<double>(items) / static_cast<double>(rounds * sms) - 0.03 * (sp - 1);   // a split costs a partial + ticket
    if (eff > best_eff) {
      best_eff = eff;
      best = sp;
    }
  }
  static int splits_opt = -1;
  if (splits_opt < 0) {
    const char* e = getenv("WB200_XATTN_SPLITS");
    splits_opt = (e && atoi(e) >= 1 && atoi(e) <= 8) ? atoi(e) : 0;
  }
  if (splits_opt && (splits_opt - 1) * ((p.total_tiles + splits_opt - 1) / splits_opt) < p.total_tiles) best = splits_opt;
  p.splits = best;
  p.tiles_per_split = (p.total_tiles + best - 1) / best;
  p.total_items = pairs * best;
  CUtensorMap mapQ, mapKV;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(d), static_cast<uint64_t>(n_audio) * n_q};
    uint64_t strides[1] = {static_cast<uint64_t>(d) * 2};
    uint32_t box[2] = {64, 16};
    if (make_tmap_16bit(&mapQ, dtype, q, 2, dims, strides, box)) return 46;
  }
  {
    uint64_t dims[3] = {64, static_cast<uint64_t>(T), static_cast<uint64_t>(n_audio) * 2 * n_head};
    uint64_t strides[2] = {128, static_cast<uint64_t>(T) * 128};
    uint32_t box[3] = {64, kX2TileKeys, 1};
    if (make_tmap_16bit(&mapKV, dtype, kv, 3, dims, strides, box)) return 47;
  }
  const int grid = p.total_items < sms ? p.total_items : sms;
  ProfileScope prof(PROF_CROSS_ATTN, s);
  cudaError_t le;
  if (dtype == DT_BF16) {
    static SmemOptIn optin;
    auto kern = cross_attention_tma_kernel<__nv_bfloat16>;
    if (!optin.ensure(kern, kX2SmemBytes)) return 48;
    le = launch_pdl(kern, dim3(grid), dim3(kX2Threads), kX2SmemBytes, s, p, mapQ, mapKV);
  } else {
    static SmemOptIn optin;
    auto kern = cross_attention_tma_kernel<__half>;
    if (!optin.ensure(kern, kX2SmemBytes)) return 48;
    le = launch_pdl(kern, dim3(grid), dim3(kX2Threads), kX2SmemBytes, s, p, mapQ, mapKV);
  }
  count_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 49;
}

// Step-mode self attention of all beams of an audio together (beam-window kv layout, 2 <= G <= 8 and enough (audio,
// head) items to occupy the SMs: the session picks the layout from self_attention_tma_covers at create).
bool self_attention_tma_covers(int n_audio, int G, int n_head, int max_ctx) {
  return G >= 2 && G <= kS2MaxG && max_ctx * 4 <= kS2IndBytes && n_audio * n_head * 2 >= sm_count();
}

int launch_self_attention_tma(int dtype, const void* qkv, void* kcache, void* vcache, void* out, const int* indir,
                              const int* len_ptr, const int* skip_flag, int n_audio, int G, int n_head, int max_ctx,
                              cudaStream_t s) {
  const int sms = sm_count();
  if (!self_attention_tma_covers(n_audio, G, n_head, max_ctx)) return 56;
  const int d = n_head * 64;
  Self2Params p;
  p.out = out;
  p.kcache = kcache;
  p.vcache = vcache;
  p.indir = indir;
  p.len_ptr = len_ptr;
  p.skip_flag = skip_flag;
  p.n_audio = n_audio;
  p.n_head = n_head;
  p.G = G;
  p.ctx = max_ctx;
  p.d = d;
  p.g_magic = (65536 + G - 1) / G;
  CUtensorMap mapQKV, mapK, mapV;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(n_audio) * G};
    uint64_t strides[1] = {static_cast<uint64_t>(3 * d) * 2};
    uint32_t box[2] = {64, 16};
    if (make_tmap_16bit(&mapQKV, dtype, qkv, 2, dims, strides, box)) return 57;
  }
  {
    uint64_t dims[3] = {64, static_cast<uint64_t>(max_ctx) * G, static_cast<uint64_t>(n_audio) * n_head};
    uint64_t strides[2] = {128, static_cast<uint64_t>(max_ctx) * G * 128};
    uint32_t box[3] = {64, kS2TileRows, 1};
    if (make_tmap_16bit(&mapK, dtype, kcache, 3, dims, strides, box)) return 58;
    if (make_tmap_16bit(&mapV, dtype, vcache, 3, dims, strides, box)) return 58;
  }
  const int items = n_audio * n_head;
  const int grid = items < sms ? items : sms;
  ProfileScope prof(PROF_SELF_ATTN, s);
  cudaError_t le;
  if (dtype == DT_BF16) {
    static SmemOptIn optin;
    auto kern = self_attention_tma_kernel<__nv_bfloat16>;
    if (!optin.ensure(kern, kS2SmemBytes)) return 59;
    le = launch_pdl(kern, dim3(grid), dim3(kS2Threads), kS2SmemBytes, s, p, mapQKV, mapK, mapV);
  } else {
    static SmemOptIn optin;
    auto kern = self_attention_tma_kernel<__half>;
    if (!optin.ensure(kern, kS2SmemBytes)) return 59;
    le = launch_pdl(kern, dim3(grid), dim3(kS2Threads), kS2SmemBytes, s, p, mapQKV, mapK, mapV);
  }
  count_launch();
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 59;
}

}  // namespace wb
