// Non-causal multi-head attention of the audio encoder on tcgen05 (reference
// whisper/model.py:114-139, SDPA branch: softmax(q k^T / sqrt(64)) v, T = 1500, head dim 64).
//
// One CTA owns 256 query rows of one (batch, head): two 128-row tiles, each with its own softmax
// warpgroup, sharing one K/V stream.  While warpgroup 0 is exponentiating tile 0's scores the
// tensor core runs tile 1's MMAs and vice versa (ping-pong), which is what hides the MUFU-bound
// softmax behind the tensor pipe.
//
//   warps 0-3 : softmax warpgroup for tile 0  (thread = one query row = one TMEM lane)
//   warps 4-7 : softmax warpgroup for tile 1
//   warp  8   : TMA producer - Q (2 x 128x64), then K/V blocks of 128 keys through a 3-stage ring
//   warp  9   : tcgen05.mma issuer + TMEM allocator
//
// Per key block j and tile w:   S_w = Q_w K_j^T  (UMMA 128x128x16, K-major A and B)
//                               P_w = exp2(S_w * scale - m)  -> 16-bit, written to smem in the
//                                     128B-swizzled K-major layout UMMA expects for an A operand
//                               O_w = P_w V_j  (UMMA 128x64x16, B = V tile read MN-major)
// The running output is kept in registers (64 fp32 per row) and rescaled there, so the PV product
// of each block is a fresh (non-accumulating) TMEM tile and no TMEM read-modify-write is needed.
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.cuh"

namespace wb {

constexpr int kAttThreads = 320;
constexpr int kKvStages = 3;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: 128 rows x 64 x 16-bit
constexpr int kAttSmem = 2 * kTileBytes /*Q*/ + kKvStages * 2 * kTileBytes /*K,V*/ +
                         2 * 2 * kTileBytes /*P: 2 tiles x 2 sub-tiles*/ + 1024 + 256;

template <typename T>
__global__ void __launch_bounds__(kAttThreads, 1)
enc_attention_kernel(const __grid_constant__ CUtensorMap map_qkv, T* __restrict__ out, int Tn,
                     int n_head, int d_model) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * kTileBytes;
  uint8_t* sP = sKV + kKvStages * 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kTileBytes);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // kKvStages
  uint64_t* kv_empty = kv_full + kKvStages; // kKvStages
  uint64_t* s_full = kv_empty + kKvStages;  // 2
  uint64_t* p_full = s_full + 2;            // 2
  uint64_t* o_full = p_full + 2;            // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_blocks = (Tn + 127) / 128;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

  if (warp == 8) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(sQ, &map_qkv, q_full, h * 64, q0, b);
      tma_load_3d(sQ + kTileBytes, &map_qkv, q_full, h * 64, q0 + 128, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + stage * 2 * kTileBytes;
        mbar_expect_tx(&kv_full[stage], 2 * kTileBytes);
        tma_load_3d(sk, &map_qkv, &kv_full[stage], d_model + h * 64, j * 128, b);
        tma_load_3d(sk + kTileBytes, &map_qkv, &kv_full[stage], 2 * d_model + h * 64, j * 128, b);
        if (++stage == kKvStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = umma_idesc(Cvt<T>::kUmmaFmt, 128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc(Cvt<T>::kUmmaFmt, 128, 64, 0, 1);  // B = V, MN-major
      auto issue_s = [&](int w, int stage) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(sQ + w * kTileBytes), 16, 1024);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(sKV + stage * 2 * kTileBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + w * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[w]);
      };
      auto issue_pv = [&](int w, int stage) {
        const uint32_t pbase = smem_u32(sP + w * 2 * kTileBytes);
        const uint32_t vbase = smem_u32(sKV + stage * 2 * kTileBytes + kTileBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: P, K-major, key block of 64 per sub-tile, 16 keys = 32 B inside the swizzle row
          const uint64_t adesc =
              umma_desc_sw128(pbase + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
          // B: V[key][dh] read MN-major: 16 keys = 16 rows of 128 B
          const uint64_t bdesc = umma_desc_sw128(vbase + k * 16 * 128, 1024, 1024);
          umma_f16(tmem_base + 256 + w * 64, adesc, bdesc, idesc_o, k != 0);
        }
        umma_commit(&o_full[w]);
      };

      mbar_wait(q_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < n_blocks; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == kKvStages) {
          nstage = 0;
          nphase ^= 1;
        }
        if (j + 1 < n_blocks) mbar_wait(&kv_full[nstage], nphase);
        for (int w = 0; w < 2; ++w) {
          mbar_wait(&p_full[w], j & 1);
          tc_fence_after();
          if (j + 1 < n_blocks) issue_s(w, nstage);
          issue_pv(w, stage);
        }
        umma_commit(&kv_empty[stage]);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int w = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int qrow = q0 + w * 128 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + w * 128;
    const uint32_t o_addr = tmem_base + lane_off + 256 + w * 64;
    uint8_t* myP = sP + w * 2 * kTileBytes + row * 128;
    const float scale_log2 = 0.125f * 1.4426950408889634f;

    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

    for (int j = 0; j < n_blocks; ++j) {
      mbar_wait(&s_full[w], j & 1);
      tc_fence_after();
      const int kvalid = Tn - j * 128;  // keys [0,kvalid) of this block are real
      const bool full_block = kvalid >= 128;   // every block but the last: no per-element masking
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_addr + c * 32, r);
        tmem_ld_wait();
        if (full_block) {
          // four independent running maxima keep the FMNMX chain short
          float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(r[i]));
            m1 = fmaxf(m1, __uint_as_float(r[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(r[i + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * scale_log2);
      const float alpha = fast_exp2(m_run - m_new);
      float rs0 = 0.f, rs1 = 0.f;
      if (j > 0) {
        // fold in the previous block's P V before P's smem buffer is overwritten below
        mbar_wait(&o_full[w], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(o_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha_prev + __uint_as_float(r[i]);
        }
      }
      alpha_prev = alpha;
      // P = exp2(S * scale - m) for this row, 32 keys at a time, packed to 16 bits and stored straight into
      // the 128B-swizzled K-major smem tile (16-byte chunks XOR-ed with row % 8) the PV MMA reads
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_addr + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
        if (full_block) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, -m_new));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
            rs0 += p0;
            rs1 += p1;
            pk[i / 2] = Cvt<T>::pack2(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, -m_new));
            float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m_new));
            if (c * 32 + i >= kvalid) p0 = 0.f;
            if (c * 32 + i + 1 >= kvalid) p1 = 0.f;
            rs0 += p0;
            rs1 += p1;
            pk[i / 2] = Cvt<T>::pack2(p0, p1);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ch = c * 4 + q;
          const int sub = ch >> 3, c16 = ch & 7;
          *reinterpret_cast<uint4*>(myP + sub * kTileBytes + ((c16 ^ (row & 7)) << 4)) =
              make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      l_run = l_run * alpha + (rs0 + rs1);
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[w]);
    }
    // last block's P V
    mbar_wait(&o_full[w], (n_blocks - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(o_addr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        o_acc[c * 32 + i] = o_acc[c * 32 + i] * alpha_prev + __uint_as_float(r[i]);
    }
    if (qrow < Tn) {
      const float inv = 1.0f / l_run;
      T* orow = out + (static_cast<long long>(b) * Tn + qrow) * d_model + h * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 u;
        u.x = Cvt<T>::pack2(o_acc[q * 8 + 0] * inv, o_acc[q * 8 + 1] * inv);
        u.y = Cvt<T>::pack2(o_acc[q * 8 + 2] * inv, o_acc[q * 8 + 3] * inv);
        u.z = Cvt<T>::pack2(o_acc[q * 8 + 4] * inv, o_acc[q * 8 + 5] * inv);
        u.w = Cvt<T>::pack2(o_acc[q * 8 + 6] * inv, o_acc[q * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + q * 8) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

int launch_enc_attention(int dtype, const void* qkv, void* out, int B, int T, int n_head,
                         cudaStream_t s) {
  if (B <= 0 || T <= 0) return 0;
  const int d = n_head * 64;
  CUtensorMap map;
  uint64_t dims[3] = {static_cast<uint64_t>(3 * d), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {static_cast<uint64_t>(3 * d) * 2, static_cast<uint64_t>(T) * 3 * d * 2};
  uint32_t box[3] = {64, 128, 1};
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return 30;
  if (make_tmap_16bit(&map, dtype, qkv, 3, dims, strides, box)) return 31;
  dim3 grid((T + 255) / 256, n_head, B);
  ProfileScope prof(PROF_ENC_ATTN, s);
  static SmemOptIn optin[2];
  if (dtype == DT_BF16) {
    auto kern = enc_attention_kernel<__nv_bfloat16>;
    if (!optin[0].ensure(kern, kAttSmem)) return 32;
    kern<<<grid, kAttThreads, kAttSmem, s>>>(map, static_cast<__nv_bfloat16*>(out), T, n_head, d);
  } else {
    auto kern = enc_attention_kernel<__half>;
    if (!optin[1].ensure(kern, kAttSmem)) return 32;
    kern<<<grid, kAttThreads, kAttSmem, s>>>(map, static_cast<__half*>(out), T, n_head, d);
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 33;
}

}  // namespace wb
