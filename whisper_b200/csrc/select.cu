// Token selection on the device: the logit filters, log-softmax, top-k and the greedy / beam
// bookkeeping of reference whisper/decoding.py:272-505, which the reference runs as Python loops
// with a host sync per row and per beam candidate.
//
//   filter_topk_kernel : one CTA per row.  Applies SuppressBlank (decoding.py:423-430),
//                        SuppressTokens (:433-438) and ApplyTimestampRules (:441-505) as masks
//                        computed on the fly (the logits buffer is never rewritten), then the
//                        log-softmax normaliser and the top-K (K = 1 greedy, beam+1 beam search)
//                        by warp/CTA reductions.  Integer rules are exact; ties in the top-K go
//                        to the lower token id.  With a temperature (GreedyDecoder, decoding.py:283:
//                        Categorical(logits / T).sample()) the K = 1 selection becomes a Gumbel-max
//                        draw with a counter-based generator, see gumbel_noise() for the contract.
//   greedy_update_kernel / beam_update_kernel : one CTA.  Append the chosen tokens, update
//                        sum_logprobs, EOT bookkeeping / finished-hypothesis store, beam parent
//                        table (the kv-cache "reorder"), completion flag.
#include "kernels.h"
#include "ptx.cuh"

namespace wb {

constexpr int kSelThreads = 512;
constexpr int kMaxTopK = 17;  // beam <= 16
constexpr int kSelCluster = 8; // CTAs per row when there are few rows (portable cluster size)
constexpr int kSelUnroll = 4;  // 16-byte loads in flight per thread in the vocabulary scans

// total order used everywhere: larger value first, then smaller index
__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  if (m2 == -INFINITY) return;
  if (m == -INFINITY) {
    m = m2;
    s = s2;
    return;
  }
  if (m2 > m) {
    s = s * __expf(m - m2) + s2;
    m = m2;
  } else {
    s += s2 * __expf(m2 - m);
  }
}


// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): counter-based, so the
// noise of (row, step, token) needs no state and does not depend on launch geometry.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// RNG contract of the sampling path (restated by oracle/decoding.py: gumbel_noise):
//   words  = Philox4x32-10(counter = (v >> 2, row, L, 0), key = (seed_lo, seed_hi)),  L = tokens in the row
//   u      = ((words[v & 3] >> 8) + 0.5) * 2^-24            in (0, 1)
//   g      = -log(-log(u))                                  fp32
//   sample = argmax_v( logit_v / T + g_v ) over the tokens the filters leave, ties to the lower id,
// which is an exact draw from Categorical(softmax(logits / T)) (decoding.py:283).
__device__ __forceinline__ float gumbel_noise(uint32_t seed_lo, uint32_t seed_hi, int row, int L, int v) {
  uint32_t c[4] = {static_cast<uint32_t>(v) >> 2, static_cast<uint32_t>(row), static_cast<uint32_t>(L), 0u};
  philox4x32_10(c, seed_lo, seed_hi);
  const uint32_t w = c[v & 3];
  const float u = (static_cast<float>(w >> 8) + 0.5f) * 5.9604644775390625e-8f;
  return -logf(-logf(u));
}

// Per-thread sorted list of the KM best (value, index) pairs, kept in registers: every index below is a compile-time
// constant after unrolling (a runtime-indexed array would live in local memory - the first version of this kernel did,
// and spent most of its 95 us there).
template <int KM>
struct TopList {
  float v[KM];
  int i[KM];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      v[k] = -INFINITY;
      i[k] = 0x7fffffff;
    }
  }
  __device__ __forceinline__ void push(float val, int idx) {
    if (!better(val, idx, v[KM - 1], i[KM - 1])) return;
    v[KM - 1] = val;
    i[KM - 1] = idx;
#pragma unroll
    for (int k = KM - 1; k > 0; --k) {
      if (better(v[k], i[k], v[k - 1], i[k - 1])) {
        const float tv = v[k];
        const int ti = i[k];
        v[k] = v[k - 1];
        i[k] = i[k - 1];
        v[k - 1] = tv;
        i[k - 1] = ti;
      }
    }
  }
  // remove and return the head (static shifts)
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int k = 0; k + 1 < KM; ++k) {
      v[k] = v[k + 1];
      i[k] = i[k + 1];
    }
    v[KM - 1] = -INFINITY;
    i[KM - 1] = 0x7fffffff;
  }
};

// K rounds of "pop the best head among the 32 lanes' sorted lists"; lane 0 hands every winner to emit(k, value, index)
template <int KM, typename Emit>
__device__ __forceinline__ void warp_merge_lists(TopList<KM>& l, int K, Emit emit) {
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < K; ++k) {
    const float bv = l.v[0];
    const int bi = l.i[0];
    float cv = bv;
    int ci = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
      if (better(ov, oi, cv, ci)) {
        cv = ov;
        ci = oi;
      }
    }
    if (ci == bi && cv == bv && bi != 0x7fffffff) l.pop();   // this lane's head was taken
    if (lane == 0) emit(k, cv, ci);
  }
}

// thread-block cluster helpers (CL CTAs share one row: partial results are read from the peers' shared memory)
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t peer_u32(const void* own_smem, uint32_t rank) {
  uint32_t ra, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(own_smem)), "r"(rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ float peer_f32(const void* own_smem, uint32_t rank) { return __uint_as_float(peer_u32(own_smem, rank)); }

// CL = 1: one CTA per row.  CL = 8 (few rows: one audio decoded greedily would otherwise scan its 51 866 logits with a
// single CTA - 48 us measured, profiles/r2_launches_turbo_b1.csv): a cluster of 8 CTAs per row, each scanning an eighth
// of the vocabulary; the (max, sum-exp) partials and the per-CTA top-K lists are merged through distributed shared memory.
template <int KM, int CL>
__global__ void __launch_bounds__(kSelThreads) filter_topk_kernel(const FilterParams p) {
  if (p.skip_flag && *p.skip_flag) return;
  const int r = CL > 1 ? blockIdx.x / CL : blockIdx.x;
  const uint32_t crank = CL > 1 ? cluster_rank() : 0u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = *p.len_ptr;
  const float* x = p.logits + static_cast<long long>(r / p.row_div) * p.ld;
  const int tb = p.timestamp_begin;

  __shared__ int s_rule[4];      // ts_lo (mask [tb, ts_lo)), mask_all_ts, mask_text_below_eot, drop_text
  __shared__ float s_red[kSelThreads / 32][4];
  __shared__ float s_fin[2];
  __shared__ float s_tv[kSelThreads / 32][KM];
  __shared__ int s_ti[kSelThreads / 32][KM];
  __shared__ float s_part[4];    // this CTA's (m_txt, s_txt, m_ts, s_ts), read by the cluster peers
  __shared__ float s_cv[KM];     // this CTA's top-K, read by rank 0
  __shared__ int s_ci[KM];

  const bool first = (L == p.sample_begin);
  if (tid == 0) {
    int ts_lo = tb, all_ts = 0, text_lt_eot = 0;
    if (p.ts_rules) {
      const int* row = p.tokens + static_cast<long long>(r) * p.max_ctx;
      const int n = L - p.sample_begin;                       // sampled tokens so far
      const bool last_ts = n >= 1 && row[L - 1] >= tb;        // decoding.py:461-463
      const bool pen_ts = n < 2 || row[L - 2] >= tb;          // decoding.py:464-466
      if (last_ts) {
        if (pen_ts) all_ts = 1; else text_lt_eot = 1;         // decoding.py:468-472
      }
      int last_stamp = -1;                                    // decoding.py:474-484
      for (int i = L - 1; i >= p.sample_begin; --i)
        if (row[i] >= tb) { last_stamp = row[i]; break; }
      if (last_stamp >= 0) ts_lo = (last_ts && !pen_ts) ? last_stamp : last_stamp + 1;
    }
    s_rule[0] = ts_lo;
    s_rule[1] = all_ts;
    s_rule[2] = text_lt_eot;
  }
  __syncthreads();
  const int ts_lo = s_rule[0];
  const bool all_ts = s_rule[1], text_lt_eot = s_rule[2];
  const int ts_hi = (p.ts_rules && first && p.max_initial_ts >= 0) ? tb + p.max_initial_ts : p.V - 1;
  const bool use_blank = p.suppress_blank && first;

  // rule mask of token v given the suppress / blank bitmap words that hold it
  auto masked = [&](int v, uint32_t sup_word, uint32_t blank_word) -> bool {
    if ((sup_word >> (v & 31)) & 1u) return true;
    if (use_blank && ((blank_word >> (v & 31)) & 1u)) return true;
    if (p.ts_rules) {
      if (v >= tb) {
        if (all_ts || v < ts_lo || v > ts_hi) return true;
      } else {
        if (first) return true;                               // decoding.py:486-488
        if (text_lt_eot && v < p.eot) return true;
      }
    }
    return false;
  };
  // The row is walked four tokens per thread per step (one 16-byte load; the four tokens share their bitmap words);
  // the <= 3 tokens past the last multiple of four are handled by the first threads.
  const int V4 = p.V >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int q_lo = CL > 1 ? static_cast<int>(static_cast<long long>(V4) * crank / CL) : 0;
  const int q_hi = CL > 1 ? static_cast<int>(static_cast<long long>(V4) * (crank + 1) / CL) : V4;
  const bool has_tail = crank == CL - 1;

  // ---- pass 1: (max, sum-exp) of the text part [0, tb) and the timestamp part [tb, V)
  float m_txt = -INFINITY, s_txt = 0.f, m_ts = -INFINITY, s_ts = 0.f;
  auto acc1 = [&](int v, float val, uint32_t sw, uint32_t bw) {
    if (val == -INFINITY || masked(v, sw, bw)) return;
    if (v < tb) lse_merge(m_txt, s_txt, val, 1.f); else lse_merge(m_ts, s_ts, val, 1.f);
  };
  // four independent 16-byte loads in flight per thread: the scan is latency-bound (one CTA reads a 207 KB row)
  for (int q = q_lo + tid; q < q_hi; q += kSelUnroll * kSelThreads) {
    float4 f[kSelUnroll];
    uint32_t sw[kSelUnroll], bw[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int qq = q + u * kSelThreads;
      if (qq < q_hi) {
        f[u] = x4[qq];
        sw[u] = __ldg(p.suppress_mask + (qq >> 3));
        bw[u] = use_blank ? __ldg(p.blank_mask + (qq >> 3)) : 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int qq = q + u * kSelThreads;
      if (qq < q_hi) {
        const int v = qq << 2;
        acc1(v, f[u].x, sw[u], bw[u]);
        acc1(v + 1, f[u].y, sw[u], bw[u]);
        acc1(v + 2, f[u].z, sw[u], bw[u]);
        acc1(v + 3, f[u].w, sw[u], bw[u]);
      }
    }
  }
  if (has_tail && tid < (p.V & 3)) {
    const int v = (V4 << 2) + tid;
    acc1(v, x[v], __ldg(p.suppress_mask + (v >> 5)), use_blank ? __ldg(p.blank_mask + (v >> 5)) : 0u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lse_merge(m_txt, s_txt, __shfl_xor_sync(0xffffffffu, m_txt, o), __shfl_xor_sync(0xffffffffu, s_txt, o));
    lse_merge(m_ts, s_ts, __shfl_xor_sync(0xffffffffu, m_ts, o), __shfl_xor_sync(0xffffffffu, s_ts, o));
  }
  if (lane == 0) {
    s_red[warp][0] = m_txt;
    s_red[warp][1] = s_txt;
    s_red[warp][2] = m_ts;
    s_red[warp][3] = s_ts;
  }
  __syncthreads();
  if (CL > 1) {
    if (tid == 0) {
      float a = -INFINITY, b = 0.f, c = -INFINITY, d = 0.f;
      for (int w = 0; w < kSelThreads / 32; ++w) {
        lse_merge(a, b, s_red[w][0], s_red[w][1]);
        lse_merge(c, d, s_red[w][2], s_red[w][3]);
      }
      s_part[0] = a;
      s_part[1] = b;
      s_part[2] = c;
      s_part[3] = d;
    }
    cluster_sync_all();
  }
  if (tid == 0) {
    float a = -INFINITY, b = 0.f, c = -INFINITY, d = 0.f;
    if (CL > 1) {
      for (uint32_t k = 0; k < CL; ++k) {          // the same order in every CTA of the cluster: identical results
        lse_merge(a, b, peer_f32(&s_part[0], k), peer_f32(&s_part[1], k));
        lse_merge(c, d, peer_f32(&s_part[2], k), peer_f32(&s_part[3], k));
      }
    } else {
      for (int w = 0; w < kSelThreads / 32; ++w) {
        lse_merge(a, b, s_red[w][0], s_red[w][1]);
        lse_merge(c, d, s_red[w][2], s_red[w][3]);
      }
    }
    float mt = a, st = b;
    lse_merge(mt, st, c, d);                                   // everything
    const float lse_all = mt + __logf(st);
    const float lse_ts = c == -INFINITY ? -INFINITY : c + __logf(d);
    int drop_text = 0;
    if (p.ts_rules) {
      // decoding.py:498-505: logsumexp(logprobs[tb:]) > max(logprobs[:tb])
      const float ts_lp = lse_ts - lse_all;
      const float txt_lp = a - lse_all;
      drop_text = ts_lp > txt_lp;
    }
    s_fin[0] = drop_text ? lse_ts : lse_all;
    s_rule[3] = drop_text;
  }
  __syncthreads();
  const float lse = s_fin[0];
  const bool drop_text = s_rule[3];

  // ---- pass 2: top-K of the surviving logits (the row is L2 / L1 resident from pass 1)
  TopList<KM> mine;
  mine.init();
  const int K = p.K;
  const bool sampling = p.inv_temp > 0.f && K == 1;
  auto acc2 = [&](int v, float val, uint32_t sw, uint32_t bw) {
    if (val == -INFINITY || masked(v, sw, bw) || (drop_text && v < tb)) return;
    if (sampling) val = fmaf(val, p.inv_temp, gumbel_noise(p.seed_lo, p.seed_hi, r, L, v));
    mine.push(val, v);
  };
  for (int q = q_lo + tid; q < q_hi; q += kSelUnroll * kSelThreads) {
    float4 f[kSelUnroll];
    uint32_t sw[kSelUnroll], bw[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int qq = q + u * kSelThreads;
      if (qq < q_hi) {
        f[u] = x4[qq];
        sw[u] = __ldg(p.suppress_mask + (qq >> 3));
        bw[u] = use_blank ? __ldg(p.blank_mask + (qq >> 3)) : 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int qq = q + u * kSelThreads;
      if (qq < q_hi) {
        const int v = qq << 2;
        acc2(v, f[u].x, sw[u], bw[u]);
        acc2(v + 1, f[u].y, sw[u], bw[u]);
        acc2(v + 2, f[u].z, sw[u], bw[u]);
        acc2(v + 3, f[u].w, sw[u], bw[u]);
      }
    }
  }
  if (has_tail && tid < (p.V & 3)) {
    const int v = (V4 << 2) + tid;
    acc2(v, x[v], __ldg(p.suppress_mask + (v >> 5)), use_blank ? __ldg(p.blank_mask + (v >> 5)) : 0u);
  }
  // warp merge, then lane w of warp 0 holds warp w's sorted list and the same merge runs across the 16 warps
  warp_merge_lists<KM>(mine, K, [&](int k, float cv, int ci) {
    s_tv[warp][k] = cv;
    s_ti[warp][k] = ci;
  });
  __syncthreads();
  auto emit_final = [&](int k, float cv, int ci) {
    // log-probability of the chosen token under the UN-tempered distribution (decoding.py:285-287)
    const float chosen = sampling ? ((ci >= 0 && ci < p.V) ? x[ci] : -INFINITY) : cv;
    p.top_val[static_cast<long long>(r) * K + k] = chosen - lse;
    p.top_idx[static_cast<long long>(r) * K + k] = ci;
  };
  if (warp == 0) {
    TopList<KM> l;
    l.init();
    if (lane < kSelThreads / 32) {
#pragma unroll
      for (int k = 0; k < KM; ++k)
        if (k < K) {
          l.v[k] = s_tv[lane][k];
          l.i[k] = s_ti[lane][k];
        }
    }
    if (CL > 1)
      warp_merge_lists<KM>(l, K, [&](int k, float cv, int ci) {
        s_cv[k] = cv;
        s_ci[k] = ci;
      });
    else
      warp_merge_lists<KM>(l, K, emit_final);
  }
  if (CL > 1) {
    cluster_sync_all();
    if (crank == 0 && warp == 0) {
      // lane c holds the sorted list of cluster rank c
      TopList<KM> l;
      l.init();
      if (lane < CL) {
#pragma unroll
        for (int k = 0; k < KM; ++k)
          if (k < K) {
            l.v[k] = peer_f32(&s_cv[k], lane);
            l.i[k] = static_cast<int>(peer_u32(&s_ci[k], lane));
          }
      }
      warp_merge_lists<KM>(l, K, emit_final);
    }
    cluster_sync_all();          // nobody leaves while rank 0 still reads its shared memory
  }
}

// -------------------------------------------------------------------------------------------------
// no-speech probability at the <|startoftranscript|> position (decoding.py:689-693)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads) no_speech_kernel(const float* logits, long long ld, int V,
                                                                int no_speech, float* out, int row_stride,
                                                                int row_offset) {
  const float* x = logits + (static_cast<long long>(blockIdx.x) * row_stride + row_offset) * ld;
  __shared__ float s_red[kSelThreads / 32][2];
  float m = -INFINITY, s = 0.f;
  for (int v = threadIdx.x; v < V; v += kSelThreads) lse_merge(m, s, x[v], 1.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    lse_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
  if ((threadIdx.x & 31) == 0) {
    s_red[threadIdx.x >> 5][0] = m;
    s_red[threadIdx.x >> 5][1] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = -INFINITY, b = 0.f;
    for (int w = 0; w < kSelThreads / 32; ++w) lse_merge(a, b, s_red[w][0], s_red[w][1]);
    out[blockIdx.x] = __expf(x[no_speech] - a) / b;
  }
}

// -------------------------------------------------------------------------------------------------
// greedy update (decoding.py:277-293, temperature 0)
// -------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(1024) greedy_update_kernel(const GreedyParams p) {
  if (p.skip_flag && *p.skip_flag) return;
  __shared__ int s_all;
  if (threadIdx.x == 0) s_all = 1;
  __syncthreads();
  const int L = *p.len_ptr;
  int all_eot = 1;
  for (int r = threadIdx.x; r < p.R; r += blockDim.x) {
    int* row = p.tokens + static_cast<long long>(r) * p.max_ctx;
    const int last = row[L - 1];
    int nxt = p.top_idx[r];
    if (last != p.eot) p.sum_logprobs[r] += p.top_val[r];      // decoding.py:287
    else nxt = p.eot;                                          // decoding.py:289
    row[L] = nxt;
    if (nxt != p.eot) all_eot = 0;
  }
  if (!all_eot) atomicAnd(&s_all, 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    *p.len_ptr = L + 1;
    if (s_all) *p.done_flag = 1;
  }
}

// -------------------------------------------------------------------------------------------------
// beam-search update (decoding.py:323-382); one warp per audio, lane 0 runs the (tiny) sequential
// candidate logic, all lanes help with prefix comparison and row copies.
// -------------------------------------------------------------------------------------------------

constexpr int kMaxBeam = 16;
constexpr int kMaxCand = kMaxBeam * (kMaxBeam + 1);

constexpr int kBeamThreads = 256;

// One CTA per audio.  The candidate logic of decoding.py:335-360 is inherently sequential (dict insertion order, "last
// writer wins", first G non-EOT survivors) and runs on thread 0 over shared memory; everything around it is parallel:
// the G x (G+1) candidates are gathered by as many threads, finished hypotheses and the new token / parent-table rows
// are copied by one warp each.  Whether two beams hold the same token prefix - what the reference's dict keys decide -
// is NOT recomputed by comparing rows: it follows exactly from the previous step,
//     same'[j1][j2] = same[src(j1)][src(j2)]  and  token(j1) == token(j2),
// and is carried in a G x G byte matrix per audio (all ones after the prefill: every beam starts from the prompt).
__global__ void __launch_bounds__(kBeamThreads) beam_update_kernel(const BeamParams p) {
  if (p.skip_flag && *p.skip_flag) return;
  __shared__ int s_all_done;
  __shared__ float s_score[kMaxCand];
  __shared__ short s_tok_row[kMaxCand];   // owning beam j
  __shared__ int s_tok[kMaxCand];
  __shared__ short s_order[kMaxCand];
  __shared__ unsigned char s_dead[kMaxCand];
  __shared__ unsigned char s_same[kMaxBeam][kMaxBeam];
  __shared__ int s_newsrc[kMaxBeam];
  __shared__ int s_newtok[kMaxBeam];
  __shared__ int s_fin_src[kMaxBeam + 1];  // newly finished hypotheses of this step: source row, slot
  __shared__ int s_fin_slot[kMaxBeam + 1];
  __shared__ int s_n_fin;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = *p.len_ptr;
  const int G = p.G, K = G + 1, N = G * K;
  const int a = blockIdx.x;
  const int r0 = a * G;
  if (tid == 0) s_all_done = 1;
  // candidates in insertion order (j, then top-k rank); score = fp32(sum_lp + lp)   (decoding.py:339-346)
  for (int c = tid; c < N; c += kBeamThreads) {
    const int j = c / K;
    s_score[c] = p.sum_logprobs[r0 + j] + p.top_val[static_cast<long long>(r0 + j) * K + c % K];
    s_tok[c] = p.top_idx[static_cast<long long>(r0 + j) * K + c % K];
    s_tok_row[c] = static_cast<short>(j);
    s_dead[c] = 0;
  }
  if (tid < G * G) s_same[tid / G][tid % G] = p.same_in[static_cast<long long>(a) * kMaxBeam * kMaxBeam + (tid / G) * kMaxBeam + tid % G];
  __syncthreads();
  if (tid == 0) {
    // dict semantics: a later duplicate (same prefix, same token) overwrites value and source but
    // keeps the FIRST insertion position.  s_dead[c] marks removed later duplicates.
    for (int c = 0; c < N; ++c) {
      if (s_dead[c]) continue;
      for (int c2 = c + 1; c2 < N; ++c2) {
        if (s_dead[c2]) continue;
        const int j1 = s_tok_row[c], j2 = s_tok_row[c2];
        if (j1 != j2 && s_tok[c] == s_tok[c2] && s_same[j1][j2]) {
          s_score[c] = s_score[c2];       // last writer's value ...
          s_tok_row[c] = s_tok_row[c2];   // ... and source
          s_dead[c2] = 1;
        }
      }
    }
    // stable descending sort of the live candidates (insertion sort; N <= 272)
    int n_live = 0;
    for (int c = 0; c < N; ++c) {
      if (s_dead[c]) continue;
      int pos = n_live++;
      while (pos > 0 && s_score[s_order[pos - 1]] < s_score[c]) {
        s_order[pos] = s_order[pos - 1];
        --pos;
      }
      s_order[pos] = static_cast<short>(c);
    }
    // decoding.py:349-360: walk the ranking; EOT candidates met before the beam is full are finished
    int fin_n = p.fin_count[a], n_kept = 0, n_fin = 0;
    for (int q = 0; q < n_live && n_kept < G; ++q) {
      const int c = s_order[q];
      const int tok = s_tok[c];
      const int src = r0 + s_tok_row[c];
      if (tok == p.eot) {
        if (fin_n < p.max_candidates) {                    // decoding.py:372-375
          const int slot = a * p.max_candidates + fin_n;
          p.fin_score[slot] = s_score[c];
          s_fin_src[n_fin] = src;
          s_fin_slot[n_fin] = slot;
          ++n_fin;
          ++fin_n;
        }
      } else {
        s_newsrc[n_kept] = src;
        s_newtok[n_kept] = tok;
        p.sum_logprobs[r0 + n_kept] = s_score[c];          // decoding.py:354 (safe: scores were gathered above)
        ++n_kept;
      }
    }
    s_n_fin = n_fin;
    p.fin_count[a] = fin_n;
    if (fin_n < p.max_candidates) s_all_done = 0;
  }
  __syncthreads();
  // prefix-equality matrix of the new beams
  if (tid < G * G) {
    const int j1 = tid / G, j2 = tid % G;
    const int s1 = s_newsrc[j1] - r0, s2 = s_newsrc[j2] - r0;
    p.same_out[static_cast<long long>(a) * kMaxBeam * kMaxBeam + j1 * kMaxBeam + j2] =
        (j1 == j2) ? 1 : (s_same[s1][s2] && s_newtok[j1] == s_newtok[j2]);
  }
  // work items for the warps: newly finished hypotheses (prefix of the source row + EOT), then the new beams (token
  // row + kv-cache parent table)
  const int n_fin = s_n_fin;
  for (int item = warp; item < n_fin + G; item += kBeamThreads / 32) {
    if (item < n_fin) {
      const int* srow = p.tokens_in + static_cast<long long>(s_fin_src[item]) * p.max_ctx;
      int* drow = p.fin_tokens + static_cast<long long>(s_fin_slot[item]) * p.max_ctx;
      for (int i = lane; i < L; i += 32) drow[i] = srow[i];
      if (lane == 0) {
        drow[L] = p.eot;
        p.fin_len[s_fin_slot[item]] = L + 1;
      }
    } else {
      const int j = item - n_fin;
      const int src = s_newsrc[j];
      const int* srow = p.tokens_in + static_cast<long long>(src) * p.max_ctx;
      const int* sind = p.indir_in + static_cast<long long>(src) * p.max_ctx;
      int* drow = p.tokens_out + static_cast<long long>(r0 + j) * p.max_ctx;
      int* dind = p.indir_out + static_cast<long long>(r0 + j) * p.max_ctx;
      for (int i = lane; i < L; i += 32) {
        drow[i] = srow[i];
        // position L-1 was written by row `src` in the step that produced these logits - unless it is
        // still a prompt position (first select after the prefill), which lives in the audio's row 0
        dind[i] = (i == L - 1 && i >= p.n_init) ? src : sind[i];
      }
      if (lane == 0) {
        drow[L] = s_newtok[j];
        p.source_out[r0 + j] = src;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // last CTA to arrive publishes the new length / buffer index and the completion flag: every CTA
    // read *len_ptr before taking its ticket, so bumping it here cannot race with them.
    if (s_all_done) atomicAdd(&p.tickets[1], 1);
    __threadfence();
    const int t = atomicAdd(&p.tickets[0], 1);
    if (t == p.n_audio - 1) {
      __threadfence();
      const int n_full = atomicAdd(&p.tickets[1], 0);
      p.tickets[0] = 0;
      p.tickets[1] = 0;
      *p.len_ptr = L + 1;
      *p.cur_out_ptr = p.out_index;
      if (n_full == p.n_audio) *p.done_flag = 1;               // decoding.py:377-381
    }
  }
}

// -------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------
int launch_filter_topk(const FilterParams& p, int R, cudaStream_t s) {
  if (p.K < 1 || p.K > kMaxTopK) return 50;
  ProfileScope prof(PROF_SELECT, s);
  if ((reinterpret_cast<uintptr_t>(p.logits) & 15) || (p.ld & 3)) return 50;      // rows are read with 16-byte loads
  if (R * kSelCluster <= 2 * sm_count()) {
    // few rows (one wave at two CTAs per SM): a cluster of CTAs per row
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(R * kSelCluster);
    cfg.blockDim = dim3(kSelThreads);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kSelCluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (p.K == 1) e = cudaLaunchKernelEx(&cfg, filter_topk_kernel<1, kSelCluster>, p);
    else if (p.K <= 6) e = cudaLaunchKernelEx(&cfg, filter_topk_kernel<6, kSelCluster>, p);
    else e = cudaLaunchKernelEx(&cfg, filter_topk_kernel<kMaxTopK, kSelCluster>, p);
    count_launch();
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : 51;
  }
  if (p.K == 1) filter_topk_kernel<1, 1><<<R, kSelThreads, 0, s>>>(p);
  else if (p.K <= 6) filter_topk_kernel<6, 1><<<R, kSelThreads, 0, s>>>(p);
  else filter_topk_kernel<kMaxTopK, 1><<<R, kSelThreads, 0, s>>>(p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 51;
}
int launch_no_speech(const float* logits, long long ld, int V, int no_speech, float* out, int rows,
                     int row_stride, int row_offset, cudaStream_t s) {
  no_speech_kernel<<<rows, kSelThreads, 0, s>>>(logits, ld, V, no_speech, out, row_stride, row_offset);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 52;
}
// softmax over the token range [first, first + n) of every row (everything outside is treated as masked: the
// `logits[:, mask] = -inf` of decoding.py:60-62, the `[:, :eot]` slice of timing.py:198-201).  One CTA per row;
// writes any of: the n probabilities, the arg-max token id, the probability of one given token per row.
__global__ void __launch_bounds__(256) range_softmax_kernel(const float* __restrict__ logits, long long ld, int first, int n,
                                                            float* __restrict__ probs, int* __restrict__ argmax,
                                                            const int* __restrict__ gather_tok, float* __restrict__ gather_out) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const float* row = logits + static_cast<long long>(blockIdx.x) * ld + first;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int i = tid; i < n; i += 256) {
    const float v = row[i];
    if (v > mx) {          // strided scan in increasing i: the first maximum of this thread
      mx = v;
      mi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) {      // ties -> lower id, like torch.argmax
      mx = ov;
      mi = oi;
    }
  }
  if (lane == 0) {
    s_val[warp] = mx;
    s_idx[warp] = mi;
  }
  __syncthreads();
  mx = s_val[0];
  mi = s_idx[0];
  for (int w = 1; w < 8; ++w)
    if (s_val[w] > mx || (s_val[w] == mx && s_idx[w] < mi)) {
      mx = s_val[w];
      mi = s_idx[w];
    }
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < n; i += 256) sum += expf(row[i] - mx);
  sum = warp_sum(sum);
  if (lane == 0) s_val[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < 8; ++w) sum += s_val[w];
  const float inv = 1.0f / sum;
  if (probs)
    for (int i = tid; i < n; i += 256) probs[static_cast<long long>(blockIdx.x) * n + i] = expf(row[i] - mx) * inv;
  if (tid == 0) {
    if (argmax) argmax[blockIdx.x] = first + mi;
    if (gather_tok && gather_out) {
      const int t = gather_tok[blockIdx.x] - first;
      gather_out[blockIdx.x] = (t >= 0 && t < n) ? expf(row[t] - mx) * inv : 0.f;
    }
  }
}

int launch_range_softmax(const float* logits, long long ld, int first, int n, int rows, float* probs, int* argmax,
                         const int* gather_tok, float* gather_out, cudaStream_t s) {
  if (rows <= 0 || n <= 0) return 0;
  range_softmax_kernel<<<rows, 256, 0, s>>>(logits, ld, first, n, probs, argmax, gather_tok, gather_out);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 56;
}

int launch_greedy_update(const GreedyParams& p, cudaStream_t s) {
  greedy_update_kernel<<<1, 1024, 0, s>>>(p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 53;
}
int launch_beam_update(const BeamParams& p, cudaStream_t s) {
  if (p.G > kMaxBeam) return 54;
  beam_update_kernel<<<p.n_audio, kBeamThreads, 0, s>>>(p);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? 0 : 55;
}

}  // namespace wb
