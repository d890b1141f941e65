// Launch description of the fused decoder-layer kernel (dec_layer.cu); built once per decoder session by engine.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb {

constexpr int kDLMaxPhases = 4;
constexpr int kDLUnit = 16;        // output columns per work unit (one tcgen05.ld .x16, one 16-row weight box)
constexpr int kDLMaxUnits = 12;    // widest per-CTA tile: 192 columns

enum { DL_FOLD = 1, DL_GELU = 2, DL_RESID = 4, DL_STATS = 8, DL_OUTF32 = 16 };   // OUTF32: fp32 output (the logits)
// phase kinds (few-rows form only; the tile form knows Linear phases): Linear, self-attention + kv append over the
// rows' lineages, cross-attention partials over a slice of the audio keys, merge of those partials
enum { DS_LINEAR = 0, DS_SELF = 1, DS_CROSS = 2, DS_COMBINE = 3, DS_LN = 4 };   // LN: out = LayerNorm(a) with c1 = gamma, c2 = beta

// One Linear of the chain.  DL_FOLD: the input is LayerNorm(x) - W holds W (.) gamma, c1 / c2 the fold vectors, the row
// statistics come from the LN partials left by the producer of x.  DL_RESID: out is the residual stream,
// out = round(acc + bias) + out (model.py:165-170).  DL_STATS: also leave LN partials of the rows written.
struct DLPhase {
  int N, K;
  int flags;
  int units_box;          // rows of the weight box / 16 = the widest per-CTA share of the phase
  int bm;                 // rows per row block: 64 (UMMA M = 64) or 128 (UMMA M = 128, full tensor-pipe rate; wide phases)
  int kouter;             // 1: the tensor maps are 3-D {64 k, rows, K / 64} and one request brings a whole stage of an operand
  const void* bias;       // T[N]   (phases without DL_FOLD)
  const float* c1;        // fp32 [N] (DL_FOLD)
  const float* c2;
  void* out;              // T [R, ldo]
  long long ldo;
  const void* a;          // T [R, lda] input rows and T [N, K] weights as plain pointers (few-rows form; the tile form
  long long lda;          // reads them through the tensor maps)
  const void* w;
  int type;               // DS_*
  void* kc;               // DS_SELF: this layer's self K / V caches [row][H][ctx][64]; DS_CROSS: kc = the layer's cross K/V
  void* vc;               //          block [audio][2H][T][64]
};

struct DLParams {
  int n_phases;
  int R, m_tiles;
  int ln_slots_in;        // > 0: LN partial slots left for the FIRST phase by an earlier kernel with another split (1 after embed)
  float4* ln_part;        // [slot][ln_ld] (count, mean, M2, -) partial statistics of the residual stream
  int ln_ld;
  unsigned int* sync;     // [0] grid-barrier counter, [1] exit counter; zero between launches
  const int* skip_flag;
  // optional trace (tools/bench_dec_layer.cu): [cta][phase][8] SM-clock stamps - 0 producer start, 1 first stage landed,
  // 2 last MMA committed, 3 accumulator seen by the epilogue, 4 stores done, 5 arrived on the grid barrier,
  // 6 next phase released (poller), 7 LN statistics gathered
  unsigned long long* trace;
  int dr_a_off, dr_tail_off;   // few-rows form: shared-memory offsets of the input rows and of the scratch tail
  // few-rows form, whole decoder stack in one launch: the phases come from a table in global memory (n_phases entries)
  // and include the attention of every layer
  const DLPhase* table;
  const void* qkv;             // [R, 3d] q | k | v of the new position (output of the QKV phases)
  const void* q;               // [R, d]  cross-attention queries
  void* att;                   // [R, d]  attention output
  const int* indir;            // [R, ctx] position -> physical cache row
  const int* len_ptr;          // tokens per row including the new one
  float* xpart;                // [R * H][splits][66] cross-attention partials (m, l, o[64])
  int n_head, ctx, T, G, splits, d;
  DLPhase ph[kDLMaxPhases];
};

struct alignas(64) DLMaps {
  // 3-D views {64 k, rows, K / 64}: one box = `ks` consecutive [rows x 64] k-blocks
  CUtensorMap a[kDLMaxPhases];        // activations [R, K], box 64 x 64 rows x ks
  CUtensorMap b[kDLMaxPhases];        // weights [N, K], box 64 x 16 * units_box rows x ks
};

struct DLLaunch {
  int dtype = 0;
  int grid = 0;
  int rows_smem = 0;      // > 0: launch the few-rows form (dec_rows_kernel) with this much dynamic shared memory
  DLParams p;
  DLMaps maps;
};

constexpr int kDRMaxRows = 32;     // the few-rows form covers R <= 32 (1, 2 or 4 mma.sync n = 8 operand tiles) where shared memory allows

extern int g_fused_layer;
extern int g_fused_stack;          // few-rows sessions: whole stack in one launch (wb200_set_fused_decoder_stack / WB200_FUSED_STACK: 0 off, 1 default, 2 incl. logits)
extern int g_fused_rows;           // few-rows form for R <= kDRMaxRows (wb200_set_fused_decoder_rows / WB200_FUSED_ROWS, default on)
int dl_grid_size();
bool dl_supported(int R, int d, int grid);
void dl_init_launch(DLLaunch& L, int dtype, int R, int grid, float4* ln_part, int ln_ld, unsigned int* sync,
                    const int* skip_flag, int ln_slots_in);
// appends nothing by itself: fills phase `idx` (caller sets p.n_phases)
int dl_fill_phase(DLLaunch& L, int idx, int dtype, int R, int grid, const void* A, long long lda, const void* W, int N, int K,
                  const void* bias, const float* c1, const float* c2, int flags, void* out, long long ldo, int bm = 0);
// after every phase is filled: switch the launch to the few-rows form if it applies (R <= kDRMaxRows, slab + rows fit);
// returns true when it did
bool dl_use_rows_form(DLLaunch& L);
// whole-stack launch of the few-rows form: `host_table` (n entries, Linear phases filled with dl_fill_phase-compatible
// fields, attention phases with type / kc / vc) is checked for shared-memory fit; returns false if the form does not apply
bool dl_plan_stack(DLLaunch& L, const DLPhase* host_table, int n, const DLPhase* device_table);
int dl_launch(const DLLaunch& L, cudaStream_t s);

}  // namespace wb
