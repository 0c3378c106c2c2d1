// Persistent tcgen05 GEMM with fused epilogues:  C = epilogue(A[M,K] * W[N,K]^T).
// A and W are fp16, K-contiguous ("K-major"), staged by TMA (128-byte swizzle) through a 4-stage
// mbarrier ring; accumulators are fp32 in TMEM (two stages, so the epilogue of tile i overlaps
// the MMAs of tile i+1); one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16).
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..11 = epilogue (warp w reads TMEM lanes 32*(w%4).., column half w/4).
#pragma once
#include "common.cuh"

namespace made {

struct GemmEpilogue {
  const float* bias = nullptr;          // [N]
  const float* row_table = nullptr;     // [row_mod, N] fp32, added by ((row_src ? row_src[row] : row) % row_mod)
  int row_mod = 1;
  const int32_t* row_src = nullptr;     // ragged batches: tok_src (position of a packed row = tok_src % L)
  const int32_t* h_row_idx = nullptr;   // ragged batches: out_h row r is written to row h_row_idx[r] (scatter)
  const void* residual = nullptr;       // [M, N], row stride res_ld elements
  const op_t* residual_lo = nullptr;    // fp16 residual given as a (hi, lo) pair: low halves, same stride
  int residual_f32 = 0;
  int64_t res_ld = 0;
  int act = 0;                          // 0 none, 1 GELU(erf), 2 ReLU
  const float* ln_gamma = nullptr;      // LayerNorm over the full row (needs N == 256)
  const float* ln_beta = nullptr;
  float ln_eps = 1e-5f;
  int l2norm = 0;                       // F.normalize(p=2, eps=1e-12) over the row (N == 256)
  const float* row_mask = nullptr;      // [M]; rows with mask == 0 are written as 0
  op_t* out_h = nullptr;
  int64_t ld_h = 0;
  op_t* out_lo = nullptr;               // fp16(result - float(out_h)): the low half of a (hi, lo) pair, stride ld_h
                                        // (TMA-store path: exclusive with out_f32, whose staging boxes it uses)
  float* out_f32 = nullptr;
  int64_t ld_f32 = 0;
  const op_t* add2 = nullptr;  // second output: fp16(result + add2[row, col])
  int64_t add2_ld = 0;
  op_t* out2_h = nullptr;
  int64_t ld_out2 = 0;
};

struct GemmParams {
  int64_t M = 0;
  int N = 0, K = 0;
  int m_stride = 128;   // rows between consecutive M tiles (96 for per-track batched tiles)
  int m_valid = 128;    // rows of each tile that are stored
  int b_batched = 0;    // 1: B rows start at tile_m * m_stride (per-tile B, e.g. Gram matrix)
  const int32_t* m_dev = nullptr;   // device scalar: actual row count (<= M); M then only sizes the grid / TMA map
  int tma_store = 0;                // set by the launcher: out_h / out_f32 leave through TMA bulk stores
  int split = 0;                    // fp16 (hi, lo) operand pairs, the lo halves K columns to the right of the hi halves:
                                    // 1: W = [W_hi | W_lo] (ldb >= 2K)            C = A W_hi^T + A W_lo^T
                                    // 2: also A = [A_hi | A_lo] (lda >= 2K)        C = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T
  long long* trace = nullptr;       // diagnostics (scripts/diag_gemm_trace.py): clock64 stamps of CTA 0, null in production
  int debug = 0;                    // bench-only ablations (MADE_GEMM_DEBUG): 1 no bulk stores, 2 no staging writes either,
                                    // 4 no TMEM loads (results are wrong with any bit set)
  int l2_prefetch = 0;              // the producer prefetches the A tiles of its next output tile into L2 (MADE_GEMM_L2_PREFETCH=1)
  int n_store = 0;                  // > 0: only the first n_store (< N) output columns exist; W rows >= w_rows read as
                                    // zero and the TMA store clips the rest (needs the TMA-store path, else EUNSUPPORTED)
  GemmEpilogue epi;
};

// Host launcher. A: [M, K] fp16 with row stride lda; W: [N(or M for batched), K] fp16, stride ldb.
int gemm_f16_tc(const op_t* A, int64_t lda, const op_t* W, int64_t ldb,
                 int64_t w_rows, const GemmParams& p, int block_n, cudaStream_t stream);
// whether outputs may leave through TMA bulk stores (MADE_GEMM_TMA_STORE=0 turns them off)
bool gemm_tma_store_enabled();

}  // namespace made
