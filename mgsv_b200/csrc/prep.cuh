// Internal launchers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace made {

// prep.cu
int cast_mask_rows(const void* in, int in_dtype, const float* mask, int64_t rows, int dim,
                   op_t* out, cudaStream_t st);
int layernorm_rows(const void* in, int in_is_op, int64_t ld_in, int64_t rows, const float* gamma,
                   const float* beta, op_t* out_h, float* out_f32, cudaStream_t st);
int pool_norm(const float* seq, const float* mask, int64_t B, int L, float* pooled, cudaStream_t st);
int detr_prep(const op_t* frame_out, const float* frame_mask, const op_t* seg_out,
              const float* seg_mask, const int32_t* track_idx, int64_t seq_offset, const float* inv_dim_t,
              int64_t B, op_t* src, op_t* pos, op_t* srcpos, float* mask_out,
              cudaStream_t st);
int heads_final(const float* hs, const op_t* h2, int64_t rows, const float* w_cls,
                const float* b_cls, const float* w_sp, const float* b_sp, float* logits, float* spans,
                cudaStream_t st);
int mask_bits(const float* mask, int64_t n, uint32_t* bits, cudaStream_t st);
int vhat_rows(const float* v, int64_t rows, __half* out, cudaStream_t st);

// attn.cu
int mha_core(const op_t* Q, int64_t ldq, const op_t* K, int64_t ldk,
             const op_t* V, int64_t ldv, const float* key_mask, int64_t B, int L,
             op_t* O, int64_t ldo, cudaStream_t st);
int dec_attn_folded(const float* qt, const op_t* mp, const op_t* mem, const float* key_mask, int64_t B,
                    int L, op_t* out, cudaStream_t st);

// xpool.cu
int xpool_set_constants(const float* bias_prime, const float* gamma3, const float* beta3, cudaStream_t st);
int xpool_score(const op_t* q, const __half* vhat, int64_t n_queries, const op_t* kz,
                int64_t ldkz, int z_col, const op_t* gram, const uint32_t* maskbits,
                int64_t n_tracks, float* sim, int64_t ld, int64_t col_offset, cudaStream_t st);

}  // namespace made
