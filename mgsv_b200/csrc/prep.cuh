// Internal launchers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace made {

// ragged.cu — token-packed batches (see the header comment there)
struct Ragged {
  const int32_t* seq_len = nullptr;   // [B]
  const int32_t* seq_off = nullptr;   // [B]
  const int32_t* total = nullptr;     // [1] device scalar
  const int32_t* tok_src = nullptr;   // [total] = b * L + t
  int64_t B = 0;
  int L = 0;
};
size_t ragged_index_words(int64_t B, int L);
int ragged_build(const float* mask, int64_t B, int L, int32_t* idx, Ragged* out, cudaStream_t st);
// split: out rows are [hi | lo] pairs of 2*dim fp16 (lo = what the fp16 rounding of an fp32 input dropped)
int ingest_gather(const void* in, int in_dtype, const Ragged& rb, int dim, op_t* out, bool split, cudaStream_t st);
int pool_norm_ragged(const float* seq_packed, const Ragged& rb, float* pooled, cudaStream_t st);
int scatter_rows_f32(const float* packed, const Ragged& rb, float* padded, cudaStream_t st);
int scatter_rows_f32_nozero(const float* packed, const Ragged& rb, float* padded, cudaStream_t st);
// padded[tok_src[i], :] = float(hi[i, :]) + float(lo[i, :])  (256-wide fp16 (hi, lo) pair rows; lo nullable)
int scatter_rows_pair_nozero(const op_t* hi, const op_t* lo, const Ragged& rb, float* padded, cudaStream_t st);
int offset_rows(const Ragged& rb, int32_t base, int32_t* row_off, int32_t* row_len, cudaStream_t st);
int detr_mask(const float* frame_mask, const float* seg_mask, const int32_t* track_idx, int64_t seq_offset,
              int64_t B, float* mask_out, cudaStream_t st);
int detr_prep_ragged(const op_t* frame_out, const op_t* seg_out, const int32_t* track_idx, int64_t seq_offset,
                     const Ragged& rb, const float* inv_dim_t, op_t* src, op_t* pos, op_t* srcpos,
                     cudaStream_t st);

// prep.cu
// out_h [rows, ld_out] fp16 (nullable), out_lo (nullable) = fp16(result - float(out_h)) with the same stride,
// out_f32 [rows, 256] (nullable)
int layernorm_rows(const void* in, int in_is_op, int64_t ld_in, int64_t rows, const float* gamma,
                   const float* beta, op_t* out_h, int64_t ld_out, op_t* out_lo, float* out_f32, cudaStream_t st);
int heads_final(const float* hs, const op_t* h2, int64_t rows, const float* w_cls,
                const float* b_cls, const float* w_sp, const float* b_sp, float* logits, float* spans,
                cudaStream_t st);
int mask_bits(const float* mask, int64_t n, uint32_t* bits, cudaStream_t st);
int vhat_rows(const float* v, int64_t rows, float* out, cudaStream_t st);

// attn.cu
int mha_core(const op_t* Q, int64_t ldq, const op_t* K, int64_t ldk,
             const op_t* V, int64_t ldv, const float* key_mask, int64_t B, int L,
             op_t* O, int64_t ldo, cudaStream_t st, const int32_t* seq_off = nullptr,
             const int32_t* seq_len = nullptr);
int dec_attn_folded(const float* qt, const op_t* mp, const op_t* mem, const float* key_mask, int64_t B,
                    int L, op_t* out, cudaStream_t st, const int32_t* seq_off = nullptr,
                    const int32_t* seq_len = nullptr);

// ffn_fused.cu — out = epilogue(act(x W1^T + b1) W2^T + b2 + residual), hidden activation kept on chip
int ffn_fused(const op_t* x, int64_t ldx, const op_t* w1, const float* b1, const op_t* w2, const float* b2, int act,
              const op_t* res_hi, const op_t* res_lo, int64_t res_ld, const float* ln_gamma, const float* ln_beta,
              op_t* out_hi, op_t* out_lo, int64_t ld_out, const op_t* add2, int64_t add2_ld, op_t* out2,
              int64_t ld_out2, int64_t M, const int32_t* m_dev, cudaStream_t st);

// ca_fusion.cu — cross-attention core of the mml_fusion "CA" variant (model_Base.py:127-165)
int ca_attention(const op_t* q, const op_t* kv, const float* q_mask, const float* kv_mask, int64_t B, op_t* out,
                 cudaStream_t st);

// exact_f32.cu — fp32 CUDA-core path (MADE_PREC_FP32 and the materialised Transformer_XA output)
struct ExactEncW {      // raw fp32 weights of one temporal encoder (device pointers)
  int L = 0, din = 0;
  const float *proj_w = nullptr, *proj_b = nullptr, *pe = nullptr, *ln1_g = nullptr, *ln1_b = nullptr, *in_w = nullptr,
              *in_b = nullptr, *out_w = nullptr, *out_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr, *ff1_w = nullptr,
              *ff1_b = nullptr, *ff2_w = nullptr, *ff2_b = nullptr, *fin_w = nullptr, *fin_b = nullptr;
};
struct ExactXpW {       // raw fp32 weights of Transformer_XA
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr, *ln3_g = nullptr, *ln3_b = nullptr,
              *q_w = nullptr, *q_b = nullptr, *k_w = nullptr, *k_b = nullptr, *v_w = nullptr, *v_b = nullptr, *o_w = nullptr,
              *o_b = nullptr, *l_w = nullptr, *l_b = nullptr;
};
size_t exact_encode_ws_floats(int64_t B, int L, int din);
int exact_encode(const ExactEncW& w, const void* feats, int feats_dtype, const float* masks, int64_t B, float* ws,
                 float* seq_f32, float* pooled, cudaStream_t st);
size_t exact_xpool_ws_floats(int64_t n_q, int64_t n_m, int L);
int exact_xpool(const ExactXpW& w, const float* video, int64_t n_q, const float* seg, const float* seg_mask, int64_t n_m,
                int L, float* ws, float* pooled, cudaStream_t st);
int exact_pooled_cosine(const float* video, const float* pooled, int64_t n_q, int64_t n_m, float* sim, int64_t ld,
                        int64_t col0, cudaStream_t st);

// xpool.cu
struct XpoolConsts {    // folded X-Pool constants of one checkpoint (per made_ctx)
  float bias[256];      // b' = (I + Wl) beta2 + bl
  float gamma[256];     // gamma3
  float gamma2[256];    // gamma3^2
  float beta[256];      // beta3
  // sums over the 256 features
  float B1, B2;         // sum b', sum b'^2
  float G2, G2b2, G2b;  // sum g^2, sum g^2 b'^2, sum g^2 b'
  float Gbb, Gb, Bb;    // sum g beta b', sum g beta, sum beta^2
};
void xpool_fill_constants(const float* bias_prime, const float* gamma3, const float* beta3, XpoolConsts* h, float* c5);
int xpool_w5(const op_t* z, int64_t ldz, int64_t rows, const float* c5_dev, op_t* gw, cudaStream_t st);
int xpool_score(const XpoolConsts& consts, const op_t* q, const float* vhat, int64_t n_queries, const op_t* kz,
                int64_t ldkz, int z_col, const op_t* gram, const uint32_t* maskbits,
                int64_t n_tracks, float* sim, int64_t ld, int64_t col_offset, cudaStream_t st);

}  // namespace made
