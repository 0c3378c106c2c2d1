/* made_b200 — C ABI of the B200-native MaDe inference + scoring hot path.
 *
 * The reference (xxayt/MGSV) is pure Python/PyTorch, so the "FFI" a maintainer binds is ctypes
 * from the Python modules named below (see INTEGRATION.md for the stubs).  Every entry point:
 *   - takes plain device pointers (contiguous, row-major, 16-byte aligned) and sizes,
 *   - enqueues on the given CUDA stream (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) and never synchronises,
 *   - returns MADE_OK or a negative MADE_E* code; made_last_error_string() gives the text
 *     (thread-local).  The Python shim raises ValueError for MADE_EINVAL / MADE_EUNSUPPORTED and
 *     RuntimeError otherwise, matching the reference's exception conventions
 *     (model_Uni.py:275, span_utils.py:107-108).
 * All 16-bit activation/operand buffers are IEEE fp16 (the path's GEMM operand type: fp16 operands,
 * fp32 accumulation; DESIGN.md "numerics").
 * The caller owns every buffer.  A made_ctx owns packed weights and a workspace; one ctx per
 * (process, device), calls on it serialised by the caller (the reference is single-threaded).
 */
#ifndef MADE_B200_H_
#define MADE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MADE_ABI_VERSION 2

#define MADE_OK 0
#define MADE_EINVAL (-1)        /* bad shape / pointer / argument            */
#define MADE_ECUDA (-2)         /* CUDA runtime or driver error              */
#define MADE_ENOMEM (-3)        /* workspace allocation failed               */
#define MADE_EUNSUPPORTED (-4)  /* config outside the shipped MaDe config    */
#define MADE_ESTATE (-5)        /* weights not loaded / ctx misuse           */

#define MADE_DTYPE_F32 0
#define MADE_DTYPE_BF16 1
#define MADE_DTYPE_F16 2

#define MADE_VIDEO 0
#define MADE_MUSIC 1

/* Arithmetic of the similarity path (temporal encoders + X-Pool projections), DESIGN.md "numerics".
 * FP16  : every GEMM operand is one fp16 value (fastest; similarity error ~2e-4 rms of the scale).
 * SPLIT : the weights and token activations whose rounding dominates the similarity error travel as
 *         fp16 (hi, lo) pairs and their GEMMs run 2-3 tcgen05 passes into one fp32 accumulator
 *         (default; meets the 1e-3 similarity bar on the full 2000 x 4000 job).
 * FP32  : the temporal encoders run on the CUDA cores in the reference's own fp32 arithmetic (made_encode only;
 *         with made_xpool_pooled + made_pooled_cosine for the X-Pool similarity this is the "1e-5 in fp32" mode:
 *         exact, ~50x slower than SPLIT). */
#define MADE_PREC_FP16 0
#define MADE_PREC_SPLIT 1
#define MADE_PREC_FP32 2

typedef struct made_ctx made_ctx;

const char* made_last_error_string(void);
int made_abi_version(void);
/* MADE_OK iff `device` is an sm_100 part. */
int made_device_check(int device);

/* Kernel-family timing (bench.py's roofline): while enabled, the launchers of the tensor-core kernels bracket
 * every launch with CUDA events on the launching stream.  made_prof_collect synchronises the device, sums the
 * elapsed milliseconds and launch counts per family and clears the window.
 * Families: 0 = gemm_tc_kernel, 1 = ffn_fused_kernel, 2 = xpool_score_kernel, 3 = attention kernels, 4 = rank / top-k. */
#define MADE_PROF_KINDS 5
int made_prof_enable(int on);
int made_prof_collect(double* ms_by_kind, int64_t* launches_by_kind, int n_kinds);

/* ---------------------------------------------------------------------------------------------
 * span utilities — music_detr/span_utils.py, bit-exact fp32
 * ------------------------------------------------------------------------------------------- */
/* span_cw_to_se (span_utils.py:15-24): cw [n,2] -> se [n,2] */
int made_span_cw_to_se(const float* cw, float* se, int64_t n, void* stream);
/* span_se_to_cw (span_utils.py:4-13): se [n,2] -> cw [n,2] */
int made_span_se_to_cw(const float* se, float* cw, int64_t n, void* stream);
/* detr_iou + individual_IoU_tensor (span_utils.py:119-170) for n predicted spans in seconds:
 * clamp start >= 0, end <= max_m_duration then <= m_duration[i]; 0 when the ground truth is
 * degenerate or the union is <= 0.  gt_moment [n,2]. */
int made_span_iou(const float* pred_st, const float* pred_ed, const float* gt_moment, const float* m_duration,
                  float max_m_duration, int64_t n, float* iou, void* stream);
/* generalized_temporal_iou (span_utils.py:86-115): spans1_se [n,2], spans2_se [m,2] -> giou [n,m] */
int made_giou(const float* spans1_se, int64_t n, const float* spans2_se, int64_t m, float* giou,
              void* stream);
/* temporal_iou (span_utils.py:39-66): -> iou [n,m], union [n,m] */
int made_temporal_iou(const float* spans1_se, int64_t n, const float* spans2_se, int64_t m,
                      float* iou, float* uni, void* stream);
/* HungarianMatcher cost matrix (matcher.py:66-88) for (c,w) spans given the foreground softmax
 * probabilities: C = w_span*L1 + w_giou*(-gIoU) + w_class*(-p_fg), [n,m].  The caller has already
 * dropped zero-width targets (matcher.py:59). */
int made_matcher_cost(const float* prob_fg, const float* out_spans_cw, int64_t n,
                      const float* tgt_spans_cw, int64_t m, float w_span, float w_giou,
                      float w_class, float* cost, void* stream);
/* Driver post-processing (test-MaDe.py:306-316) fused with detr_iou (span_utils.py:147-170,
 * 119-145): logits [n,2], spans_cw [n,2], gt_moment [n,2] seconds, m_duration [n]
 * -> pred_st, pred_ed (seconds), score (foreground prob), iou (nullable). */
int made_moment_postproc(const float* pred_logits, const float* pred_spans_cw,
                         const float* gt_moment, const float* m_duration, float max_m_duration,
                         int64_t n, float* pred_st, float* pred_ed, float* score, float* iou,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * ranking — utils/util_test.py:32-97 (Recall_metrics, dedup=True) + test-MaDe.py:401-403
 * ------------------------------------------------------------------------------------------- */
/* Scores are double(single[r,c]) + double(dual[r,c]) (dual nullable).  Per row r:
 *   rank_out[r]  = number of distinct music ids strictly ahead of the ground truth, where the GT
 *                  score is the best column of the GT id: gt_col[r] must be the LAST column
 *                  carrying that id and prev_same[c] the previous column with the same id as c
 *                  (-1 if none; prev_same nullable = all ids distinct).  If gt_score_in is given
 *                  (multi-GPU: the GT lives on another shard) it replaces the gt_col walk.
 *   topk_idx/topk_score [n_rows,k] = exact top-k, score descending, lower column first on ties;
 *                  indices are column + col_offset. k <= 256. */
int made_rank_topk(const float* single, const float* dual, int64_t ld, int64_t n_rows,
                   int64_t n_cols, const int32_t* gt_col, const double* gt_score_in,
                   const int32_t* prev_same, int32_t col_offset, int k, int32_t* topk_idx,
                   double* topk_score, int32_t* rank_out, double* gt_score_out, void* stream);
/* Merge per-shard candidates [n_rows, n_cand] (idx -1 = empty) into the global top-k. */
int made_topk_merge(const double* cand_score, const int32_t* cand_idx, int64_t n_rows, int n_cand,
                    int k, int32_t* out_idx, double* out_score, void* stream);
/* cal_distance(..., "COS") (modules/loss.py:52-56): out[i,j] = <a_i/|a_i|, b_j/|b_j|>, fp32. */
int made_cosine_sim(const float* a, int64_t n, const float* b, int64_t m, int d, float* out,
                    int64_t ld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * model context — model/model_Uni.py:14 (Uni_model), shipped config only
 * ------------------------------------------------------------------------------------------- */
int made_ctx_create(made_ctx** out, int device);
int made_ctx_destroy(made_ctx* ctx);
/* Weights by reference state_dict key (util_train.py:51-53): n host fp32 arrays.
 * Packs fp16 GEMM operands and the folded X-Pool / decoder weights. */
int made_ctx_load_weights(made_ctx* ctx, int n, const char* const* names, const float* const* host_ptrs,
                          const int64_t* numels, void* stream);
/* MADE_PREC_* (default MADE_PREC_SPLIT); takes effect for the calls that follow. */
int made_ctx_set_precision(made_ctx* ctx, int mode);
/* Columns of a packed operand row of `dim` features in the context's precision mode (2*dim for
 * (hi | lo) pairs): the row width of made_ingest_ragged's output. */
int made_ctx_operand_width(const made_ctx* ctx, int dim);

/* ---------------------------------------------------------------------------------------------
 * Ragged (token-packed) batches.  The reference computes every zero-padded position and masks it
 * afterwards (model_Base.py:533-541); padded keys never reach a softmax and padded query rows never
 * reach an output, so this path only materialises the VALID tokens of a batch: the rows with
 * mask != 0, packed in order into a dense [total, features] matrix.
 *   seq_len[b] = valid tokens of sequence b      seq_off[b] = exclusive prefix sum of seq_len
 *   total[0]   = sum(seq_len), a DEVICE scalar   tok_src[i] = b * L + t of packed row i
 * All pointers are device memory inside the caller's idx_workspace of made_ragged_index_words(B, L)
 * int32 words; the host never learns `total` (no sync) — kernels read it from the device.
 * ------------------------------------------------------------------------------------------- */
typedef struct made_ragged {
  const int32_t* seq_len;
  const int32_t* seq_off;
  const int32_t* total;
  const int32_t* tok_src;
  int64_t B;
  int32_t L;
} made_ragged;
int64_t made_ragged_index_words(int64_t B, int L);
/* masks [B, L] float {0,1} (device) -> descriptor (3 small kernels on `stream`). */
int made_ragged_build(const float* masks, int64_t B, int L, int32_t* idx_workspace, made_ragged* out, void* stream);
/* Feature ingest = the masked_fill + cast at the top of forward_{video,audio}_encoder_feature
 * (model_Base.py:556 / :595): feats [B, L, dim] (fp32 / bf16 / fp16) -> out16_packed
 * [<= B*L, made_ctx_operand_width(ctx, dim)] fp16 holding the valid rows only (in MADE_PREC_SPLIT every row is
 * [hi | lo]: the fp16 rounding of the features and what that rounding dropped); rows with mask == 0 are NEVER READ.  `feats` may be device memory or
 * pinned host memory (unified addressing: the kernel then pulls just the valid rows over PCIe).
 * dim % 8 == 0. */
int made_ingest_ragged(made_ctx* ctx, const void* feats, int feats_dtype, const made_ragged* rb, int dim,
                       void* out16_packed, void* stream);

/* Host -> device transfer of a zero-padded feature tensor, valid rows only (the reference copies the
 * whole padded tensor per batch, test-MaDe.py:268-271).  host_feats [B, L, dim] and host_masks
 * [B, L] are HOST pointers (host_feats pinned for the copies to be asynchronous); for every
 * sequence the rows [0, last row with mask != 0] are copied to the same offsets of dev_staging
 * [B, L, dim] by the copy engines (one batched cudaMemcpyBatchAsync, no SM involved); the other
 * rows of dev_staging are left untouched and must be treated as garbage (made_encode /
 * made_ingest_ragged never read rows whose mask is 0).  bytes_copied (nullable) = bytes queued.
 * host_stage16 (nullable): a pinned host fp16 tensor of the same [B, L, dim] shape.  When given
 * (fp32 features only) n_threads host threads first round the valid rows to fp16 into it — the
 * same rounding the ingest kernel would apply on the device — and the fp16 rows are what crosses
 * PCIe; dev_staging is then an fp16 [B, L, dim] tensor.  The caller must not reuse host_stage16
 * before the copies queued here have completed. */
int made_h2d_valid_rows(const void* host_feats, int feats_dtype, const float* host_masks, int64_t B, int L,
                        int dim, void* host_stage16, int n_threads, void* dev_staging, int64_t* bytes_copied,
                        void* stream);

/* forward_{video,audio}_encoder_feature (model_Base.py:544-617): feats [B,L,Din] (fp32, bf16 or fp16; rows with mask 0 are
 * never read),
 * masks [B,L] float {0,1}; L,Din = 50,512 (MADE_VIDEO) or 96,768 (MADE_MUSIC).
 * -> seq16 [B,L,256] fp16, seq_f32 [B,L,256] (nullable), pooled [B,256] fp32 (L2-normalised). */
int made_encode(made_ctx* ctx, int modality, const void* feats, int feats_dtype, const float* masks,
                int64_t B, void* seq16, float* seq_f32, float* pooled, void* stream);
/* Same on a batch that was already ingested: x16_packed = output of made_ingest_ragged for `rb`
 * (lets the caller run the ingest of the next chunk on another stream).  Outputs keep the padded
 * [B,L,256] layout with zero rows at the padded positions (model_Base.py:541). */
int made_encode_ragged(made_ctx* ctx, int modality, const void* x16_packed, const made_ragged* rb, void* seq16,
                       float* seq_f32, float* pooled, void* stream);

/* Per-track X-Pool operands from encoded segments (modules/transformer.py:165, 102-106 folded):
 * seg16 [N,96,256], seg_masks [N,96] -> kz [N*96,768] fp16 (K | V'' | Z''),
 * gram [N*96,112] fp16 = [G = V''V''^T (96) | W5 = Z''.{1,b',g3^2,g3^2 b',g3 beta3} (5) | 0],
 * maskbits [N,4] u32. */
int made_gallery_prepare(made_ctx* ctx, const void* seg16, const float* seg_masks, int64_t N,
                         void* kz, void* gram, uint32_t* maskbits, void* stream);
/* Per-query X-Pool operands (modules/transformer.py:164, 98; metrics.py:19):
 * video_feats [N,256] fp32 -> q [N,256] fp16 (q_proj(LN1(v))/16), vhat [N,256] fp32 (v / |v|). */
int made_query_prepare(made_ctx* ctx, const float* video_feats, int64_t N, void* q, float* vhat,
                       void* stream);
/* Transformer_XA + sim_matrix_music_pooling fused (modules/transformer.py:156-180,
 * modules/metrics.py:10-24): sim[v, col_offset + m] for v < N_v, m < N_m; sim row stride ld. */
int made_xpool_score(made_ctx* ctx, const void* q, const float* vhat, int64_t n_queries,
                     const void* kz, const void* gram, const uint32_t* maskbits, int64_t n_tracks,
                     float* sim, int64_t ld, int64_t col_offset, void* stream);

/* Transformer_XA.forward MATERIALISED (modules/transformer.py:156-180), fp32 CUDA-core arithmetic:
 * which = MADE_MUSIC: model.video_guided_to_music_pooling_cross_transformer — video_feats [N_v,256] (the guides),
 *         seg_f32 [N_m,96,256] fp32, seg_masks [N_m,96] -> pooled [N_m, N_v, 256] fp32;
 * which = MADE_VIDEO: model.music_guided_to_video_pooling_cross_transformer (vmr_fusion "XA-music-video",
 *         model_Uni.py:24-28,203-204) — guides = music_feats [N_m,256] passed as `video_feats`, keys = frame features
 *         [N_v,50,256] + masks [N_v,50] passed as `seg_f32` / `seg_masks` -> pooled [N_v, N_m, 256].
 * The product path (made_xpool_score) never forms this tensor; this entry exists for callers that want it (the
 * compat view of model.video_guided_to_music_pooling_cross_transformer) and for the fp32 precision mode.
 * Chunk the tracks: scratch is N_m * N_v * 2.4 KB. */
int made_xpool_pooled(made_ctx* ctx, int which, const float* video_feats, int64_t n_q, const float* seg_f32,
                      const float* seg_masks, int64_t n_m, float* pooled, void* stream);
/* mml_fusion "CA" (model_Uni.py:209-211; CrossTransformer, model/model_Base.py:169-213, CrossAttention :93-165,
 * FeedForward :22-46): the music segments attend to the frames of the paired video before DETR.
 * seg_f32 [B,96,256] / seg_masks [B,96] = encoded segments of the B pairs (queries), frame_f32 [B,50,256] /
 * frame_masks [B,50] = encoded frames (keys, values) -> fused16 [B,96,256] fp16 (+ fused_f32, nullable), rows with
 * seg_masks == 0 written as 0 (the masked_fill of :210).  The DETR input of this variant is `fused16` alone:
 * call made_detr_detect with all-zero frame masks.  Needs a checkpoint that carries
 * video_music_fusion_cross_transformer.* (MADE_ESTATE otherwise). */
int made_ca_fuse(made_ctx* ctx, const float* seg_f32, const float* seg_masks, const float* frame_f32,
                 const float* frame_masks, int64_t B, void* fused16, float* fused_f32, void* stream);

/* sim_matrix_music_pooling (modules/metrics.py:10-24) on a materialised pooled tensor:
 * sim[v, col_offset + m] = < video[v]/|video[v]|, pooled[m,v]/|pooled[m,v]| >, fp32, sim row stride ld. */
int made_pooled_cosine(const float* video_feats, const float* pooled, int64_t n_q, int64_t n_m, float* sim, int64_t ld,
                       int64_t col_offset, void* stream);

/* Moment detection for B (query, track) pairs — model_Uni.py:207-227 + calc_output :117-150:
 * concat fusion, PositionEmbeddingSine, DETR encoder x2 / decoder x6, heads.
 * frame16 [Bv,50,256], frame_masks [Bv,50], seg16 [Nm,96,256], seg_masks [Nm,96],
 * track_idx [B] (nullable = identity) picks the track paired with query b, video_feats [B,256].
 * -> hs [6,B,256] (through decoder.norm), pred_logits [6,B,2], pred_spans [6,B,2] (sigmoid (c,w)),
 *    proj_queries [6,B,256] (nullable), proj_vid_mem [B,50,256] (nullable), memory [B,146,256]
 *    fp32 (nullable). */
int made_detr_detect(made_ctx* ctx, const void* frame16, const float* frame_masks,
                     const void* seg16, const float* seg_masks, const int32_t* track_idx,
                     const float* video_feats, int64_t B, float* hs, float* pred_logits,
                     float* pred_spans, float* proj_queries, float* proj_vid_mem, float* memory,
                     void* stream);

/* Evaluation-time loss scalars of Uni_model.forward (forward values only).
 * SetCriterion (loss_detr.py:74-169) for n_layers decoder outputs with one moment query:
 * pred_logits/pred_spans [n_layers,B,2], proj_queries [n_layers,B,256] and proj_vid_mem [B,50,256]
 * (both nullable), targets_cw [B,2]; w_fg/w_bg = criterion.empty_weight; out [n_layers,5] =
 * {loss_span, loss_giou, loss_label, class_error, loss_contrastive_align}. */
int made_detr_losses(const float* pred_logits, const float* pred_spans, const float* proj_queries,
                     const float* proj_vid_mem, const float* targets_cw, int64_t B, int n_layers,
                     float w_fg, float w_bg, float temperature, float* out, void* stream);
/* InfoNCELoss(dual) + CLIPLoss(single) (modules/loss.py:5-24,66-123; model_Uni.py:255-262) on
 * in-batch [n,n] similarity matrices with row stride ld -> out[0]. */
int made_retrieval_loss(const float* dual, const float* single, int64_t ld, int n, float logit_scale,
                        float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * building blocks exported for the parity tests
 * ------------------------------------------------------------------------------------------- */
/* C[M,N] = act(A[M,K] W[N,K]^T + bias (+ residual)) with optional LayerNorm over N (N == 256).
 * A, W fp16; bias/gamma/beta fp32 (nullable); residual fp32 [M,N] (nullable);
 * act: 0 none, 1 GELU(erf), 2 ReLU; out16 / out_f32 nullable (at least one). */
int made_gemm_f16(const void* A, const void* W, int64_t M, int N, int K, const float* bias,
                   const float* residual, int act, const float* ln_gamma, const float* ln_beta,
                   void* out16, float* out_f32, void* stream);
/* The split-precision GEMM behind MADE_PREC_SPLIT.  W [N, 2K] = rows of [hi | lo] fp16 pairs; split = 1: A [M, K]
 * plain fp16, C = A W_hi^T + A W_lo^T; split = 2: A [M, 2K] pairs too, C = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T
 * (three tcgen05 passes into one fp32 accumulator).  residual_pair [M, 2N] (nullable) = (hi | lo) pair added as
 * hi + lo; exactly one of out_pair [M, 2N] ((hi | lo) pair of the result) / out_f32 [M, N]. */
int made_gemm_f16_split(const void* A, const void* W, int64_t M, int N, int K, int split, const float* bias,
                        const void* residual_pair, int act, const float* ln_gamma, const float* ln_beta,
                        void* out_pair, float* out_f32, void* stream);
/* Same product with ONE plain fp16 output [M, N] and no row-wide epilogue: the shape of the encoder in_proj and the X-Pool
 * operand projections.  MADE_GEMM_WS128=1 selects an experimental weight-stationary form on 128-column tiles (both
 * halves of the weight pair of a tile column resident in shared memory; bit-identical results, measured slower). */
int made_gemm_f16_split_h(const void* A, const void* W, int64_t M, int N, int K, int split, const float* bias, int act,
                          void* out16, void* stream);
/* The fused feed-forward block (256 -> 1024 -> 256, hidden activation kept on chip):
 * out = [LayerNorm](act(x W1^T + b1) W2^T + b2 + residual).  x [M,256] fp16 (row stride ldx), W1 [1024,256],
 * W2 [256,1024] fp16; act 1 = GELU(erf), 2 = ReLU; residual_pair (nullable) and out_pair are rows of
 * [hi(256) | lo(256)] fp16 when has_lo, else plain [M,256] fp16 rows, with row strides res_ld / ld_out. */
int made_ffn_fused(const void* x, int64_t ldx, const void* w1, const float* b1, const void* w2, const float* b2,
                   int act, const void* residual_pair, int64_t res_ld, const float* ln_gamma, const float* ln_beta,
                   void* out_pair, int64_t ld_out, int has_lo, int64_t M, void* stream);
/* softmax(Q K^T / sqrt(32) + key mask) V for 8 heads of 32: q,k,v,out [B*L, 256] fp16. */
int made_mha_core(const void* q, const void* k, const void* v, const float* key_mask, int64_t B,
                  int L, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MADE_B200_H_ */
