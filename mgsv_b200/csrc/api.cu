// C ABI entry points that own state: the model context (packed weights + workspace) and the
// host-side orchestration of the encoder / X-Pool / DETR kernel sequences.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "prep.cuh"

using namespace made;

namespace {

constexpr int D = 256, DFF = 1024, LV = 50, LM = 96, LD = 146, NENC = 2, NDEC = 6;

uint16_t f2h(float f) {  // fp32 -> IEEE fp16 bits, round-to-nearest-even, saturating to +-65504
  if (f != f) return 0x7E00;
  const float lim = 65504.0f;
  f = f > lim ? lim : (f < -lim ? -lim : f);
  const __half h = __float2half_rn(f);   // host path of cuda_fp16.h: RNE incl. subnormals
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

struct Lin {            // y = x W^T + b, W [out,in] fp16 on device, b fp32
  op_t* w = nullptr;
  op_t* w2 = nullptr;   // [out, 2*in] = [fp16(W) | fp16(W - fp16(W))]: the (hi, lo) pair of the split-precision GEMMs
  float* b = nullptr;
};
struct LNp {
  float* g = nullptr;
  float* b = nullptr;
};

struct EncW {
  int L = 0, din = 0;
  Lin proj, in_proj, out_proj, ff1, ff2, fin;
  LNp ln1, ln2;
  float* pe = nullptr;  // [L,256]
};
struct DetrEncW {
  Lin qk, v, out, ff1, ff2;
  LNp n1, n2;
};
struct DetrDecW {
  Lin sa;        // folded single-query self-attention
  Lin qfold;     // [8*256, 256]: per head Wk_h^T Wq_h / sqrt(32)  -> q~ (scores = q~_h . (mem+pos)_t)
  Lin ovfold;    // [256, 8*256]: per head Wo[:, h] Wv_h            -> out = W_ov vec(mbar) + b_ov
  Lin ff1, ff2;
  LNp n1, n2, n3;
};

}  // namespace

struct made_ctx {
  int device = 0;
  bool loaded = false;
  std::unordered_map<std::string, std::vector<float>> host;
  std::vector<void*> allocs;
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0, ws_off = 0;
  bool ws_dry = false;          // sizing pass of with_arena(): take() only advances the offset
  int precision = MADE_PREC_SPLIT;
  ExactEncW xenc[2];            // raw fp32 weights for the fp32 CUDA-core path (exact_f32.cu)
  ExactXpW xxp;                 // video_guided_to_music_pooling_cross_transformer (keys = 96 segments)
  ExactXpW xxp_video;           // music_guided_to_video_pooling_cross_transformer (vmr_fusion "XA-music-video"; keys = 50 frames)
  bool has_xp_video = false;
  struct CaW {                  // video_music_fusion_cross_transformer (mml_fusion "CA", model_Base.py:169-213)
    Lin to_q, to_kv, to_out, ff1, ff2, fin;     // to_q / to_kv have no bias (b stays null)
    LNp ln_q, ln_c, ln_ff;
  } ca;
  bool has_ca = false;
  XpoolConsts xp_consts;        // folded X-Pool constants of THIS context's checkpoint
  float* xp_c5 = nullptr;       // [5][256] weight vectors of the W5 columns (device)

  EncW enc[2];
  LNp xp_ln1;
  Lin xp_q, xp_kvz;
  DetrEncW denc[NENC];
  DetrDecW ddec[NDEC];
  LNp dec_norm;
  Lin span0, span1, pq, pv;
  float *span2_w = nullptr, *span2_b = nullptr, *cls_w = nullptr, *cls_b = nullptr;
  float* inv_dim_t = nullptr;

  // ---- workspace arena ----
  int reserve(size_t bytes) {
    ws_off = 0;
    if (bytes <= ws_bytes) return MADE_OK;
    if (ws) cudaFree(ws);
    ws = nullptr;
    ws_bytes = 0;
    size_t want = bytes + (bytes >> 3);
    cudaError_t e = cudaMalloc(&ws, want);
    if (e != cudaSuccess) {
      set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return MADE_ENOMEM;
    }
    ws_bytes = want;
    return MADE_OK;
  }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = ws_dry ? nullptr : reinterpret_cast<T*>(ws + ws_off);
    ws_off += bytes;
    return p;
  }
  // Lays out a call's scratch buffers: `layout` (a sequence of take() calls) runs once as a sizing pass and
  // once for real, so the reservation can never drift from what is taken.
  template <typename F>
  int with_arena(F&& layout) {
    ws_dry = true;
    ws_off = 0;
    layout();
    const size_t need = ws_off;
    ws_dry = false;
    int rc = reserve(need);
    if (rc != MADE_OK) return rc;
    layout();
    if (ws_off > ws_bytes) {     // cannot happen: same take() sequence as the sizing pass
      set_error("workspace arena overrun: %zu > %zu bytes", ws_off, ws_bytes);
      return MADE_ESTATE;
    }
    return MADE_OK;
  }
};

namespace {

size_t padded(size_t count, size_t elem) { return (count * elem + 255) & ~size_t(255); }

// ---- weight upload helpers -------------------------------------------------------------------
int get(made_ctx* c, const std::string& key, size_t numel, const std::vector<float>** out) {
  auto it = c->host.find(key);
  if (it == c->host.end()) {
    set_error("load_weights: missing state_dict key '%s'", key.c_str());
    return MADE_EINVAL;
  }
  if (it->second.size() != numel) {
    set_error("load_weights: key '%s' has %zu elements, expected %zu", key.c_str(), it->second.size(), numel);
    return MADE_EINVAL;
  }
  *out = &it->second;
  return MADE_OK;
}

int up_f32(made_ctx* c, const float* src, size_t n, float** dst) {
  void* p = nullptr;
  MADE_CUDA(cudaMalloc(&p, n * 4));
  c->allocs.push_back(p);
  MADE_CUDA(cudaMemcpy(p, src, n * 4, cudaMemcpyHostToDevice));
  *dst = static_cast<float*>(p);
  return MADE_OK;
}

int up_op(made_ctx* c, const float* src, size_t n, op_t** dst) {
  std::vector<uint16_t> tmp(n);
  for (size_t i = 0; i < n; ++i) tmp[i] = f2h(src[i]);
  void* p = nullptr;
  MADE_CUDA(cudaMalloc(&p, n * 2));
  c->allocs.push_back(p);
  MADE_CUDA(cudaMemcpy(p, tmp.data(), n * 2, cudaMemcpyHostToDevice));
  *dst = static_cast<op_t*>(p);
  return MADE_OK;
}

int up_lin(made_ctx* c, const float* w, const float* b, size_t out_f, size_t in_f, Lin* lin, bool split = false) {
  MADE_TRY(up_op(c, w, out_f * in_f, &lin->w));
  MADE_TRY(up_f32(c, b, out_f, &lin->b));
  if (split) {     // rows of [hi | lo]: W = hi + lo to ~22 bits
    std::vector<uint16_t> tmp(out_f * in_f * 2);
    for (size_t r = 0; r < out_f; ++r)
      for (size_t k = 0; k < in_f; ++k) {
        const float x = w[r * in_f + k];
        const uint16_t hi = f2h(x);
        __half hh;
        memcpy(&hh, &hi, 2);
        tmp[r * 2 * in_f + k] = hi;
        tmp[r * 2 * in_f + in_f + k] = f2h(x - __half2float(hh));
      }
    void* p = nullptr;
    MADE_CUDA(cudaMalloc(&p, tmp.size() * 2));
    c->allocs.push_back(p);
    MADE_CUDA(cudaMemcpy(p, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
    lin->w2 = static_cast<op_t*>(p);
  }
  return MADE_OK;
}

int load_lin(made_ctx* c, const std::string& prefix, size_t out_f, size_t in_f, Lin* lin,
             const char* wname = ".weight", const char* bname = ".bias", bool split = false) {
  const std::vector<float>*w, *b;
  MADE_TRY(get(c, prefix + wname, out_f * in_f, &w));
  MADE_TRY(get(c, prefix + bname, out_f, &b));
  return up_lin(c, w->data(), b->data(), out_f, in_f, lin, split);
}

int load_ln(made_ctx* c, const std::string& prefix, LNp* ln) {
  const std::vector<float>*g, *b;
  MADE_TRY(get(c, prefix + ".weight", D, &g));
  MADE_TRY(get(c, prefix + ".bias", D, &b));
  MADE_TRY(up_f32(c, g->data(), D, &ln->g));
  MADE_TRY(up_f32(c, b->data(), D, &ln->b));
  return MADE_OK;
}

// C[n x m] = A[n x k] * B[k x m] in double
std::vector<double> matmul(const std::vector<double>& A, const std::vector<double>& B, int n, int k, int m) {
  std::vector<double> C(static_cast<size_t>(n) * m, 0.0);
  for (int i = 0; i < n; ++i)
    for (int kk = 0; kk < k; ++kk) {
      const double a = A[static_cast<size_t>(i) * k + kk];
      const double* br = &B[static_cast<size_t>(kk) * m];
      double* cr = &C[static_cast<size_t>(i) * m];
      for (int j = 0; j < m; ++j) cr[j] += a * br[j];
    }
  return C;
}
std::vector<double> to_d(const float* p, size_t n) { return std::vector<double>(p, p + n); }
std::vector<float> to_f(const std::vector<double>& v) { return std::vector<float>(v.begin(), v.end()); }

int up_key(made_ctx* c, const std::string& key, size_t numel, const float** dst) {
  const std::vector<float>* v;
  MADE_TRY(get(c, key, numel, &v));
  float* p = nullptr;
  MADE_TRY(up_f32(c, v->data(), numel, &p));
  *dst = p;
  return MADE_OK;
}

int load_exact(made_ctx* c) {
  for (int modality = 0; modality < 2; ++modality) {
    ExactEncW& w = c->xenc[modality];
    const bool vid = modality == MADE_VIDEO;
    w.L = vid ? LV : LM;
    w.din = vid ? 512 : 768;
    const std::string tr = vid ? "video_transformer" : "audio_transformer";
    const std::string pj = vid ? "vit_proj" : "ast_proj";
    MADE_TRY(up_key(c, pj + ".weight", static_cast<size_t>(D) * w.din, &w.proj_w));
    MADE_TRY(up_key(c, pj + ".bias", D, &w.proj_b));
    w.pe = c->enc[modality].pe;
    w.ln1_g = c->enc[modality].ln1.g; w.ln1_b = c->enc[modality].ln1.b;
    w.ln2_g = c->enc[modality].ln2.g; w.ln2_b = c->enc[modality].ln2.b;
    MADE_TRY(up_key(c, tr + ".layers.0.1.in_proj_weight", 3 * D * D, &w.in_w));
    MADE_TRY(up_key(c, tr + ".layers.0.1.in_proj_bias", 3 * D, &w.in_b));
    MADE_TRY(up_key(c, tr + ".layers.0.1.out_proj.weight", D * D, &w.out_w));
    MADE_TRY(up_key(c, tr + ".layers.0.1.out_proj.bias", D, &w.out_b));
    MADE_TRY(up_key(c, tr + ".layers.0.3.0.weight", DFF * D, &w.ff1_w));
    MADE_TRY(up_key(c, tr + ".layers.0.3.0.bias", DFF, &w.ff1_b));
    MADE_TRY(up_key(c, tr + ".layers.0.3.3.weight", D * DFF, &w.ff2_w));
    MADE_TRY(up_key(c, tr + ".layers.0.3.3.bias", D, &w.ff2_b));
    MADE_TRY(up_key(c, tr + ".final_linear.weight", D * D, &w.fin_w));
    MADE_TRY(up_key(c, tr + ".final_linear.bias", D, &w.fin_b));
  }
  auto load_xa = [&](const std::string& x, ExactXpW& w) -> int {
    MADE_TRY(up_key(c, x + ".layer_norm1.weight", D, &w.ln1_g));
    MADE_TRY(up_key(c, x + ".layer_norm1.bias", D, &w.ln1_b));
    MADE_TRY(up_key(c, x + ".layer_norm2.weight", D, &w.ln2_g));
    MADE_TRY(up_key(c, x + ".layer_norm2.bias", D, &w.ln2_b));
    MADE_TRY(up_key(c, x + ".layer_norm3.weight", D, &w.ln3_g));
    MADE_TRY(up_key(c, x + ".layer_norm3.bias", D, &w.ln3_b));
    MADE_TRY(up_key(c, x + ".cross_attn.q_proj.weight", D * D, &w.q_w));
    MADE_TRY(up_key(c, x + ".cross_attn.q_proj.bias", D, &w.q_b));
    MADE_TRY(up_key(c, x + ".cross_attn.k_proj.weight", D * D, &w.k_w));
    MADE_TRY(up_key(c, x + ".cross_attn.k_proj.bias", D, &w.k_b));
    MADE_TRY(up_key(c, x + ".cross_attn.v_proj.weight", D * D, &w.v_w));
    MADE_TRY(up_key(c, x + ".cross_attn.v_proj.bias", D, &w.v_b));
    MADE_TRY(up_key(c, x + ".cross_attn.out_proj.weight", D * D, &w.o_w));
    MADE_TRY(up_key(c, x + ".cross_attn.out_proj.bias", D, &w.o_b));
    MADE_TRY(up_key(c, x + ".linear_proj.weight", D * D, &w.l_w));
    MADE_TRY(up_key(c, x + ".linear_proj.bias", D, &w.l_b));
    return MADE_OK;
  };
  MADE_TRY(load_xa("video_guided_to_music_pooling_cross_transformer", c->xxp));
  // the second X-Pool module exists only in checkpoints trained with vmr_fusion "XA-music-video" (model_Uni.py:24-28)
  c->has_xp_video = c->host.count("music_guided_to_video_pooling_cross_transformer.linear_proj.weight") != 0;
  if (c->has_xp_video) MADE_TRY(load_xa("music_guided_to_video_pooling_cross_transformer", c->xxp_video));
  return MADE_OK;
}

// mml_fusion "CA": the cross-attention fusion block, present only in checkpoints trained with that flag
int load_ca(made_ctx* c) {
  const std::string x = "video_music_fusion_cross_transformer";
  c->has_ca = c->host.count(x + ".final_linear.weight") != 0;
  if (!c->has_ca) return MADE_OK;
  const std::vector<float>* w;
  MADE_TRY(get(c, x + ".layers.0.0.to_q.weight", static_cast<size_t>(4 * D) * D, &w));
  MADE_TRY(up_op(c, w->data(), w->size(), &c->ca.to_q.w));
  MADE_TRY(get(c, x + ".layers.0.0.to_kv.weight", static_cast<size_t>(8 * D) * D, &w));
  MADE_TRY(up_op(c, w->data(), w->size(), &c->ca.to_kv.w));
  MADE_TRY(load_lin(c, x + ".layers.0.0.to_out.0", D, 4 * D, &c->ca.to_out));
  MADE_TRY(load_lin(c, x + ".layers.0.1.net.0", DFF, D, &c->ca.ff1));
  MADE_TRY(load_lin(c, x + ".layers.0.1.net.3", D, DFF, &c->ca.ff2));
  MADE_TRY(load_lin(c, x + ".final_linear", D, D, &c->ca.fin));
  MADE_TRY(load_ln(c, x + ".attention_query_layer_norms.0", &c->ca.ln_q));
  MADE_TRY(load_ln(c, x + ".attention_context_layer_norms.0", &c->ca.ln_c));
  MADE_TRY(load_ln(c, x + ".ff_layer_norms.0", &c->ca.ln_ff));
  return MADE_OK;
}

int load_encoder(made_ctx* c, int modality) {
  EncW& e = c->enc[modality];
  const bool vid = modality == MADE_VIDEO;
  e.L = vid ? LV : LM;
  e.din = vid ? 512 : 768;
  const std::string tr = vid ? "video_transformer" : "audio_transformer";
  // the four linears whose operand rounding dominates the similarity error carry (hi, lo) weight pairs
  // (tests/tools/precision_pipeline.py); the FF branch is a small perturbation of the residual stream
  MADE_TRY(load_lin(c, vid ? "vit_proj" : "ast_proj", D, e.din, &e.proj, ".weight", ".bias", true));
  MADE_TRY(load_ln(c, tr + ".layers.0.0", &e.ln1));
  MADE_TRY(load_lin(c, tr + ".layers.0.1", 3 * D, D, &e.in_proj, ".in_proj_weight", ".in_proj_bias", true));
  MADE_TRY(load_lin(c, tr + ".layers.0.1.out_proj", D, D, &e.out_proj, ".weight", ".bias", true));
  MADE_TRY(load_ln(c, tr + ".layers.0.2", &e.ln2));
  MADE_TRY(load_lin(c, tr + ".layers.0.3.0", DFF, D, &e.ff1));
  MADE_TRY(load_lin(c, tr + ".layers.0.3.3", D, DFF, &e.ff2));
  MADE_TRY(load_lin(c, tr + ".final_linear", D, D, &e.fin, ".weight", ".bias", true));
  const std::vector<float>* pe;
  const size_t pe_len = vid ? 250 : 300;
  MADE_TRY(get(c, vid ? "video_position_embedding.pe" : "audio_position_embedding.pe", pe_len * D, &pe));
  MADE_TRY(up_f32(c, pe->data(), static_cast<size_t>(e.L) * D, &e.pe));  // first L rows (model_Base.py:533)
  return MADE_OK;
}

int load_xpool(made_ctx* c, cudaStream_t st) {
  const std::string x = "video_guided_to_music_pooling_cross_transformer";
  MADE_TRY(load_ln(c, x + ".layer_norm1", &c->xp_ln1));
  const std::vector<float>*wq, *bq, *wk, *bk, *wv, *bv, *wo, *bo, *wl, *bl, *g2, *b2, *g3, *b3;
  MADE_TRY(get(c, x + ".cross_attn.q_proj.weight", D * D, &wq));
  MADE_TRY(get(c, x + ".cross_attn.q_proj.bias", D, &bq));
  MADE_TRY(get(c, x + ".cross_attn.k_proj.weight", D * D, &wk));
  MADE_TRY(get(c, x + ".cross_attn.k_proj.bias", D, &bk));
  MADE_TRY(get(c, x + ".cross_attn.v_proj.weight", D * D, &wv));
  MADE_TRY(get(c, x + ".cross_attn.v_proj.bias", D, &bv));
  MADE_TRY(get(c, x + ".cross_attn.out_proj.weight", D * D, &wo));
  MADE_TRY(get(c, x + ".cross_attn.out_proj.bias", D, &bo));
  MADE_TRY(get(c, x + ".linear_proj.weight", D * D, &wl));
  MADE_TRY(get(c, x + ".linear_proj.bias", D, &bl));
  MADE_TRY(get(c, x + ".layer_norm2.weight", D, &g2));
  MADE_TRY(get(c, x + ".layer_norm2.bias", D, &b2));
  MADE_TRY(get(c, x + ".layer_norm3.weight", D, &g3));
  MADE_TRY(get(c, x + ".layer_norm3.bias", D, &b3));
  // q = (Wq LN1(v) + bq) / sqrt(256)   (modules/transformer.py:98,111) — exact power-of-two scale
  {
    std::vector<float> w(D * D), b(D);
    for (int i = 0; i < D * D; ++i) w[i] = (*wq)[i] * 0.0625f;
    for (int i = 0; i < D; ++i) b[i] = (*bq)[i] * 0.0625f;
    MADE_TRY(up_lin(c, w.data(), b.data(), D, D, &c->xp_q, true));
  }
  // V'' = centre_d( Wo (Wv s + bv) + bo ),  Z'' = W' V'',  W' = (I + Wl) diag(gamma2)
  std::vector<double> Wo = to_d(wo->data(), D * D), Wv = to_d(wv->data(), D * D);
  std::vector<double> Wvo = matmul(Wo, Wv, D, D, D);
  std::vector<double> bvo(D, 0.0);
  for (int i = 0; i < D; ++i) {
    double s = (*bo)[i];
    for (int k = 0; k < D; ++k) s += Wo[static_cast<size_t>(i) * D + k] * (*bv)[k];
    bvo[i] = s;
  }
  for (int j = 0; j < D; ++j) {
    double m = 0;
    for (int i = 0; i < D; ++i) m += Wvo[static_cast<size_t>(i) * D + j];
    m /= D;
    for (int i = 0; i < D; ++i) Wvo[static_cast<size_t>(i) * D + j] -= m;
  }
  {
    double m = 0;
    for (int i = 0; i < D; ++i) m += bvo[i];
    m /= D;
    for (int i = 0; i < D; ++i) bvo[i] -= m;
  }
  std::vector<double> Wp(static_cast<size_t>(D) * D);
  std::vector<float> bprime(D);
  for (int n = 0; n < D; ++n) {
    double s = (*b2)[n] + (*bl)[n];
    for (int d = 0; d < D; ++d) {
      Wp[static_cast<size_t>(n) * D + d] = ((n == d ? 1.0 : 0.0) + (*wl)[static_cast<size_t>(n) * D + d]) * (*g2)[d];
      s += static_cast<double>((*wl)[static_cast<size_t>(n) * D + d]) * (*b2)[d];
    }
    bprime[n] = static_cast<float>(s);
  }
  std::vector<double> Wz = matmul(Wp, Wvo, D, D, D);
  std::vector<double> bz(D, 0.0);
  for (int n = 0; n < D; ++n) {
    double s = 0;
    for (int d = 0; d < D; ++d) s += Wp[static_cast<size_t>(n) * D + d] * bvo[d];
    bz[n] = s;
  }
  std::vector<float> w(static_cast<size_t>(3) * D * D), b(3 * D);
  memcpy(w.data(), wk->data(), sizeof(float) * D * D);
  for (int i = 0; i < D * D; ++i) {
    w[static_cast<size_t>(D) * D + i] = static_cast<float>(Wvo[i]);
    w[static_cast<size_t>(2) * D * D + i] = static_cast<float>(Wz[i]);
  }
  for (int i = 0; i < D; ++i) {
    b[i] = (*bk)[i];
    b[D + i] = static_cast<float>(bvo[i]);
    b[2 * D + i] = static_cast<float>(bz[i]);
  }
  MADE_TRY(up_lin(c, w.data(), b.data(), 3 * D, D, &c->xp_kvz, true));
  // folded constants live in the context (a second context with another checkpoint keeps its own)
  std::vector<float> c5(5 * D);
  xpool_fill_constants(bprime.data(), g3->data(), b3->data(), &c->xp_consts, c5.data());
  MADE_TRY(up_f32(c, c5.data(), c5.size(), &c->xp_c5));
  (void)st;
  return MADE_OK;
}

int load_detr(made_ctx* c) {
  for (int l = 0; l < NENC; ++l) {
    const std::string p = "detr_transformer.encoder.layers." + std::to_string(l);
    const std::vector<float>*w, *b;
    MADE_TRY(get(c, p + ".self_attn.in_proj_weight", 3 * D * D, &w));
    MADE_TRY(get(c, p + ".self_attn.in_proj_bias", 3 * D, &b));
    MADE_TRY(up_lin(c, w->data(), b->data(), 2 * D, D, &c->denc[l].qk));
    MADE_TRY(up_lin(c, w->data() + 2 * D * D, b->data() + 2 * D, D, D, &c->denc[l].v));
    MADE_TRY(load_lin(c, p + ".self_attn.out_proj", D, D, &c->denc[l].out));
    MADE_TRY(load_lin(c, p + ".linear1", DFF, D, &c->denc[l].ff1));
    MADE_TRY(load_lin(c, p + ".linear2", D, DFF, &c->denc[l].ff2));
    MADE_TRY(load_ln(c, p + ".norm1", &c->denc[l].n1));
    MADE_TRY(load_ln(c, p + ".norm2", &c->denc[l].n2));
  }
  const std::vector<float>* qpos;
  MADE_TRY(get(c, "decoder_query_embed.weight", D, &qpos));
  constexpr int H = 8, DH = 32;
  for (int l = 0; l < NDEC; ++l) {
    const std::string p = "detr_transformer.decoder.layers." + std::to_string(l);
    const std::vector<float>*sw, *sb, *so, *sob, *cw, *cb, *ow, *ob;
    MADE_TRY(get(c, p + ".self_attn.in_proj_weight", 3 * D * D, &sw));
    MADE_TRY(get(c, p + ".self_attn.in_proj_bias", 3 * D, &sb));
    MADE_TRY(get(c, p + ".self_attn.out_proj.weight", D * D, &so));
    MADE_TRY(get(c, p + ".self_attn.out_proj.bias", D, &sob));
    // single query: softmax over one key = 1, so SA(tgt) = Wo (Wv tgt + bv) + bo  
    std::vector<double> Wsa = matmul(to_d(so->data(), D * D), to_d(sw->data() + 2 * D * D, D * D), D, D, D);
    std::vector<float> bsa(D);
    for (int i = 0; i < D; ++i) {
      double s = (*sob)[i];
      for (int k = 0; k < D; ++k) s += static_cast<double>((*so)[static_cast<size_t>(i) * D + k]) * (*sb)[2 * D + k];
      bsa[i] = static_cast<float>(s);
    }
    std::vector<float> wsa = to_f(Wsa);
    MADE_TRY(up_lin(c, wsa.data(), bsa.data(), D, D, &c->ddec[l].sa));
    // ---- cross-attention with ONE query per sequence, K/V projections folded away ----
    // scores_h[t] = q_h . (Wk_h mp_t + bk_h) / sqrt(32) = (Wk_h^T q_h / sqrt(32)) . mp_t + const_h
    //   (const_h is the same for every key t and cancels in the softmax), with
    //   q = Wq (tgt + query_pos) + bq  and  mp = memory + pos;
    // out = Wo concat_h(Wv_h mbar_h + bv_h) + bo,  mbar_h = sum_t a_{h,t} memory_t.
    MADE_TRY(get(c, p + ".multihead_attn.in_proj_weight", 3 * D * D, &cw));
    MADE_TRY(get(c, p + ".multihead_attn.in_proj_bias", 3 * D, &cb));
    MADE_TRY(get(c, p + ".multihead_attn.out_proj.weight", D * D, &ow));
    MADE_TRY(get(c, p + ".multihead_attn.out_proj.bias", D, &ob));
    const float* Wq = cw->data();
    const float* Wk = cw->data() + D * D;
    const float* Wv = cw->data() + 2 * D * D;
    const double inv_sqrt_dh = 1.0 / std::sqrt(static_cast<double>(DH));
    std::vector<float> wqf(static_cast<size_t>(H) * D * D), bqf(static_cast<size_t>(H) * D);
    std::vector<float> wov(static_cast<size_t>(D) * H * D), bov(D);
    std::vector<double> qb(D);   // Wq qpos + bq
    for (int i = 0; i < D; ++i) {
      double s = (*cb)[i];
      for (int k = 0; k < D; ++k) s += static_cast<double>(Wq[static_cast<size_t>(i) * D + k]) * (*qpos)[k];
      qb[i] = s;
    }
    for (int h = 0; h < H; ++h) {
      for (int d = 0; d < D; ++d) {          // output row h*256+d of W~q: sum_j Wk[h*32+j, d] * Wq[h*32+j, :]
        float* row = &wqf[(static_cast<size_t>(h) * D + d) * D];
        std::vector<double> acc(D, 0.0);
        double bacc = 0.0;
        for (int j = 0; j < DH; ++j) {
          const double kjd = Wk[static_cast<size_t>(h * DH + j) * D + d];
          const float* wqrow = Wq + static_cast<size_t>(h * DH + j) * D;
          for (int e = 0; e < D; ++e) acc[e] += kjd * wqrow[e];
          bacc += kjd * qb[h * DH + j];
        }
        for (int e = 0; e < D; ++e) row[e] = static_cast<float>(acc[e] * inv_sqrt_dh);
        bqf[static_cast<size_t>(h) * D + d] = static_cast<float>(bacc * inv_sqrt_dh);
      }
      for (int i = 0; i < D; ++i) {          // W_ov[i, h*256+d] = sum_j Wo[i, h*32+j] * Wv[h*32+j, d]
        float* row = &wov[static_cast<size_t>(i) * H * D + static_cast<size_t>(h) * D];
        std::vector<double> acc(D, 0.0);
        for (int j = 0; j < DH; ++j) {
          const double oij = (*ow)[static_cast<size_t>(i) * D + h * DH + j];
          const float* wvrow = Wv + static_cast<size_t>(h * DH + j) * D;
          for (int d = 0; d < D; ++d) acc[d] += oij * wvrow[d];
        }
        for (int d = 0; d < D; ++d) row[d] = static_cast<float>(acc[d]);
      }
    }
    for (int i = 0; i < D; ++i) {
      double s = (*ob)[i];
      for (int k = 0; k < D; ++k) s += static_cast<double>((*ow)[static_cast<size_t>(i) * D + k]) * (*cb)[2 * D + k];
      bov[i] = static_cast<float>(s);
    }
    MADE_TRY(up_lin(c, wqf.data(), bqf.data(), static_cast<size_t>(H) * D, D, &c->ddec[l].qfold));
    MADE_TRY(up_lin(c, wov.data(), bov.data(), D, static_cast<size_t>(H) * D, &c->ddec[l].ovfold));
    MADE_TRY(load_lin(c, p + ".linear1", DFF, D, &c->ddec[l].ff1));
    MADE_TRY(load_lin(c, p + ".linear2", D, DFF, &c->ddec[l].ff2));
    MADE_TRY(load_ln(c, p + ".norm1", &c->ddec[l].n1));
    MADE_TRY(load_ln(c, p + ".norm2", &c->ddec[l].n2));
    MADE_TRY(load_ln(c, p + ".norm3", &c->ddec[l].n3));
  }
  MADE_TRY(load_ln(c, "detr_transformer.decoder.norm", &c->dec_norm));
  MADE_TRY(load_lin(c, "span_embed.layers.0", D, D, &c->span0));
  MADE_TRY(load_lin(c, "span_embed.layers.1", D, D, &c->span1));
  MADE_TRY(load_lin(c, "contrastive_align_projection_query", D, D, &c->pq));
  MADE_TRY(load_lin(c, "contrastive_align_projection_vid", D, D, &c->pv));
  const std::vector<float>*w, *b;
  MADE_TRY(get(c, "span_embed.layers.2.weight", 2 * D, &w));
  MADE_TRY(get(c, "span_embed.layers.2.bias", 2, &b));
  MADE_TRY(up_f32(c, w->data(), 2 * D, &c->span2_w));
  MADE_TRY(up_f32(c, b->data(), 2, &c->span2_b));
  MADE_TRY(get(c, "class_embed.weight", 2 * D, &w));
  MADE_TRY(get(c, "class_embed.bias", 2, &b));
  MADE_TRY(up_f32(c, w->data(), 2 * D, &c->cls_w));
  MADE_TRY(up_f32(c, b->data(), 2, &c->cls_b));
  // position_encoding.py:65-66: dim_t = 10000 ** (2 * (i // 2) / 256), fp32
  std::vector<float> inv(D);
  for (int i = 0; i < D; ++i) inv[i] = 1.0f / powf(10000.0f, static_cast<float>(2 * (i / 2)) / 256.0f);
  MADE_TRY(up_f32(c, inv.data(), D, &c->inv_dim_t));
  return MADE_OK;
}

__global__ void cast_f32_op_kernel(const float* __restrict__ in, op_t* __restrict__ out, int64_t n) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f2op(in[i]);
}
// plain linear helper: out = act(A W^T + b)
int linear(const op_t* A, int64_t lda, const Lin& w, int64_t M, int N, int K, GemmEpilogue epi,
           cudaStream_t st) {
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  epi.bias = w.b;
  p.epi = epi;
  return gemm_f16_tc(A, lda, w.w, K, N, p, 256, st);
}

bool fused_ffn_enabled() {
  static const bool enabled = [] {
    const char* v = getenv("MADE_FUSED_FFN");
    return !(v && v[0] == '0');
  }();
  return enabled;
}

#define CTX_READY(ctx)                                                       \
  do {                                                                       \
    if (!(ctx)) { set_error("null made_ctx"); return MADE_EINVAL; }          \
    if (!(ctx)->loaded) { set_error("made_ctx has no weights loaded"); return MADE_ESTATE; } \
    MADE_CUDA(cudaSetDevice((ctx)->device));                                 \
  } while (0)

}  // namespace

extern "C" {

int made_ctx_create(made_ctx** out, int device) {
  MADE_REQUIRE(out, "ctx_create: null out");
  MADE_TRY(made_device_check(device));
  MADE_CUDA(cudaSetDevice(device));
  made_ctx* c = new made_ctx();
  c->device = device;
  *out = c;
  return MADE_OK;
}

int made_ctx_destroy(made_ctx* c) {
  if (!c) return MADE_OK;
  cudaSetDevice(c->device);
  for (void* p : c->allocs) cudaFree(p);
  if (c->ws) cudaFree(c->ws);
  delete c;
  return MADE_OK;
}

int made_ctx_load_weights(made_ctx* c, int n, const char* const* names, const float* const* host_ptrs,
                          const int64_t* numels, void* stream) {
  MADE_REQUIRE(c && names && host_ptrs && numels, "load_weights: null argument");
  MADE_CUDA(cudaSetDevice(c->device));
  for (void* p : c->allocs) cudaFree(p);
  c->allocs.clear();
  c->host.clear();
  c->loaded = false;
  for (int i = 0; i < n; ++i) c->host[names[i]] = std::vector<float>(host_ptrs[i], host_ptrs[i] + numels[i]);
  MADE_TRY(load_encoder(c, MADE_VIDEO));
  MADE_TRY(load_encoder(c, MADE_MUSIC));
  MADE_TRY(load_xpool(c, static_cast<cudaStream_t>(stream)));
  MADE_TRY(load_detr(c));
  MADE_TRY(load_exact(c));
  MADE_TRY(load_ca(c));
  c->host.clear();
  c->loaded = true;
  return MADE_OK;
}

static Ragged to_ragged(const made_ragged* rb) {
  Ragged r;
  r.seq_len = rb->seq_len;
  r.seq_off = rb->seq_off;
  r.total = rb->total;
  r.tok_src = rb->tok_src;
  r.B = rb->B;
  r.L = rb->L;
  return r;
}

int64_t made_ragged_index_words(int64_t B, int L) {
  if (B <= 0 || L <= 0) return 0;
  return static_cast<int64_t>(ragged_index_words(B, L));
}

int made_ragged_build(const float* masks, int64_t B, int L, int32_t* idx_workspace, made_ragged* out, void* stream) {
  MADE_REQUIRE(out, "ragged_build: null descriptor");
  Ragged r;
  MADE_TRY(ragged_build(masks, B, L, idx_workspace, &r, static_cast<cudaStream_t>(stream)));
  out->seq_len = r.seq_len;
  out->seq_off = r.seq_off;
  out->total = r.total;
  out->tok_src = r.tok_src;
  out->B = r.B;
  out->L = r.L;
  return MADE_OK;
}

int made_ingest_ragged(made_ctx* c, const void* feats, int feats_dtype, const made_ragged* rb, int dim,
                       void* out16_packed, void* stream) {
  MADE_REQUIRE(c, "ingest_ragged: null made_ctx");
  MADE_REQUIRE(feats && rb && out16_packed, "ingest_ragged: null pointer");
  MADE_REQUIRE(feats_dtype >= MADE_DTYPE_F32 && feats_dtype <= MADE_DTYPE_F16, "ingest_ragged: bad dtype %d",
               feats_dtype);
  MADE_REQUIRE(dim > 0 && dim % 8 == 0, "ingest_ragged: dim=%d must be a positive multiple of 8", dim);
  if (rb->B == 0) return MADE_OK;
  return ingest_gather(feats, feats_dtype, to_ragged(rb), dim, static_cast<op_t*>(out16_packed),
                       c->precision == MADE_PREC_SPLIT, static_cast<cudaStream_t>(stream));
}

int made_ctx_set_precision(made_ctx* c, int mode) {
  MADE_REQUIRE(c, "set_precision: null made_ctx");
  MADE_REQUIRE(mode == MADE_PREC_FP16 || mode == MADE_PREC_SPLIT || mode == MADE_PREC_FP32, "set_precision: unknown mode %d",
               mode);
  c->precision = mode;
  return MADE_OK;
}

int made_ctx_operand_width(const made_ctx* c, int dim) {
  return c && c->precision == MADE_PREC_SPLIT ? 2 * dim : dim;
}

// Scratch of one encoder pass over T packed rows.  Split precision: token activations that feed a
// K = 256 / K = din GEMM are (hi, lo) fp16 pairs, [T, 512] with the lo halves 256 columns to the right,
// and double as the residual stream (hi + lo carries ~22 bits); fp16 mode keeps an fp32 copy instead.
struct EncBufs {
  op_t *x1 = nullptr, *qkv = nullptr, *att = nullptr, *x2 = nullptr, *h = nullptr, *x3 = nullptr;
  float *x1f = nullptr, *x2f = nullptr, *seqf = nullptr;
};
static void encode_layout(made_ctx* c, int64_t T, EncBufs* b) {
  const bool sp = c->precision == MADE_PREC_SPLIT;
  const int64_t W = sp ? 2 * D : D;
  b->x1 = c->take<op_t>(T * W);
  b->x1f = sp ? nullptr : c->take<float>(T * D);
  b->qkv = c->take<op_t>(T * 3 * D);
  b->att = c->take<op_t>(T * D);
  b->x2 = c->take<op_t>(T * W);
  b->x2f = sp ? nullptr : c->take<float>(T * D);
  b->h = c->take<op_t>(T * DFF);
  b->x3 = c->take<op_t>(T * W);
  b->seqf = c->take<float>(T * D);
}

// forward_{video,audio}_encoder_feature on a token-packed batch (valid tokens only).
static int encode_packed(made_ctx* c, int modality, const op_t* x0, const Ragged& rb, const EncBufs& bf, void* seq16,
                         float* seq_f32, float* pooled, cudaStream_t st) {
  const EncW& e = c->enc[modality];
  const int64_t B = rb.B;
  const int64_t T = B * e.L;   // upper bound of the packed row count (the real count lives on the device)
  const bool sp = c->precision == MADE_PREC_SPLIT;
  const int64_t W = sp ? 2 * D : D;          // row stride of the (hi | lo) activation buffers

  // A [T, K] (row stride lda; split_a: a (hi | lo) pair) x w -> N columns
  auto lin = [&](const op_t* A, int64_t lda, bool split_a, const Lin& w, bool split_w, int N, int K, GemmEpilogue ep) {
    GemmParams p;
    p.M = T;
    p.N = N;
    p.K = K;
    p.m_dev = rb.total;
    p.split = split_w ? (split_a ? 2 : 1) : 0;
    ep.bias = w.b;
    p.epi = ep;
    return gemm_f16_tc(A, lda, split_w ? w.w2 : w.w, split_w ? 2 * K : K, N, p, 256, st);
  };
  auto residual_from = [&](GemmEpilogue& ep, const op_t* pair, const float* f32) {
    if (sp) {
      ep.residual = pair;
      ep.residual_lo = pair + D;
      ep.residual_f32 = 0;
      ep.res_ld = W;
    } else {
      ep.residual = f32;
      ep.residual_f32 = 1;
      ep.res_ld = D;
    }
  };
  auto out_pair = [&](GemmEpilogue& ep, op_t* pair, float* f32) {
    ep.out_h = pair;
    ep.ld_h = W;
    if (sp) {
      ep.out_lo = pair + D;
    } else {
      ep.out_f32 = f32;
      ep.ld_f32 = D;
    }
  };
  {  // :559/598 projection, :533 += pe[position], Transformer_enhancement norm1 (:86)
    GemmEpilogue ep;
    ep.row_table = e.pe;
    ep.row_mod = e.L;
    ep.row_src = rb.tok_src;
    ep.ln_gamma = e.ln1.g;
    ep.ln_beta = e.ln1.b;
    out_pair(ep, bf.x1, bf.x1f);
    MADE_TRY(lin(x0, sp ? 2 * e.din : e.din, sp, e.proj, sp, D, e.din, ep));
  }
  {  // packed in_proj (nn.MultiheadAttention)
    GemmEpilogue ep;
    ep.out_h = bf.qkv;
    ep.ld_h = 3 * D;
    MADE_TRY(lin(bf.x1, W, sp, e.in_proj, sp, 3 * D, D, ep));
  }
  MADE_TRY(mha_core(bf.qkv, 3 * D, bf.qkv + D, 3 * D, bf.qkv + 2 * D, 3 * D, nullptr, B, e.L, bf.att, D, st,
                    rb.seq_off, rb.seq_len));
  {  // out_proj + residual (from the normed tensor, Q2) + norm2 (:87-88)
    GemmEpilogue ep;
    residual_from(ep, bf.x1, bf.x1f);
    ep.ln_gamma = e.ln2.g;
    ep.ln_beta = e.ln2.b;
    out_pair(ep, bf.x2, bf.x2f);
    MADE_TRY(lin(bf.att, D, false, e.out_proj, sp, D, D, ep));
  }
  if (fused_ffn_enabled()) {
    // FF: Linear -> GELU(erf) -> Linear + residual (:89) in ONE kernel, hidden activation kept on chip; plain
    // fp16 operands in both precision modes (the FF branch is a small perturbation of the residual stream)
    if (sp) {
      MADE_TRY(ffn_fused(bf.x2, W, e.ff1.w, e.ff1.b, e.ff2.w, e.ff2.b, 1, bf.x2, bf.x2 + D, W, nullptr, nullptr, bf.x3,
                         bf.x3 + D, W, nullptr, 0, nullptr, 0, T, rb.total, st));
    } else {
      // fp16 mode keeps the residual stream in fp32 (x2f): the residual add stays in the unfused FF2 GEMM
      GemmEpilogue ep;
      ep.act = 1;
      ep.out_h = bf.h;
      ep.ld_h = DFF;
      MADE_TRY(lin(bf.x2, W, false, e.ff1, false, DFF, D, ep));
      GemmEpilogue ep2;
      residual_from(ep2, bf.x2, bf.x2f);
      ep2.out_h = bf.x3;
      ep2.ld_h = W;
      MADE_TRY(lin(bf.h, DFF, false, e.ff2, false, D, DFF, ep2));
    }
  } else {
    {  // FF: Linear -> GELU(erf); plain fp16 operands in both modes
      GemmEpilogue ep;
      ep.act = 1;
      ep.out_h = bf.h;
      ep.ld_h = DFF;
      MADE_TRY(lin(bf.x2, W, false, e.ff1, false, DFF, D, ep));
    }
    {  // FF: Linear + residual (:89)
      GemmEpilogue ep;
      residual_from(ep, bf.x2, bf.x2f);
      ep.out_h = bf.x3;
      ep.ld_h = W;
      if (sp) ep.out_lo = bf.x3 + D;
      MADE_TRY(lin(bf.h, DFF, false, e.ff2, false, D, DFF, ep));
    }
  }
  {  // final_linear (:91); masked_fill (:541) = padded rows of the output stay zero
    MADE_CUDA(cudaMemsetAsync(seq16, 0, static_cast<size_t>(T) * D * 2, st));
    GemmEpilogue ep;
    ep.out_h = static_cast<op_t*>(seq16);
    ep.ld_h = D;
    ep.h_row_idx = rb.tok_src;      // scatter packed rows back to [B, L, 256]
    ep.out_f32 = bf.seqf;
    ep.ld_f32 = D;
    MADE_TRY(lin(bf.x3, W, sp, e.fin, sp, D, D, ep));
  }
  if (seq_f32) MADE_TRY(scatter_rows_f32(bf.seqf, rb, seq_f32, st));
  // masked mean + F.normalize (:579-580 / :615-616)
  MADE_TRY(pool_norm_ragged(bf.seqf, rb, pooled, st));
  return MADE_OK;
}

int made_encode_ragged(made_ctx* c, int modality, const void* x16_packed, const made_ragged* rb, void* seq16,
                       float* seq_f32, float* pooled, void* stream) {
  CTX_READY(c);
  MADE_REQUIRE(modality == MADE_VIDEO || modality == MADE_MUSIC, "encode_ragged: bad modality %d", modality);
  MADE_REQUIRE(rb, "encode_ragged: null descriptor");
  if (rb->B == 0) return MADE_OK;
  MADE_REQUIRE(x16_packed && seq16 && pooled, "encode_ragged: null pointer");
  const EncW& e = c->enc[modality];
  MADE_REQUIRE(rb->L == e.L, "encode_ragged: descriptor built for L=%d, modality needs L=%d", rb->L, e.L);
  const int64_t T = rb->B * e.L;
  MADE_REQUIRE(T < (1LL << 31), "encode: batch too large (%lld tokens); chunk the call", (long long)T);
  EncBufs bf;
  MADE_TRY(c->with_arena([&] { encode_layout(c, T, &bf); }));
  return encode_packed(c, modality, static_cast<const op_t*>(x16_packed), to_ragged(rb), bf, seq16, seq_f32, pooled,
                       static_cast<cudaStream_t>(stream));
}

int made_encode(made_ctx* c, int modality, const void* feats, int feats_dtype, const float* masks, int64_t B,
                void* seq16, float* seq_f32, float* pooled, void* stream) {
  CTX_READY(c);
  MADE_REQUIRE(modality == MADE_VIDEO || modality == MADE_MUSIC, "encode: bad modality %d", modality);
  MADE_REQUIRE(feats_dtype >= MADE_DTYPE_F32 && feats_dtype <= MADE_DTYPE_F16, "encode: bad feature dtype %d",
               feats_dtype);
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(feats && masks && seq16 && pooled, "encode: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const EncW& e = c->enc[modality];
  const int64_t T = B * e.L;
  MADE_REQUIRE(T < (1LL << 31), "encode: batch too large (%lld tokens); chunk the call", (long long)T);
  if (c->precision == MADE_PREC_FP32) {
    // fp32 CUDA-core path: the reference's arithmetic (exact_f32.cu); seq16 = fp16 cast of the fp32 output (DETR input)
    float* ws = nullptr;
    float* seqf = nullptr;
    MADE_TRY(c->with_arena([&] {
      ws = c->take<float>(exact_encode_ws_floats(B, e.L, e.din));
      seqf = seq_f32 ? nullptr : c->take<float>(T * D);
    }));
    float* out32 = seq_f32 ? seq_f32 : seqf;
    MADE_TRY(exact_encode(c->xenc[modality], feats, feats_dtype, masks, B, ws, out32, pooled, st));
    cast_f32_op_kernel<<<static_cast<unsigned>(ceil_div64(T * D, 256)), 256, 0, st>>>(out32, static_cast<op_t*>(seq16), T * D);
    MADE_CHECK_LAUNCH();
    return MADE_OK;
  }
  const bool sp = c->precision == MADE_PREC_SPLIT;
  int32_t* idx = nullptr;
  op_t* x0 = nullptr;
  EncBufs bf;
  MADE_TRY(c->with_arena([&] {
    idx = c->take<int32_t>(ragged_index_words(B, e.L));
    x0 = c->take<op_t>(T * e.din * (sp ? 2 : 1));
    encode_layout(c, T, &bf);
  }));
  Ragged rb;
  MADE_TRY(ragged_build(masks, B, e.L, idx, &rb, st));
  // model_Base.py:556/595 masked_fill + cast: only the valid rows are read and packed
  MADE_TRY(ingest_gather(feats, feats_dtype, rb, e.din, x0, sp, st));
  return encode_packed(c, modality, x0, rb, bf, seq16, seq_f32, pooled, st);
}

int made_gallery_prepare(made_ctx* c, const void* seg16, const float* seg_masks, int64_t N, void* kz,
                         void* gram, uint32_t* maskbits, void* stream) {
  CTX_READY(c);
  if (N == 0) return MADE_OK;
  MADE_REQUIRE(seg16 && seg_masks && kz && gram && maskbits, "gallery_prepare: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t T = N * LM;
  MADE_REQUIRE(T < (1LL << 31), "gallery_prepare: too many tracks in one call; chunk it");
  const bool sp = c->precision == MADE_PREC_SPLIT;
  const int64_t W = sp ? 2 * D : D;
  op_t* spn = nullptr;
  MADE_TRY(c->with_arena([&] { spn = c->take<op_t>(T * W); }));
  // shared LayerNorm1 on the segments (modules/transformer.py:165)
  MADE_TRY(layernorm_rows(seg16, 1, D, T, c->xp_ln1.g, c->xp_ln1.b, spn, W, sp ? spn + D : nullptr, nullptr, st));
  op_t* kzb = static_cast<op_t*>(kz);
  {
    GemmParams p;
    p.M = T;
    p.N = 3 * D;
    p.K = D;
    p.split = sp ? 2 : 0;
    p.epi.bias = c->xp_kvz.b;
    p.epi.out_h = kzb;
    p.epi.ld_h = 3 * D;
    MADE_TRY(gemm_f16_tc(spn, W, sp ? c->xp_kvz.w2 : c->xp_kvz.w, sp ? 2 * D : D, 3 * D, p, 256, st));
  }
  {  // per-track Gram matrix G = V'' V''^T (96 x 96), batched over tracks -> columns 0..95 of [G | W5 | 0]
    GemmParams p;
    p.M = T;
    p.N = LM;
    p.K = D;
    p.m_stride = LM;
    p.m_valid = LM;
    p.b_batched = 1;
    p.epi.out_h = static_cast<op_t*>(gram);
    p.epi.ld_h = 112;
    MADE_TRY(gemm_f16_tc(kzb + D, 3 * D, kzb + D, 3 * D, T, p, 96, st));
  }
  // W5 = Z'' . {1, b', g3^2, g3^2 b', g3 beta3} -> columns 96..100 (xpool.cu)
  MADE_TRY(xpool_w5(kzb + 2 * D, 3 * D, T, c->xp_c5, static_cast<op_t*>(gram), st));
  MADE_TRY(mask_bits(seg_masks, N, maskbits, st));
  return MADE_OK;
}

int made_query_prepare(made_ctx* c, const float* video_feats, int64_t N, void* q, float* vhat, void* stream) {
  CTX_READY(c);
  if (N == 0) return MADE_OK;
  MADE_REQUIRE(video_feats && q && vhat, "query_prepare: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool sp = c->precision == MADE_PREC_SPLIT;
  const int64_t W = sp ? 2 * D : D;
  op_t* vp = nullptr;
  MADE_TRY(c->with_arena([&] { vp = c->take<op_t>(N * W); }));
  MADE_TRY(layernorm_rows(video_feats, 0, D, N, c->xp_ln1.g, c->xp_ln1.b, vp, W, sp ? vp + D : nullptr, nullptr, st));
  GemmParams p;
  p.M = N;
  p.N = D;
  p.K = D;
  p.split = sp ? 2 : 0;
  p.epi.bias = c->xp_q.b;
  p.epi.out_h = static_cast<op_t*>(q);
  p.epi.ld_h = D;
  MADE_TRY(gemm_f16_tc(vp, W, sp ? c->xp_q.w2 : c->xp_q.w, sp ? 2 * D : D, D, p, 256, st));
  MADE_TRY(vhat_rows(video_feats, N, vhat, st));
  return MADE_OK;
}

int made_xpool_score(made_ctx* c, const void* q, const float* vhat, int64_t n_queries, const void* kz,
                     const void* gram, const uint32_t* maskbits, int64_t n_tracks, float* sim, int64_t ld,
                     int64_t col_offset, void* stream) {
  CTX_READY(c);
  MADE_REQUIRE(ld >= col_offset + n_tracks, "xpool_score: ld=%lld too small", (long long)ld);
  return xpool_score(c->xp_consts, static_cast<const op_t*>(q), vhat, n_queries,
                     static_cast<const op_t*>(kz), 3 * D, 2 * D, static_cast<const op_t*>(gram),
                     maskbits, n_tracks, sim, ld, col_offset, static_cast<cudaStream_t>(stream));
}

int made_detr_detect(made_ctx* c, const void* frame16, const float* frame_masks, const void* seg16,
                     const float* seg_masks, const int32_t* track_idx, const float* video_feats, int64_t B,
                     float* hs, float* pred_logits, float* pred_spans, float* proj_queries, float* proj_vid_mem,
                     float* memory, void* stream) {
  CTX_READY(c);
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(frame16 && frame_masks && seg16 && seg_masks && video_feats && hs && pred_logits && pred_spans,
               "detr_detect: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MADE_REQUIRE(B * LD < (1LL << 31), "detr_detect: batch too large; chunk the call");
  // The encoder (token-wise GEMMs over B*146 rows) runs in chunks of kEncChunk sequences so that its
  // scratch stays small and L2-friendly; the decoder (one query per sequence: tiny, latency-bound
  // GEMMs) runs once over all B sequences.
  constexpr int64_t kEncChunk = 2048;
  const int64_t Bc = B < kEncChunk ? B : kEncChunk;
  const int64_t Tc = Bc * LD, Tall = B * LD;
  const int64_t R = NDEC * B;
  const int64_t n_chunks = (B + Bc - 1) / Bc;
  const size_t idx_words = ragged_index_words(Bc, LD);
  // Encoder output of every sequence, token-packed per encoder chunk: chunk k owns rows
  // [k * Tc, ...) of mem_all / mp_all, sequence b its chunk's rows [seq_off[b], + seq_len[b]).
  op_t *mem_all, *mp_all, *src, *pos, *srcpos, *qk, *v, *att, *s1, *hbuf, *src2, *srcpos2, *lo1, *lo2, *tgt, *t1, *mbar, *t2,
      *hdec, *t3, *hsb, *sp0, *sp1;
  float *mask, *t1f, *qt, *t2f, *t3all;
  int32_t *row_off, *row_len;
  std::vector<int32_t*> idx_chunk(static_cast<size_t>(n_chunks));
  MADE_TRY(c->with_arena([&] {
    mem_all = c->take<op_t>(Tall * D);     // memory, fp16
    mp_all = c->take<op_t>(Tall * D);      // memory + pos
    mask = c->take<float>(Tall);           // concatenated key mask [B, 146]
    row_off = c->take<int32_t>(B);         // first row of sequence b inside mem_all / mp_all
    row_len = c->take<int32_t>(B);
    src = c->take<op_t>(Tc * D);
    pos = c->take<op_t>(Tc * D);
    srcpos = c->take<op_t>(Tc * D);
    qk = c->take<op_t>(Tc * 2 * D);
    v = c->take<op_t>(Tc * D);
    att = c->take<op_t>(Tc * D);
    s1 = c->take<op_t>(Tc * 2 * D);        // (hi | lo) pairs
    hbuf = fused_ffn_enabled() ? nullptr : c->take<op_t>(Tc * DFF);
    src2 = c->take<op_t>(Tc * D);
    srcpos2 = c->take<op_t>(Tc * D);
    lo1 = c->take<op_t>(Tc * D);           // low halves of the layer outputs (ping-pong)
    lo2 = c->take<op_t>(Tc * D);
    tgt = c->take<op_t>(B * D);
    t1 = c->take<op_t>(B * D);
    t1f = c->take<float>(B * D);
    qt = c->take<float>(B * 8 * D);
    mbar = c->take<op_t>(B * 8 * D);
    t2 = c->take<op_t>(B * D);
    t2f = c->take<float>(B * D);
    hdec = c->take<op_t>(B * DFF);
    t3 = c->take<op_t>(B * D);
    t3all = c->take<float>(R * D);
    hsb = c->take<op_t>(R * D);
    sp0 = c->take<op_t>(R * D);
    sp1 = c->take<op_t>(R * D);
    for (auto& ip : idx_chunk) ip = c->take<int32_t>(idx_words);
  }));

  const op_t* fr = static_cast<const op_t*>(frame16);
  static_assert(NENC % 2 == 0, "the encoder ping-pong below ends in the (src, srcpos) buffers");
  if (memory) MADE_CUDA(cudaMemsetAsync(memory, 0, static_cast<size_t>(Tall) * D * 4, st));
  // ---------------- encoder (forward_post, music_detr/transformer.py:191-210), chunked, valid tokens only ----
  for (int64_t b0 = 0, ck = 0; b0 < B; b0 += Bc, ++ck) {
    const int64_t nb = (B - b0) < Bc ? (B - b0) : Bc;
    const int64_t T = nb * LD;          // upper bound of this chunk's packed rows
    float* mask_c = mask + b0 * LD;
    int32_t* idx = idx_chunk[static_cast<size_t>(ck)];
    Ragged rb;
    MADE_TRY(detr_mask(frame_masks + b0 * LV, seg_masks, track_idx ? track_idx + b0 : nullptr, b0, nb, mask_c, st));
    MADE_TRY(ragged_build(mask_c, nb, LD, idx, &rb, st));
    MADE_TRY(detr_prep_ragged(fr + b0 * LV * D, static_cast<const op_t*>(seg16), track_idx ? track_idx + b0 : nullptr,
                              b0, rb, c->inv_dim_t, src, pos, srcpos, st));
    MADE_TRY(offset_rows(rb, static_cast<int32_t>(b0 * LD), row_off + b0, row_len + b0, st));
    auto lin = [&](const op_t* A, int64_t lda, const Lin& w, int N, int K, GemmEpilogue ep) {
      GemmParams p;
      p.M = T;
      p.N = N;
      p.K = K;
      p.m_dev = rb.total;
      ep.bias = w.b;
      p.epi = ep;
      return gemm_f16_tc(A, lda, w.w, K, N, p, 256, st);
    };
    // The residual stream travels as fp16 (hi, lo) pairs: s1 = [hi | lo] rows of 512, the layer output as the
    // fp16 tensor the next GEMMs read (hi) plus a separate low-half buffer (~22 bits together; no fp32 copies).
    op_t *cur = src, *curpos = srcpos, *nxt = src2, *nxtpos = srcpos2;
    op_t *cur_lo = nullptr, *nxt_lo = lo1;
    for (int l = 0; l < NENC; ++l) {
      const DetrEncW& w = c->denc[l];
      if (l == NENC - 1) {   // the last layer writes the memory straight into the all-sequence buffers
        nxt = mem_all + b0 * LD * D;
        nxtpos = mp_all + b0 * LD * D;
      }
      {
        GemmEpilogue ep;
        ep.out_h = qk;
        ep.ld_h = 2 * D;
        MADE_TRY(lin(curpos, D, w.qk, 2 * D, D, ep));   // q = k = src + pos
      }
      {
        GemmEpilogue ep;
        ep.out_h = v;
        ep.ld_h = D;
        MADE_TRY(lin(cur, D, w.v, D, D, ep));           // value = src
      }
      MADE_TRY(mha_core(qk, 2 * D, qk + D, 2 * D, v, D, nullptr, nb, LD, att, D, st, rb.seq_off, rb.seq_len));
      {
        GemmEpilogue ep;
        ep.residual = cur;
        ep.residual_lo = cur_lo;      // layer 0: the fp16 input features are exact
        ep.residual_f32 = 0;
        ep.res_ld = D;
        ep.ln_gamma = w.n1.g;
        ep.ln_beta = w.n1.b;
        ep.out_h = s1;
        ep.out_lo = s1 + D;
        ep.ld_h = 2 * D;
        MADE_TRY(lin(att, D, w.out, D, D, ep));
      }
      if (fused_ffn_enabled()) {
        // linear2(relu(linear1(src))) + src -> norm2 (+ second output "+ pos") in one kernel, hidden on chip
        MADE_TRY(ffn_fused(s1, 2 * D, w.ff1.w, w.ff1.b, w.ff2.w, w.ff2.b, 2, s1, s1 + D, 2 * D, w.n2.g, w.n2.b, nxt, nxt_lo,
                           D, pos, D, nxtpos, D, T, rb.total, st));
      } else {
        {
          GemmEpilogue ep;
          ep.act = 2;
          ep.out_h = hbuf;
          ep.ld_h = DFF;
          MADE_TRY(lin(s1, 2 * D, w.ff1, DFF, D, ep));
        }
        {
          GemmEpilogue ep;
          ep.residual = s1;
          ep.residual_lo = s1 + D;
          ep.residual_f32 = 0;
          ep.res_ld = 2 * D;
          ep.ln_gamma = w.n2.g;
          ep.ln_beta = w.n2.b;
          ep.out_h = nxt;
          ep.out_lo = nxt_lo;
          ep.ld_h = D;
          ep.add2 = pos;
          ep.add2_ld = D;
          ep.out2_h = nxtpos;
          ep.ld_out2 = D;
          MADE_TRY(lin(hbuf, DFF, w.ff2, D, DFF, ep));
        }
      }
      op_t* t = cur; cur = nxt; nxt = t;
      t = curpos; curpos = nxtpos; nxtpos = t;
      cur_lo = nxt_lo;
      nxt_lo = nxt_lo == lo1 ? lo2 : lo1;
    }
    if (memory) MADE_TRY(scatter_rows_pair_nozero(cur, cur_lo, rb, memory + b0 * LD * D, st));
  }
  // ---------------- decoder (forward_post :273-307), one moment query per sequence ----------------
  cast_f32_op_kernel<<<static_cast<unsigned>(ceil_div64(B * D, 256)), 256, 0, st>>>(video_feats, tgt, B * D);
  MADE_CHECK_LAUNCH();
  const op_t* tin = tgt;
  const float* tinf = video_feats;
  for (int l = 0; l < NDEC; ++l) {
    const DetrDecW& w = c->ddec[l];
    float* t3f = t3all + static_cast<size_t>(l) * B * D;
    op_t* tout = (l & 1) ? tgt : t3;   // the layer's fp16 output ping-pongs between two buffers: the next layer reads it as `tin`
    {
      GemmEpilogue ep;   // self-attention on a single query (folded) + norm1
      ep.residual = tinf;
      ep.residual_f32 = 1;
      ep.res_ld = D;
      ep.ln_gamma = w.n1.g;
      ep.ln_beta = w.n1.b;
      ep.out_h = t1;
      ep.ld_h = D;
      ep.out_f32 = t1f;
      ep.ld_f32 = D;
      MADE_TRY(linear(tin, D, w.sa, B, D, D, ep, st));
    }
    {
      GemmEpilogue ep;   // q~ = per-head Wk_h^T q_h / sqrt(32)
      ep.out_f32 = qt;
      ep.ld_f32 = 8 * D;
      MADE_TRY(linear(t1, D, w.qfold, B, 8 * D, D, ep, st));
    }
    MADE_TRY(dec_attn_folded(qt, mp_all, mem_all, nullptr, B, LD, mbar, st, row_off, row_len));
    {
      GemmEpilogue ep;   // out_proj o v_proj folded, + residual + norm2
      ep.residual = t1f;
      ep.residual_f32 = 1;
      ep.res_ld = D;
      ep.ln_gamma = w.n2.g;
      ep.ln_beta = w.n2.b;
      ep.out_h = t2;
      ep.ld_h = D;
      ep.out_f32 = t2f;
      ep.ld_f32 = D;
      MADE_TRY(linear(mbar, 8 * D, w.ovfold, B, D, 8 * D, ep, st));
    }
    {
      GemmEpilogue ep;
      ep.act = 2;
      ep.out_h = hdec;
      ep.ld_h = DFF;
      MADE_TRY(linear(t2, D, w.ff1, B, DFF, D, ep, st));
    }
    {
      GemmEpilogue ep;
      ep.residual = t2f;
      ep.residual_f32 = 1;
      ep.res_ld = D;
      ep.ln_gamma = w.n3.g;
      ep.ln_beta = w.n3.b;
      ep.out_h = tout;
      ep.ld_h = D;
      ep.out_f32 = t3f;
      ep.ld_f32 = D;
      MADE_TRY(linear(hdec, DFF, w.ff2, B, D, DFF, ep, st));
    }
    tin = tout;
    tinf = t3f;
  }
  // decoder.norm on every layer's output (:136), then the heads (model_Uni.py:131-149)
  MADE_TRY(layernorm_rows(t3all, 0, D, R, c->dec_norm.g, c->dec_norm.b, hsb, D, nullptr, hs, st));
  {
    GemmEpilogue ep;
    ep.act = 2;
    ep.out_h = sp0;
    ep.ld_h = D;
    MADE_TRY(linear(hsb, D, c->span0, R, D, D, ep, st));
  }
  {
    GemmEpilogue ep;
    ep.act = 2;
    ep.out_h = sp1;
    ep.ld_h = D;
    MADE_TRY(linear(sp0, D, c->span1, R, D, D, ep, st));
  }
  MADE_TRY(heads_final(hs, sp1, R, c->cls_w, c->cls_b, c->span2_w, c->span2_b, pred_logits, pred_spans, st));
  if (proj_queries) {
    GemmEpilogue ep;
    ep.l2norm = 1;
    ep.out_f32 = proj_queries;
    ep.ld_f32 = D;
    MADE_TRY(linear(hsb, D, c->pq, R, D, D, ep, st));
  }
  if (proj_vid_mem) {
    GemmEpilogue ep;
    ep.l2norm = 1;
    ep.out_f32 = proj_vid_mem;
    ep.ld_f32 = D;
    MADE_TRY(linear(fr, D, c->pv, B * LV, D, D, ep, st));
  }
  return MADE_OK;
}

int made_gemm_f16(const void* A, const void* W, int64_t M, int N, int K, const float* bias, const float* residual,
                   int act, const float* ln_gamma, const float* ln_beta, void* out_h, float* out_f32,
                   void* stream) {
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.epi.bias = bias;
  p.epi.residual = residual;
  p.epi.residual_f32 = 1;
  p.epi.res_ld = N;
  p.epi.act = act;
  p.epi.ln_gamma = ln_gamma;
  p.epi.ln_beta = ln_beta;
  p.epi.out_h = static_cast<op_t*>(out_h);
  p.epi.ld_h = N;
  p.epi.out_f32 = out_f32;
  p.epi.ld_f32 = N;
  return gemm_f16_tc(static_cast<const op_t*>(A), K, static_cast<const op_t*>(W), K, N, p, 256,
                      static_cast<cudaStream_t>(stream));
}

int made_xpool_pooled(made_ctx* c, int which, const float* video_feats, int64_t n_q, const float* seg_f32,
                      const float* seg_masks, int64_t n_m, float* pooled, void* stream) {
  CTX_READY(c);
  MADE_REQUIRE(which == MADE_VIDEO || which == MADE_MUSIC, "xpool_pooled: which must be MADE_MUSIC (video-guided pooling of "
               "music segments) or MADE_VIDEO (music-guided pooling of video frames)");
  MADE_REQUIRE(which == MADE_MUSIC || c->has_xp_video, "xpool_pooled: the loaded checkpoint has no "
               "music_guided_to_video_pooling_cross_transformer (vmr_fusion 'XA-music-video')");
  if (n_q == 0 || n_m == 0) return MADE_OK;
  MADE_REQUIRE(video_feats && seg_f32 && seg_masks && pooled, "xpool_pooled: null pointer");
  const int Lk = which == MADE_MUSIC ? LM : LV;
  MADE_REQUIRE(n_m * n_q < (1LL << 31) / 256 * 8, "xpool_pooled: %lld x %lld pairs in one call; chunk the tracks",
               (long long)n_m, (long long)n_q);
  float* ws = nullptr;
  MADE_TRY(c->with_arena([&] { ws = c->take<float>(exact_xpool_ws_floats(n_q, n_m, Lk)); }));
  return exact_xpool(which == MADE_MUSIC ? c->xxp : c->xxp_video, video_feats, n_q, seg_f32, seg_masks, n_m, Lk, ws, pooled,
                     static_cast<cudaStream_t>(stream));
}

int made_ca_fuse(made_ctx* c, const float* seg_f32, const float* seg_masks, const float* frame_f32, const float* frame_masks,
                 int64_t B, void* fused16, float* fused_f32, void* stream) {
  CTX_READY(c);
  MADE_REQUIRE(c->has_ca, "ca_fuse: the loaded checkpoint has no video_music_fusion_cross_transformer (mml_fusion 'CA')");
  if (B == 0) return MADE_OK;
  MADE_REQUIRE(seg_f32 && seg_masks && frame_f32 && frame_masks && fused16, "ca_fuse: null pointer");
  MADE_REQUIRE(B * LM < (1LL << 31) / (4 * D), "ca_fuse: batch too large; chunk the call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t Tq = B * LM, Tk = B * LV;
  op_t *nq, *nc, *q, *kv, *att, *nf, *h, *x2;
  float* ax;
  MADE_TRY(c->with_arena([&] {
    nq = c->take<op_t>(Tq * D);
    nc = c->take<op_t>(Tk * D);
    q = c->take<op_t>(Tq * 4 * D);
    kv = c->take<op_t>(Tk * 8 * D);
    att = c->take<op_t>(Tq * 4 * D);
    ax = c->take<float>(Tq * D);
    nf = c->take<op_t>(Tq * D);
    h = c->take<op_t>(Tq * DFF);
    x2 = c->take<op_t>(Tq * D);
  }));
  const made_ctx::CaW& w = c->ca;
  // CrossTransformer.forward (model_Base.py:199-213), depth 1
  MADE_TRY(layernorm_rows(seg_f32, 0, D, Tq, w.ln_q.g, w.ln_q.b, nq, D, nullptr, nullptr, st));          // norm_x
  MADE_TRY(layernorm_rows(frame_f32, 0, D, Tk, w.ln_c.g, w.ln_c.b, nc, D, nullptr, nullptr, st));        // norm_context
  {
    GemmEpilogue e;
    e.out_h = q; e.ld_h = 4 * D;
    MADE_TRY(linear(nq, D, w.to_q, Tq, 4 * D, D, e, st));                                                  // to_q (no bias)
  }
  {
    GemmEpilogue e;
    e.out_h = kv; e.ld_h = 8 * D;
    MADE_TRY(linear(nc, D, w.to_kv, Tk, 8 * D, D, e, st));                                                 // to_kv (no bias)
  }
  MADE_TRY(ca_attention(q, kv, seg_masks, frame_masks, B, att, st));
  {
    GemmEpilogue e;                                                                                        // to_out + x
    e.residual = seg_f32; e.residual_f32 = 1; e.res_ld = D;
    e.out_f32 = ax; e.ld_f32 = D;
    MADE_TRY(linear(att, 4 * D, w.to_out, Tq, D, 4 * D, e, st));
  }
  MADE_TRY(layernorm_rows(ax, 0, D, Tq, w.ln_ff.g, w.ln_ff.b, nf, D, nullptr, nullptr, st));              // ff_layer_norms
  {
    GemmEpilogue e;
    e.act = 1;                                                                                             // nn.GELU()
    e.out_h = h; e.ld_h = DFF;
    MADE_TRY(linear(nf, D, w.ff1, Tq, DFF, D, e, st));
  }
  {
    GemmEpilogue e;                                                                                        // ff(norm_x) + attn_x
    e.residual = ax; e.residual_f32 = 1; e.res_ld = D;
    e.out_h = x2; e.ld_h = D;
    MADE_TRY(linear(h, DFF, w.ff2, Tq, D, DFF, e, st));
  }
  {
    GemmEpilogue e;                                                                                        // final_linear,
    e.row_mask = seg_masks;                                                                                // masked_fill (model_Uni.py:210)
    e.out_h = static_cast<op_t*>(fused16); e.ld_h = D;
    e.out_f32 = fused_f32; e.ld_f32 = D;
    MADE_TRY(linear(x2, D, w.fin, Tq, D, D, e, st));
  }
  return MADE_OK;
}

int made_pooled_cosine(const float* video_feats, const float* pooled, int64_t n_q, int64_t n_m, float* sim, int64_t ld,
                       int64_t col_offset, void* stream) {
  MADE_REQUIRE(video_feats && pooled && sim && ld >= col_offset + n_m, "pooled_cosine: bad arguments");
  return exact_pooled_cosine(video_feats, pooled, n_q, n_m, sim, ld, col_offset, static_cast<cudaStream_t>(stream));
}

int made_gemm_f16_split(const void* A, const void* W, int64_t M, int N, int K, int split, const float* bias,
                        const void* residual_pair, int act, const float* ln_gamma, const float* ln_beta, void* out_pair,
                        float* out_f32, void* stream) {
  MADE_REQUIRE(split == 1 || split == 2, "gemm_f16_split: split must be 1 (W pairs) or 2 (A and W pairs)");
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.split = split;
  p.epi.bias = bias;
  if (residual_pair) {
    p.epi.residual = residual_pair;
    p.epi.residual_lo = static_cast<const op_t*>(residual_pair) + N;
    p.epi.residual_f32 = 0;
    p.epi.res_ld = 2 * N;
  }
  p.epi.act = act;
  p.epi.ln_gamma = ln_gamma;
  p.epi.ln_beta = ln_beta;
  if (out_pair) {
    p.epi.out_h = static_cast<op_t*>(out_pair);
    p.epi.out_lo = static_cast<op_t*>(out_pair) + N;
    p.epi.ld_h = 2 * N;
  }
  if (out_f32 && !out_pair) {
    p.epi.out_f32 = out_f32;
    p.epi.ld_f32 = N;
  }
  MADE_REQUIRE(out_pair || out_f32, "gemm_f16_split: no output");
  MADE_REQUIRE(!(out_pair && out_f32), "gemm_f16_split: one output");
  return gemm_f16_tc(static_cast<const op_t*>(A), split == 2 ? 2 * K : K, static_cast<const op_t*>(W), 2 * K, N, p, 256,
                     static_cast<cudaStream_t>(stream));
}

int made_gemm_f16_split_h(const void* A, const void* W, int64_t M, int N, int K, int split, const float* bias, int act,
                          void* out16, void* stream) {
  MADE_REQUIRE(split == 1 || split == 2, "gemm_f16_split_h: split must be 1 (W pairs) or 2 (A and W pairs)");
  MADE_REQUIRE(out16, "gemm_f16_split_h: no output");
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.split = split;
  p.epi.bias = bias;
  p.epi.act = act;
  p.epi.out_h = static_cast<op_t*>(out16);
  p.epi.ld_h = N;
  return gemm_f16_tc(static_cast<const op_t*>(A), split == 2 ? 2 * K : K, static_cast<const op_t*>(W), 2 * K, N, p, 256,
                     static_cast<cudaStream_t>(stream));
}

int made_ffn_fused(const void* x, int64_t ldx, const void* w1, const float* b1, const void* w2, const float* b2, int act,
                   const void* residual_pair, int64_t res_ld, const float* ln_gamma, const float* ln_beta, void* out_pair,
                   int64_t ld_out, int has_lo, int64_t M, void* stream) {
  const op_t* res = static_cast<const op_t*>(residual_pair);
  op_t* out = static_cast<op_t*>(out_pair);
  return ffn_fused(static_cast<const op_t*>(x), ldx, static_cast<const op_t*>(w1), b1, static_cast<const op_t*>(w2), b2,
                   act, res, res && has_lo ? res + D : nullptr, res_ld, ln_gamma, ln_beta, out, has_lo ? out + D : nullptr,
                   ld_out, nullptr, 0, nullptr, 0, M, nullptr, static_cast<cudaStream_t>(stream));
}

int made_mha_core(const void* q, const void* k, const void* v, const float* key_mask, int64_t B, int L, void* out,
                  void* stream) {
  return mha_core(static_cast<const op_t*>(q), D, static_cast<const op_t*>(k), D,
                  static_cast<const op_t*>(v), D, key_mask, B, L, static_cast<op_t*>(out), D,
                  static_cast<cudaStream_t>(stream));
}

}  // extern "C"
