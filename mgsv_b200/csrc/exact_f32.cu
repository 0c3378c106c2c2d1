// fp32 CUDA-core path (MADE_PREC_FP32): the temporal encoders and X-Pool in the reference's own arithmetic —
// fp32 operands, fp32 FFMA accumulation, erff / expf / exact divisions, the reference's operation order per
// module (model_Base.py:82-91, 520-617; modules/transformer.py:87-180; modules/metrics.py:10-24) — for the north
// star's "1e-5 in fp32" similarity bar, and to MATERIALISE Transformer_XA's [N_m, N_v, 256] output when a caller
// really wants that tensor (the compat view of model.video_guided_to_music_pooling_cross_transformer).
// Throughput is not the point here (SIMT tiles, ~10 TFLOP/s): the tcgen05 path is the product; this is its
// high-precision mode.
#include <type_traits>

#include "common.cuh"
#include "prep.cuh"

namespace made {

namespace {

constexpr int kTM = 64, kTN = 64, kTK = 16;

// C[b] = act(alpha * A[b] op(B[b]) + bias) + residual[b];  A [M,K] row-major; B [N,K] (b_nk) or [K,N]
struct SgemmP {
  const float* A; int64_t lda, sA;
  const float* B; int64_t ldb, sB; int b_nk;
  float* C; int64_t ldc, sC;
  int M, N, K;
  const float* bias;            // [N]
  const float* residual; int64_t ldr, sR;
  int act;                      // 0 none, 1 GELU(erf), 2 ReLU
  float alpha;
};

__global__ void __launch_bounds__(256)
sgemm_kernel(const SgemmP p) {
  __shared__ float sA[kTK][kTM + 4];
  __shared__ float sB[kTK][kTN + 4];
  const int b = blockIdx.z;
  const float* A = p.A + b * p.sA;
  const float* B = p.B + b * p.sB;
  float* C = p.C + b * p.sC;
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += kTK) {
    // A tile: 64 rows x 16 k; 256 threads x 4 elements
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      const int r = e >> 4, kk = e & 15;
      const int gm = m0 + r, gk = k0 + kk;
      sA[kk][r] = (gm < p.M && gk < p.K) ? A[static_cast<int64_t>(gm) * p.lda + gk] : 0.f;
    }
    if (p.b_nk) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + i * 256;
        const int r = e >> 4, kk = e & 15;
        const int gn = n0 + r, gk = k0 + kk;
        sB[kk][r] = (gn < p.N && gk < p.K) ? B[static_cast<int64_t>(gn) * p.ldb + gk] : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + i * 256;
        const int kk = e >> 6, c = e & 63;
        const int gn = n0 + c, gk = k0 + kk;
        sB[kk][c] = (gn < p.N && gk < p.K) ? B[static_cast<int64_t>(gk) * p.ldb + gn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = sA[kk][ty * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) bb[v] = sB[kk][tx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], bb[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int gm = m0 + ty * 4 + u;
    if (gm >= p.M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gn = n0 + tx * 4 + v;
      if (gn >= p.N) continue;
      float x = acc[u][v] * p.alpha;
      if (p.bias) x += p.bias[gn];
      if (p.act == 1) x = 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
      else if (p.act == 2) x = fmaxf(x, 0.f);
      if (p.residual) x += p.residual[b * p.sR + static_cast<int64_t>(gm) * p.ldr + gn];
      C[static_cast<int64_t>(gm) * p.ldc + gn] = x;
    }
  }
}

int sgemm(const SgemmP& p, int batch, cudaStream_t st) {
  if (p.M == 0 || p.N == 0 || batch == 0) return MADE_OK;
  dim3 grid((p.N + kTN - 1) / kTN, (p.M + kTM - 1) / kTM, batch);
  sgemm_kernel<<<grid, 256, 0, st>>>(p);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

SgemmP linear_p(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc, int M, int N, int K) {
  SgemmP p{};
  p.A = A; p.lda = lda; p.B = W; p.ldb = K; p.b_nk = 1; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.alpha = 1.f;
  return p;
}

// rows of 256: x[r] = (mask ? x[r] : 0) + add[(r % L)]   (model_Base.py:556 masked_fill is on the INPUT; :533 += pe)
__global__ void add_pe_kernel(float* __restrict__ x, const float* __restrict__ pe, int64_t rows, int L) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * 256) return;
  const int64_t r = i >> 8;
  x[i] += pe[(r % L) * 256 + (i & 255)];
}

// masked copy of raw features: out[r, :] = mask[r] ? in[r, :] : 0  (fp32 / bf16 / fp16 in)
template <typename T>
__global__ void mask_rows_kernel(const T* __restrict__ in, const float* __restrict__ mask, int64_t rows, int dim,
                                 float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * dim) return;
  const int64_t r = i / dim;
  float v = 0.f;
  if (mask[r] != 0.f) {
    if constexpr (sizeof(T) == 4) v = in[i];
    else if constexpr (std::is_same<T, __half>::value) v = __half2float(in[i]);
    else v = __bfloat162float(in[i]);
  }
  out[i] = v;
}

// nn.MultiheadAttention core in fp32: one block per (sequence, head), thread = query row; key padding mask.
// qkv [B*L, 768] (q | k | v), out [B*L, 256].
__global__ void __launch_bounds__(128)
mha_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ mask, int L, float* __restrict__ out) {
  extern __shared__ float sm[];        // K [L][32], V [L][32], valid [L]
  float* sK = sm;
  float* sV = sm + L * 32;
  float* sM = sV + L * 32;
  const int b = blockIdx.x, h = blockIdx.y;
  for (int i = threadIdx.x; i < L * 32; i += blockDim.x) {
    const int t = i >> 5, d = i & 31;
    const float* row = qkv + (static_cast<int64_t>(b) * L + t) * 768;
    sK[i] = row[256 + h * 32 + d];
    sV[i] = row[512 + h * 32 + d];
  }
  for (int t = threadIdx.x; t < L; t += blockDim.x) sM[t] = mask[static_cast<int64_t>(b) * L + t];
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= L) return;
  float q[32];
  const float* qrow = qkv + (static_cast<int64_t>(b) * L + t) * 768 + h * 32;
  const float scale = 0.17677669529663687f;     // 1 / sqrt(32); torch scales q before q k^T
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] = qrow[d] * scale;
  float mx = -INFINITY;
  for (int s = 0; s < L; ++s) {
    if (sM[s] == 0.f) continue;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) a = fmaf(q[d], sK[s * 32 + d], a);
    mx = fmaxf(mx, a);
  }
  float acc[32] = {};
  float den = 0.f;
  for (int s = 0; s < L; ++s) {
    if (sM[s] == 0.f) continue;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) a = fmaf(q[d], sK[s * 32 + d], a);
    const float e = expf(a - mx);
    den += e;
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = fmaf(e, sV[s * 32 + d], acc[d]);
  }
  float* o = out + (static_cast<int64_t>(b) * L + t) * 256 + h * 32;
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = acc[d] / den;
}

// masked mean over time + L2 normalise; also zeroes the padded rows of seq in place (model_Base.py:541, 579-580)
__global__ void __launch_bounds__(256)
pool_f32_kernel(float* __restrict__ seq, const float* __restrict__ mask, int L, float* __restrict__ pooled) {
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  const int d = threadIdx.x;
  float acc = 0.f, n = 0.f;
  for (int t = 0; t < L; ++t) {
    const float mk = mask[b * L + t];
    float* p = seq + (b * L + t) * 256 + d;
    if (mk == 0.f) *p = 0.f;
    else acc += *p;
    n += mk;
  }
  const float v = acc / n;
  const float s = warp_sum(v * v);
  if ((d & 31) == 0) red[d >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  pooled[b * 256 + d] = v / fmaxf(sqrtf(tot), 1e-12f);
}

// rows of L <= 96 logits per (key sequence, query): masked softmax in place.  logits [n_seq * n_q, L], mask [n_seq, L]
__global__ void softmax_keys_kernel(float* __restrict__ logits, const float* __restrict__ mask, int64_t n_q, int64_t rows,
                                    int L) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t m = row / n_q;
  float v[3];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int t = lane + 32 * j;
    v[j] = (t < L && mask[m * L + t] != 0.f) ? logits[row * L + t] : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) { v[j] = expf(v[j] - mx); s += v[j]; }
  s = warp_sum(s);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (lane + 32 * j < L) logits[row * L + lane + 32 * j] = v[j] / s;
}

// sim[v, col0 + m] = < video[v] / |video[v]|, pooled[m, v] / |pooled[m, v]| >   (modules/metrics.py:19-23); warp per pair
__global__ void pooled_cos_kernel(const float* __restrict__ video, const float* __restrict__ pooled, int64_t n_q,
                                  int64_t n_m, float* __restrict__ sim, int64_t ld, int64_t col0) {
  const int lane = threadIdx.x & 31;
  const int64_t pair = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= n_q * n_m) return;
  const int64_t m = pair / n_q, v = pair % n_q;
  const float4* a = reinterpret_cast<const float4*>(video + v * 256 + lane * 8);
  const float4* b = reinterpret_cast<const float4*>(pooled + pair * 256 + lane * 8);
  const float4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float na = 0.f, nb = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { na = fmaf(av[j], av[j], na); nb = fmaf(bv[j], bv[j], nb); }
  na = sqrtf(warp_sum(na));
  nb = sqrtf(warp_sum(nb));
  float d = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) d = fmaf(av[j] / na, bv[j] / nb, d);
  d = warp_sum(d);
  if (lane == 0) sim[v * ld + col0 + m] = d;
}

}  // namespace

// ---- host orchestration ------------------------------------------------------------------------------------
int exact_encode(const ExactEncW& w, const void* feats, int feats_dtype, const float* masks, int64_t B, float* ws,
                 float* seq_f32, float* pooled, cudaStream_t st) {
  const int L = w.L, din = w.din;
  const int64_t T = B * L;
  float* x0 = ws;                       // [T, din]
  float* x1 = x0 + T * din;             // [T, 256]
  float* qkv = x1 + T * 256;            // [T, 768]
  float* att = qkv + T * 768;           // [T, 256]
  float* x2 = att + T * 256;            // [T, 256]
  float* hb = x2 + T * 256;             // [T, 1024]
  float* x3 = hb + T * 1024;            // [T, 256]
  const unsigned blk = static_cast<unsigned>(ceil_div64(T * din, 256));
  if (feats_dtype == MADE_DTYPE_F32)
    mask_rows_kernel<float><<<blk, 256, 0, st>>>(static_cast<const float*>(feats), masks, T, din, x0);
  else if (feats_dtype == MADE_DTYPE_F16)
    mask_rows_kernel<__half><<<blk, 256, 0, st>>>(static_cast<const __half*>(feats), masks, T, din, x0);
  else
    mask_rows_kernel<__nv_bfloat16><<<blk, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(feats), masks, T, din, x0);
  MADE_CHECK_LAUNCH();
  const int M = static_cast<int>(T);
  MADE_TRY(sgemm(linear_p(x0, din, w.proj_w, w.proj_b, x1, 256, M, 256, din), 1, st));
  add_pe_kernel<<<static_cast<unsigned>(ceil_div64(T * 256, 256)), 256, 0, st>>>(x1, w.pe, T, L);
  MADE_CHECK_LAUNCH();
  MADE_TRY(layernorm_rows(x1, 0, 256, T, w.ln1_g, w.ln1_b, nullptr, 256, nullptr, x1, st));
  MADE_TRY(sgemm(linear_p(x1, 256, w.in_w, w.in_b, qkv, 768, M, 768, 256), 1, st));
  mha_f32_kernel<<<dim3(static_cast<unsigned>(B), 8), 128, (2 * L * 32 + L) * sizeof(float), st>>>(qkv, masks, L, att);
  MADE_CHECK_LAUNCH();
  {
    SgemmP p = linear_p(att, 256, w.out_w, w.out_b, x2, 256, M, 256, 256);
    p.residual = x1; p.ldr = 256;                 // residual from the NORMED tensor (Q2)
    MADE_TRY(sgemm(p, 1, st));
  }
  MADE_TRY(layernorm_rows(x2, 0, 256, T, w.ln2_g, w.ln2_b, nullptr, 256, nullptr, x2, st));
  {
    SgemmP p = linear_p(x2, 256, w.ff1_w, w.ff1_b, hb, 1024, M, 1024, 256);
    p.act = 1;
    MADE_TRY(sgemm(p, 1, st));
  }
  {
    SgemmP p = linear_p(hb, 1024, w.ff2_w, w.ff2_b, x3, 256, M, 256, 1024);
    p.residual = x2; p.ldr = 256;
    MADE_TRY(sgemm(p, 1, st));
  }
  MADE_TRY(sgemm(linear_p(x3, 256, w.fin_w, w.fin_b, seq_f32, 256, M, 256, 256), 1, st));
  pool_f32_kernel<<<static_cast<unsigned>(B), 256, 0, st>>>(seq_f32, masks, L, pooled);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

size_t exact_encode_ws_floats(int64_t B, int L, int din) {
  const size_t T = static_cast<size_t>(B) * L;
  return T * (static_cast<size_t>(din) + 256 + 768 + 256 + 256 + 1024 + 256);
}

// Transformer_XA.forward (modules/transformer.py:156-180) materialised: pooled [n_m, n_q, 256] for n_m key sequences
// of L <= 96 tokens (L = 96 music segments for the video-guided module, 50 frames for the music-guided one).
// ws: n_q*512 + n_m*L*768 + n_m*n_q*(L + 256 + 256) floats.
int exact_xpool(const ExactXpW& w, const float* video, int64_t n_q, const float* seg, const float* seg_mask, int64_t n_m,
                int L, float* ws, float* pooled, cudaStream_t st) {
  float* vln = ws;                          // [n_q, 256]
  float* q = vln + n_q * 256;               // [n_q, 256]
  float* sln = q + n_q * 256;               // [n_m*96, 256]
  float* kk = sln + n_m * L * 256;         // [n_m*96, 256]
  float* vv = kk + n_m * L * 256;          // [n_m*96, 256]
  float* logit = vv + n_m * L * 256;       // [n_m, n_q, 96]
  float* att = logit + n_m * n_q * L;      // [n_m, n_q, 256]
  float* o = att + n_m * n_q * 256;         // [n_m, n_q, 256]
  const int NQ = static_cast<int>(n_q);
  const int64_t R = n_m * n_q;
  MADE_TRY(layernorm_rows(video, 0, 256, n_q, w.ln1_g, w.ln1_b, nullptr, 256, nullptr, vln, st));
  MADE_TRY(layernorm_rows(seg, 0, 256, n_m * L, w.ln1_g, w.ln1_b, nullptr, 256, nullptr, sln, st));
  MADE_TRY(sgemm(linear_p(vln, 256, w.q_w, w.q_b, q, 256, NQ, 256, 256), 1, st));
  MADE_TRY(sgemm(linear_p(sln, 256, w.k_w, w.k_b, kk, 256, static_cast<int>(n_m * L), 256, 256), 1, st));
  MADE_TRY(sgemm(linear_p(sln, 256, w.v_w, w.v_b, vv, 256, static_cast<int>(n_m * L), 256, 256), 1, st));
  {   // logits[m] = q K_m^T / sqrt(256)   (:111)
    SgemmP p{};
    p.A = q; p.lda = 256; p.sA = 0; p.B = kk; p.ldb = 256; p.sB = static_cast<int64_t>(L) * 256; p.b_nk = 1;
    p.C = logit; p.ldc = L; p.sC = n_q * L; p.M = NQ; p.N = L; p.K = 256; p.alpha = 0.0625f;
    MADE_TRY(sgemm(p, static_cast<int>(n_m), st));
  }
  softmax_keys_kernel<<<static_cast<unsigned>(ceil_div64(R, 8)), 256, 0, st>>>(logit, seg_mask, n_q, R, L);
  MADE_CHECK_LAUNCH();
  {   // attention[m] = softmax V_m
    SgemmP p{};
    p.A = logit; p.lda = L; p.sA = n_q * L; p.B = vv; p.ldb = 256; p.sB = static_cast<int64_t>(L) * 256; p.b_nk = 0;
    p.C = att; p.ldc = 256; p.sC = n_q * 256; p.M = NQ; p.N = 256; p.K = L; p.alpha = 1.f;
    MADE_TRY(sgemm(p, static_cast<int>(n_m), st));
  }
  MADE_TRY(sgemm(linear_p(att, 256, w.o_w, w.o_b, o, 256, static_cast<int>(R), 256, 256), 1, st));     // out_proj (:122)
  MADE_TRY(layernorm_rows(o, 0, 256, R, w.ln2_g, w.ln2_b, nullptr, 256, nullptr, o, st));              // layer_norm2 (:174)
  {   // attn_out + linear_proj(attn_out)  (:176-177)
    SgemmP p = linear_p(o, 256, w.l_w, w.l_b, pooled, 256, static_cast<int>(R), 256, 256);
    p.residual = o; p.ldr = 256;
    MADE_TRY(sgemm(p, 1, st));
  }
  MADE_TRY(layernorm_rows(pooled, 0, 256, R, w.ln3_g, w.ln3_b, nullptr, 256, nullptr, pooled, st));   // layer_norm3 (:178)
  return MADE_OK;
}

size_t exact_xpool_ws_floats(int64_t n_q, int64_t n_m, int L) {
  return static_cast<size_t>(n_q) * 512 + static_cast<size_t>(n_m) * L * 768 + static_cast<size_t>(n_m) * n_q * (L + 256 + 256);
}

int exact_pooled_cosine(const float* video, const float* pooled, int64_t n_q, int64_t n_m, float* sim, int64_t ld,
                        int64_t col0, cudaStream_t st) {
  if (n_q == 0 || n_m == 0) return MADE_OK;
  pooled_cos_kernel<<<static_cast<unsigned>(ceil_div64(n_q * n_m, 8)), 256, 0, st>>>(video, pooled, n_q, n_m, sim, ld, col0);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
