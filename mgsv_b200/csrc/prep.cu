// Small memory-bound kernels around the GEMMs: row LayerNorm, span/class heads, segment-mask bit
// packing and v_hat preparation (the token-packing kernels live in ragged.cu).  All are vectorised (16-byte accesses), one warp
// per 256-wide row with shuffle reductions.
#include "common.cuh"
#include "prep.cuh"

namespace made {

// Row LayerNorm over 256 features: warp per row, 8 features per lane. in fp32 or fp16 -> fp16/fp32.
template <typename TIn>
__global__ void layernorm_rows_kernel(const TIn* __restrict__ in, int64_t ld_in, int64_t rows,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, op_t* __restrict__ out_h, int64_t ld_out,
                                      op_t* __restrict__ out_lo, float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[8];
  if constexpr (sizeof(TIn) == 4) {
    const float4* p = reinterpret_cast<const float4*>(in + row * ld_in + lane * 8);
    float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 a = *reinterpret_cast<const uint4*>(in + row * ld_in + lane * 8);
    const op2_t* h = reinterpret_cast<const op2_t*>(&a);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = op2_to_f2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = warp_sum(s) * (1.0f / 256);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { float d = v[j] - mean; sq = fmaf(d, d, sq); }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / 256) + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + lane * 8), g1 = *reinterpret_cast<const float4*>(gamma + lane * 8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + lane * 8), b1 = *reinterpret_cast<const float4*>(beta + lane * 8 + 4);
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j] + bb[j];
  if (out_h) {
    const uint4 hi = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
    *reinterpret_cast<uint4*>(out_h + row * ld_out + lane * 8) = hi;
    if (out_lo) {
      const op2_t* hh = reinterpret_cast<const op2_t*>(&hi);
      uint32_t lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = op2_to_f2(hh[j]);
        lo[j] = pack_op2(v[2 * j] - f.x, v[2 * j + 1] - f.y);
      }
      *reinterpret_cast<uint4*>(out_lo + row * ld_out + lane * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out_f32 + row * 256 + lane * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// Output heads on hs rows (model_Uni.py:131-135): logits = class_embed(hs) [2];
// spans = sigmoid(span_embed.layers.2(h2)) [2] with h2 = the MLP's second hidden layer (fp16).
// Warp per row.
__global__ void heads_final_kernel(const float* __restrict__ hs, const op_t* __restrict__ h2,
                                   int64_t rows, const float* __restrict__ w_cls, const float* __restrict__ b_cls,
                                   const float* __restrict__ w_sp, const float* __restrict__ b_sp,
                                   float* __restrict__ logits, float* __restrict__ spans) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    const float x = hs[row * 256 + c];
    const float y = op2f(h2[row * 256 + c]);
    a[0] = fmaf(x, w_cls[c], a[0]);
    a[1] = fmaf(x, w_cls[256 + c], a[1]);
    a[2] = fmaf(y, w_sp[c], a[2]);
    a[3] = fmaf(y, w_sp[256 + c], a[3]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = warp_sum(a[k]);
  if (lane == 0) {
    logits[row * 2] = a[0] + b_cls[0];
    logits[row * 2 + 1] = a[1] + b_cls[1];
    spans[row * 2] = 1.f / (1.f + expf(-(a[2] + b_sp[0])));
    spans[row * 2 + 1] = 1.f / (1.f + expf(-(a[3] + b_sp[1])));
  }
}

// segment mask [n,96] float -> 3 x 32 validity bits (+1 pad word) per track
__global__ void mask_bits_kernel(const float* __restrict__ mask, int64_t n, uint32_t* __restrict__ bits) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * 4) return;
  const int64_t tr = i >> 2;
  const int w = static_cast<int>(i & 3);
  uint32_t v = 0;
  if (w < 3)
    for (int k = 0; k < 32; ++k) v |= (mask[tr * 96 + w * 32 + k] != 0.f ? 1u : 0u) << k;
  bits[i] = v;
}

// v_hat = v / |v| (modules/metrics.py:19), fp32.  Warp per row.
__global__ void vhat_kernel(const float* __restrict__ v, int64_t rows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x[8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = v[row * 256 + lane * 8 + j]; s = fmaf(x[j], x[j], s); }
  const float nrm = sqrtf(warp_sum(s));
  float4* o = reinterpret_cast<float4*>(out + row * 256 + lane * 8);
  o[0] = make_float4(x[0] / nrm, x[1] / nrm, x[2] / nrm, x[3] / nrm);
  o[1] = make_float4(x[4] / nrm, x[5] / nrm, x[6] / nrm, x[7] / nrm);
}

// ---- launchers ---------------------------------------------------------------------------
int layernorm_rows(const void* in, int in_is_op, int64_t ld_in, int64_t rows, const float* gamma,
                   const float* beta, op_t* out_h, int64_t ld_out, op_t* out_lo, float* out_f32, cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  const unsigned blocks = static_cast<unsigned>(ceil_div64(rows, 8));
  if (in_is_op)
    layernorm_rows_kernel<op_t><<<blocks, 256, 0, st>>>(static_cast<const op_t*>(in), ld_in,
                                                                 rows, gamma, beta, 1e-5f, out_h, ld_out, out_lo, out_f32);
  else
    layernorm_rows_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(in), ld_in, rows, gamma,
                                                         beta, 1e-5f, out_h, ld_out, out_lo, out_f32);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int heads_final(const float* hs, const op_t* h2, int64_t rows, const float* w_cls,
                const float* b_cls, const float* w_sp, const float* b_sp, float* logits, float* spans,
                cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  heads_final_kernel<<<static_cast<unsigned>(ceil_div64(rows, 8)), 256, 0, st>>>(hs, h2, rows, w_cls, b_cls,
                                                                                 w_sp, b_sp, logits, spans);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int mask_bits(const float* mask, int64_t n, uint32_t* bits, cudaStream_t st) {
  if (n == 0) return MADE_OK;
  mask_bits_kernel<<<static_cast<unsigned>(ceil_div64(n * 4, 256)), 256, 0, st>>>(mask, n, bits);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int vhat_rows(const float* v, int64_t rows, float* out, cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  vhat_kernel<<<static_cast<unsigned>(ceil_div64(rows, 8)), 256, 0, st>>>(v, rows, out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
