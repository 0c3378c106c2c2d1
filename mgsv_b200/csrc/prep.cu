// Small memory-bound kernels around the GEMMs: input cast + mask-zero, row LayerNorm, masked mean
// pooling + L2 norm, DETR input assembly with the sine position embedding, span/class heads,
// segment-mask bit packing and v_hat preparation.  All are vectorised (16-byte accesses), one warp
// per 256-wide row with shuffle reductions.
#include "common.cuh"
#include "prep.cuh"

namespace made {

// x[t, :] = mask[t] ? in[t, :] : 0  -> fp16     (model_Base.py:556 / :595 masked_fill + cast)
// Rows with mask == 0 are never read (the reference's dataloader zero-pads them and the model
// overwrites them with 0 anyway), so a padded feature tensor costs only its valid rows of traffic.
// kIn: 0 = fp32, 1 = bf16, 2 = fp16 (MADE_DTYPE_*).
template <int kIn>
__global__ void cast_mask_rows_kernel(const void* __restrict__ in_, const float* __restrict__ mask,
                                      int64_t rows, int dim, op_t* __restrict__ out) {
  const int vec = dim / 8;
  const int64_t total = rows * vec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / vec;
    const int c = static_cast<int>(i % vec) * 8;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (mask[row] != 0.f) {
      float v[8];
      if constexpr (kIn == MADE_DTYPE_F32) {
        const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(in_) + row * dim + c);
        float4 a = __ldcs(p), b = __ldcs(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else if constexpr (kIn == MADE_DTYPE_BF16) {
        uint4 a = __ldcs(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(in_) + row * dim + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
      } else {
        uint4 a = __ldcs(reinterpret_cast<const uint4*>(static_cast<const __half*>(in_) + row * dim + c));
        const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
      }
      o = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
    }
    *reinterpret_cast<uint4*>(out + row * dim + c) = o;
  }
}

// Row LayerNorm over 256 features: warp per row, 8 features per lane. in fp32 or fp16 -> fp16/fp32.
template <typename TIn>
__global__ void layernorm_rows_kernel(const TIn* __restrict__ in, int64_t ld_in, int64_t rows,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, op_t* __restrict__ out_h,
                                      float* __restrict__ out_f32) {
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
  if (out_h)
    *reinterpret_cast<uint4*>(out_h + row * 256 + lane * 8) =
        make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out_f32 + row * 256 + lane * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// pooled[b] = normalize( sum_t seq[b,t,:] / sum_t mask[b,t] )   (model_Base.py:579-580 / :615-616)
// seq rows of padded positions are already zero.  One CTA (256 threads = features) per sequence.
__global__ void __launch_bounds__(256)
pool_norm_kernel(const float* __restrict__ seq, const float* __restrict__ mask, int L,
                 float* __restrict__ pooled) {
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  const int d = threadIdx.x;
  float acc = 0.f, cnt = 0.f;
  for (int t = 0; t < L; ++t) {
    acc += seq[(b * L + t) * 256 + d];
    cnt += mask[b * L + t];
  }
  float v = acc / cnt;
  float s = warp_sum(v * v);
  if ((d & 31) == 0) red[d >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  pooled[b * 256 + d] = v / fmaxf(sqrtf(tot), 1e-12f);
}

// DETR input assembly (model_Uni.py:207-216 + position_encoding.py:51-71):
//   src[b]  = cat(frame_out[b] (50), seg_out[track_idx[b]] (96))           fp16 [B,146,256]
//   mask[b] = cat(frame_mask[b], seg_mask[track_idx[b]])                   f32  [B,146]
//   pos[b,t,2j] = sin(x/dim_t), pos[b,t,2j+1] = cos(x/dim_t),  x = cumsum(mask)/(total+1e-6)*2pi
// One CTA per sequence; warp per token row.
__global__ void __launch_bounds__(256)
detr_prep_kernel(const op_t* __restrict__ frame_out, const float* __restrict__ frame_mask,
                 const op_t* __restrict__ seg_out, const float* __restrict__ seg_mask,
                 const int32_t* __restrict__ track_idx, int64_t seq_offset,
                 const float* __restrict__ inv_dim_t, op_t* __restrict__ src, op_t* __restrict__ pos,
                 op_t* __restrict__ srcpos, float* __restrict__ mask_out) {
  constexpr int LV = 50, LM = 96, L = 146;
  __shared__ float sx[L];
  const int64_t b = blockIdx.x;
  const int64_t tr = track_idx ? track_idx[b] : seq_offset + b;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < L) {
    float mk = tid < LV ? frame_mask[b * LV + tid] : seg_mask[tr * LM + (tid - LV)];
    sx[tid] = mk;
    mask_out[b * L + tid] = mk;
  }
  __syncthreads();
  if (tid == 0) {
    float c = 0.f;
    for (int t = 0; t < L; ++t) { c += sx[t]; sx[t] = c; }   // cumsum in fp32, sequential like torch
    const float denom = c + 1e-6f;
    for (int t = 0; t < L; ++t) sx[t] = sx[t] / denom * 6.283185307179586f;
  }
  __syncthreads();
  for (int t = warp; t < L; t += 8) {
    const op_t* sp = t < LV ? frame_out + (b * LV + t) * 256 : seg_out + (tr * LM + (t - LV)) * 256;
    uint4 raw = *reinterpret_cast<const uint4*>(sp + lane * 8);
    const op2_t* h = reinterpret_cast<const op2_t*>(&raw);
    const float x = sx[t];
    float pv[8], sv[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = x * inv_dim_t[lane * 8 + 2 * j];   // same dim_t for the (sin, cos) pair
      pv[2 * j] = sinf(a);
      pv[2 * j + 1] = cosf(a);
      float2 f = op2_to_f2(h[j]);
      sv[2 * j] = f.x;
      sv[2 * j + 1] = f.y;
    }
    const int64_t o = (b * L + t) * 256 + lane * 8;
    *reinterpret_cast<uint4*>(src + o) = raw;
    *reinterpret_cast<uint4*>(pos + o) = make_uint4(pack_op2(pv[0], pv[1]), pack_op2(pv[2], pv[3]),
                                                    pack_op2(pv[4], pv[5]), pack_op2(pv[6], pv[7]));
    *reinterpret_cast<uint4*>(srcpos + o) =
        make_uint4(pack_op2(sv[0] + pv[0], sv[1] + pv[1]), pack_op2(sv[2] + pv[2], sv[3] + pv[3]),
                   pack_op2(sv[4] + pv[4], sv[5] + pv[5]), pack_op2(sv[6] + pv[6], sv[7] + pv[7]));
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

// v_hat = v / |v| (modules/metrics.py:19) -> fp16.  Warp per row.
__global__ void vhat_kernel(const float* __restrict__ v, int64_t rows, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x[8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = v[row * 256 + lane * 8 + j]; s = fmaf(x[j], x[j], s); }
  const float nrm = sqrtf(warp_sum(s));
  __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(x[2 * j] / nrm, x[2 * j + 1] / nrm);
  *reinterpret_cast<uint4*>(out + row * 256 + lane * 8) = *reinterpret_cast<uint4*>(h);
}

// ---- launchers ---------------------------------------------------------------------------
int cast_mask_rows(const void* in, int in_dtype, const float* mask, int64_t rows, int dim,
                   op_t* out, cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  MADE_REQUIRE(dim % 8 == 0, "cast_mask_rows: dim must be a multiple of 8");
  const int64_t total = rows * (dim / 8);
  int64_t blocks = ceil_div64(total, 256);
  int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  // Pinned host memory (UVA): the kernel pulls the valid rows over PCIe itself.  Two CTAs per SM
  // keep ~2 MB of reads in flight (far more than the link needs) while leaving the SMs' thread and
  // shared-memory slots free for the GEMM kernels of the previous chunk running on another stream.
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, in) == cudaSuccess && attr.type == cudaMemoryTypeHost)
    cap = static_cast<int64_t>(sm_count()) * 2;
  else
    (void)cudaGetLastError();
  if (blocks > cap) blocks = cap;
  const unsigned g = static_cast<unsigned>(blocks);
  if (in_dtype == MADE_DTYPE_F32)
    cast_mask_rows_kernel<MADE_DTYPE_F32><<<g, 256, 0, st>>>(in, mask, rows, dim, out);
  else if (in_dtype == MADE_DTYPE_BF16)
    cast_mask_rows_kernel<MADE_DTYPE_BF16><<<g, 256, 0, st>>>(in, mask, rows, dim, out);
  else
    cast_mask_rows_kernel<MADE_DTYPE_F16><<<g, 256, 0, st>>>(in, mask, rows, dim, out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int layernorm_rows(const void* in, int in_is_op, int64_t ld_in, int64_t rows, const float* gamma,
                   const float* beta, op_t* out_h, float* out_f32, cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  const unsigned blocks = static_cast<unsigned>(ceil_div64(rows, 8));
  if (in_is_op)
    layernorm_rows_kernel<op_t><<<blocks, 256, 0, st>>>(static_cast<const op_t*>(in), ld_in,
                                                                 rows, gamma, beta, 1e-5f, out_h, out_f32);
  else
    layernorm_rows_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(in), ld_in, rows, gamma,
                                                         beta, 1e-5f, out_h, out_f32);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int pool_norm(const float* seq, const float* mask, int64_t B, int L, float* pooled, cudaStream_t st) {
  if (B == 0) return MADE_OK;
  pool_norm_kernel<<<static_cast<unsigned>(B), 256, 0, st>>>(seq, mask, L, pooled);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int detr_prep(const op_t* frame_out, const float* frame_mask, const op_t* seg_out,
              const float* seg_mask, const int32_t* track_idx, int64_t seq_offset, const float* inv_dim_t,
              int64_t B, op_t* src, op_t* pos, op_t* srcpos, float* mask_out,
              cudaStream_t st) {
  if (B == 0) return MADE_OK;
  detr_prep_kernel<<<static_cast<unsigned>(B), 256, 0, st>>>(frame_out, frame_mask, seg_out, seg_mask,
                                                             track_idx, seq_offset, inv_dim_t, src, pos, srcpos,
                                                             mask_out);
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

int vhat_rows(const float* v, int64_t rows, __half* out, cudaStream_t st) {
  if (rows == 0) return MADE_OK;
  vhat_kernel<<<static_cast<unsigned>(ceil_div64(rows, 8)), 256, 0, st>>>(v, rows, out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
