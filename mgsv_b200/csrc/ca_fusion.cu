// Cross-attention core of the `mml_fusion = "CA"` variant: CrossAttention.forward of model/model_Base.py:127-165
// (8 heads x 128, queries = the 96 music segments, keys / values = the 50 video frames of the paired video).
//   dots = q k^T * 128^-0.5 ; keys with kv_mask == 0 -> -inf (before the softmax) ; attn = softmax(dots) ;
//   query rows with q_mask == 0 -> attn = 0 (after the softmax) ; out = attn v, heads concatenated.
// The projections around it (to_q, to_kv, to_out, FeedForward, final_linear) are tcgen05 GEMMs issued by
// made_ca_fuse (api.cu); this kernel is 0.02 GFLOP per pair and runs on the CUDA cores: one CTA per (pair, head),
// K and V of the head staged in shared memory as fp16, lane = key for the logits (two keys per lane, the query row
// broadcast from shared memory), lane = 4 output features for attn v.
#include "common.cuh"
#include "prep.cuh"

namespace made {

namespace {

constexpr int kCaLq = 96, kCaLk = 50, kCaHeads = 8, kCaDh = 128, kCaInner = kCaHeads * kCaDh;   // 1024
constexpr int kCaKStride = kCaDh + 2;      // halves per staged K row: 65 words -> lanes (= keys) hit distinct banks

__global__ void __launch_bounds__(128)
ca_attention_kernel(const op_t* __restrict__ q, const op_t* __restrict__ kv, const float* __restrict__ q_mask,
                    const float* __restrict__ kv_mask, op_t* __restrict__ out) {
  __shared__ op_t sK[kCaLk * kCaKStride];
  __shared__ __align__(16) op_t sV[kCaLk * kCaDh];
  __shared__ float sQ[4][kCaDh];
  __shared__ float sMask[kCaLk];
  const int64_t b = blockIdx.x;
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const op_t* kbase = kv + b * kCaLk * (2 * kCaInner) + h * kCaDh;
  for (int i = threadIdx.x; i < kCaLk * (kCaDh / 2); i += 128) {      // pairs of halves
    const int j = i / (kCaDh / 2), d2 = i % (kCaDh / 2);
    const op2_t kk = *reinterpret_cast<const op2_t*>(kbase + static_cast<int64_t>(j) * (2 * kCaInner) + 2 * d2);
    const op2_t vv = *reinterpret_cast<const op2_t*>(kbase + static_cast<int64_t>(j) * (2 * kCaInner) + kCaInner + 2 * d2);
    *reinterpret_cast<op2_t*>(sK + j * kCaKStride + 2 * d2) = kk;
    *reinterpret_cast<op2_t*>(sV + j * kCaDh + 2 * d2) = vv;
  }
  if (threadIdx.x < kCaLk) sMask[threadIdx.x] = kv_mask[b * kCaLk + threadIdx.x];
  __syncthreads();
  const float scale = 0.08838834764831845f;     // 128^-0.5
  const int j0 = lane, j1 = lane + 32;          // this lane's keys
  const bool ok0 = sMask[j0] != 0.f;
  const bool ok1 = j1 < kCaLk && sMask[j1] != 0.f;
  for (int i = warp; i < kCaLq; i += 4) {
    const int64_t row = b * kCaLq + i;
    op_t* o = out + row * kCaInner + h * kCaDh + lane * 4;
    if (q_mask[row] == 0.f) {                   // attn.masked_fill(q_mask == 0, 0): the row's output is 0
      *reinterpret_cast<uint2*>(o) = make_uint2(0u, 0u);
      continue;
    }
    {   // query row -> shared memory (fp32), 4 features per lane
      const uint2 raw = *reinterpret_cast<const uint2*>(q + row * kCaInner + h * kCaDh + lane * 4);
      const float2 a = op2_to_f2(*reinterpret_cast<const op2_t*>(&raw.x)), c = op2_to_f2(*reinterpret_cast<const op2_t*>(&raw.y));
      *reinterpret_cast<float4*>(&sQ[warp][lane * 4]) = make_float4(a.x, a.y, c.x, c.y);
    }
    __syncwarp();
    float s0 = 0.f, s1 = 0.f;
    const op_t* k0 = sK + j0 * kCaKStride;
    const op_t* k1 = sK + (j1 < kCaLk ? j1 : 0) * kCaKStride;
#pragma unroll 8
    for (int d = 0; d < kCaDh; d += 2) {
      const float2 qa = *reinterpret_cast<const float2*>(&sQ[warp][d]);
      const float2 ka = op2_to_f2(*reinterpret_cast<const op2_t*>(k0 + d));
      const float2 kb = op2_to_f2(*reinterpret_cast<const op2_t*>(k1 + d));
      s0 = fmaf(qa.x, ka.x, fmaf(qa.y, ka.y, s0));
      s1 = fmaf(qa.x, kb.x, fmaf(qa.y, kb.y, s1));
    }
    s0 = ok0 ? s0 * scale : -INFINITY;
    s1 = ok1 ? s1 * scale : -INFINITY;
    const float mx = warp_max(fmaxf(s0, s1));
    float p0 = expf(s0 - mx), p1 = expf(s1 - mx);
    const float inv = 1.0f / warp_sum(p0 + p1);
    p0 *= inv;
    p1 *= inv;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 5
    for (int j = 0; j < kCaLk; ++j) {
      const float pj = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
      const uint2 raw = *reinterpret_cast<const uint2*>(sV + j * kCaDh + lane * 4);
      const float2 a = op2_to_f2(*reinterpret_cast<const op2_t*>(&raw.x)), c = op2_to_f2(*reinterpret_cast<const op2_t*>(&raw.y));
      acc[0] = fmaf(pj, a.x, acc[0]);
      acc[1] = fmaf(pj, a.y, acc[1]);
      acc[2] = fmaf(pj, c.x, acc[2]);
      acc[3] = fmaf(pj, c.y, acc[3]);
    }
    *reinterpret_cast<uint2*>(o) = make_uint2(pack_op2(acc[0], acc[1]), pack_op2(acc[2], acc[3]));
    __syncwarp();
  }
}

}  // namespace

// q [B*96, 1024], kv [B*50, 2048] (k | v), q_mask [B,96], kv_mask [B,50] -> out [B*96, 1024]
int ca_attention(const op_t* q, const op_t* kv, const float* q_mask, const float* kv_mask, int64_t B, op_t* out,
                 cudaStream_t st) {
  if (B == 0) return MADE_OK;
  ca_attention_kernel<<<dim3(static_cast<unsigned>(B), kCaHeads), 128, 0, st>>>(q, kv, q_mask, kv_mask, out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // namespace made
