// Evaluation-time loss scalars that Uni_model.forward returns (model_Uni.py:254-262, 287-289):
//   * SetCriterion (music_detr/loss_detr.py:74-128,130-169) for the final and the 5 auxiliary
//     decoder layers, specialised to ONE moment query per sample, where the Hungarian matching of
//     matcher.py:36-92 is the identity on samples whose target has w != 0;
//   * InfoNCELoss(dual) + CLIPLoss(single) (modules/loss.py:5-24, 66-123, audio_id=None).
// Forward values only (no backward in this build).
#include "common.cuh"

namespace made {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  __syncthreads();
  return t;
}

// grid = n_layers, block = 256.  out[l] = {loss_span, loss_giou, loss_label, class_error, loss_contrastive_align}
__global__ void __launch_bounds__(256)
detr_losses_kernel(const float2* __restrict__ logits, const float2* __restrict__ spans,
                   const float* __restrict__ proj_q, const float* __restrict__ proj_v,
                   const float2* __restrict__ targets, int64_t B, float w_fg, float w_bg,
                   float inv_temperature, float* __restrict__ out) {
  __shared__ float red[8];
  const int l = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float l1 = 0.f, gi = 0.f, ce = 0.f, nm = 0.f, correct = 0.f, nce = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += 256) {
    const float2 lg = logits[l * B + b];
    const float2 sp = spans[l * B + b];
    const float2 tg = targets[b];
    const bool matched = tg.y != 0.f;                       // matcher.py:59
    if (matched) {
      l1 += fabsf(sp.x - tg.x) + fabsf(sp.y - tg.y);        // loss_detr.py:88-89
      const float s1 = sp.x - 0.5f * sp.y, e1 = sp.x + 0.5f * sp.y;
      const float s2 = tg.x - 0.5f * tg.y, e2 = tg.x + 0.5f * tg.y;
      const float inter = fmaxf(fminf(e1, e2) - fmaxf(s1, s2), 0.f);
      const float uni = (e1 - s1) + (e2 - s2) - inter;
      const float enc = fmaxf(fmaxf(e1, e2) - fminf(s1, s2), 0.f);
      gi += 1.f - (inter / uni - (enc - uni) / enc);         // :90
      nm += 1.f;
      correct += lg.x >= lg.y ? 1.f : 0.f;                   // misc.py accuracy, foreground = 0
    }
    const float mx = fmaxf(lg.x, lg.y);
    const float lse = mx + logf(expf(lg.x - mx) + expf(lg.y - mx));
    ce += matched ? w_fg * (lse - lg.x) : w_bg * (lse - lg.y);   // :106 weighted CE, reduction none
  }
  if (proj_q && proj_v) {
    // loss_contrastive_align :112-128 with one query: -pos/num_pos + logsumexp over the single query
    for (int64_t b = warp; b < B; b += 8) {
      float acc = 0.f;
      const float* q = proj_q + (static_cast<int64_t>(l) * B + b) * 256;
      for (int f = 0; f < 50; ++f) {
        const float* v = proj_v + (b * 50 + f) * 256;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(q[lane + 32 * j], v[lane + 32 * j], acc);
      }
      acc = warp_sum(acc) * inv_temperature;
      if (lane == 0) {
        const bool matched = targets[b].y != 0.f;
        const float pos = matched ? acc : 0.f, npos = matched ? 1.f : 0.f;
        nce += -pos / npos + acc;
      }
    }
  }
  l1 = block_sum_256(l1, red);
  gi = block_sum_256(gi, red);
  ce = block_sum_256(ce, red);
  nm = block_sum_256(nm, red);
  correct = block_sum_256(correct, red);
  nce = block_sum_256(nce, red);
  if (threadIdx.x == 0) {
    out[l * 5 + 0] = l1 / (2.f * nm);
    out[l * 5 + 1] = gi / nm;
    out[l * 5 + 2] = ce / static_cast<float>(B);
    out[l * 5 + 3] = 100.f - correct * (100.f / nm);
    out[l * 5 + 4] = nce / static_cast<float>(B);
  }
}

// Symmetric cross-entropy with diagonal labels on logits = sims * exp(logit_scale):
// out[0] += (mean_i(lse_row_i - d_i) + mean_j(lse_col_j - d_j)) / 2.   grid = 2 (dual, single).
__global__ void __launch_bounds__(256)
retrieval_loss_kernel(const float* __restrict__ dual, const float* __restrict__ single, int64_t ld, int n,
                      float scale, float* __restrict__ out) {
  __shared__ float red[8];
  const float* s = blockIdx.x == 0 ? dual : single;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float mr = -INFINITY, mc = -INFINITY;
    for (int j = 0; j < n; ++j) {
      mr = fmaxf(mr, s[i * ld + j] * scale);
      mc = fmaxf(mc, s[j * ld + i] * scale);
    }
    float sr = 0.f, sc = 0.f;
    for (int j = 0; j < n; ++j) {
      sr += expf(s[i * ld + j] * scale - mr);
      sc += expf(s[j * ld + i] * scale - mc);
    }
    const float d = s[i * ld + i] * scale;
    acc += (mr + logf(sr) - d) + (mc + logf(sc) - d);
  }
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, acc / (2.f * n));
}

}  // namespace made

using namespace made;

extern "C" {

int made_detr_losses(const float* pred_logits, const float* pred_spans, const float* proj_queries,
                     const float* proj_vid_mem, const float* targets_cw, int64_t B, int n_layers,
                     float w_fg, float w_bg, float temperature, float* out, void* stream) {
  MADE_REQUIRE(pred_logits && pred_spans && targets_cw && out, "detr_losses: null pointer");
  MADE_REQUIRE(B > 0 && n_layers > 0 && temperature > 0.f, "detr_losses: bad sizes");
  detr_losses_kernel<<<n_layers, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(pred_logits), reinterpret_cast<const float2*>(pred_spans), proj_queries,
      proj_vid_mem, reinterpret_cast<const float2*>(targets_cw), B, w_fg, w_bg, 1.f / temperature, out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

int made_retrieval_loss(const float* dual, const float* single, int64_t ld, int n, float logit_scale,
                        float* out, void* stream) {
  MADE_REQUIRE(dual && single && out && n > 0 && ld >= n, "retrieval_loss: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MADE_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  retrieval_loss_kernel<<<2, 256, 0, st>>>(dual, single, ld, n, expf(logit_scale), out);
  MADE_CHECK_LAUNCH();
  return MADE_OK;
}

}  // extern "C"
