/* CPU ORACLE (test infrastructure only, never shipped): plain-C restatement of the span arithmetic
 * of music_detr/span_utils.py and the matcher cost of music_detr/matcher.py, one IEEE fp32
 * operation per reference tensor op, in the reference's order.  Build with -ffp-contract=off.
 *   span_cw_to_se            span_utils.py:15-24
 *   temporal_iou             span_utils.py:39-66
 *   generalized_temporal_iou span_utils.py:86-115
 *   matcher cost             matcher.py:66-88 (cdist p=1, -gIoU, -p_fg; weights left to right)
 *   detr_iou                 span_utils.py:147-170 + individual_IoU_tensor :119-145
 */
#include <math.h>
#include <stdint.h>

static float fmaxf_(float a, float b) { return a > b ? a : b; }
static float fminf_(float a, float b) { return a < b ? a : b; }

void oracle_cw_to_se(const float* cw, float* se, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    float h = 0.5f * cw[2 * i + 1];
    se[2 * i] = cw[2 * i] - h;
    se[2 * i + 1] = cw[2 * i] + h;
  }
}

static float giou_pair(float s1, float e1, float s2, float e2, float* iou_out, float* uni_out) {
  float a1 = e1 - s1, a2 = e2 - s2;
  float left = fmaxf_(s1, s2), right = fminf_(e1, e2);
  float inter = fmaxf_(right - left, 0.0f);
  float uni = (a1 + a2) - inter;
  float iou = inter / uni;
  float enc = fmaxf_(fmaxf_(e1, e2) - fminf_(s1, s2), 0.0f);
  if (iou_out) *iou_out = iou;
  if (uni_out) *uni_out = uni;
  return iou - (enc - uni) / enc;
}

void oracle_giou(const float* a, int64_t n, const float* b, int64_t m, float* out) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j)
      out[i * m + j] = giou_pair(a[2 * i], a[2 * i + 1], b[2 * j], b[2 * j + 1], 0, 0);
}

void oracle_temporal_iou(const float* a, int64_t n, const float* b, int64_t m, float* iou, float* uni) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < m; ++j)
      giou_pair(a[2 * i], a[2 * i + 1], b[2 * j], b[2 * j + 1], &iou[i * m + j], &uni[i * m + j]);
}

void oracle_matcher_cost(const float* prob_fg, const float* out_cw, int64_t n, const float* tgt_cw, int64_t m,
                         float w_span, float w_giou, float w_class, float* cost) {
  for (int64_t i = 0; i < n; ++i) {
    float c1 = out_cw[2 * i], w1 = out_cw[2 * i + 1];
    float h1 = 0.5f * w1, s1 = c1 - h1, e1 = c1 + h1;
    for (int64_t j = 0; j < m; ++j) {
      float c2 = tgt_cw[2 * j], w2 = tgt_cw[2 * j + 1];
      float h2 = 0.5f * w2, s2 = c2 - h2, e2 = c2 + h2;
      float l1 = fabsf(c1 - c2) + fabsf(w1 - w2);
      float g = giou_pair(s1, e1, s2, e2, 0, 0);
      float t = w_span * l1 + w_giou * (-g);
      cost[i * m + j] = t + w_class * (-prob_fg[i]);
    }
  }
}

void oracle_detr_iou(const float* pred_st, const float* pred_ed, const float* gt_moment, const float* m_duration,
                     float max_m_duration, int64_t n, float* iou) {
  for (int64_t i = 0; i < n; ++i) {
    float gs = gt_moment[2 * i], ge = gt_moment[2 * i + 1];
    float ps = fmaxf_(pred_st[i], 0.0f);
    float pe = fminf_(fminf_(pred_ed[i], max_m_duration), m_duration[i]);
    float inter = fmaxf_(fminf_(ge, pe) - fmaxf_(gs, ps), 0.0f);
    float uni = ((pe - ps) + (ge - gs)) - inter;
    float v = inter / uni;
    if (gs >= ge || uni <= 0.0f) v = 0.0f;
    iou[i] = v;
  }
}
