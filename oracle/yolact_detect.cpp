// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// YOLACT detection post-processing: anchor box decode, Fast-NMS, prototype-mask assembly.
// NOT IN THE REFERENCE: /root/reference/src/yolact.rs:1-5,93-95 says the author skipped it.
// BASELINE.json's north_star nevertheless puts it on the hot path, so the spec restated here is the
// published YOLACT algorithm (Bolya et al., ICCV 2019, §3-4; upstream dbolya/yolact
// layers/box_utils.py::decode / jaccard, layers/functions/detection.py::fast_nms,
// layers/output_utils.py::postprocess + crop).  PARITY UNPINNED: no reference output exists.
//
// Choices made where upstream is silent or non-deterministic (all documented in DESIGN.md §5):
//  * all inputs are the model's quantised u8 head outputs; exp() of a quantised logit is a function
//    of the u8 code only, so softmax / box-size exponentials are evaluated through 256-entry tables
//    (this is what lets the CUDA path be bit-exact);
//  * softmax denominators are summed class 0..80 sequentially in fp32;
//  * sorts are stable: descending score, ties by ascending (class, prior index);
//  * IoU with a zero-area union is 0;
//  * the mask logit is computed exactly: (sp*sc) * sum_k (P_k - zp_p) * (C_k - zp_c) in int32;
//  * masks are produced at prototype resolution (56x56): sigmoid, crop to the box padded by 1 px,
//    threshold at 0.5 for the binary form.
//
// Build with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "tod_oracle.h"

extern "C" {

void tod_oracle_detect_default_cfg(tod_oracle_detect_cfg* c) {
  c->num_priors = 3147;
  c->num_classes = 81;
  c->mask_dim = 32;
  c->proto_h = 56;
  c->proto_w = 56;
  c->conf_thresh = 0.05f;
  c->nms_thresh = 0.5f;
  c->top_k = 200;
  c->max_dets = 100;
}

// Priors in upstream make_priors order: per level, row-major cells, 3 aspect ratios per cell.
int tod_oracle_make_priors(float* out, int max_priors) {
  static const int kSizes[5] = {28, 14, 7, 4, 2};
  static const float kScales[5] = {12.0f, 24.0f, 48.0f, 96.0f, 192.0f};
  static const float kAspect[3] = {1.0f, 0.5f, 2.0f};
  const float max_size = 224.0f;
  int n = 0;
  for (int l = 0; l < 5; ++l) {
    const int s = kSizes[l];
    for (int j = 0; j < s; ++j)
      for (int i = 0; i < s; ++i) {
        const float x = (static_cast<float>(i) + 0.5f) / static_cast<float>(s);
        const float y = (static_cast<float>(j) + 0.5f) / static_cast<float>(s);
        for (int a = 0; a < 3; ++a) {
          if (n >= max_priors) return -1;
          const float ar = std::sqrt(kAspect[a]);
          out[4 * n + 0] = x;
          out[4 * n + 1] = y;
          out[4 * n + 2] = kScales[l] * ar / max_size;
          out[4 * n + 3] = kScales[l] / ar / max_size;
          ++n;
        }
      }
  }
  return n;
}

namespace {
struct Cand {
  float score;
  int cls;    // 0-based foreground class
  int prior;
  int rank;   // position inside its class' sorted list
};

inline float iou(const float* a, const float* b) {
  const float ix = std::min(a[2], b[2]) - std::max(a[0], b[0]);
  const float iy = std::min(a[3], b[3]) - std::max(a[1], b[1]);
  const float inter = (ix > 0.0f ? ix : 0.0f) * (iy > 0.0f ? iy : 0.0f);
  const float area_a = (a[2] - a[0]) * (a[3] - a[1]);
  const float area_b = (b[2] - b[0]) * (b[3] - b[1]);
  const float uni = area_a + area_b - inter;
  return uni > 0.0f ? inter / uni : 0.0f;
}
}  // namespace

int tod_oracle_detect(const tod_oracle_detect_cfg* cfg, const float* priors, const uint8_t* cls_q, float cls_scale,
                      int cls_zp, const uint8_t* box_q, float box_scale, int box_zp, const uint8_t* coef_q,
                      float coef_scale, int coef_zp, const uint8_t* proto_q, float proto_scale, int proto_zp,
                      float* det_box, float* det_score, int32_t* det_class, int32_t* det_prior, float* masks,
                      uint8_t* masks_bin) {
  (void)cls_zp;
  const int P = cfg->num_priors, C = cfg->num_classes, K = cfg->mask_dim;
  // tables over the u8 code (see header comment)
  float exp_diff[256];  // exp(cls_scale * (d - 255)), d = q - qmax + 255
  for (int d = 0; d < 256; ++d) exp_diff[d] = std::exp(cls_scale * static_cast<float>(d - 255));
  float box_deq[256], box_exp[256];
  for (int q = 0; q < 256; ++q) {
    box_deq[q] = box_scale * static_cast<float>(q - box_zp);
    box_exp[q] = std::exp(box_deq[q] * 0.2f);
  }
  // decode (box_utils.decode, variances 0.1 / 0.2) -> (x1,y1,x2,y2)
  std::vector<float> boxes(static_cast<size_t>(P) * 4);
  for (int p = 0; p < P; ++p) {
    const float* pr = priors + 4 * p;
    const uint8_t* l = box_q + 4 * p;
    const float cx = pr[0] + box_deq[l[0]] * 0.1f * pr[2];
    const float cy = pr[1] + box_deq[l[1]] * 0.1f * pr[3];
    const float w = pr[2] * box_exp[l[2]];
    const float h = pr[3] * box_exp[l[3]];
    const float x1 = cx - w / 2.0f, y1 = cy - h / 2.0f;
    boxes[4 * p + 0] = x1;
    boxes[4 * p + 1] = y1;
    boxes[4 * p + 2] = w + x1;
    boxes[4 * p + 3] = h + y1;
  }
  // softmax scores, candidate filter (detection.py: max over fg classes > conf_thresh)
  std::vector<float> score(static_cast<size_t>(P) * C);
  std::vector<int> kept;
  for (int p = 0; p < P; ++p) {
    const uint8_t* q = cls_q + static_cast<size_t>(p) * C;
    int qmax = 0;
    for (int c = 0; c < C; ++c) qmax = std::max(qmax, static_cast<int>(q[c]));
    float sum = 0.0f;
    for (int c = 0; c < C; ++c) sum += exp_diff[q[c] - qmax + 255];
    float best = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float s = exp_diff[q[c] - qmax + 255] / sum;
      score[static_cast<size_t>(p) * C + c] = s;
      if (c >= 1 && s > best) best = s;
    }
    if (best > cfg->conf_thresh) kept.push_back(p);
  }
  // Fast-NMS per class
  std::vector<Cand> all;
  std::vector<int> order;
  for (int c = 1; c < C; ++c) {
    order = kept;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return score[static_cast<size_t>(a) * C + c] > score[static_cast<size_t>(b) * C + c];
    });
    const int n = std::min<int>(cfg->top_k, static_cast<int>(order.size()));
    for (int j = 0; j < n; ++j) {
      float mx = 0.0f;  // column max of the strictly-upper-triangular IoU matrix
      for (int i = 0; i < j; ++i) mx = std::max(mx, iou(&boxes[4 * order[i]], &boxes[4 * order[j]]));
      const float s = score[static_cast<size_t>(order[j]) * C + c];
      if (mx <= cfg->nms_thresh && s > cfg->conf_thresh) all.push_back({s, c - 1, order[j], j});
    }
  }
  std::stable_sort(all.begin(), all.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });
  const int nd = std::min<int>(cfg->max_dets, static_cast<int>(all.size()));
  const int ph = cfg->proto_h, pw = cfg->proto_w;
  const float ls = proto_scale * coef_scale;
  for (int d = 0; d < nd; ++d) {
    const Cand& cd = all[d];
    std::memcpy(det_box + 4 * d, &boxes[4 * cd.prior], 4 * sizeof(float));
    det_score[d] = cd.score;
    det_class[d] = cd.cls;
    det_prior[d] = cd.prior;
    if (!masks && !masks_bin) continue;
    // output_utils.crop / sanitize_coordinates(padding = 1, cast = False)
    const float* b = det_box + 4 * d;
    float x1 = b[0] * pw, x2 = b[2] * pw, y1 = b[1] * ph, y2 = b[3] * ph;
    float xa = std::min(x1, x2) - 1.0f, xb = std::max(x1, x2) + 1.0f;
    float ya = std::min(y1, y2) - 1.0f, yb = std::max(y1, y2) + 1.0f;
    xa = std::max(xa, 0.0f); xb = std::min(xb, static_cast<float>(pw));
    ya = std::max(ya, 0.0f); yb = std::min(yb, static_cast<float>(ph));
    const uint8_t* cq = coef_q + static_cast<size_t>(cd.prior) * K;
    for (int y = 0; y < ph; ++y)
      for (int x = 0; x < pw; ++x) {
        const uint8_t* pq = proto_q + (static_cast<size_t>(y) * pw + x) * K;
        int32_t dot = 0;
        for (int k = 0; k < K; ++k) dot += (static_cast<int32_t>(pq[k]) - proto_zp) * (static_cast<int32_t>(cq[k]) - coef_zp);
        const float logit = static_cast<float>(dot) * ls;
        float mval = 1.0f / (1.0f + std::exp(-logit));
        const float fx = static_cast<float>(x), fy = static_cast<float>(y);
        if (!(fx >= xa && fx < xb && fy >= ya && fy < yb)) mval = 0.0f;
        const size_t o = (static_cast<size_t>(d) * ph + y) * pw + x;
        if (masks) masks[o] = mval;
        if (masks_bin) masks_bin[o] = mval > 0.5f ? 1 : 0;
      }
  }
  return nd;
}

}  // extern "C"
