// Launchers of the kernels around the int8 graph (post.cu): the reference's literal
// post-processing (yolact.rs:90-131), the 640x480 <-> 448x224 Triangle resampling of classify()
// (yolact.rs:192-234), and the YOLACT detection head (decode, Fast-NMS, mask assembly) that the
// north-star adds.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tod {

// ---------------------------------------------------------------- literal post-processing
struct SegPost {
  int gh, gw, ch;   // seg output grid and channels (28, 28, 81)
  float scale;      // output dequantisation (yolact.rs:177)
  int zp;
  int id_mode;      // 0 literal, 1 intent
  int up;           // nearest upsample factor (8)
};
// seg: u8 [tiles][gh][gw][ch] (tile stride in bytes); out: u32 [tiles][gh*up][gw*up];
// diverges: int[tiles], set to 1 where the reference's flood fill would not terminate.
void launch_seg_postprocess(const uint8_t* seg, int64_t tile_stride, int tiles, const SegPost& p, uint32_t* out,
                            uint32_t* cells_out, int* diverges, cudaStream_t s);

// ---------------------------------------------------------------- Triangle resampling
constexpr int kMaxTaps = 8;
struct ResampleAxis {   // device tables for one axis of image 0.24.1's sampler
  const int* left;      // [out]
  const int* count;     // [out]
  const float* weight;  // [out][kMaxTaps]
  int in_size, out_size;
};
// frames u32 [n][H][W] (r<<24|g<<16|b<<8) -> tiles u8 [2n][th][tw][3] through a (2*tw) x th canvas
void launch_classify_pre(const uint32_t* frames, int n, int W, int H, const ResampleAxis& vert, const ResampleAxis& horz,
                         float* tmp, uint8_t* tiles, int tw, int th, cudaStream_t s);
// tile results u32 [2n][th][tw] -> frames u32 [n][H][W]; target (may be null) = low 16 bits (scene.rs:93)
void launch_classify_post(const uint32_t* tile_px, int n, int W, int H, const ResampleAxis& vert, const ResampleAxis& horz,
                          float* tmp, uint32_t* frames, uint16_t* target, int tw, int th, cudaStream_t s);

// ---------------------------------------------------------------- detection
struct DetectCfg {
  int P, C, K;        // priors, classes incl. background, mask coefficients
  int ph, pw;         // prototype resolution
  float conf_thresh, nms_thresh;
  int top_k, max_dets;
  int box_zp, coef_zp, proto_zp;
  float mask_scale;   // proto_scale * coef_scale
};
struct DetectBuffers {        // all device memory, sized for max tiles
  const float* priors;        // [P][4]
  const float* exp_diff;      // [256] exp(cls_scale * (d - 255))
  const float* box_deq;       // [256]
  const float* box_exp;       // [256]
  float* boxes;               // [tiles][P][4] decoded x1,y1,x2,y2
  unsigned long long* cand;   // [tiles][C-1][P] candidate keys
  int* cand_count;            // [tiles][C-1]
  unsigned long long* surv;   // [tiles][(C-1)*top_k]
  int* surv_count;            // [tiles]
  // results
  int* det_count;             // [tiles]
  float* det_box;             // [tiles][max_dets][4]
  float* det_score;           // [tiles][max_dets]
  int* det_class;             // [tiles][max_dets]
  int* det_prior;             // [tiles][max_dets]
  float* masks;               // [tiles][max_dets][ph*pw]
  uint8_t* masks_bin;         // [tiles][max_dets][ph*pw]
  uint32_t* masks_bits;       // [tiles][max_dets][ceil(ph*pw / 32)]  bit-packed masks_bin
};
// cls/box/coef/proto: the model's u8 outputs, with their byte strides between tiles
int launch_detect(const DetectCfg& c, const DetectBuffers& b, const uint8_t* cls, int64_t cls_ts, const uint8_t* box,
                  int64_t box_ts, const uint8_t* coef, int64_t coef_ts, const uint8_t* proto, int64_t proto_ts,
                  int tiles, bool want_masks, cudaStream_t s);
// the two halves of launch_detect: boxes / scores / Fast-NMS / top-k need only the class and box heads, so the graph
// executor runs them beside the protonet branch; mask assembly needs the prototypes and the selected detections
int launch_detect_boxes(const DetectCfg& c, const DetectBuffers& b, const uint8_t* cls, int64_t cls_ts, const uint8_t* box,
                        int64_t box_ts, int tiles, cudaStream_t s, cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join);
int launch_detect_masks(const DetectCfg& c, const DetectBuffers& b, const uint8_t* coef, int64_t coef_ts, const uint8_t* proto,
                        int64_t proto_ts, int tiles, cudaStream_t s);
// output dequantisation (yolact.rs:179-186) of `tiles` u8 tensors of `elems` elements each into float [tiles][elems]
int launch_dequant_u8(const uint8_t* in, int64_t tile_stride, int c, int c_store, int64_t elems, int tiles, float scale, int zp,
                      float* out, cudaStream_t s);
// bilinear resize (PyTorch align_corners = false) of the cropped float masks to th x tw, > 0.5, bit-packed [tiles][max_dets][th*tw/32]
int launch_mask_upsample(const DetectCfg& c, const DetectBuffers& b, int tiles, int th, int tw, uint32_t* out_bits, cudaStream_t s);
size_t detect_select_smem(const DetectCfg& c);
int detect_setup_kernels(const DetectCfg& c);  // opt-in dynamic shared memory; once per handle

}  // namespace tod
