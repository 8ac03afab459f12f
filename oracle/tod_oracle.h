/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C API of the CPU oracle (liboracle.so).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (tiny-object-detection_b200/) never links, imports or executes anything under oracle/.
 *
 * Every function cites the reference lines it restates in its definition.
 */
#ifndef TOD_ORACLE_H_
#define TOD_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- point cloud / weights (pt_cloud.cpp) ---------------- */
typedef struct tod_oracle_scene_params {
  int width, height;
  float max_depth_in;
  float x_fov, y_fov;
  float bot_avoidance_const;
  int bot_norm_const;
  int terrain_norm_const;
  float bump_err;
  int sample_shift; /* 0 or 1 */
  int weights_mode; /* 0 literal, 1 intent */
} tod_oracle_scene_params;

void tod_oracle_scene_default_params(tod_oracle_scene_params* p);
void tod_oracle_bump_table(float val, int s, float bump_err, uint32_t* out);
void tod_oracle_pt_cloud(const tod_oracle_scene_params* P, const uint16_t* depth, const uint16_t* target,
                         uint32_t* map, float* balls4);
void tod_oracle_pt_cloud_weights(const tod_oracle_scene_params* P, const uint32_t* map, float* world4,
                                 float* conn0, float* conn1);
void tod_oracle_scene_materialize(int npx, const uint32_t* map, const float* world4, const float* conn0,
                                  const float* conn1, const float* balls4, float* height, float* pos3,
                                  int32_t* balls2, float* connections8);

/* ---------------- path.rs (path.cpp): the Scene's consumer ---------------- */
int tod_oracle_path_literal_panics(void);
int tod_oracle_path_modify(int W, int H, const float* height, const float* pos3, const int32_t* balls2, const float* conn8,
                           float* cost, int32_t* pred, float* dirs, int cap);
int tod_oracle_path_serialize(uint64_t created_secs, const float* dirs, int n, uint8_t* out);

/* ---------------- yolact.rs literal pre/post-processing (yolact_post.cpp) ---------------- */
/* yolact.rs:169-188 */
void tod_oracle_dequant_u8(const uint8_t* q, int n, float scale, int zero_point, float* out);
/* yolact.rs:108-118: f32[784*81] -> u8[784] */
void tod_oracle_cell_classes(const float* seg, int cells, int channels, uint8_t* classes);
/* yolact.rs:52-88. mode 0 literal (returns 1 if the reference would not terminate; ids all -1),
 * mode 1 intent (4-connected CCL, raster-order ids). */
int tod_oracle_terrible_id(const uint8_t* classes, int mode, int8_t* ids);
/* yolact.rs:127-128 (literal `&` pack, or intent `|` pack) + 8x nearest replicate: -> u32[224*224] */
void tod_oracle_pack_upsample(const uint8_t* classes, const int8_t* ids, int mode, uint32_t* out);
/* image 0.24.1 imageops::resize(FilterType::Triangle) on RGB8 (third-party, unpinned). */
void tod_oracle_resize_triangle_rgb8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
/* yolact.rs:192-234 pre half: u32 frame -> two 224x224x3 u8 tiles */
void tod_oracle_classify_pre(const uint32_t* frame, int width, int height, uint8_t* tiles2);
/* yolact.rs:219-233 post half: two u32[224*224] tile results -> u32 frame (in place semantics) */
void tod_oracle_classify_post(const uint32_t* t1, const uint32_t* t2, int width, int height, uint32_t* frame);
/* scene.rs:93 */
void tod_oracle_target_from_frame(const uint32_t* frame, int n, uint16_t* target);

/* ---------------- YOLACT decode / Fast-NMS / masks (yolact_detect.cpp; not in the reference) -------- */
typedef struct tod_oracle_detect_cfg {
  int num_priors;      /* 3147 */
  int num_classes;     /* 81 incl. background */
  int mask_dim;        /* 32 */
  int proto_h, proto_w; /* 56 */
  float conf_thresh;   /* 0.05 */
  float nms_thresh;    /* 0.5 */
  int top_k;           /* 200 */
  int max_dets;        /* 100 */
} tod_oracle_detect_cfg;
void tod_oracle_detect_default_cfg(tod_oracle_detect_cfg* c);
/* priors for the 224x224 head: levels 28,14,7,4,2; 3 aspect ratios. out f32[3147*4] (cx,cy,w,h) */
int tod_oracle_make_priors(float* out, int max_priors);
/* quantised head outputs (u8) + their (scale, zp) -> detections.  Returns number of detections.
 * det_box f32[max_dets*4] (x1,y1,x2,y2 in [0,1]), det_score f32, det_class i32 (0-based fg class),
 * det_prior i32 (prior index == the bit-exact "keep index"), masks f32[max_dets*ph*pw] (sigmoid+crop),
 * masks_bin u8 (thresholded 0.5). */
int tod_oracle_detect(const tod_oracle_detect_cfg* cfg, const float* priors, const uint8_t* cls_q, float cls_scale,
                      int cls_zp, const uint8_t* box_q, float box_scale, int box_zp, const uint8_t* coef_q,
                      float coef_scale, int coef_zp, const uint8_t* proto_q, float proto_scale, int proto_zp,
                      float* det_box, float* det_score, int32_t* det_class, int32_t* det_prior, float* masks,
                      uint8_t* masks_bin);

/* ---------------- TFLite int8 graph (tflite_model.cpp, tflite_ops.cpp) ---------------- */
typedef struct tod_oracle_model tod_oracle_model;
/* returns NULL on error; tod_oracle_last_error() explains */
tod_oracle_model* tod_oracle_model_load(const char* path);
void tod_oracle_model_free(tod_oracle_model* m);
const char* tod_oracle_last_error(void);
int tod_oracle_model_num_tensors(const tod_oracle_model* m);
int tod_oracle_model_num_ops(const tod_oracle_model* m);
int tod_oracle_model_num_inputs(const tod_oracle_model* m);
int tod_oracle_model_num_outputs(const tod_oracle_model* m);
int tod_oracle_model_output_tensor(const tod_oracle_model* m, int i);
int tod_oracle_model_input_tensor(const tod_oracle_model* m, int i);
int tod_oracle_model_op_code(const tod_oracle_model* m, int op);            /* builtin code */
int tod_oracle_model_op_output(const tod_oracle_model* m, int op, int k);   /* tensor index */
/* shape4 padded with leading 1s; returns element count; type: tflite TensorType code */
int64_t tod_oracle_model_tensor_info(const tod_oracle_model* m, int t, int* shape4, int* type, float* scale,
                                     int* zero_point);
/* run one [1,H,W,3] u8 input through the whole graph (threads = OpenMP threads to use).
 * Every intermediate tensor is kept so tests can compare any of them. */
int tod_oracle_model_invoke(tod_oracle_model* m, const uint8_t* input, int threads);
/* pointer to tensor t's data after invoke (int8/uint8 bytes, or int32/float) */
const void* tod_oracle_model_tensor_data(const tod_oracle_model* m, int t);
/* total multiply-accumulates of CONV_2D + DEPTHWISE_CONV_2D for one invoke */
int64_t tod_oracle_model_macs(const tod_oracle_model* m);

/* fixed point helpers exposed for property tests */
int32_t tod_oracle_srdhm(int32_t a, int32_t b);
int32_t tod_oracle_rdivpot(int32_t x, int e);
int32_t tod_oracle_mbqm(int32_t x, int32_t q, int shift);
void tod_oracle_quantize_multiplier(double m, int32_t* q, int* shift);

#ifdef __cplusplus
}
#endif
#endif
