/* tod.h — C ABI of libtod_b200.so, the B200-native (sm_100a) implementation of the per-frame
 * perception hot path of icf3ver/tiny-object-detection.
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface it
 * replaces (file:line under /root/reference).  The reference has no FFI of its own for this path:
 * the boundary is three crate-internal Rust signatures, so these are the symbols a
 * thin `extern "C"` Rust crate binds (ffi/src/lib.rs shows that binding; INTEGRATION.md explains
 * where it is spliced into src/yolact.rs and src/scene.rs).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types;
 *   - every function returns 0 (TOD_OK) or a negative tod_status; it never aborts.  The Rust shim
 *     `.expect()`s the code to keep the reference's panic-on-error behaviour;
 *   - tod_last_error() returns a thread-local human-readable message for the last failure;
 *   - handles are NOT thread-safe (mirrors `&mut self`, yolact.rs:39): one handle per thread / GPU;
 *   - buffers passed to the non-`_device` entry points are HOST memory owned by the caller (pinned
 *     memory recommended); buffers passed to `_device` entry points are device pointers on the
 *     handle's GPU and the call is asynchronous on the given CUDA stream (NULL = handle's stream);
 *   - there is NO CPU fallback: without a CUDA device every create call fails with
 *     TOD_ERR_NO_DEVICE.
 */
#ifndef TOD_H_
#define TOD_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOD_ABI_VERSION 1

typedef enum tod_status {
  TOD_OK = 0,
  TOD_ERR_INVALID_ARG = -1,
  TOD_ERR_IO = -2,          /* model file unreadable */
  TOD_ERR_MODEL = -3,       /* malformed / unsupported .tflite (e.g. edgetpu-custom-op) */
  TOD_ERR_CUDA = -4,        /* CUDA runtime / driver error */
  TOD_ERR_NO_DEVICE = -5,   /* no usable sm_100 device */
  TOD_ERR_CAPACITY = -6,    /* batch larger than the handle was created for */
  TOD_ERR_UNSUPPORTED = -7
} tod_status;

/* positive return: the call completed but the *reference* would not have (the literal
 * flood fill of yolact.rs:60-77 never terminates when two class-3 cells touch). */
#define TOD_WARN_REFERENCE_DIVERGES 1

const char* tod_last_error(void);
int tod_abi_version(void);
int tod_device_count(int* count);

/* ======================================================================================
 * Scene: depth -> point cloud / height map / connection weights.
 * Replaces the GPU half of `append_scene` (src/scene.rs:147-331): the Vulkan dispatch of
 * shaders/pt_cloud.comp (scene.rs:238-247) and shaders/pt_cloud_weights.comp (scene.rs:249-260),
 * the uploads (scene.rs:197-198) and the four read-backs (scene.rs:246,257-259).
 * ====================================================================================== */
typedef struct tod_scene tod_scene;

typedef struct tod_scene_params {
  int32_t width;                 /* pt_cloud.comp:23  (640; 320 for Carmine, README.md:18) */
  int32_t height;                /* pt_cloud.comp:24  (480; 240) */
  float max_depth_in;            /* pt_cloud.comp:25 */
  float x_fov;                   /* pt_cloud.comp:28 */
  float y_fov;                   /* pt_cloud.comp:27 */
  float bot_avoidance_const;     /* pt_cloud.comp:32 */
  int32_t bot_norm_const;        /* pt_cloud.comp:36 */
  int32_t terrain_norm_const;    /* pt_cloud.comp:37 */
  float bump_err;                /* pt_cloud.comp:39 */
  int32_t sample_shift;          /* 0 = texelFetch(x,y) (default), 1 = what the Pi's V3D did */
  int32_t weights_mode;          /* 0 = literal (pack()==0), 1 = intent (true neighbour distances) */
  int32_t max_batch;             /* frames per call the handle is sized for */
} tod_scene_params;

/* fills the reference's constants (pt_cloud.comp:23-39) */
void tod_scene_default_params(tod_scene_params* p);

/* replaces the per-call Vulkan object creation of scene.rs:152-224 (hoisted: done once) */
int tod_scene_create(int device, const tod_scene_params* params, tod_scene** out);
void tod_scene_destroy(tod_scene* s);

/* One call = n independent frames through both shaders (scene.rs:274-282).
 *   depth   u16[n][H][W]   (R16_UINT upload, scene.rs:197)
 *   target  u16[n][H][W]   low byte = class, high byte = id (R8G8_UINT upload, scene.rs:198)
 *   map     u32[n][H][W]   height map        (map_dest,          scene.rs:231,246)
 *   world4  f32[n][H][W][4]                  (world_dest,        scene.rs:232,257)
 *   conn0   f32[n][H][W][4]                  (connections0_dest, scene.rs:233,258)
 *   conn1   f32[n][H][W][4]                  (connections1_dest, scene.rs:234,259)
 *   balls4  f32[n][100][4]                   (ball_buffer,       scene.rs:211)
 * Any output pointer may be NULL to skip that read-back. Blocks until the results are in place
 * (scene.rs:282 `future.wait`). */
int tod_scene_append_batch(tod_scene* s, const uint16_t* depth, const uint16_t* target, int n, uint32_t* map,
                           float* world4, float* conn0, float* conn1, float* balls4);

/* Same, device pointers, asynchronous on `stream` (a cudaStream_t; NULL = the handle's stream).  Output pointers that
 * are NULL are skipped, except that with ALL five NULL the results are kept in the handle's own images (that is what
 * tod_scene_materialize reads).  One handle serves one stream at a time (shared scratch buffers). */
int tod_scene_append_batch_device(tod_scene* s, const uint16_t* d_depth, const uint16_t* d_target, int n,
                                  uint32_t* d_map, float* d_world4, float* d_conn0, float* d_conn1,
                                  float* d_balls4, void* stream);

/* `Scene` materialisation (scene.rs:312-327), fused on the device: from the results the last append call left in the
 * handle's own images (tod_scene_append_batch, or tod_scene_append_batch_device with all outputs NULL; after a device
 * call that wrote to caller buffers this returns TOD_ERR_INVALID_ARG) for frame `frame`, fills host buffers
 *   height f32[H*W], pos f32[H*W][3], balls i32[100][2], connections f32[H*W][8]. */
int tod_scene_materialize(tod_scene* s, int frame, float* height, float* pos3, int32_t* balls2,
                          float* connections8);

/* average device time (ms) of the dominant kernel (pt_stamp) over the last append call */
int tod_scene_last_kernel_ms(tod_scene* s, float* stamp_ms, float* weights_ms);

/* ======================================================================================
 * Yolact: int8 YOLACT inference + post-processing.
 * Replaces `Yolact::init` (src/yolact.rs:17-37) and `Yolact::classify` (yolact.rs:39-41,192-234),
 * i.e. interpreter.invoke() (yolact.rs:163) and everything around it.
 * ====================================================================================== */
typedef struct tod_yolact tod_yolact;

typedef struct tod_yolact_options {
  int32_t max_tiles;        /* 224x224 tiles per call the handle is sized for (2 per camera frame) */
  int32_t id_mode;          /* 0 = literal terrible_id + `&` pack, 1 = intent (CCL + `|`) */
  float conf_thresh;        /* 0.05  YOLACT detection score threshold */
  float nms_thresh;         /* 0.5   Fast-NMS IoU threshold */
  int32_t top_k;            /* 200   per-class candidates */
  int32_t max_dets;         /* 100 */
  int32_t use_cuda_graph;   /* 1 = capture the per-batch pipeline in a CUDA graph */
  int32_t conv_impl;        /* 0 = tcgen05 int8 implicit GEMM wherever the layer shape allows (default),
                               1 = CUDA-core direct convolution everywhere (on-device cross-check),
                               2 = tcgen05 with the general (literal-arithmetic) epilogue and the first-generation
                                   depthwise kernel (cross-check of the fast requantisation forms) */
  int32_t fusion;           /* 0 = every TFLite tensor is materialised (each one can be fetched and compared),
                               1 = PAD folded into the following convolution, QUANTIZE / RELU / TANH byte maps and
                                   residual / FPN ADDs fused into the producing convolution's epilogue (default) */
  int32_t use_pdl;          /* 1 = programmatic dependent launch between consecutive kernels of a graph branch (default) */
  int32_t batches_in_flight;/* how many handles work on this GPU at the same time (frame-loop double buffering, tod_pool_*).
                               <= 1 (default): this handle has the GPU to itself - every convolution launch spreads over as many
                               SMs as it has tiles (lowest latency).  >= 2: launches that fit one round give each CTA two tiles
                               and half the SMs go to the other batches' kernels (a convolution CTA holds its SM exclusively, so
                               its fixed cost is paid per CTA): +3 % throughput with three batches in flight, +1.5 % latency.
                               tod_pool_create sets it to the pool's depth. */
} tod_yolact_options;

void tod_yolact_default_options(tod_yolact_options* o);

/* <- Yolact::init (yolact.rs:17-37).  `tflite_path` replaces the hard-coded path of yolact.rs:19; a
 * model holding the edgetpu custom op is rejected with TOD_ERR_MODEL (use the CPU model,
 * data/FRC_model.tflite). */
int tod_yolact_create(const char* tflite_path, int device, const tod_yolact_options* opts, tod_yolact** out);
void tod_yolact_destroy(tod_yolact* y);

/* <- Yolact::classify (yolact.rs:39): in place on a caller-owned 640x480 frame, px in =
 * r<<24|g<<16|b<<8 (scene.rs:86), px out = [resized class, 0, 0, 0] big-endian (yolact.rs:231-233). */
int tod_yolact_classify(tod_yolact* y, uint32_t* frame, int width, int height);
/* batched form: n frames, contiguous */
int tod_yolact_classify_batch(tod_yolact* y, uint32_t* frames, int n, int width, int height);
/* device-resident batched form: d_frames in/out on the GPU, optional d_target u16[n][H][W] receives the
 * `(px & 0xFFFF) as u16` extraction of scene.rs:93 so the scene path can consume it without a host trip */
int tod_yolact_classify_batch_device(tod_yolact* y, uint32_t* d_frames, int n, int width, int height,
                                     uint16_t* d_target, void* stream);

/* Parses a .tflite without touching a GPU (what FlatBufferModel::build_from_file + InterpreterBuilder check,
 * yolact.rs:18-29): operator / tensor counts and the multiply-accumulates of one invoke. */
int tod_model_inspect(const char* tflite_path, int32_t* num_ops, int32_t* num_tensors, int64_t* macs);

/* Anchor priors f32[n][4] (cx,cy,w,h in [0,1]) for the detection head.  The handle computes the YOLACT priors
 * for the FRC head (3147 = 3 x (28^2+14^2+7^2+4^2+2^2)) itself; any other head must be given its priors. */
int tod_yolact_set_priors(tod_yolact* y, const float* priors, int n);

/* 1 if some tile of the last classify / infer call would have hung the reference's flood fill (id_mode 0);
 * synchronises the handle's stream. */
int tod_yolact_last_diverged(tod_yolact* y, int* diverged);

/* Model introspection (what interpreter.tensor_info gives the reference, yolact.rs:150,169-176) */
int tod_yolact_num_outputs(const tod_yolact* y);
/* shape of the model input [1, tile_h, tile_w, 3] (the dims classify_tile checks, yolact.rs:149-158) */
int tod_yolact_input_info(const tod_yolact* y, int32_t shape4[4]);
int tod_yolact_output_info(const tod_yolact* y, int index, int32_t shape4[4], float* scale, int32_t* zero_point,
                           int32_t* elems);
int tod_yolact_num_tensors(const tod_yolact* y);
int tod_yolact_num_ops(const tod_yolact* y);
int tod_yolact_tensor_info(const tod_yolact* y, int tensor, int32_t shape4[4], int32_t* type, float* scale,
                           int32_t* zero_point, int32_t* elems);

typedef struct tod_detections {
  int32_t max_dets;       /* capacity per tile of the arrays below */
  int32_t* count;         /* [n]                       detections per tile */
  float* boxes;           /* [n][max_dets][4]          x1,y1,x2,y2 in [0,1] */
  float* scores;          /* [n][max_dets] */
  int32_t* classes;       /* [n][max_dets]             0-based foreground class */
  int32_t* priors;        /* [n][max_dets]             prior index == NMS keep index */
  float* masks;           /* [n][max_dets][56][56]     sigmoid + crop, may be NULL */
  uint8_t* masks_bin;     /* [n][max_dets][56][56]     > 0.5, may be NULL */
  uint32_t* masks_bits;   /* [n][max_dets][ceil(56*56/32)]  the same binary masks, 1 bit per prototype pixel (bit i of word
                             w = pixel 32*w + i): an eighth of the read-back of masks_bin; may be NULL */
  uint32_t* masks_tile_bits; /* [n][max_dets][ceil(224*224/32)]  YOLACT postprocess: the cropped float masks resized to the tile
                             (bilinear, align_corners = false) and thresholded at 0.5, 1 bit per tile pixel; may be NULL.
                             Implies the float-mask computation; rows of detections >= count[tile] are zero. */
} tod_detections;

/* Runs the int8 graph on n RGB tiles u8[n][224][224][3] (== interpreter.invoke(), yolact.rs:161-163).
 *   outputs_u8[k]  (k < num_outputs, may be NULL) receives output tensor k, u8[n][elems_k]
 *   tile_classes   (may be NULL) u32[n][224][224]: the literal postprocess() result (yolact.rs:90-131)
 *   dets           (may be NULL) YOLACT decode + Fast-NMS + mask assembly (north-star; not in the reference)
 * Returns TOD_OK, or TOD_WARN_REFERENCE_DIVERGES if id_mode == 0 and some tile would hang the reference. */
int tod_yolact_infer_tiles(tod_yolact* y, const uint8_t* rgb_tiles, int n, uint8_t* const* outputs_u8,
                           uint32_t* tile_classes, tod_detections* dets);
/* Same call with one more output: cell_classes (may be NULL) u32[n][28][28] is the packed (class, id) grid *before* the
 * 8x8 nearest replication of yolact.rs:127-128 - tile_classes[t][y][x] == cell_classes[t][y / 8][x / 8] - i.e. the same
 * information in 1/64 of the read-back (12.8 MB -> 0.2 MB per 64 tiles).  A frame loop that only needs the classes
 * passes tile_classes = NULL. */
int tod_yolact_infer_tiles_cells(tod_yolact* y, const uint8_t* rgb_tiles, int n, uint8_t* const* outputs_u8,
                                 uint32_t* tile_classes, uint32_t* cell_classes, tod_detections* dets);

/* Output k of the last call dequantised on the GPU exactly as the reference does for its `results: Vec<Vec<f32>>`
 * (yolact.rs:169-188): out[t][e] = scale * ((u8 as i32 - zero_point) as f32), f32[n][elems_k]. */
int tod_yolact_fetch_output_f32(tod_yolact* y, int index, int n, float* out);

/* Device-resident form used by the fused RGB-D pipeline and the benchmark: input tiles already on the GPU;
 * results stay on the GPU (fetch with tod_yolact_fetch_*).  Asynchronous on `stream` (NULL = the handle's own).
 * A handle serves ONE stream at a time (its activation arena and scratch buffers are shared): do not drive one handle
 * from two streams concurrently.  tod_yolact_fetch_* copy on the handle's own stream, which the library orders behind
 * the caller's stream (an event recorded at the end of this call); they block until the data is on the host. */
int tod_yolact_infer_tiles_device(tod_yolact* y, const uint8_t* d_rgb_tiles, int n, void* stream);
int tod_yolact_fetch_output(tod_yolact* y, int index, int n, uint8_t* out);
int tod_yolact_fetch_tensor(tod_yolact* y, int tensor, int n, void* out, size_t out_bytes);
int tod_yolact_fetch_tile_classes(tod_yolact* y, int n, uint32_t* out);
int tod_yolact_fetch_detections(tod_yolact* y, int n, tod_detections* dets);

/* accounting for bench.py: kernels launched per infer call, MACs per tile, and per-kernel device time */
int tod_yolact_stats(const tod_yolact* y, int64_t* macs_per_tile, int32_t* launches_per_call,
                     int32_t* tc_conv_layers);
/* times every op of the graph once with CUDA events for n resident tiles: fills up to `cap` entries of
 * ms[] (per op, graph order) and returns the op count */
int tod_yolact_profile_ops(tod_yolact* y, int n, float* ms, int32_t* kinds, int cap);

/* ======================================================================================
 * Pool: the frame sharder.  Independent frames / tiles are cut into contiguous ranges, ceil(n / G) per
 * GPU; every GPU runs `depth` Yolact handles, one host thread each, that take the range's chunks (<= max_tiles tiles)
 * in turn, so a chunk's copies and tail overlap the next chunk's backbone.  Results land in the caller's buffers at
 * the frame's index; a frame's bytes do not depend on the number of GPUs or handles.  No collective, no NCCL.
 * Replaces the reference's one-frame-at-a-time loops (scene.rs:77-119 `process_scene`, main.rs:78-96 `manage`)
 * for a box with several GPUs; with one GPU it is the pipelined (double-buffered) form of the single-handle calls.
 * ====================================================================================== */
typedef struct tod_pool tod_pool;
/* devices == NULL or n_devices <= 0: every visible sm_100 device.  opts as for tod_yolact_create (max_tiles = chunk size). */
int tod_pool_create(const char* tflite_path, const int32_t* devices, int n_devices, int depth,
                    const tod_yolact_options* opts, tod_pool** out);
void tod_pool_destroy(tod_pool* p);
int tod_pool_num_devices(const tod_pool* p);
int tod_pool_num_handles(const tod_pool* p);
/* tod_yolact_infer_tiles_cells over any n (same buffers, n tiles long) */
int tod_pool_infer_tiles(tod_pool* p, const uint8_t* rgb_tiles, int n, uint8_t* const* outputs_u8, uint32_t* tile_classes,
                         uint32_t* cell_classes, tod_detections* dets);
/* tod_yolact_classify_batch over any n: <- Yolact::classify (yolact.rs:39) for every frame, in place */
int tod_pool_classify_batch(tod_pool* p, uint32_t* frames, int n, int width, int height);
/* The fused RGB-D frame loop for n frames (scene.rs:84-97 feeding scene.rs:147-331): classify in place, target = low
 * 16 bits (scene.rs:93) kept on the GPU, then pt_cloud + pt_cloud_weights.  frames u32[n][H][W] in/out, depth
 * u16[n][H][W]; outputs as tod_scene_append_batch (any may be NULL). */
int tod_pool_rgbd_batch(tod_pool* p, const tod_scene_params* scene_params, uint32_t* frames, const uint16_t* depth, int n,
                        uint32_t* map, float* world4, float* conn0, float* conn1, float* balls4);

/* ======================================================================================
 * Path: the Scene's consumer.  <- `path::modify_path` (src/path.rs:25-120) and `Path::serialize`
 * (path.rs:17-21).  The reference function cannot complete on any input (224*224-element arrays indexed with 640x480
 * node numbers, path.rs:29-30,38,99; cost[usize::MAX - 2] on the last extraction step, path.rs:105) and is never
 * reached (main.rs:92): tod_path_reference_panics() == 1 states that, and tod_path_modify computes the evident intent
 * (rules in DESIGN.md §5.5 / oracle/path.cpp): multi-target shortest paths over Scene::neighbors' 4-neighbourhood with
 * edge weight connections[node][cn] + |height[node] - height[neighbor]| (path.rs:64), then the (magnitude, rotation)
 * list walked from START_NODE = W*H - H/2 as path.rs:99-117.
 *   height f32[W*H], pos3 f32[W*H][3], balls2 i32[>=3][2], connections8 f32[W*H][8]: the Scene (tod_scene_materialize)
 *   cost_out f32[W*H] / pred_out i32[W*H] (may be NULL): converged costs (f32::MAX = unreached) and predecessors
 *                 (-2 = target, -1 = none)
 *   directions f32[cap][2], *n_directions = entries of the path (may exceed cap; -1 = start cannot reach a target)
 * ====================================================================================== */
int tod_path_reference_panics(void);
int tod_path_modify(int device, int width, int height_px, const float* height, const float* pos3, const int32_t* balls2,
                    const float* connections8, float* cost_out, int32_t* pred_out, float* directions, int cap,
                    int32_t* n_directions);
/* Path::serialize: u64 seconds since the epoch, big-endian, then big-endian f32 (magnitude, rotation) pairs - what
 * handle_path_request writes to the RoboRIO socket (path.rs:158-162).  *bytes = 8 + 8 n. */
int tod_path_serialize(uint64_t created_secs, const float* directions, int n, uint8_t* out, size_t cap, size_t* bytes);

/* ======================================================================================
 * Micro-benchmark used as the int8 roofline denominator (MEASURED_PEAKS.json has no int8 peak,
 * SURVEY §8d): a plain tcgen05 kind::i8 GEMM  C[M,N] s32 = A[M,K] s8 * B[N,K]^T s8.
 * ====================================================================================== */
int tod_i8_gemm_selftest(int device, int M, int N, int K, int iters, float* ms_per_iter, double* max_abs_err);
/* Tensor-pipe peak of tcgen05.mma.kind::i8: every SM issues n_mma M=128 x N=256 x K=32 MMAs from resident shared
 * memory (no loads, no epilogue); *tops = best of `iters` launches in TOP/s.  The conv roofline's denominator. */
int tod_i8_mma_peak(int device, int n_mma, int iters, double* tops);
/* multiply-accumulates per tile of every planned step, in tod_yolact_profile_ops order; returns the step count */
int tod_yolact_step_macs(const tod_yolact* y, int64_t* macs, int cap);

/* Diagnostics: one step enqueued on the executor's lane streams without a CUDA graph, a timing event behind every launch.
 * end_ms[i] = end of planned step i relative to the start of the step, lanes[i] = the lane stream it ran on, kinds[i] as
 * tod_yolact_profile_ops; the last three entries are the segmentation pass, box decode / NMS / top-k, and mask assembly.
 * Returns the number of entries. */
int tod_yolact_trace_steps(tod_yolact* y, int n, float* end_ms, int32_t* lanes, int32_t* kinds, int cap);

/* One KxK (K = 1 or 3), stride-1, SAME convolution [tiles,H,W,IC] -> [tiles,H,W,OC] on seeded random data through
 * both the tcgen05 implicit-GEMM kernel and the CUDA-core direct kernel: times both and counts differing bytes. */
int tod_conv_selftest(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, float* ms_tc,
                      float* ms_direct, long long* mismatches);
/* Same, selecting the epilogue variant under test.  flags: 1 = ReLU6-style clamp instead of the full int8 range,
 * 2 = a fused 256-entry byte map behind the requantisation, 4 = force the general epilogue, 8 = multipliers >= 0.5
 * (the planner must itself fall back to the general epilogue), 16 = activation minimum above -128. */
int tod_conv_selftest_ex(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, int flags,
                         float* ms_tc, float* ms_direct, long long* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* TOD_H_ */
