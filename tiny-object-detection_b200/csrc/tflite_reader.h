// Product-side .tflite container reader (host only).
//
// Replaces FlatBufferModel::build_from_file + InterpreterBuilder of the reference
// (/root/reference/src/yolact.rs:18-29): parses the TFL3 flatbuffer into a flat graph description
// the CUDA planner (yolact_plan.cu) consumes.  No flatbuffers library exists in this image, so the
// container is walked by hand; table/field numbering is the public tensorflow/lite/schema/schema.fbs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tod {

enum TfType : int { kF32 = 0, kI32 = 2, kU8 = 3, kI64 = 4, kI8 = 9 };

enum TfOp : int {
  kAdd = 0,
  kConcat = 2,
  kConv2D = 3,
  kDepthwise = 4,
  kRelu = 19,
  kReshape = 22,
  kResizeBilinear = 23,
  kTanh = 28,
  kPad = 34,
  kQuantize = 114
};

enum TfAct : int { kActNone = 0, kActRelu = 1, kActReluN1To1 = 2, kActRelu6 = 3 };
enum TfPadding : int { kSame = 0, kValid = 1 };

struct GTensor {
  std::string name;
  int type = 0;
  int rank = 0;
  int dims[4] = {1, 1, 1, 1};  // right-aligned NHWC
  std::vector<float> scales;   // 1 (per tensor) or C (per channel)
  std::vector<int64_t> zero_points;
  int quant_dim = 0;
  const uint8_t* const_data = nullptr;  // into Graph::blob
  size_t const_bytes = 0;

  int64_t elems() const { return int64_t(dims[0]) * dims[1] * dims[2] * dims[3]; }
  int elem_size() const { return (type == kI8 || type == kU8) ? 1 : (type == kI64 ? 8 : 4); }
  float scale() const { return scales.empty() ? 0.f : scales[0]; }
  int32_t zp() const { return zero_points.empty() ? 0 : int32_t(zero_points[0]); }
  bool is_const() const { return const_data != nullptr; }
};

struct GOp {
  int code = -1;
  std::vector<int> inputs, outputs;
  int padding = kSame;
  int stride_h = 1, stride_w = 1, dil_h = 1, dil_w = 1;
  int depth_multiplier = 1;
  int activation = kActNone;
  int axis = 0;
  bool align_corners = false, half_pixel_centers = false;
};

struct Graph {
  std::vector<uint8_t> blob;  // the file image; const tensors point into it
  std::vector<GTensor> tensors;
  std::vector<GOp> ops;
  std::vector<int> inputs, outputs;
  std::string description;
};

// Returns TOD_OK or a negative tod_status (TOD_ERR_IO / TOD_ERR_MODEL) with tod_last_error() set.
int read_tflite(const char* path, Graph* g);

}  // namespace tod
