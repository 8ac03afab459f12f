// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// CPU restatement of what `interpreter.invoke()` (/root/reference/src/yolact.rs:163) computes for
// the ten builtin operators of the FRC model (/root/reference/data/FRC_model_edgetpu.log:7-19):
// CONV_2D, DEPTHWISE_CONV_2D, ADD, QUANTIZE, PAD, RELU, CONCATENATION, RESHAPE, TANH,
// RESIZE_BILINEAR.  The arithmetic lives in a third-party dependency that is NOT under
// /root/reference: TensorFlow Lite C++ via crate `tflite 0.9.0`
// (git littletitan/tflite-rs@abcaeab4, Cargo.lock:1106-1108) wrapped by `edgetpu 0.1.0`
// (git littleTitan/edgetpu-rs@23311e02, Cargo.lock:314-316).  What follows restates TFLite's
// published *reference* integer kernels (reference_integer_ops::ConvPerChannel,
// DepthwiseConvPerChannel, Add, Requantize, ReluX, PadImpl, Concatenation, LUT Tanh,
// ResizeBilinearInteger) per the rules in SURVEY.md §10.
//
// PARITY UNPINNED: neither model blob (.MISSING_LARGE_BLOBS:1-2) nor any TFLite runtime is
// available, and the reference has no golden vector for the interpreter.  The restatement is
// cross-checked against an independent numpy restatement (oracle/synth_model.py) only.
//
// The .tflite container is parsed with a hand-written FlatBuffers reader (no flatbuffers
// library in this image); field ids follow tensorflow/lite/schema/schema.fbs (v3).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "fixedpoint.h"
#include "tod_oracle.h"

namespace oracle {

static thread_local std::string g_err;

// ---------------------------------------------------------------- flatbuffer reader
struct FB {
  const uint8_t* b = nullptr;
  size_t n = 0;
  bool ok(size_t pos, size_t len) const { return pos <= n && len <= n - pos; }
  template <class T>
  T rd(size_t pos) const {
    T v{};
    if (ok(pos, sizeof(T))) std::memcpy(&v, b + pos, sizeof(T));
    return v;
  }
  // absolute position of field `id` of table at `t`, or 0 if absent
  size_t field(size_t t, int id) const {
    const int32_t so = rd<int32_t>(t);
    const size_t vt = static_cast<size_t>(static_cast<int64_t>(t) - so);
    const uint16_t vsz = rd<uint16_t>(vt);
    const size_t slot = 4 + 2 * static_cast<size_t>(id);
    if (slot + 2 > vsz) return 0;
    const uint16_t off = rd<uint16_t>(vt + slot);
    return off ? t + off : 0;
  }
  template <class T>
  T scalar(size_t t, int id, T def) const {
    const size_t p = field(t, id);
    return p ? rd<T>(p) : def;
  }
  size_t indirect(size_t t, int id) const {  // table / vector / string target
    const size_t p = field(t, id);
    return p ? p + rd<uint32_t>(p) : 0;
  }
  uint32_t vlen(size_t v) const { return v ? rd<uint32_t>(v) : 0; }
  size_t vtab(size_t v, uint32_t i) const {  // i-th table of a vector of tables
    const size_t e = v + 4 + 4 * static_cast<size_t>(i);
    return e + rd<uint32_t>(e);
  }
  template <class T>
  std::vector<T> vec(size_t t, int id) const {
    std::vector<T> out;
    const size_t v = indirect(t, id);
    const uint32_t len = vlen(v);
    if (v && ok(v + 4, static_cast<size_t>(len) * sizeof(T))) {
      out.resize(len);
      if (len) std::memcpy(out.data(), b + v + 4, len * sizeof(T));
    }
    return out;
  }
  std::string str(size_t t, int id) const {
    const size_t v = indirect(t, id);
    const uint32_t len = vlen(v);
    if (!v || !ok(v + 4, len)) return {};
    return std::string(reinterpret_cast<const char*>(b + v + 4), len);
  }
};

enum { T_FLOAT32 = 0, T_INT32 = 2, T_UINT8 = 3, T_INT64 = 4, T_INT8 = 9 };
enum {
  OP_ADD = 0, OP_CONCATENATION = 2, OP_CONV_2D = 3, OP_DEPTHWISE_CONV_2D = 4, OP_RELU = 19, OP_RESHAPE = 22,
  OP_RESIZE_BILINEAR = 23, OP_TANH = 28, OP_PAD = 34, OP_QUANTIZE = 114
};
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU_N1_TO_1 = 2, ACT_RELU6 = 3 };
enum { PAD_SAME = 0, PAD_VALID = 1 };

struct Tensor {
  std::vector<int> shape;
  int type = 0;
  std::string name;
  std::vector<float> scale;
  std::vector<int64_t> zp;
  int qdim = 0;
  const uint8_t* cdata = nullptr;  // constant data (points into the file image)
  size_t cbytes = 0;
  std::vector<uint8_t> data;       // runtime data
  int64_t elems() const {
    int64_t e = 1;
    for (int d : shape) e *= d;
    return e;
  }
  int esize() const { return (type == T_INT8 || type == T_UINT8) ? 1 : (type == T_INT64 ? 8 : 4); }
  int dim4(int i) const {  // shape right-aligned to 4-D
    const int r = static_cast<int>(shape.size());
    const int j = i - (4 - r);
    return j < 0 ? 1 : shape[j];
  }
  const uint8_t* ptr() const { return cdata ? cdata : data.data(); }
  float s0() const { return scale.empty() ? 0.0f : scale[0]; }
  int32_t z0() const { return zp.empty() ? 0 : static_cast<int32_t>(zp[0]); }
};

struct Op {
  int code = -1;
  std::vector<int> in, out;
  int padding = 0, stride_w = 1, stride_h = 1, dil_w = 1, dil_h = 1, depth_mult = 1, act = 0, axis = 0;
  bool align_corners = false, half_pixel = false;
};

struct Model {
  std::vector<uint8_t> file;
  std::vector<Tensor> tensors;
  std::vector<Op> ops;
  std::vector<int> inputs, outputs;
  int64_t macs = 0;
};

static bool load(const char* path, Model* m) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { g_err = std::string("cannot open ") + path; return false; }
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  m->file.resize(sz > 0 ? sz : 0);
  if (sz <= 0 || std::fread(m->file.data(), 1, sz, f) != static_cast<size_t>(sz)) {
    std::fclose(f);
    g_err = "short read";
    return false;
  }
  std::fclose(f);
  FB fb{m->file.data(), m->file.size()};
  if (sz < 8 || std::memcmp(m->file.data() + 4, "TFL3", 4) != 0) { g_err = "not a TFL3 flatbuffer"; return false; }
  const size_t root = fb.rd<uint32_t>(0);
  const size_t opcodes = fb.indirect(root, 1), subgraphs = fb.indirect(root, 2), buffers = fb.indirect(root, 4);
  if (!subgraphs || fb.vlen(subgraphs) < 1) { g_err = "no subgraph"; return false; }
  std::vector<int> codes;
  for (uint32_t i = 0; i < fb.vlen(opcodes); ++i) {
    const size_t oc = fb.vtab(opcodes, i);
    const int dep = fb.scalar<int8_t>(oc, 0, 0), cur = fb.scalar<int32_t>(oc, 3, 0);
    if (fb.indirect(oc, 1)) { g_err = "custom op '" + fb.str(oc, 1) + "' (e.g. edgetpu-custom-op) is not supported"; return false; }
    codes.push_back(std::max(dep, cur));
  }
  const size_t sg = fb.vtab(subgraphs, 0);
  const size_t tens = fb.indirect(sg, 0), opsv = fb.indirect(sg, 3);
  m->inputs = fb.vec<int32_t>(sg, 1);
  m->outputs = fb.vec<int32_t>(sg, 2);
  m->tensors.resize(fb.vlen(tens));
  for (uint32_t i = 0; i < fb.vlen(tens); ++i) {
    const size_t t = fb.vtab(tens, i);
    Tensor& T = m->tensors[i];
    T.shape = fb.vec<int32_t>(t, 0);
    T.type = fb.scalar<int8_t>(t, 1, 0);
    T.name = fb.str(t, 3);
    const uint32_t bi = fb.scalar<uint32_t>(t, 2, 0);
    const size_t q = fb.indirect(t, 4);
    if (q) {
      T.scale = fb.vec<float>(q, 2);
      T.zp = fb.vec<int64_t>(q, 3);
      T.qdim = fb.scalar<int32_t>(q, 6, 0);
    }
    if (bi > 0 && buffers && bi < fb.vlen(buffers)) {
      const size_t bt = fb.vtab(buffers, bi);
      const size_t dv = fb.indirect(bt, 0);
      if (dv && fb.vlen(dv) > 0) { T.cdata = fb.b + dv + 4; T.cbytes = fb.vlen(dv); }
    }
  }
  m->ops.resize(fb.vlen(opsv));
  for (uint32_t i = 0; i < fb.vlen(opsv); ++i) {
    const size_t o = fb.vtab(opsv, i);
    Op& P = m->ops[i];
    const uint32_t oi = fb.scalar<uint32_t>(o, 0, 0);
    if (oi >= codes.size()) { g_err = "bad opcode index"; return false; }
    P.code = codes[oi];
    P.in = fb.vec<int32_t>(o, 1);
    P.out = fb.vec<int32_t>(o, 2);
    const size_t bo = fb.indirect(o, 4);
    switch (P.code) {
      case OP_CONV_2D:
        if (bo) {
          P.padding = fb.scalar<int8_t>(bo, 0, 0); P.stride_w = fb.scalar<int32_t>(bo, 1, 0); P.stride_h = fb.scalar<int32_t>(bo, 2, 0);
          P.act = fb.scalar<int8_t>(bo, 3, 0); P.dil_w = fb.scalar<int32_t>(bo, 4, 1); P.dil_h = fb.scalar<int32_t>(bo, 5, 1);
        }
        break;
      case OP_DEPTHWISE_CONV_2D:
        if (bo) {
          P.padding = fb.scalar<int8_t>(bo, 0, 0); P.stride_w = fb.scalar<int32_t>(bo, 1, 0); P.stride_h = fb.scalar<int32_t>(bo, 2, 0);
          P.depth_mult = fb.scalar<int32_t>(bo, 3, 0); P.act = fb.scalar<int8_t>(bo, 4, 0);
          P.dil_w = fb.scalar<int32_t>(bo, 5, 1); P.dil_h = fb.scalar<int32_t>(bo, 6, 1);
        }
        break;
      case OP_ADD: if (bo) P.act = fb.scalar<int8_t>(bo, 0, 0); break;
      case OP_CONCATENATION: if (bo) { P.axis = fb.scalar<int32_t>(bo, 0, 0); P.act = fb.scalar<int8_t>(bo, 1, 0); } break;
      case OP_RESIZE_BILINEAR: if (bo) { P.align_corners = fb.scalar<uint8_t>(bo, 2, 0) != 0; P.half_pixel = fb.scalar<uint8_t>(bo, 3, 0) != 0; } break;
      case OP_RELU: case OP_RESHAPE: case OP_TANH: case OP_PAD: case OP_QUANTIZE: break;
      default: g_err = "unsupported builtin operator code " + std::to_string(P.code); return false;
    }
  }
  return true;
}

// ---------------------------------------------------------------- kernels
static void act_range(int act, const Tensor& out, int32_t* lo, int32_t* hi) {
  // TFLite CalculateActivationRangeQuantized
  const int32_t qmin = out.type == T_UINT8 ? 0 : -128, qmax = out.type == T_UINT8 ? 255 : 127;
  const float scale = out.s0();
  const int32_t zp = out.z0();
  auto quant = [&](float f) { return zp + static_cast<int32_t>(std::round(f / scale)); };
  *lo = qmin; *hi = qmax;
  if (act == ACT_RELU) *lo = std::max(qmin, quant(0.0f));
  else if (act == ACT_RELU6) { *lo = std::max(qmin, quant(0.0f)); *hi = std::min(qmax, quant(6.0f)); }
  else if (act == ACT_RELU_N1_TO_1) { *lo = std::max(qmin, quant(-1.0f)); *hi = std::min(qmax, quant(1.0f)); }
}

static void conv_padding(int padding, int in, int k, int stride, int dil, int* out, int* pad) {
  // TFLite ComputeOutSize / ComputePaddingWithOffset
  const int eff = (k - 1) * dil + 1;
  if (padding == PAD_SAME) *out = (in + stride - 1) / stride;
  else *out = (in + stride - eff) / stride;
  const int total = std::max(0, (*out - 1) * stride + eff - in);
  *pad = padding == PAD_SAME ? total / 2 : 0;
}

static void per_channel_mult(const Tensor& in, const Tensor& w, const Tensor& out, int channels, std::vector<int32_t>* q,
                             std::vector<int>* sh) {
  // TFLite PopulateConvolutionQuantizationParams
  q->resize(channels); sh->resize(channels);
  for (int c = 0; c < channels; ++c) {
    const float ws = w.scale.size() > 1 ? w.scale[c] : w.s0();
    const double eff = static_cast<double>(in.s0()) * static_cast<double>(ws) / static_cast<double>(out.s0());
    QuantizeMultiplier(eff, &(*q)[c], &(*sh)[c]);
  }
}

static bool run_conv(Model& m, const Op& op, bool depthwise) {
  const Tensor& in = m.tensors[op.in[0]];
  const Tensor& w = m.tensors[op.in[1]];
  const Tensor* bias = op.in.size() > 2 && op.in[2] >= 0 ? &m.tensors[op.in[2]] : nullptr;
  Tensor& out = m.tensors[op.out[0]];
  if (in.type != T_INT8 || w.type != T_INT8 || out.type != T_INT8) { g_err = "conv: only int8 supported"; return false; }
  const int B = in.dim4(0), IH = in.dim4(1), IW = in.dim4(2), IC = in.dim4(3);
  const int OC = depthwise ? w.dim4(3) : w.dim4(0), KH = w.dim4(1), KW = w.dim4(2);
  int OH, OW, ph, pw;
  conv_padding(op.padding, IH, KH, op.stride_h, op.dil_h, &OH, &ph);
  conv_padding(op.padding, IW, KW, op.stride_w, op.dil_w, &OW, &pw);
  if (out.dim4(1) != OH || out.dim4(2) != OW || out.dim4(3) != OC) { g_err = "conv: output shape mismatch for " + out.name; return false; }
  std::vector<int32_t> q; std::vector<int> sh;
  per_channel_mult(in, w, out, OC, &q, &sh);
  int32_t lo, hi;
  act_range(op.act, out, &lo, &hi);
  const int32_t in_off = -in.z0(), out_off = out.z0();
  const int8_t* I = reinterpret_cast<const int8_t*>(in.ptr());
  const int8_t* Wt = reinterpret_cast<const int8_t*>(w.ptr());
  const int32_t* Bs = bias ? reinterpret_cast<const int32_t*>(bias->ptr()) : nullptr;
  int8_t* O = reinterpret_cast<int8_t*>(out.data.data());
  const int dm = op.depth_mult;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int oy = 0; oy < OH; ++oy) {
      std::vector<int32_t> acc(OC);
      for (int ox = 0; ox < OW; ++ox) {
        std::fill(acc.begin(), acc.end(), 0);
        const int iy0 = oy * op.stride_h - ph, ix0 = ox * op.stride_w - pw;
        for (int fy = 0; fy < KH; ++fy) {
          const int iy = iy0 + op.dil_h * fy;
          if (iy < 0 || iy >= IH) continue;
          for (int fx = 0; fx < KW; ++fx) {
            const int ix = ix0 + op.dil_w * fx;
            if (ix < 0 || ix >= IW) continue;  // out-of-range taps are skipped
            const int8_t* ip = I + ((static_cast<size_t>(b) * IH + iy) * IW + ix) * IC;
            if (depthwise) {
              const int8_t* wp = Wt + (static_cast<size_t>(fy) * KW + fx) * OC;
              for (int ic = 0; ic < IC; ++ic) {
                const int32_t v = ip[ic] + in_off;
                for (int k = 0; k < dm; ++k) acc[ic * dm + k] += wp[ic * dm + k] * v;
              }
            } else {
              for (int oc = 0; oc < OC; ++oc) {
                const int8_t* wp = Wt + ((static_cast<size_t>(oc) * KH + fy) * KW + fx) * IC;
                int32_t a = 0;
                for (int ic = 0; ic < IC; ++ic) a += wp[ic] * (ip[ic] + in_off);
                acc[oc] += a;
              }
            }
          }
        }
        int8_t* op_ = O + ((static_cast<size_t>(b) * OH + oy) * OW + ox) * OC;
        for (int oc = 0; oc < OC; ++oc) {
          int32_t a = acc[oc] + (Bs ? Bs[oc] : 0);
          a = MBQM(a, q[oc], sh[oc]) + out_off;
          a = std::min(std::max(a, lo), hi);
          op_[oc] = static_cast<int8_t>(a);
        }
      }
    }
  const int64_t per_out = depthwise ? static_cast<int64_t>(KH) * KW : static_cast<int64_t>(KH) * KW * IC;
  m.macs += static_cast<int64_t>(B) * OH * OW * OC * per_out;
  return true;
}

static inline int32_t load_q(const Tensor& t, int64_t i) {
  return t.type == T_UINT8 ? static_cast<int32_t>(t.ptr()[i]) : static_cast<int32_t>(reinterpret_cast<const int8_t*>(t.ptr())[i]);
}
static inline void store_q(Tensor& t, int64_t i, int32_t v) {
  if (t.type == T_UINT8) t.data[i] = static_cast<uint8_t>(v);
  else reinterpret_cast<int8_t*>(t.data.data())[i] = static_cast<int8_t>(v);
}

static bool run_add(Model& m, const Op& op) {
  const Tensor& a = m.tensors[op.in[0]];
  const Tensor& b = m.tensors[op.in[1]];
  Tensor& out = m.tensors[op.out[0]];
  if (a.elems() != b.elems() || a.elems() != out.elems()) { g_err = "add: broadcasting unsupported"; return false; }
  // TFLite add.cc Prepare
  const int left_shift = 20;
  const double twice_max = 2 * std::max(a.s0(), b.s0());
  int32_t m1, m2, mo; int s1, s2, so;
  QuantizeMultiplier(a.s0() / twice_max, &m1, &s1);
  QuantizeMultiplier(b.s0() / twice_max, &m2, &s2);
  QuantizeMultiplier(twice_max / ((1 << left_shift) * out.s0()), &mo, &so);
  int32_t lo, hi;
  act_range(op.act, out, &lo, &hi);
  const int64_t n = out.elems();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int32_t x1 = (load_q(a, i) - a.z0()) * (1 << left_shift);
    const int32_t x2 = (load_q(b, i) - b.z0()) * (1 << left_shift);
    const int32_t y1 = MBQM(x1, m1, s1), y2 = MBQM(x2, m2, s2);
    int32_t r = MBQM(y1 + y2, mo, so) + out.z0();
    r = std::min(std::max(r, lo), hi);
    store_q(out, i, r);
  }
  return true;
}

static bool run_quantize(Model& m, const Op& op, bool relu) {
  const Tensor& in = m.tensors[op.in[0]];
  Tensor& out = m.tensors[op.out[0]];
  const int32_t qmin = out.type == T_UINT8 ? 0 : -128, qmax = out.type == T_UINT8 ? 255 : 127;
  const int64_t n = out.elems();
  if (in.type == T_FLOAT32) {  // SURVEY §10.5 float -> int8
    const float* f = reinterpret_cast<const float*>(in.ptr());
    for (int64_t i = 0; i < n; ++i) {
      int32_t v = static_cast<int32_t>(std::round(f[i] / out.s0())) + out.z0();
      store_q(out, i, std::min(std::max(v, qmin), qmax));
    }
    return true;
  }
  int32_t q; int sh;
  QuantizeMultiplier(static_cast<double>(in.s0()) / static_cast<double>(out.s0()), &q, &sh);
  int32_t lo = qmin, hi = qmax;
  if (relu) lo = std::max(qmin, out.z0() + static_cast<int32_t>(std::round(0.0f / out.s0())));  // SURVEY §10.5 RELU
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int32_t v = MBQM(load_q(in, i) - in.z0(), q, sh) + out.z0();
    store_q(out, i, std::min(std::max(v, lo), hi));
  }
  return true;
}

static bool run_pad(Model& m, const Op& op) {
  const Tensor& in = m.tensors[op.in[0]];
  const Tensor& pd = m.tensors[op.in[1]];
  Tensor& out = m.tensors[op.out[0]];
  const int r = static_cast<int>(in.shape.size());
  if (r != 4 || pd.type != T_INT32 || pd.elems() != 8) { g_err = "pad: only 4-D int32 paddings"; return false; }
  const int32_t* p = reinterpret_cast<const int32_t*>(pd.ptr());
  if (p[0] || p[1] || p[6] || p[7]) { g_err = "pad: batch/channel padding unsupported"; return false; }
  const int B = in.dim4(0), H = in.dim4(1), W = in.dim4(2), C = in.dim4(3);
  const int OH = out.dim4(1), OW = out.dim4(2);
  if (OH != H + p[2] + p[3] || OW != W + p[4] + p[5]) { g_err = "pad: shape mismatch"; return false; }
  std::memset(out.data.data(), static_cast<uint8_t>(out.z0()), out.data.size());  // pad value = zero point (§10.5)
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      std::memcpy(out.data.data() + ((static_cast<size_t>(b) * OH + y + p[2]) * OW + p[4]) * C,
                  in.ptr() + (static_cast<size_t>(b) * H + y) * W * C, static_cast<size_t>(W) * C);
  return true;
}

static bool run_concat(Model& m, const Op& op) {
  Tensor& out = m.tensors[op.out[0]];
  const int r = static_cast<int>(out.shape.size());
  const int axis = op.axis < 0 ? op.axis + r : op.axis;
  int64_t outer = 1, inner = 1;
  for (int i = 0; i < axis; ++i) outer *= out.shape[i];
  for (int i = axis + 1; i < r; ++i) inner *= out.shape[i];
  int64_t off = 0;
  for (int ti : op.in) {
    const Tensor& in = m.tensors[ti];
    if (in.s0() != out.s0() || in.z0() != out.z0()) { g_err = "concat: inputs must share the output quantisation"; return false; }
    const int64_t chunk = in.shape[axis] * inner;
    for (int64_t o = 0; o < outer; ++o)
      std::memcpy(out.data.data() + (o * out.shape[axis] * inner + off), in.ptr() + o * chunk, chunk);
    off += chunk;
  }
  return true;
}

static bool run_tanh(Model& m, const Op& op) {
  const Tensor& in = m.tensors[op.in[0]];
  Tensor& out = m.tensors[op.out[0]];
  // TFLite LUT tanh
  const int32_t qmin = in.type == T_UINT8 ? 0 : -128, qmax = in.type == T_UINT8 ? 255 : 127;
  const float inverse_scale = 1 / out.s0();
  uint8_t lut[256];
  for (int32_t v = qmin; v <= qmax; ++v) {
    const float deq = in.s0() * (v - in.z0());
    const float tr = std::tanh(deq);
    const float resc = std::round(tr * inverse_scale);
    const int32_t qv = static_cast<int32_t>(resc + out.z0());
    lut[static_cast<uint8_t>(v)] = static_cast<uint8_t>(std::max(std::min(qmax, qv), qmin));
  }
  const int64_t n = out.elems();
  for (int64_t i = 0; i < n; ++i) out.data[i] = lut[in.ptr()[i]];
  return true;
}

static bool run_resize(Model& m, const Op& op) {
  const Tensor& in = m.tensors[op.in[0]];
  Tensor& out = m.tensors[op.out[0]];
  // TFLite reference_ops::ResizeBilinearInteger: 10-bit fixed-point weights (chosen variant)
  const int B = in.dim4(0), IH = in.dim4(1), IW = in.dim4(2), C = in.dim4(3), OH = out.dim4(1), OW = out.dim4(2);
  int32_t hs = ((1 << 10) * IH + OH / 2) / OH, ws = ((1 << 10) * IW + OW / 2) / OW;
  if (op.align_corners && OH > 1) hs = ((1 << 10) * (IH - 1) + (OH - 1) / 2) / (OH - 1);
  if (op.align_corners && OW > 1) ws = ((1 << 10) * (IW - 1) + (OW - 1) / 2) / (OW - 1);
  auto interp = [&](int32_t v, int32_t s10, int32_t size, int32_t* sv, int32_t* lo, int32_t* hi) {
    *sv = v * s10;
    if (op.half_pixel) *sv += s10 / 2 - (1 << 9);
    *lo = std::max(*sv / (1 << 10), 0);
    *hi = std::min((*sv + (1 << 10) - 1) / (1 << 10), size - 1);
  };
  const int8_t* I = reinterpret_cast<const int8_t*>(in.ptr());
  int8_t* O = reinterpret_cast<int8_t*>(out.data.data());
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < OH; ++y) {
      int32_t iy, y0, y1;
      interp(y, hs, IH, &iy, &y0, &y1);
      for (int x = 0; x < OW; ++x) {
        int32_t ix, x0, x1;
        interp(x, ws, IW, &ix, &x0, &x1);
        const int64_t wy1 = iy - (1 << 10) * y0, wy0 = (1 << 10) - wy1;
        const int64_t wx1 = ix - (1 << 10) * x0, wx0 = (1 << 10) - wx1;
        for (int c = 0; c < C; ++c) {
          auto at = [&](int yy, int xx) { return static_cast<int64_t>(I[((static_cast<size_t>(b) * IH + yy) * IW + xx) * C + c]); };
          const int64_t o20 = at(y0, x0) * wy0 * wx0 + at(y1, x0) * wy1 * wx0 + at(y0, x1) * wy0 * wx1 + at(y1, x1) * wy1 * wx1;
          const int64_t round = o20 > 0 ? (1 << 19) : -(1 << 19);
          O[((static_cast<size_t>(b) * OH + y) * OW + x) * C + c] = static_cast<int8_t>((o20 + round) / (1 << 20));
        }
      }
    }
  return true;
}

static bool invoke(Model& m, const uint8_t* input, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  m.macs = 0;
  for (Tensor& t : m.tensors)
    if (!t.cdata && t.data.size() != static_cast<size_t>(t.elems() * t.esize())) t.data.assign(t.elems() * t.esize(), 0);
  Tensor& in = m.tensors[m.inputs[0]];
  std::memcpy(in.data.data(), input, in.data.size());  // yolact.rs:161-162
  for (const Op& op : m.ops) {
    bool ok = true;
    switch (op.code) {
      case OP_CONV_2D: ok = run_conv(m, op, false); break;
      case OP_DEPTHWISE_CONV_2D: ok = run_conv(m, op, true); break;
      case OP_ADD: ok = run_add(m, op); break;
      case OP_QUANTIZE: ok = run_quantize(m, op, false); break;
      case OP_RELU: ok = run_quantize(m, op, true); break;
      case OP_PAD: ok = run_pad(m, op); break;
      case OP_CONCATENATION: ok = run_concat(m, op); break;
      case OP_RESHAPE: {
        const Tensor& a = m.tensors[op.in[0]];
        Tensor& o = m.tensors[op.out[0]];
        if (a.elems() != o.elems()) { g_err = "reshape: element count mismatch"; ok = false; break; }
        std::memcpy(o.data.data(), a.ptr(), o.data.size());
        break;
      }
      case OP_TANH: ok = run_tanh(m, op); break;
      case OP_RESIZE_BILINEAR: ok = run_resize(m, op); break;
      default: g_err = "unsupported op"; ok = false;
    }
    if (!ok) return false;
  }
  return true;
}

}  // namespace oracle

using oracle::Model;

extern "C" {

tod_oracle_model* tod_oracle_model_load(const char* path) {
  Model* m = new Model();
  if (!oracle::load(path, m)) { delete m; return nullptr; }
  return reinterpret_cast<tod_oracle_model*>(m);
}
void tod_oracle_model_free(tod_oracle_model* m) { delete reinterpret_cast<Model*>(m); }
const char* tod_oracle_last_error(void) { return oracle::g_err.c_str(); }
int tod_oracle_model_num_tensors(const tod_oracle_model* m) { return static_cast<int>(reinterpret_cast<const Model*>(m)->tensors.size()); }
int tod_oracle_model_num_ops(const tod_oracle_model* m) { return static_cast<int>(reinterpret_cast<const Model*>(m)->ops.size()); }
int tod_oracle_model_num_inputs(const tod_oracle_model* m) { return static_cast<int>(reinterpret_cast<const Model*>(m)->inputs.size()); }
int tod_oracle_model_num_outputs(const tod_oracle_model* m) { return static_cast<int>(reinterpret_cast<const Model*>(m)->outputs.size()); }
int tod_oracle_model_output_tensor(const tod_oracle_model* m, int i) { return reinterpret_cast<const Model*>(m)->outputs[i]; }
int tod_oracle_model_input_tensor(const tod_oracle_model* m, int i) { return reinterpret_cast<const Model*>(m)->inputs[i]; }
int tod_oracle_model_op_code(const tod_oracle_model* m, int op) { return reinterpret_cast<const Model*>(m)->ops[op].code; }
int tod_oracle_model_op_output(const tod_oracle_model* m, int op, int k) { return reinterpret_cast<const Model*>(m)->ops[op].out[k]; }
int64_t tod_oracle_model_tensor_info(const tod_oracle_model* m, int t, int* shape4, int* type, float* scale, int* zero_point) {
  const oracle::Tensor& T = reinterpret_cast<const Model*>(m)->tensors[t];
  for (int i = 0; i < 4; ++i) shape4[i] = T.dim4(i);
  *type = T.type;
  *scale = T.s0();
  *zero_point = T.z0();
  return T.elems();
}
int tod_oracle_model_invoke(tod_oracle_model* m, const uint8_t* input, int threads) {
  return oracle::invoke(*reinterpret_cast<Model*>(m), input, threads) ? 0 : -1;
}
const void* tod_oracle_model_tensor_data(const tod_oracle_model* m, int t) { return reinterpret_cast<const Model*>(m)->tensors[t].ptr(); }
int64_t tod_oracle_model_macs(const tod_oracle_model* m) { return reinterpret_cast<const Model*>(m)->macs; }

int32_t tod_oracle_srdhm(int32_t a, int32_t b) { return oracle::SRDHM(a, b); }
int32_t tod_oracle_rdivpot(int32_t x, int e) { return oracle::RDivPOT(x, e); }
int32_t tod_oracle_mbqm(int32_t x, int32_t q, int shift) { return oracle::MBQM(x, q, shift); }
void tod_oracle_quantize_multiplier(double m, int32_t* q, int* shift) { oracle::QuantizeMultiplier(m, q, shift); }

}  // extern "C"
