// See tflite_reader.h.  Hand-rolled, bounds-checked FlatBuffers walk of a TFL3 model.
#include "tflite_reader.h"

#include <cstdio>
#include <cstring>

#include "common.h"

namespace tod {
namespace {

// A bounds-checked view of the file.  Every accessor returns a neutral value on an out-of-range
// read and latches `bad`, so a truncated / hostile file ends in TOD_ERR_MODEL, never in a crash.
class Blob {
 public:
  Blob(const uint8_t* p, size_t n) : p_(p), n_(n) {}
  bool bad = false;

  template <class T>
  T at(size_t pos) {
    T v{};
    if (pos > n_ || sizeof(T) > n_ - pos) {
      bad = true;
      return v;
    }
    std::memcpy(&v, p_ + pos, sizeof(T));
    return v;
  }
  const uint8_t* ptr(size_t pos, size_t len) {
    if (pos > n_ || len > n_ - pos) {
      bad = true;
      return nullptr;
    }
    return p_ + pos;
  }
  size_t follow(size_t pos) { return pos + at<uint32_t>(pos); }  // uoffset

 private:
  const uint8_t* p_;
  size_t n_;
};

struct Table {
  Blob* b = nullptr;
  size_t pos = 0;  // 0 == absent
  explicit operator bool() const { return pos != 0; }

  size_t slot(int id) const {
    if (!pos) return 0;
    const int32_t back = b->at<int32_t>(pos);
    const int64_t vt = int64_t(pos) - back;
    if (vt < 0) {
      b->bad = true;
      return 0;
    }
    const uint16_t vt_bytes = b->at<uint16_t>(size_t(vt));
    const size_t entry = 4 + 2 * size_t(id);
    if (entry + 2 > vt_bytes) return 0;
    const uint16_t off = b->at<uint16_t>(size_t(vt) + entry);
    return off ? pos + off : 0;
  }
  template <class T>
  T get(int id, T fallback) const {
    const size_t s = slot(id);
    return s ? b->at<T>(s) : fallback;
  }
  Table child(int id) const {
    const size_t s = slot(id);
    return Table{b, s ? b->follow(s) : 0};
  }
  // vector start (position of the u32 length) or 0
  size_t vec(int id) const {
    const size_t s = slot(id);
    return s ? b->follow(s) : 0;
  }
};

uint32_t vec_len(Blob& b, size_t v) { return v ? b.at<uint32_t>(v) : 0; }
Table vec_table(Blob& b, size_t v, uint32_t i) { return Table{&b, b.follow(v + 4 + 4 * size_t(i))}; }

template <class T>
std::vector<T> vec_scalars(Blob& b, size_t v) {
  std::vector<T> out;
  const uint32_t n = vec_len(b, v);
  if (!n) return out;
  const uint8_t* p = b.ptr(v + 4, size_t(n) * sizeof(T));
  if (!p) return out;
  out.resize(n);
  std::memcpy(out.data(), p, size_t(n) * sizeof(T));
  return out;
}

std::string vec_string(Blob& b, size_t v) {
  const uint32_t n = vec_len(b, v);
  const uint8_t* p = n ? b.ptr(v + 4, n) : nullptr;
  return p ? std::string(reinterpret_cast<const char*>(p), n) : std::string();
}

}  // namespace

int read_tflite(const char* path, Graph* g) {
  if (!path || !g) return fail(TOD_ERR_INVALID_ARG, "read_tflite: null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(TOD_ERR_IO, "cannot open model file '%s'", path);
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 16) {
    std::fclose(f);
    return fail(TOD_ERR_MODEL, "'%s' is too small to be a .tflite model", path);
  }
  g->blob.resize(size_t(sz));
  const size_t got = std::fread(g->blob.data(), 1, size_t(sz), f);
  std::fclose(f);
  if (got != size_t(sz)) return fail(TOD_ERR_IO, "short read on '%s'", path);
  if (std::memcmp(g->blob.data() + 4, "TFL3", 4) != 0)
    return fail(TOD_ERR_MODEL, "'%s' has no TFL3 file identifier", path);

  Blob b(g->blob.data(), g->blob.size());
  const Table model{&b, b.follow(0)};
  g->description = vec_string(b, model.vec(3));

  // operator codes; a custom code (the edgetpu-custom-op of FRC_model_edgetpu.tflite,
  // yolact.rs:19,23) cannot run here and is rejected.
  std::vector<int> codes;
  const size_t oc = model.vec(1);
  for (uint32_t i = 0; i < vec_len(b, oc); ++i) {
    const Table t = vec_table(b, oc, i);
    const std::string custom = vec_string(b, t.vec(1));
    if (!custom.empty())
      return fail(TOD_ERR_MODEL,
                  "model uses custom operator '%s' (an EdgeTPU-compiled model?); load the CPU model "
                  "data/FRC_model.tflite instead",
                  custom.c_str());
    const int legacy = t.get<int8_t>(0, 0);
    const int full = t.get<int32_t>(3, 0);
    codes.push_back(full > legacy ? full : legacy);
  }

  const size_t sgs = model.vec(2);
  if (vec_len(b, sgs) != 1) return fail(TOD_ERR_MODEL, "expected exactly one subgraph, found %u", vec_len(b, sgs));
  const Table sg = vec_table(b, sgs, 0);
  const size_t bufs = model.vec(4);

  g->inputs = vec_scalars<int32_t>(b, sg.vec(1));
  g->outputs = vec_scalars<int32_t>(b, sg.vec(2));

  const size_t tv = sg.vec(0);
  const uint32_t nt = vec_len(b, tv);
  // every vector element occupies at least four bytes of the file: a count beyond that is a corrupt length, not a model
  if (size_t(nt) > g->blob.size() / 4) return fail(TOD_ERR_MODEL, "tensor count %u exceeds what a %zu-byte file can hold", nt, g->blob.size());
  g->tensors.resize(nt);
  for (uint32_t i = 0; i < nt && !b.bad; ++i) {
    const Table t = vec_table(b, tv, i);
    GTensor& T = g->tensors[i];
    const std::vector<int32_t> shape = vec_scalars<int32_t>(b, t.vec(0));
    if (shape.size() > 4) return fail(TOD_ERR_MODEL, "tensor %u has rank %zu > 4", i, shape.size());
    T.rank = int(shape.size());
    for (size_t k = 0; k < shape.size(); ++k) {
      if (shape[k] <= 0) return fail(TOD_ERR_MODEL, "tensor %u has a non-positive / dynamic dimension", i);
      T.dims[4 - shape.size() + k] = shape[k];
    }
    T.type = t.get<int8_t>(1, 0);
    T.name = vec_string(b, t.vec(3));
    const Table q = t.child(4);
    if (q) {
      T.scales = vec_scalars<float>(b, q.vec(2));
      T.zero_points = vec_scalars<int64_t>(b, q.vec(3));
      T.quant_dim = q.get<int32_t>(6, 0);
    }
    const uint32_t bi = t.get<uint32_t>(2, 0);
    if (bi != 0) {
      if (bi >= vec_len(b, bufs)) return fail(TOD_ERR_MODEL, "tensor %u references buffer %u of %u", i, bi, vec_len(b, bufs));
      const size_t dv = vec_table(b, bufs, bi).vec(0);
      const uint32_t len = vec_len(b, dv);
      if (len) {
        T.const_data = b.ptr(dv + 4, len);
        T.const_bytes = len;
        if (T.const_data && size_t(T.elems()) * T.elem_size() != len)
          return fail(TOD_ERR_MODEL, "tensor %u ('%s'): buffer holds %u bytes, shape needs %lld", i, T.name.c_str(), len,
                      (long long)(T.elems() * T.elem_size()));
      }
    }
  }

  const size_t ov = sg.vec(3);
  const uint32_t no = vec_len(b, ov);
  if (size_t(no) > g->blob.size() / 4) return fail(TOD_ERR_MODEL, "operator count %u exceeds what a %zu-byte file can hold", no, g->blob.size());
  g->ops.resize(no);
  for (uint32_t i = 0; i < no && !b.bad; ++i) {
    const Table o = vec_table(b, ov, i);
    GOp& P = g->ops[i];
    const uint32_t ci = o.get<uint32_t>(0, 0);
    if (ci >= codes.size()) return fail(TOD_ERR_MODEL, "operator %u: opcode index %u out of range", i, ci);
    P.code = codes[ci];
    P.inputs = vec_scalars<int32_t>(b, o.vec(1));
    P.outputs = vec_scalars<int32_t>(b, o.vec(2));
    for (int t : P.inputs)
      if (t >= int(nt)) return fail(TOD_ERR_MODEL, "operator %u: input tensor %d out of range", i, t);
    for (int t : P.outputs)
      if (t < 0 || t >= int(nt)) return fail(TOD_ERR_MODEL, "operator %u: output tensor %d out of range", i, t);
    const Table opt = o.child(4);
    switch (P.code) {
      case kConv2D:
        P.padding = opt.get<int8_t>(0, 0);
        P.stride_w = opt.get<int32_t>(1, 1);
        P.stride_h = opt.get<int32_t>(2, 1);
        P.activation = opt.get<int8_t>(3, 0);
        P.dil_w = opt.get<int32_t>(4, 1);
        P.dil_h = opt.get<int32_t>(5, 1);
        break;
      case kDepthwise:
        P.padding = opt.get<int8_t>(0, 0);
        P.stride_w = opt.get<int32_t>(1, 1);
        P.stride_h = opt.get<int32_t>(2, 1);
        P.depth_multiplier = opt.get<int32_t>(3, 1);
        P.activation = opt.get<int8_t>(4, 0);
        P.dil_w = opt.get<int32_t>(5, 1);
        P.dil_h = opt.get<int32_t>(6, 1);
        break;
      case kAdd:
        P.activation = opt.get<int8_t>(0, 0);
        break;
      case kConcat:
        P.axis = opt.get<int32_t>(0, 0);
        P.activation = opt.get<int8_t>(1, 0);
        break;
      case kResizeBilinear:
        P.align_corners = opt.get<uint8_t>(2, 0) != 0;
        P.half_pixel_centers = opt.get<uint8_t>(3, 0) != 0;
        break;
      case kRelu:
      case kReshape:
      case kTanh:
      case kPad:
      case kQuantize:
        break;
      default:
        return fail(TOD_ERR_MODEL, "operator %u: builtin code %d is not one of the ten the FRC model uses", i, P.code);
    }
  }
  if (b.bad) return fail(TOD_ERR_MODEL, "'%s' is truncated or malformed (offset outside the file)", path);
  if (g->inputs.size() != 1) return fail(TOD_ERR_MODEL, "expected one input tensor, found %zu", g->inputs.size());
  return TOD_OK;
}

}  // namespace tod
