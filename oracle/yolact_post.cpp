// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// CPU restatement of the Rust pre/post-processing around the interpreter:
//   /root/reference/src/yolact.rs:52-88    terrible_id
//   /root/reference/src/yolact.rs:90-131   postprocess
//   /root/reference/src/yolact.rs:133-190  classify_tile (RGB unpack, dequantisation)
//   /root/reference/src/yolact.rs:192-234  classify (resize, tiling, stitch, resize back)
//   /root/reference/src/scene.rs:86,93     pixel packing in / target extraction out
// The Triangle resize is third-party: crate `image 0.24.1` (Cargo.lock:481-483),
// imageops::resize -> vertical_sample then horizontal_sample with an f32 RGBA
// intermediate and round-to-nearest on the way out.  Not under /root/reference:
// PARITY UNPINNED (restated from the crate's published algorithm).
//
// Pinning: the reference holds no test or golden vector for any of this; the
// known-answer vectors in tests/test_oracle_yolact.py are hand-derived from the
// cited lines.
//
// Build with -ffp-contract=off (Rust never fuses mul+add).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "tod_oracle.h"

extern "C" {

// yolact.rs:177  scale * (((x as i32) - zero_point) as f32)
void tod_oracle_dequant_u8(const uint8_t* q, int n, float scale, int zero_point, float* out) {
  for (int i = 0; i < n; ++i) out[i] = scale * static_cast<float>(static_cast<int32_t>(q[i]) - zero_point);
}

// yolact.rs:108-118
void tod_oracle_cell_classes(const float* seg, int cells, int channels, uint8_t* classes) {
  for (int c = 0; c < cells; ++c) {
    const float* chunk = seg + static_cast<size_t>(c) * channels;
    float mx = 0.0f;  // :109
    bool f[4];
    for (int i = 0; i < 4; ++i) {  // :110  *a > max && {max = *a; true}
      f[i] = chunk[i] > mx;
      if (f[i]) mx = chunk[i];
    }
    uint8_t cls;  // :112-117
    if (!f[0] && f[1] && !f[2] && !f[3]) cls = 1;
    else if (!f[0] && f[2] && !f[3]) cls = 2;
    else if (!f[0] && f[3]) cls = 3;
    else cls = 0;
    classes[c] = cls;
  }
}

// yolact.rs:52-88.  Release-build semantics: `px - 1` / `px - 28` wrap on usize underflow and
// Vec::get returns None for the wrapped index.
int tod_oracle_terrible_id(const uint8_t* classes, int mode, int8_t* ids) {
  const int N = 28 * 28;
  for (int i = 0; i < N; ++i) ids[i] = -1;  // :54
  if (mode == 0) {
    // Literal: flood_fill has no visited set (:60-77).  It terminates iff the popped cell has no
    // class-3 neighbour at flat index +-1 / +-28; then nothing is ever labelled.  If any class-3
    // cell has such a neighbour the `while let` loop ping-pongs forever -> report divergence.
    int diverges = 0;
    for (int px = 0; px < N && !diverges; ++px) {
      if (classes[px] != 3) continue;
      const int nb[4] = {px - 1, px + 1, px - 28, px + 28};
      for (int k = 0; k < 4; ++k)
        if (nb[k] >= 0 && nb[k] < N && classes[nb[k]] == 3) diverges = 1;
    }
    return diverges;  // ids all -1 either way
  }
  // Intent: proper 4-connected components of class-3 cells on the 28x28 grid, ids in raster
  // order of each component's first cell, stored in an i8 that wraps like `id += 1` would.
  int next = -1;
  std::vector<int> stack;
  for (int px = 0; px < N; ++px) {
    if (classes[px] != 3 || ids[px] != -1) continue;
    ++next;
    const int8_t id = static_cast<int8_t>(next & 0x7F);  // keep ids non-negative (-1 == none)
    ids[px] = id;
    stack.assign(1, px);
    while (!stack.empty()) {
      const int p = stack.back();
      stack.pop_back();
      const int x = p % 28, y = p / 28;
      const int cand[4] = {x > 0 ? p - 1 : -1, x < 27 ? p + 1 : -1, y > 0 ? p - 28 : -1, y < 27 ? p + 28 : -1};
      for (int k = 0; k < 4; ++k) {
        const int q = cand[k];
        if (q >= 0 && classes[q] == 3 && ids[q] == -1) { ids[q] = id; stack.push_back(q); }
      }
    }
  }
  return 0;
}

// yolact.rs:127-128
void tod_oracle_pack_upsample(const uint8_t* classes, const int8_t* ids, int mode, uint32_t* out) {
  for (int cy = 0; cy < 28; ++cy)
    for (int cx = 0; cx < 28; ++cx) {
      const uint32_t cls = classes[cy * 28 + cx];
      const uint32_t idu = static_cast<uint32_t>(static_cast<int32_t>(ids[cy * 28 + cx]));  // `id as u32` sign-extends
      uint32_t v;
      if (mode == 0) v = (cls << 24) & (idu << 16);  // literal `&`
      else v = (cls << 24) | ((idu & 0xFFu) << 16);   // intent
      for (int dy = 0; dy < 8; ++dy)
        for (int dx = 0; dx < 8; ++dx) out[(cy * 8 + dy) * 224 + cx * 8 + dx] = v;
    }
}

namespace {
inline float triangle_kernel(float x) {
  const float a = std::fabs(x);
  return a < 1.0f ? 1.0f - a : 0.0f;
}
inline int64_t clamp64(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct Taps {
  int left;
  std::vector<float> w;
};
// image 0.24.1 sample.rs: shared weight computation of vertical_sample / horizontal_sample
void make_taps(int in_size, int out_size, std::vector<Taps>* taps) {
  const float ratio = static_cast<float>(in_size) / static_cast<float>(out_size);
  const float sratio = ratio < 1.0f ? 1.0f : ratio;
  const float src_support = 1.0f * sratio;  // Triangle support = 1.0
  taps->resize(out_size);
  for (int o = 0; o < out_size; ++o) {
    float input = (static_cast<float>(o) + 0.5f) * ratio;
    int64_t left = static_cast<int64_t>(std::floor(input - src_support));
    left = clamp64(left, 0, static_cast<int64_t>(in_size) - 1);
    int64_t right = static_cast<int64_t>(std::ceil(input + src_support));
    right = clamp64(right, left + 1, static_cast<int64_t>(in_size));
    input = input - 0.5f;
    Taps& t = (*taps)[o];
    t.left = static_cast<int>(left);
    t.w.clear();
    float sum = 0.0f;
    for (int64_t i = left; i < right; ++i) {
      const float w = triangle_kernel((static_cast<float>(i) - input) / sratio);
      t.w.push_back(w);
      sum += w;
    }
    for (float& w : t.w) w /= sum;
  }
}
}  // namespace

void tod_oracle_resize_triangle_rgb8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  // vertical_sample: (sw x sh) u8 -> (sw x dh) f32
  std::vector<Taps> vt, ht;
  make_taps(sh, dh, &vt);
  make_taps(sw, dw, &ht);
  std::vector<float> tmp(static_cast<size_t>(sw) * dh * 3);
  for (int oy = 0; oy < dh; ++oy) {
    const Taps& t = vt[oy];
    for (int x = 0; x < sw; ++x) {
      float acc[3] = {0.0f, 0.0f, 0.0f};
      for (size_t i = 0; i < t.w.size(); ++i) {
        const uint8_t* p = src + (static_cast<size_t>(t.left + static_cast<int>(i)) * sw + x) * 3;
        for (int c = 0; c < 3; ++c) acc[c] += static_cast<float>(p[c]) * t.w[i];
      }
      for (int c = 0; c < 3; ++c) tmp[(static_cast<size_t>(oy) * sw + x) * 3 + c] = acc[c];
    }
  }
  // horizontal_sample: (sw x dh) f32 -> (dw x dh) u8, clamp to [0,255] then round (FloatNearest)
  for (int ox = 0; ox < dw; ++ox) {
    const Taps& t = ht[ox];
    for (int y = 0; y < dh; ++y) {
      float acc[3] = {0.0f, 0.0f, 0.0f};
      for (size_t i = 0; i < t.w.size(); ++i) {
        const float* p = &tmp[(static_cast<size_t>(y) * sw + t.left + static_cast<int>(i)) * 3];
        for (int c = 0; c < 3; ++c) acc[c] += p[c] * t.w[i];
      }
      for (int c = 0; c < 3; ++c) {
        float v = acc[c];
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        dst[(static_cast<size_t>(y) * dw + ox) * 3 + c] = static_cast<uint8_t>(std::round(v));
      }
    }
  }
}

// yolact.rs:195-214: u32 BE [r,g,b,_] -> RGB8 -> resize_exact(448,224) -> tiles (0,0) and (224,0)
void tod_oracle_classify_pre(const uint32_t* frame, int width, int height, uint8_t* tiles2) {
  std::vector<uint8_t> rgb(static_cast<size_t>(width) * height * 3);
  for (int i = 0; i < width * height; ++i) {  // :195-201 to_be_bytes()[..3]
    rgb[3 * i + 0] = static_cast<uint8_t>(frame[i] >> 24);
    rgb[3 * i + 1] = static_cast<uint8_t>(frame[i] >> 16);
    rgb[3 * i + 2] = static_cast<uint8_t>(frame[i] >> 8);
  }
  std::vector<uint8_t> canvas(448 * 224 * 3);
  tod_oracle_resize_triangle_rgb8(rgb.data(), width, height, canvas.data(), 448, 224);  // :207-208
  for (int t = 0; t < 2; ++t)                                                            // :213-214
    for (int y = 0; y < 224; ++y)
      std::memcpy(tiles2 + (static_cast<size_t>(t) * 224 * 224 + static_cast<size_t>(y) * 224) * 3,
                  canvas.data() + (static_cast<size_t>(y) * 448 + t * 224) * 3, 224 * 3);
}

// yolact.rs:219-233: stitch rows, u32 -> RGB bytes, resize_exact(width,height), repack [r,g,b,0]
void tod_oracle_classify_post(const uint32_t* t1, const uint32_t* t2, int width, int height, uint32_t* frame) {
  std::vector<uint8_t> canvas(448 * 224 * 3);
  for (int y = 0; y < 224; ++y)
    for (int x = 0; x < 448; ++x) {
      const uint32_t px = x < 224 ? t1[y * 224 + x] : t2[y * 224 + x - 224];
      uint8_t* d = &canvas[(static_cast<size_t>(y) * 448 + x) * 3];
      d[0] = static_cast<uint8_t>(px >> 24);
      d[1] = static_cast<uint8_t>(px >> 16);
      d[2] = static_cast<uint8_t>(px >> 8);
    }
  std::vector<uint8_t> out(static_cast<size_t>(width) * height * 3);
  tod_oracle_resize_triangle_rgb8(canvas.data(), 448, 224, out.data(), width, height);
  for (int i = 0; i < width * height; ++i)
    frame[i] = (static_cast<uint32_t>(out[3 * i]) << 24) | (static_cast<uint32_t>(out[3 * i + 1]) << 16) |
               (static_cast<uint32_t>(out[3 * i + 2]) << 8);
}

// scene.rs:93  ((px << 16) >> 16) as u16
void tod_oracle_target_from_frame(const uint32_t* frame, int n, uint16_t* target) {
  for (int i = 0; i < n; ++i) target[i] = static_cast<uint16_t>((frame[i] << 16) >> 16);
}

}  // extern "C"
