// tod.hpp — C++ mirror of the reference's Rust interface for the hot path, over the C ABI of tod.h.
//
// The reference is a Rust crate; no Rust toolchain exists in this repository's build image, so the host side above the
// C ABI is written in C++ with the reference's names, argument meaning and error behaviour (the reference panics on every
// failure: here every failure throws tod::Error):
//
//   tod::Yolact::init()                <- Yolact::init()                      src/yolact.rs:17-37
//   tod::Yolact::classify(frame)       <- Yolact::classify(&mut [u32])        src/yolact.rs:39-41
//   tod::append_scene(queues, scene)   <- append_scene(...)                   src/scene.rs:147-331
//   tod::Scene                         <- struct Scene                        src/scene.rs:122-143
//   tod::YolactPool                    (no counterpart: several batches in flight over the same C ABI)
#pragma once
#include <array>
#include <cstdint>
#include <future>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "tod.h"

namespace tod {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what + ": " + tod_last_error()), code(c) {}
};
inline int expect(int rc, const char* what) {
  if (rc < 0) throw Error(rc, what);
  return rc;
}

class Yolact {
 public:
  static Yolact init(const char* model = "data/FRC_model.tflite", int device = 0, int max_tiles = 2) {
    tod_yolact_options o;
    tod_yolact_default_options(&o);
    o.max_tiles = max_tiles;
    Yolact y;
    expect(tod_yolact_create(model, device, &o, &y.h_), "Yolact::init");
    return y;
  }
  // in place on 640*480 pixels r<<24|g<<16|b<<8 (scene.rs:86); returns true where the reference itself would hang
  bool classify(std::vector<uint32_t>& frame_buffer) {
    if (frame_buffer.size() != 640u * 480u) throw Error(TOD_ERR_INVALID_ARG, "classify: frame buffer must hold 640x480 pixels");
    return expect(tod_yolact_classify(h_, frame_buffer.data(), 640, 480), "Yolact::classify") == TOD_WARN_REFERENCE_DIVERGES;
  }
  Yolact(Yolact&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  Yolact& operator=(Yolact&& o) noexcept {
    std::swap(h_, o.h_);
    return *this;
  }
  Yolact(const Yolact&) = delete;
  ~Yolact() { tod_yolact_destroy(h_); }
  tod_yolact* handle() { return h_; }

 private:
  Yolact() = default;
  tod_yolact* h_ = nullptr;
};

// Several batches in flight (the frame loop's double / triple buffering; the Python mirror is tod_b200.YolactPool): `depth`
// handles take alternate batches, each call on its own host thread, so one batch's copies and latency-bound tail overlap the
// next batch's backbone.  Every batch runs through one handle exactly as tod_yolact_infer_tiles would run it alone.
class YolactPool {
 public:
  YolactPool(const char* model, int device, int max_tiles, int depth = 3) {
    tod_yolact_options o;
    tod_yolact_default_options(&o);
    o.max_tiles = max_tiles;
    o.batches_in_flight = depth < 1 ? 1 : depth;  // the handles share the GPU: convolution launches sized for throughput
    for (int i = 0; i < (depth < 1 ? 1 : depth); ++i) {
      slots_.emplace_back(new Slot());
      expect(tod_yolact_create(model, device, &o, &slots_.back()->h), "YolactPool");
    }
  }
  YolactPool(const YolactPool&) = delete;
  ~YolactPool() {
    for (auto& s : slots_) {
      std::lock_guard<std::mutex> g(s->busy);  // wait for the batch in flight
      tod_yolact_destroy(s->h);
    }
  }
  // queues one batch of u8[n][th][tw][3] tiles; the future yields tod_yolact_infer_tiles' return code (throws tod::Error on failure)
  std::future<int> submit(const uint8_t* rgb_tiles, int n, uint32_t* tile_classes, tod_detections* dets) {
    Slot* s = slots_[next_++ % slots_.size()].get();
    return std::async(std::launch::async, [=] {
      std::lock_guard<std::mutex> g(s->busy);  // a handle runs one batch at a time
      return expect(tod_yolact_infer_tiles(s->h, rgb_tiles, n, nullptr, tile_classes, dets), "YolactPool::submit");
    });
  }
  size_t depth() const { return slots_.size(); }

 private:
  struct Slot {
    tod_yolact* h = nullptr;
    std::mutex busy;
  };
  std::vector<std::unique_ptr<Slot>> slots_;
  size_t next_ = 0;
};

struct Scene {  // scene.rs:122-132
  std::vector<float> height;
  std::vector<std::tuple<float, float, float>> pos;
  std::vector<std::pair<int32_t, int32_t>> balls;
  std::vector<std::array<float, 8>> connections;

  std::vector<size_t> neighbors(size_t px) const {  // scene.rs:134-143, typo included
    std::vector<size_t> out;
    if (px > 0) out.push_back(px - 1);
    if (px < 680 * 480 - 1) out.push_back(px + 1);
    if (px / 640 > 0) out.push_back(px - 640);
    if (px / 640 < 480 - 1) out.push_back(px + 640);
    return out;
  }
};

class SceneGpu {  // owns what scene.rs:152-224 re-creates per call
 public:
  explicit SceneGpu(int device = 0) {
    tod_scene_params p;
    tod_scene_default_params(&p);
    expect(tod_scene_create(device, &p, &h_), "SceneGpu");
  }
  SceneGpu(const SceneGpu&) = delete;
  ~SceneGpu() { tod_scene_destroy(h_); }
  tod_scene* handle() { return h_; }

 private:
  tod_scene* h_ = nullptr;
};

using FrameQueue = std::vector<std::array<uint16_t, 640 * 480>>;

// append_scene (scene.rs:147): pops the newest depth / target frames (scene.rs:186-187), runs both kernels, blocks
// (scene.rs:282) and overwrites *scene (scene.rs:329-330).
inline void append_scene(FrameQueue& point_cloud_queue, FrameQueue& target_buffer_queue, Scene& scene, SceneGpu& gpu) {
  if (point_cloud_queue.empty() || target_buffer_queue.empty()) throw Error(TOD_ERR_INVALID_ARG, "append_scene: empty queue");
  const auto depth = point_cloud_queue.back();
  point_cloud_queue.pop_back();
  const auto target = target_buffer_queue.back();
  target_buffer_queue.pop_back();
  expect(tod_scene_append_batch(gpu.handle(), depth.data(), target.data(), 1, nullptr, nullptr, nullptr, nullptr, nullptr), "append_scene");
  const size_t n = 640 * 480;
  scene.height.assign(n, 0.f);
  scene.pos.assign(n, {});
  scene.balls.assign(100, {});
  scene.connections.assign(n, {});
  static_assert(sizeof(std::tuple<float, float, float>) == 12 && sizeof(std::pair<int32_t, int32_t>) == 8, "packed layout expected");
  std::vector<float> pos3(n * 3);
  std::vector<int32_t> balls2(200);
  expect(tod_scene_materialize(gpu.handle(), 0, scene.height.data(), pos3.data(), balls2.data(), &scene.connections[0][0]), "append_scene");
  for (size_t i = 0; i < n; ++i) scene.pos[i] = std::make_tuple(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
  for (size_t i = 0; i < 100; ++i) scene.balls[i] = {balls2[2 * i], balls2[2 * i + 1]};
}

}  // namespace tod
