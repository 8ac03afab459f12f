// A C++ stand-in for the reference's frame loop (src/main.rs:78-96 + src/scene.rs:77-119): synthetic RGB-D frames go
// through Yolact::classify -> target extraction -> append_scene exactly as main.rs drives them, using only include/tod.hpp.
//   g++ -std=c++17 -O2 -Iinclude tools/frame_loop.cpp -Ltiny-object-detection_b200/lib -ltod_b200 -Wl,-rpath,... -o frame_loop
//   ./frame_loop <model.tflite> [frames]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>

#include "tod.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s model.tflite [frames]\n", argv[0]);
    return 2;
  }
  const int frames = argc > 2 ? std::atoi(argv[2]) : 4;
  try {
    tod::Yolact yolact = tod::Yolact::init(argv[1]);
    tod::SceneGpu gpu;
    tod::Scene scene;
    auto depth_queue = std::make_unique<tod::FrameQueue>();
    auto target_queue = std::make_unique<tod::FrameQueue>();
    std::mt19937 rng(7);
    double checksum = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) {
      std::vector<uint32_t> buffer(640 * 480);
      for (auto& px : buffer) px = (rng() & 0xFFFFFF00u);  // scene.rs:86 r<<24|g<<16|b<<8
      yolact.classify(buffer);                              // scene.rs:92
      depth_queue->emplace_back();
      target_queue->emplace_back();
      for (size_t i = 0; i < buffer.size(); ++i) {
        target_queue->back()[i] = uint16_t((buffer[i] << 16) >> 16);  // scene.rs:93
        depth_queue->back()[i] = uint16_t(400 + (rng() % 3600));      // scene.rs:96-97 (camera depth)
      }
      tod::append_scene(*depth_queue, *target_queue, scene, gpu);     // main.rs:85-90
      for (int i = 0; i < 640 * 480; i += 997) checksum += scene.height[i] + scene.connections[i][4];
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("frame_loop: %d frames, %.2f fps, checksum %.3f, scene.height.size()=%zu\n", frames, frames / s, checksum, scene.height.size());
  } catch (const tod::Error& e) {
    std::fprintf(stderr, "frame_loop failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
