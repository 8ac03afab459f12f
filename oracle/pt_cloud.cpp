// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// CPU restatement of the reference's two compute shaders and the host half of
// append_scene:
//   /root/reference/shaders/pt_cloud.comp:23-39   constants
//   /root/reference/shaders/pt_cloud.comp:45-76   bump_img_store
//   /root/reference/shaders/pt_cloud.comp:78-82   store_ball
//   /root/reference/shaders/pt_cloud.comp:84-124  main
//   /root/reference/shaders/pt_cloud_weights.comp:25-47  pack / unpack / dist
//   /root/reference/shaders/pt_cloud_weights.comp:49-123 main (3 stages)
//   /root/reference/src/scene.rs:152-155,197-198,211  image formats / uploads
//   /root/reference/src/scene.rs:312-327               readback -> Scene
// Under-specified GLSL behaviour is pinned by the deterministic rules of
// SURVEY.md §9 (each rule is quoted where it is applied).
//
// Pinning: the only reference fixture is the depth.bmp/image.bmp/map.bmp debug trio
// (approximate, see tests/test_oracle_ptcloud.py::test_bmp_trio).  The weights shader
// has no fixture: PARITY UNPINNED for pt_cloud_weights.
//
// Build with -ffp-contract=off: GLSL/Rust never fuse a*b+c here and neither may we.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "tod_oracle.h"

namespace {

inline uint32_t float_to_uint_rz(float f) {
  // NaN y_add => stored value 0 (== __float2uint_rz semantics).
  if (!(f == f)) return 0u;
  if (f <= 0.0f) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return static_cast<uint32_t>(f);
}

struct BumpMemo {
  // y_add depends only on (val, bump_size, dx, dy) -> memoise per (val, size) slab.
  // Pure function of its key, so memoisation cannot change a result.
  std::vector<std::vector<uint32_t>> slabs;  // index = val_index
  std::vector<uint8_t> have;
};

// pt_cloud.comp:61-72 for one (val, bump_size): uint(y_add) for every (x,y) in [0,2s)^2
void bump_slab(float val, int s, float bump_err, uint32_t* out) {
  const float C_1 = val / bump_err - 1.0f;          // :61
  const float C_2 = 2.0f / static_cast<float>(s);   // :62
  for (int x = 0; x < 2 * s; ++x) {
    for (int y = 0; y < 2 * s; ++y) {
      // loc = pos - s + (x,y)  =>  pos - loc = (s - x, s - y)
      const float dx = static_cast<float>(s - x);
      const float dy = static_cast<float>(s - y);
      // :69  pow(a,2) := a*a
      const float prox = std::sqrt(dx * dx + dy * dy);
      const float y_add = val / (1.0f + std::pow(C_1, C_2 * prox - 1.0f));  // :70
      out[x * 2 * s + y] = float_to_uint_rz(y_add);                          // :72 uint(y_add)
    }
  }
}

}  // namespace

extern "C" {

void tod_oracle_scene_default_params(tod_oracle_scene_params* p) {
  p->width = 640;                 // pt_cloud.comp:23
  p->height = 480;                // :24
  p->max_depth_in = 4000.0f;      // :25
  p->y_fov = 1.01229096616f;      // :27
  p->x_fov = 1.51843644924f;      // :28
  p->bot_avoidance_const = 100.0f;  // :32
  p->bot_norm_const = 20;         // :36
  p->terrain_norm_const = 10;     // :37
  p->bump_err = 0.1f;             // :39
  p->sample_shift = 0;            // SURVEY §9.4 default (texelFetch intent)
  p->weights_mode = 0;            // literal
}

// Expose the stamp table so tests can check the product's host-built LUT.
void tod_oracle_bump_table(float val, int s, float bump_err, uint32_t* out) { bump_slab(val, s, bump_err, out); }

// pt_cloud.comp main + bump_img_store + store_ball for one frame.
// depth: u16[W*H] (R16_UINT, scene.rs:197); target: u16[W*H] uploaded as R8G8 little-endian
// (scene.rs:198) => R = low byte = class, G = high byte = id.
void tod_oracle_pt_cloud(const tod_oracle_scene_params* P, const uint16_t* depth, const uint16_t* target,
                         uint32_t* map, float* balls4) {
  const int W = P->width, H = P->height;
  // map starts at 0 every frame.
  std::memset(map, 0, sizeof(uint32_t) * static_cast<size_t>(W) * H);
  // SURVEY §9.3 deterministic store_ball: integer sums per id < 100.
  long long sx[100] = {0}, sy[100] = {0}, sn[100] = {0};

  const int ts = P->terrain_norm_const, bs = P->bot_norm_const;
  std::vector<std::vector<uint32_t>> terrain_slab(H);
  std::vector<uint32_t> bot_slab(static_cast<size_t>(4) * bs * bs);
  bump_slab(P->bot_avoidance_const, bs, P->bump_err, bot_slab.data());

  const float ty = std::tan(P->y_fov / 2.0f), tx = std::tan(P->x_fov / 2.0f);
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      // :88-91 nearest texture() at texel edges; SURVEY §9.4: texel (x-shift, y-shift) mod (W,H)
      const int sxp = ((x - P->sample_shift) % W + W) % W;
      const int syp = ((y - P->sample_shift) % H + H) % H;
      const uint32_t depth_read = depth[syp * W + sxp];
      const uint32_t cls_read = target[syp * W + sxp] & 0xFFu;
      const uint32_t id_read = target[syp * W + sxp] >> 8;
      // :93-95 (left-to-right evaluation, x / y NOT centred)
      const float cy = std::cos(std::atan(ty * static_cast<float>(y) * 2.0f / static_cast<float>(H)));
      const float cx = std::cos(std::atan(tx * static_cast<float>(x) * 2.0f / static_cast<float>(W)));
      const float d = static_cast<float>(depth_read) * cy * cx;
      // :98
      const int dz = static_cast<int>(static_cast<float>(H) * d / P->max_depth_in);
      // :108-111
      int action = static_cast<int>(cls_read);
      if (action > 1) action -= 1;
      // :114
      const int px = x, py = H - dz;
      if (action == 2) {
        // :118-120 store_ball(id, new_pos)
        if (id_read < 100) { sx[id_read] += px; sy[id_read] += py; sn[id_read] += 1; }
        continue;
      }
      const uint32_t* slab;
      int s;
      if (action == 0) {  // :116-117 terrain: val = float(img_pos.y)
        s = ts;
        if (terrain_slab[y].empty()) {
          terrain_slab[y].resize(static_cast<size_t>(4) * s * s);
          bump_slab(static_cast<float>(y), s, P->bump_err, terrain_slab[y].data());
        }
        slab = terrain_slab[y].data();
      } else {  // :121-122 robot
        s = bs;
        slab = bot_slab.data();
      }
      // :59-75
      for (int ox = 0; ox < 2 * s; ++ox) {
        const int lx = px - s + ox;
        if (!(lx > 0 && lx < W - 1)) continue;
        for (int oy = 0; oy < 2 * s; ++oy) {
          const int ly = py - s + oy;
          if (!(ly > 0 && ly < H - 1)) continue;  // :67
          uint32_t& m = map[ly * W + lx];
          const uint32_t v = slab[ox * 2 * s + oy];
          if (v > m) m = v;  // :72 imageAtomicMax
        }
      }
    }
  }
  for (int i = 0; i < 100; ++i) {
    float* b = balls4 + 4 * i;
    if (sn[i] == 0) { b[0] = b[1] = b[2] = b[3] = 0.0f; continue; }  // scene.rs:211 zero init
    b[0] = static_cast<float>(static_cast<double>(sx[i]) / static_cast<double>(sn[i]));
    b[1] = static_cast<float>(static_cast<double>(sy[i]) / static_cast<double>(sn[i]));
    b[2] = static_cast<float>(sn[i]);
    b[3] = 0.0f;
  }
}

static inline float dist3(const float* a, const float* b) {
  // pt_cloud_weights.comp:42-46 with pow(a,2) := a*a
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}

// pt_cloud_weights.comp main as three globally ordered passes.
// mode 0 = literal (pack() == 0 => every present neighbour decodes to world[0,0], §9.5)
// mode 1 = intent  (true neighbour distances, §9.5)
void tod_oracle_pt_cloud_weights(const tod_oracle_scene_params* P, const uint32_t* map, float* world4,
                                 float* conn0, float* conn1) {
  const int W = P->width, H = P->height;
  const bool literal = P->weights_mode == 0;
  // pass 1: world (+ stage 1 conn0 = pack(x,y) == 0, overwritten in stage 3)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float* w = world4 + 4 * (static_cast<size_t>(y) * W + x);
      w[0] = static_cast<float>(x);                 // :59
      w[1] = static_cast<float>(map[y * W + x]);    // :51
      w[2] = static_cast<float>(y);
      w[3] = 0.0f;                                  // :69
    }
  auto nb = [&](int x, int y) -> const float* {
    return literal ? world4 : world4 + 4 * (static_cast<size_t>(y) * W + x);
  };
  // pass 2: conn1 = (below, below-left, left, above-left)   :86-107
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const bool nxmin = x > 0, nymin = y > 0, nymax = y < H - 1;
      const float* w = world4 + 4 * (static_cast<size_t>(y) * W + x);
      float* c = conn1 + 4 * (static_cast<size_t>(y) * W + x);
      c[0] = nymax ? dist3(w, nb(x, y + 1)) : -1.0f;
      c[1] = (nxmin && nymax) ? dist3(w, nb(x - 1, y + 1)) : -1.0f;
      c[2] = nxmin ? dist3(w, nb(x - 1, y)) : -1.0f;
      c[3] = (nxmin && nymin) ? dist3(w, nb(x - 1, y - 1)) : -1.0f;
    }
  // pass 3: conn0 = (conn1[up].r, conn1[up-right].g, conn1[right].b, conn1[down-right].a)  :112-122
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const bool nxmax = x < W - 1, nymin = y > 0, nymax = y < H - 1;
      float* c = conn0 + 4 * (static_cast<size_t>(y) * W + x);
      auto c1 = [&](int xx, int yy, int k) { return conn1[4 * (static_cast<size_t>(yy) * W + xx) + k]; };
      c[0] = nymin ? c1(x, y - 1, 0) : -1.0f;
      c[1] = (nxmax && nymin) ? c1(x + 1, y - 1, 1) : -1.0f;
      c[2] = nxmax ? c1(x + 1, y, 2) : -1.0f;
      c[3] = (nxmax && nymax) ? c1(x + 1, y + 1, 3) : -1.0f;
    }
}

static inline int32_t rust_f32_as_i32(float f) {
  // Rust `as i32`: saturating, NaN -> 0
  if (!(f == f)) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return static_cast<int32_t>(-2147483647 - 1);
  return static_cast<int32_t>(f);
}

// scene.rs:312-327: Scene { height, pos, balls, connections }
void tod_oracle_scene_materialize(int npx, const uint32_t* map, const float* world4, const float* conn0,
                                  const float* conn1, const float* balls4, float* height, float* pos3,
                                  int32_t* balls2, float* connections8) {
  for (int i = 0; i < npx; ++i) {
    height[i] = static_cast<float>(map[i]);                       // :312-314
    pos3[3 * i + 0] = world4[4 * i + 0];                          // :316-318
    pos3[3 * i + 1] = world4[4 * i + 1];
    pos3[3 * i + 2] = world4[4 * i + 2];
    for (int k = 0; k < 4; ++k) {                                 // :324-327
      connections8[8 * i + k] = conn0[4 * i + k];
      connections8[8 * i + 4 + k] = conn1[4 * i + k];
    }
  }
  for (int i = 0; i < 100; ++i) {                                 // :320-322
    balls2[2 * i + 0] = rust_f32_as_i32(balls4[4 * i + 0]);
    balls2[2 * i + 1] = rust_f32_as_i32(balls4[4 * i + 1]);
  }
}

}  // extern "C"
