// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// CPU restatement of the path planner that consumes the Scene:
//   /root/reference/src/path.rs:25-120   modify_path  (relaxation from the ball targets + path extraction)
//   /root/reference/src/path.rs:17-21    Path::serialize (big-endian seconds + (magnitude, rotation) pairs)
//   /root/reference/src/scene.rs:134-143 Scene::neighbors (4-neighbourhood by flat index)
//
// LITERAL behaviour: the reference function cannot complete.  It is never reached (main.rs:92 panics first); if it were,
// `path` / `cost` are [_; 224*224] arrays indexed with 640x480 node numbers (path.rs:29-30,38), START_NODE = 640*480-240
// (path.rs:99) is far outside them, and the extraction loop dereferences cost[usize::MAX - 2] on its last step
// (path.rs:105 with path[node] == target marker): an index panic on every input.  tod_oracle_path_literal_panics()
// states exactly that; there are no literal outputs to be bit-exact with.  PARITY UNPINNED (no fixture, no runnable code).
//
// INTENT mode (what the code evidently means; every choice is listed in DESIGN.md §5.5):
//   * nodes = all W*H pixels, arrays sized W*H; neighbours as Scene::neighbors with its 680 typo read as W
//     (px-1 if px>0, px+1 if px<W*H-1, px-W if row>0, px+W if row<H-1 - left / right still wrap across row ends);
//     cn = position in that list (path.rs:60 `enumerate`), and the edge weight is literally
//     connections[node][cn] + |height[node] - height[neighbor]| evaluated left to right in f32 (path.rs:64);
//     an edge whose connection entry is negative (the shaders' -1 "no neighbour" marker) does not exist;
//   * targets = the first three balls, node = x + y*W (the reference multiplies by 480, its height), cost 0;
//   * cost[] = the least fixed point of cost[n] = min(cost[n], min_cn (cost[nb] + conn) + |dh|), which is what the
//     LIFO relaxation of path.rs:57-96 converges to when it terminates (its visiting order is an artefact of Vec::pop);
//     path[n] = the first neighbour in list order that attains the minimum;
//   * extraction from START_NODE = W*H - H/2 (= 640*480 - 240) exactly as path.rs:100-117, stopping when the next
//     node is a target (the last magnitude is then cost[node] - 0); an unreachable start yields an empty path.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>
#include "tod_oracle.h"

namespace {
inline int neighbours(int64_t px, int W, int H, int64_t out[4]) {  // scene.rs:134-143
  int n = 0;
  if (px > 0) out[n++] = px - 1;
  if (px < int64_t(W) * H - 1) out[n++] = px + 1;
  if (px / W > 0) out[n++] = px - W;
  if (px / W < H - 1) out[n++] = px + W;
  return n;
}
}  // namespace

extern "C" {

int tod_oracle_path_literal_panics(void) { return 1; }

// cost f32[W*H] (FLT_MAX = unreached), pred i32[W*H] (-1 = undefined, -2 = target); returns the number of directions
// written to dirs[2*cap] (magnitude, rotation), or -1 if the start cannot reach a target.
int tod_oracle_path_modify(int W, int H, const float* height, const float* pos3, const int32_t* balls2, const float* conn8,
                           float* cost, int32_t* pred, float* dirs, int cap) {
  const int64_t N = int64_t(W) * H;
  const float kInf = 3.402823466e+38f;  // f32::MAX (path.rs:30)
  for (int64_t i = 0; i < N; ++i) {
    cost[i] = kInf;
    pred[i] = -1;
  }
  using Item = std::pair<float, int64_t>;
  std::priority_queue<Item, std::vector<Item>, std::greater<Item>> pq;
  for (int b = 0; b < 3; ++b) {
    const int64_t x = balls2[2 * b], y = balls2[2 * b + 1];
    if (x < 0 || x >= W || y < 0 || y >= H) continue;
    const int64_t t = x + y * W;
    cost[t] = 0.0f;
    pred[t] = -2;
    pq.push({0.0f, t});
  }
  // Dijkstra over the reversed edges: settling nb improves every node n that lists nb as a neighbour
  while (!pq.empty()) {
    const Item it = pq.top();
    pq.pop();
    const int64_t nb = it.second;
    if (it.first > cost[nb]) continue;
    int64_t around[4];
    const int na = neighbours(nb, W, H, around);   // the relation is symmetric except at the two ends of the array
    for (int a = 0; a < na; ++a) {
      const int64_t n = around[a];
      if (pred[n] == -2) continue;
      int64_t mine[4];
      const int nm = neighbours(n, W, H, mine);
      for (int cn = 0; cn < nm; ++cn) {
        if (mine[cn] != nb) continue;
        const float c = conn8[8 * n + cn];
        if (c < 0.0f) continue;
        const float cand = (cost[nb] + c) + std::fabs(height[n] - height[nb]);   // path.rs:64, left to right
        if (cand < cost[n]) {
          cost[n] = cand;
          pq.push({cand, n});
        }
      }
    }
  }
  for (int64_t n = 0; n < N; ++n) {   // predecessor: first neighbour in list order attaining the fixed point
    if (pred[n] == -2 || cost[n] == kInf) continue;
    int64_t mine[4];
    const int nm = neighbours(n, W, H, mine);
    for (int cn = 0; cn < nm; ++cn) {
      const float c = conn8[8 * n + cn];
      if (c < 0.0f || cost[mine[cn]] == kInf) continue;
      if ((cost[mine[cn]] + c) + std::fabs(height[n] - height[mine[cn]]) == cost[n]) {
        pred[n] = int32_t(mine[cn]);
        break;
      }
    }
  }
  // extraction (path.rs:99-117)
  int64_t node = N - H / 2;
  if (cost[node] == kInf) return -1;
  int nd = 0;
  float rotation = 0.0f;
  while (pred[node] != -2) {
    const int64_t next = pred[node];
    if (next < 0) return -1;
    if (nd < cap) {
      dirs[2 * nd] = cost[node] - cost[next];
      dirs[2 * nd + 1] = rotation;
    }
    ++nd;
    const int64_t last = node;
    node = next;
    if (pred[node] == -2) break;
    const int64_t after = pred[node];
    const float ax = pos3[3 * last] - pos3[3 * node], ay = pos3[3 * last + 1] - pos3[3 * node + 1];
    const float bx = pos3[3 * after] - pos3[3 * node], by = pos3[3 * after + 1] - pos3[3 * node + 1];
    rotation = std::acos((ax * bx + ay * by) / (std::sqrt(ax * ax + ay * ay) * std::sqrt(bx * bx + by * by)));
  }
  return nd;
}

// path.rs:17-21: created.as_secs().to_be_bytes() ++ for each (m, r): m.to_be_bytes() ++ r.to_be_bytes()
int tod_oracle_path_serialize(uint64_t created_secs, const float* dirs, int n, uint8_t* out) {
  for (int i = 0; i < 8; ++i) out[i] = uint8_t(created_secs >> (8 * (7 - i)));
  for (int i = 0; i < 2 * n; ++i) {
    uint32_t u;
    std::memcpy(&u, dirs + i, 4);
    for (int k = 0; k < 4; ++k) out[8 + 4 * i + k] = uint8_t(u >> (8 * (3 - k)));
  }
  return 8 + 8 * n;
}

}  // extern "C"
