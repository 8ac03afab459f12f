// Scene path: depth -> bird's-eye height map ("point cloud") -> world positions + 8-neighbour
// connection weights.  B200-native replacement of the two Vulkan compute shaders the reference
// dispatches from append_scene (/root/reference/src/scene.rs:238-260):
//   shaders/pt_cloud.comp          -> land_kernel + stamp_kernel + balls_kernel
//   shaders/pt_cloud_weights.comp  -> weights_kernel
//
// Design (not a port of the shaders):
//  * The shader issues 400 (terrain) / 1600 (robot) imageAtomicMax per source pixel.  A stamp never
//    changes a pixel's column (new_pos.x = img_pos.x, pt_cloud.comp:114), so the map is cut into
//    32-column strips; one warp owns a strip, keeps it in shared memory as u16 and is the only
//    writer of those columns: lane l owns column l.  The stamps become plain shared-memory
//    read-max-write with no atomics, no races, and `map` is written to HBM exactly once, coalesced.
//  * Every transcendental of the shader (pow/sqrt of the sigmoid bump, cos(atan(tan))) depends only
//    on small integers (row, stamp offset, column), so it is tabulated once on the host at handle
//    creation; the device arithmetic is integer + IEEE mul/div and the map is bit-exact by
//    construction.
//  * Under-specified GLSL behaviour follows the deterministic rules of SURVEY.md §9
//    (zero-initialised map, integer ball sums, three ordered passes for the weights).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"

namespace tod {
namespace {

constexpr int kStripW = 32;
constexpr int kKindTerrain = 0, kKindRobot = 1, kKindBall = 2, kKindNone = 3;
constexpr int kMaxBalls = 100;  // pt_cloud.comp:17

struct SceneDev {
  int W, H, s_t, s_b, py_bias, pad_rows;
  float max_depth;
  int sample_shift, weights_mode;
};

// ------------------------------------------------------------------ land: where does each pixel stamp?
// pt_cloud.comp:84-114.  One warp = one source row of one 32-column strip.  Output: land u16 = kind<<14 | (py + py_bias),
// ball sums, a per-row "row holds a robot pixel" flag (the single-warp stamp kernels skip the 40x40 pass on robot-free
// rows) and, per (strip, source row), the range of encoded landing rows of its terrain / robot pixels (lo | hi<<16,
// lo > hi when there is none): the pruned stamp kernel only visits the source rows that can land in its band.
constexpr int kLandRows = 8;   // source rows per warp: sixteen independent loads in flight per thread

__global__ void __launch_bounds__(256) land_kernel(const uint16_t* __restrict__ depth,
                                                  const uint16_t* __restrict__ target, const float* __restrict__ cy_tab,
                                                  const float* __restrict__ cx_tab, SceneDev P,
                                                  uint16_t* __restrict__ land, unsigned int* __restrict__ row_robot,
                                                  unsigned long long* __restrict__ ball_sums,
                                                  uint32_t* __restrict__ rowinfo_t, uint32_t* __restrict__ rowinfo_b) {
  const int lane = threadIdx.x;
  const int x = blockIdx.x * kStripW + lane;
  const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * kLandRows;
  const int f = blockIdx.z;
  if (y0 >= P.H) return;  // whole warp
  const bool valid = x < P.W;
  const size_t fbase = size_t(f) * P.W * P.H;
  // nearest texture() fetch; SURVEY §9.4: texel (x - shift, y - shift) with Repeat addressing
  int sx = x - P.sample_shift;
  if (sx < 0) sx += P.W;
  unsigned dv[kLandRows], tv[kLandRows];
#pragma unroll
  for (int r = 0; r < kLandRows; ++r) {
    int sy = y0 + r - P.sample_shift;
    if (sy < 0) sy += P.H;
    const bool ok = valid && y0 + r < P.H;
    const size_t src = fbase + size_t(sy) * P.W + sx;
    dv[r] = ok ? unsigned(__ldg(depth + src)) : 0u;
    tv[r] = ok ? unsigned(__ldg(target + src)) : 0u;
  }
  const float cxv = valid ? cx_tab[x] : 0.f;
  const float fh = float(P.H);
  const unsigned full = 0xffffffffu;
  const size_t ri0 = (size_t(f) * gridDim.x + blockIdx.x) * P.H;
  uint16_t* lrow = land + fbase + size_t(y0) * P.W + x;
#pragma unroll
  for (int r = 0; r < kLandRows; ++r) {
    const int y = y0 + r;
    const bool live = valid && y < P.H;   // y < P.H is warp-uniform; no early exit: the loop body is straight-line code
    const int cls = tv[r] & 0xFF, id = tv[r] >> 8;  // R8G8 little-endian upload (scene.rs:198)
    // :93-95, evaluated left to right; the two cos(atan(tan)) factors come from the host tables
    const float de = __fmul_rn(__fmul_rn(float(dv[r]), cy_tab[min(y, P.H - 1)]), cxv);
    // :98  int(float(height) * depth / max_depth_in)
    const int dz = int(__fdiv_rn(__fmul_rn(fh, de), P.max_depth));
    const int py = P.H - dz;  // :114
    const int action = cls > 1 ? cls - 1 : cls;  // :108-111
    int kind = action == 0 ? kKindTerrain : (action == 2 ? kKindBall : kKindRobot);
    const int s = kind == kKindTerrain ? P.s_t : P.s_b;
    // rows touched: [py - s, py + s - 1]; only rows 1..H-2 are ever stored (pt_cloud.comp:67)
    if (kind != kKindBall && (py + s - 1 < 1 || py - s > P.H - 2)) kind = kKindNone;
    if (!live) kind = kKindNone;
    if (live && kind == kKindBall && id < kMaxBalls) {  // SURVEY §9.3 deterministic store_ball: integer sums (rare)
      unsigned long long* b = ball_sums + (size_t(f) * kMaxBalls + id) * 3;
      atomicAdd(b + 0, (unsigned long long)(long long)x);
      atomicAdd(b + 1, (unsigned long long)(long long)py);
      atomicAdd(b + 2, 1ull);
    }
    const int enc = max(0, min(py + P.py_bias, 0x3FFF));
    if (live) lrow[size_t(r) * P.W] = uint16_t((kind << 14) | enc);
    const unsigned t_lo = __reduce_min_sync(full, kind == kKindTerrain ? unsigned(enc) : 0xFFFFu);
    const unsigned t_hi = __reduce_max_sync(full, kind == kKindTerrain ? unsigned(enc) : 0u);
    const unsigned b_lo = __reduce_min_sync(full, kind == kKindRobot ? unsigned(enc) : 0xFFFFu);
    const unsigned b_hi = __reduce_max_sync(full, kind == kKindRobot ? unsigned(enc) : 0u);
    if (lane == 0 && y < P.H) {
      rowinfo_t[ri0 + y] = t_lo | (t_hi << 16);
      rowinfo_b[ri0 + y] = b_lo | (b_hi << 16);
      if (b_lo <= b_hi) atomicOr(row_robot + size_t(f) * P.H + y, 1u);
    }
  }
}

__global__ void balls_kernel(const unsigned long long* __restrict__ sums, float* __restrict__ balls4, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long sx = (long long)sums[3 * i + 0], sy = (long long)sums[3 * i + 1], sn = (long long)sums[3 * i + 2];
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sn > 0) {
    o.x = float(double(sx) / double(sn));
    o.y = float(double(sy) / double(sn));
    o.z = float(sn);
  }
  reinterpret_cast<float4*>(balls4)[i] = o;
}

// ------------------------------------------------------------------ stamp
// grid = (strips, frames), block = one warp.  tile rows are padded by pad_rows above and below so a
// stamp that hangs over the top/bottom edge needs no per-row test; the padding is dropped on output.
__device__ __forceinline__ void stamp_rows(uint16_t* col, const uint16_t* __restrict__ lut_row, int lo, int hi,
                                           bool active) {
  // col points at tile[(py - s + pad) * 32 + lane]; lut_row is warp-uniform
  for (int oy = lo; oy < hi; ++oy) {
    const uint16_t v = __ldg(lut_row + oy);
    if (active) {
      uint16_t* q = col + oy * kStripW;
      const uint16_t cur = *q;
      if (v > cur) *q = v;
    }
  }
}

__global__ void __launch_bounds__(32) stamp_kernel(const uint16_t* __restrict__ land,
                                                  const unsigned int* __restrict__ row_robot,
                                                  const uint16_t* __restrict__ lut_t,   // [H][2s_t][2s_t]
                                                  const uint8_t* __restrict__ span_t,   // [H][2s_t][2] non-zero oy range
                                                  const uint16_t* __restrict__ lut_b,   // [2s_b][2s_b]
                                                  const uint8_t* __restrict__ span_b,   // [2s_b][2]
                                                  SceneDev P, uint32_t* __restrict__ map) {
  extern __shared__ uint16_t tile[];
  const int lane = threadIdx.x;
  const int f = blockIdx.y;
  const int x0 = blockIdx.x * kStripW;
  const int lx = x0 + lane;
  const int rows = P.H + 2 * P.pad_rows;
  for (int i = lane; i < rows * kStripW / 2; i += 32) reinterpret_cast<uint32_t*>(tile)[i] = 0u;
  __syncwarp();
  const uint16_t* L = land + int64_t(f) * P.W * P.H;
  const unsigned int* RR = row_robot + int64_t(f) * P.H;
  const int d_t = 2 * P.s_t, d_b = 2 * P.s_b;
  for (int y = 0; y < P.H; ++y) {
    const uint16_t* Lr = L + int64_t(y) * P.W;
    // terrain: loc.x = x - s + ox  =>  source column x = lx + s - ox
    for (int ox = 0; ox < d_t; ++ox) {
      const int x = lx + P.s_t - ox;
      unsigned v = kKindNone << 14;
      if (x >= 0 && x < P.W) v = Lr[x];
      const bool active = (v >> 14) == kKindTerrain && lx < P.W;
      if (!__any_sync(0xffffffffu, active)) continue;
      const int py = int(v & 0x3FFF) - P.py_bias;
      const int e = y * d_t + ox;
      const int lo = span_t[2 * e], hi = span_t[2 * e + 1];
      stamp_rows(tile + (py - P.s_t + P.pad_rows) * kStripW + lane, lut_t + int64_t(e) * d_t, lo, hi, active);
    }
    if (RR[y]) {
      for (int ox = 0; ox < d_b; ++ox) {
        const int x = lx + P.s_b - ox;
        unsigned v = kKindNone << 14;
        if (x >= 0 && x < P.W) v = Lr[x];
        const bool active = (v >> 14) == kKindRobot && lx < P.W;
        if (!__any_sync(0xffffffffu, active)) continue;
        const int py = int(v & 0x3FFF) - P.py_bias;
        stamp_rows(tile + (py - P.s_b + P.pad_rows) * kStripW + lane, lut_b + ox * d_b, span_b[2 * ox], span_b[2 * ox + 1], active);
      }
    }
  }
  __syncwarp();
  if (lx < P.W) {
    uint32_t* M = map + int64_t(f) * P.W * P.H;
    const bool col_ok = lx > 0 && lx < P.W - 1;  // pt_cloud.comp:67
    for (int r = 0; r < P.H; ++r) {
      const bool ok = col_ok && r > 0 && r < P.H - 1;
      M[int64_t(r) * P.W + lx] = ok ? uint32_t(tile[(r + P.pad_rows) * kStripW + lane]) : 0u;
    }
  }
}

// ------------------------------------------------------------------ stamp, shared-memory atomics, eight warps per strip
// The single-warp kernels are bound by their own dependency chain (6 resident warps per SM).  Here eight warps share one
// strip: warp w takes the source rows y = w (mod 8), and since two warps may now hit the same map cell the update is a
// native shared-memory atomic max on an unpacked u32 tile (70 KB per strip, two strips per SM, 16 resident warps).
// max is commutative, so the map is bit-identical whatever the interleaving.
constexpr int kAtomWarps = 8;

__global__ void __launch_bounds__(kAtomWarps * 32) stamp_atomic_kernel(const uint16_t* __restrict__ land,
                                                                      const unsigned int* __restrict__ row_robot,
                                                                      const uint16_t* __restrict__ lut_t,   // [H][2s_t][2s_t]
                                                                      const uint8_t* __restrict__ span_t,   // [H][2s_t][2]
                                                                      const uint16_t* __restrict__ lut_b, const uint8_t* __restrict__ span_b,
                                                                      SceneDev P, uint32_t* __restrict__ map) {
  extern __shared__ uint32_t atile[];  // [rows][32] map cells, then kAtomWarps staged LUT rows (u16[d_t * d_t] each, 16-byte aligned)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.y;
  const int x0 = blockIdx.x * kStripW;
  const int lx = x0 + lane;
  const int rows = P.H + 2 * P.pad_rows;
  const int d_t = 2 * P.s_t, d_b = 2 * P.s_b;
  const int lut_words = (d_t * d_t + 7) / 8 * 4;   // u32 words per staged row, rounded to 16 bytes
  uint32_t* s_lut = atile + rows * kStripW + warp * lut_words;
  for (int i = threadIdx.x; i < rows * kStripW; i += blockDim.x) atile[i] = 0u;
  __syncthreads();
  const uint16_t* L = land + int64_t(f) * P.W * P.H;
  const unsigned int* RR = row_robot + int64_t(f) * P.H;
  const int xa = x0 - (P.s_t - 1) + lane, xb = xa + 32;   // source columns x0 - s + 1 .. x0 + s + 31 (two per lane)
  for (int y = warp; y < P.H; y += kAtomWarps) {
    const uint16_t* Lr = L + int64_t(y) * P.W;
    const unsigned r0v = (xa >= 0 && xa < P.W) ? unsigned(Lr[xa]) : unsigned(kKindNone << 14);
    const unsigned r1v = (xb >= 0 && xb < P.W) ? unsigned(Lr[xb]) : unsigned(kKindNone << 14);
    const bool any_terrain = __any_sync(0xffffffffu, (r0v >> 14) == kKindTerrain || (r1v >> 14) == kKindTerrain);
    if (any_terrain) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(lut_t + int64_t(y) * d_t * d_t);  // d_t * d_t even
      for (int i = lane; i < d_t * d_t / 2; i += 32) s_lut[i] = __ldg(src + i);
      __syncwarp();
      const uint16_t* sl = reinterpret_cast<const uint16_t*>(s_lut);
      for (int ox = 0; ox < d_t; ++ox) {
        const int e = y * d_t + ox;
        const int lo = span_t[2 * e], hi = span_t[2 * e + 1];
        if (hi <= lo) continue;                 // all-zero stamp column (warp-uniform)
        const int j = lane + (d_t - 1) - ox;    // index into the 64 staged source pixels: x = lx + s - ox
        const unsigned va = __shfl_sync(0xffffffffu, r0v, j & 31), vb = __shfl_sync(0xffffffffu, r1v, j & 31);
        const unsigned v = j < 32 ? va : vb;
        const bool active = (v >> 14) == kKindTerrain && lx < P.W;
        if (!__any_sync(0xffffffffu, active)) continue;
        uint32_t* col = atile + (int(v & 0x3FFF) - P.py_bias - P.s_t + P.pad_rows) * kStripW + lane;
        const uint16_t* sc = sl + ox * d_t;
        if (active)
          for (int oy = lo; oy < hi; ++oy) atomicMax(col + oy * kStripW, uint32_t(sc[oy]));
      }
      __syncwarp();
    }
    if (RR[y]) {
      for (int ox = 0; ox < d_b; ++ox) {
        const int x = lx + P.s_b - ox;
        unsigned v = kKindNone << 14;
        if (x >= 0 && x < P.W) v = Lr[x];
        const bool active = (v >> 14) == kKindRobot && lx < P.W;
        if (!__any_sync(0xffffffffu, active)) continue;
        const int lo = span_b[2 * ox], hi = span_b[2 * ox + 1];
        uint32_t* col = atile + (int(v & 0x3FFF) - P.py_bias - P.s_b + P.pad_rows) * kStripW + lane;
        for (int oy = lo; oy < hi; ++oy) {
          const uint32_t val = __ldg(lut_b + ox * d_b + oy);
          if (active) atomicMax(col + oy * kStripW, val);
        }
      }
    }
  }
  __syncthreads();
  if (lx < P.W) {
    uint32_t* M = map + int64_t(f) * P.W * P.H;
    const bool col_ok = lx > 0 && lx < P.W - 1;  // pt_cloud.comp:67
    for (int r = warp; r < P.H; r += kAtomWarps) {
      const bool ok = col_ok && r > 0 && r < P.H - 1;
      M[int64_t(r) * P.W + lx] = ok ? atile[(r + P.pad_rows) * kStripW + lane] : 0u;
    }
  }
}

// ------------------------------------------------------------------ stamp, packed (terrain bump size 10)
// Same ownership scheme (one warp = 32 columns, lane = column), tuned for throughput:
//  * two map rows share one 32-bit shared-memory word (lo = even row, hi = odd row), so a 20-row stamp is
//    11 word-wide read / vmaxu2 / write triples instead of 20 half-word ones;
//  * the stamp column for (source row y, column offset ox) comes pre-packed from the host in both row
//    alignments (12 words each, 3 x LDG.128, warp-uniform); a lane picks the alignment of its landing row;
//  * the 11 triples are fully unrolled and independent, which is what hides the shared-memory latency with
//    only ~6 resident warps per SM (one strip = 36 KB of shared memory);
//  * lanes without a terrain source pixel write zeros into a scratch word row instead of branching.
// Robot pixels (40x40 stamps) are rare and take the generic half-word loop on the same tile.
constexpr int kPackWords = 12;

__device__ __forceinline__ unsigned vmaxu2(unsigned a, unsigned b) { return __vmaxu2(a, b); }

__global__ void __launch_bounds__(32) stamp_packed_kernel(const uint16_t* __restrict__ land,
                                                         const unsigned int* __restrict__ row_robot,
                                                         const uint4* __restrict__ lut_pack,      // [H][20][2 alignments][3] uint4
                                                         const unsigned int* __restrict__ row_mask,  // [H] bit ox: column non-empty
                                                         const uint16_t* __restrict__ lut_b, const uint8_t* __restrict__ span_b,
                                                         SceneDev P, uint32_t* __restrict__ map) {
  extern __shared__ uint32_t wtile[];  // [(rows/2) + kPackWords scratch][32] words, then one staged LUT row
  constexpr int S = 10, D = 20, kRowVec = D * 6;  // uint4 per staged LUT row
  const int lane = threadIdx.x;
  const int f = blockIdx.y;
  const int x0 = blockIdx.x * kStripW;
  const int lx = x0 + lane;
  const int rows = P.H + 2 * P.pad_rows;
  const int pair_rows = (rows + 1) / 2;
  const int scratch = pair_rows;                   // first scratch word row
  uint4* s_lut = reinterpret_cast<uint4*>(wtile + (pair_rows + kPackWords) * kStripW);
  for (int i = lane; i < (pair_rows + kPackWords) * kStripW; i += 32) wtile[i] = 0u;
  const uint16_t* L = land + int64_t(f) * P.W * P.H;
  const unsigned int* RR = row_robot + int64_t(f) * P.H;
  uint16_t* htile = reinterpret_cast<uint16_t*>(wtile);
  // source columns needed by this strip: x0 - 9 .. x0 + 41, held as two registers per lane and read by shuffle
  const int xa = x0 - (S - 1) + lane, xb = xa + 32;
  auto ld_land = [&](int y, int x) -> unsigned {
    return (x >= 0 && x < P.W) ? unsigned(L[int64_t(y) * P.W + x]) : unsigned(kKindNone << 14);
  };
  unsigned r0v = ld_land(0, xa), r1v = ld_land(0, xb);
  for (int i = lane; i < kRowVec; i += 32) s_lut[i] = __ldg(lut_pack + i);
  __syncwarp();
  for (int y = 0; y < P.H; ++y) {
    // prefetch the next row (land + packed stamps) while this row is stamped
    const int yn = min(y + 1, P.H - 1);
    const unsigned n0v = ld_land(yn, xa), n1v = ld_land(yn, xb);
    uint4 nl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k;
      nl[k] = i < kRowVec ? __ldg(lut_pack + int64_t(yn) * kRowVec + i) : make_uint4(0, 0, 0, 0);
    }
    const unsigned mask = __ldg(row_mask + y);
#pragma unroll 2
    for (int ox = 0; ox < D; ++ox) {
      if (!((mask >> ox) & 1u)) continue;  // all-zero stamp column (warp-uniform)
      const int j = lane + (D - 1) - ox;    // index into the 64 staged source pixels: x = lx + S - ox
      const unsigned va = __shfl_sync(0xffffffffu, r0v, j & 31), vb = __shfl_sync(0xffffffffu, r1v, j & 31);
      const unsigned v = j < 32 ? va : vb;
      const bool active = (v >> 14) == kKindTerrain && lx < P.W;
      if (!__any_sync(0xffffffffu, active)) continue;
      const int r0 = int(v & 0x3FFF) - P.py_bias - S + P.pad_rows;  // first tile row of the stamp
      const int base = active ? (r0 >> 1) : scratch;
      const bool odd = active && (r0 & 1);
      const uint4* lp = s_lut + ox * 6 + (odd ? 3 : 0);
      const uint4 a0 = lp[0], a1 = lp[1], a2 = lp[2];
      const unsigned am = active ? 0xFFFFFFFFu : 0u;
      const unsigned w[11] = {a0.x & am, a0.y & am, a0.z & am, a0.w & am, a1.x & am, a1.y & am,
                              a1.z & am, a1.w & am, a2.x & am, a2.y & am, a2.z & am};
      uint32_t* q = wtile + base * kStripW + lane;
      unsigned cur[11];
#pragma unroll
      for (int k = 0; k < 11; ++k) cur[k] = q[k * kStripW];
#pragma unroll
      for (int k = 0; k < 11; ++k) q[k * kStripW] = vmaxu2(cur[k], w[k]);
    }
    if (RR[y]) {
      const uint16_t* Lr = L + int64_t(y) * P.W;
      const int d_b = 2 * P.s_b;
      for (int ox = 0; ox < d_b; ++ox) {
        const int x = lx + P.s_b - ox;
        unsigned v = kKindNone << 14;
        if (x >= 0 && x < P.W) v = Lr[x];
        const bool active = (v >> 14) == kKindRobot && lx < P.W;
        if (!__any_sync(0xffffffffu, active)) continue;
        const int r0 = int(v & 0x3FFF) - P.py_bias - P.s_b + P.pad_rows;
        const int lo = span_b[2 * ox], hi = span_b[2 * ox + 1];
        for (int oy = lo; oy < hi; ++oy) {
          const uint16_t val = __ldg(lut_b + ox * d_b + oy);
          if (active) {
            const int r = r0 + oy;
            uint16_t* hq = htile + ((r >> 1) * kStripW + lane) * 2 + (r & 1);
            if (val > *hq) *hq = val;
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k;
      if (i < kRowVec) s_lut[i] = nl[k];
    }
    r0v = n0v;
    r1v = n1v;
    __syncwarp();
  }
  if (lx < P.W) {
    uint32_t* M = map + int64_t(f) * P.W * P.H;
    const bool col_ok = lx > 0 && lx < P.W - 1;  // pt_cloud.comp:67
    for (int r = 0; r < P.H; ++r) {
      const int tr = r + P.pad_rows;
      const uint32_t wv = wtile[(tr >> 1) * kStripW + lane];
      const uint32_t val = (tr & 1) ? (wv >> 16) : (wv & 0xFFFFu);
      const bool ok = col_ok && r > 0 && r < P.H - 1;
      M[int64_t(r) * P.W + lx] = ok ? val : 0u;
    }
  }
}

// ------------------------------------------------------------------ stamp, pruned (default)
// What the shader does per source pixel - 400 imageAtomicMax (pt_cloud.comp:64-75) - is mostly redundant:
//  * the terrain bump uint(y_add) is non-decreasing in val = source row y for every stamp offset (checked element-wise
//    on the host table at create time), so of all terrain pixels of a column that land on the same map row only the
//    one with the largest y can win: 307 200 stamps per 640x480 frame become ~110 000 (column, landing row) entries;
//  * the bump depends on the offset only through d2 = dx*dx + dy*dy and is zero beyond d2 = 80 (also checked), so an
//    entry carries 36 table values, not 400.
// One warp owns one (32-column strip, 64-landing-row band) of one frame and a private shared-memory tile of the map
// around it (16-cell halo, two map rows per 32-bit word).  Phase A: walk the source rows from the bottom up (only the
// rows whose landing range meets the band, via land_kernel's row ranges); the first terrain pixel of a column that lands
// on a row is the dominant one and is appended to that lane's list.  Phase B: lane = source column; every lane walks its
// own list and applies the stamp - the generated, fully unrolled csrc/stamp_pattern.inc: 137 packed read / vmaxu2 /
// write triples with immediate offsets, no atomics.  A tile row is 64 words, so bank = (lane + dx) mod 32: conflict-free,
// and two lanes touch the same word only in different dx groups, which are separated by __syncwarp.  Robot pixels
// (constant value, 33 x 33 table) go through the same lists with a generic loop.  The tile leaves as one 12 KB block;
// merge_kernel takes the maximum of the (at most four) tiles that cover a map cell.
constexpr int kPrBand = 64, kPrHalo = 16, kPrTileW = 64;
constexpr int kPrPairs = (kPrBand + 2 * kPrHalo) / 2;   // 48 word rows
constexpr int kPrTileWords = kPrPairs * kPrTileW;        // 3072 words = 12 KB
constexpr int kPrRadius = 8, kPrD2Max = 80, kPrClasses = 36, kPrRowWords = 20;  // table row: 36 u16, padded to 80 bytes
constexpr int kPrBotCols = 2 * kPrHalo + 1, kPrBotWords = kPrHalo + 1;          // robot table: [2][33][17] words

struct PrunedDev {
  int r_min, nbands, nstrips;   // band b holds the landing rows [r_min + 64 b, r_min + 64 b + 64)
};

// Entries whose column already holds an entry one landing row further down (found earlier = larger source row) are
// "dominated below": that neighbour's stamp is at least as large on every row dy >= 1, so only the upper words of the
// stamp are needed.  The lane's list is two-ended: the other entries grow from slot 0, the dominated ones from slot 63.
__device__ __forceinline__ void pr_append(unsigned v, int y, int want_kind, int enc_lo, unsigned long long& seen, int& cnt_n, int& cnt_d,
                                          uint16_t* list, int lane) {
  const int rel = int(v & 0x3FFF) - enc_lo;
  if (int(v >> 14) == want_kind && unsigned(rel) < unsigned(kPrBand)) {
    const unsigned long long bit = 1ull << rel;
    if (!(seen & bit)) {
      const bool dominated = rel + 1 < kPrBand && ((seen >> (rel + 1)) & 1ull);
      seen |= bit;
      const int slot = dominated ? kPrBand - 1 - cnt_d : cnt_n;
      list[slot * 32 + lane] = uint16_t((rel << 10) | y);
      if (dominated) ++cnt_d;
      else ++cnt_n;
    }
  }
}

// Phase A for one pixel kind: per lane, the (landing row, largest source row) entries of its column inside the band.
// Source rows are visited bottom-up in blocks of kPrChunk rows: first every lane tests kPrChunk / 32 rows' landing ranges
// against the band (all loads of a block in flight together) and the hit rows are compacted, in descending order, into
// `rows`; then the hit rows' land values are read eight rows at a time.
constexpr int kPrChunk = 256;

__device__ __forceinline__ void pr_collect(const uint16_t* __restrict__ L, const uint32_t* __restrict__ RI, const SceneDev& P, int x, int lane,
                                           int want_kind, int y_min, int enc_lo, uint16_t* list, uint16_t* rows, int& cnt_n, int& cnt_d) {
  cnt_n = 0;
  cnt_d = 0;
  unsigned long long seen = 0ull;
  const int enc_hi = enc_lo + kPrBand - 1;
  const unsigned lt = (1u << lane) - 1u;
  for (int y_blk = P.H - 1; y_blk >= y_min; y_blk -= kPrChunk) {
    uint32_t info[kPrChunk / 32];
#pragma unroll
    for (int c = 0; c < kPrChunk / 32; ++c) {
      const int yl = y_blk - 32 * c - lane;
      info[c] = yl >= y_min ? __ldg(RI + yl) : 0x0000FFFFu;   // lo > hi: no pixel of this kind
    }
    int total = 0;
#pragma unroll
    for (int c = 0; c < kPrChunk / 32; ++c) {
      const bool hit = int(info[c] & 0xFFFFu) <= enc_hi && int(info[c] >> 16) >= enc_lo;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) rows[total + __popc(m & lt)] = uint16_t(y_blk - 32 * c - lane);
      total += __popc(m);
    }
    __syncwarp();
    for (int k = 0; k < total; k += 8) {
      int ys[8];
      unsigned vs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ys[j] = k + j < total ? int(rows[k + j]) : -1;
#pragma unroll
      for (int j = 0; j < 8; ++j) vs[j] = (ys[j] >= 0 && x < P.W) ? unsigned(L[int64_t(ys[j]) * P.W + x]) : unsigned(kKindNone << 14);
#pragma unroll
      for (int j = 0; j < 8; ++j) pr_append(vs[j], ys[j], want_kind, enc_lo, seen, cnt_n, cnt_d, list, lane);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) stamp_pruned_kernel(const uint16_t* __restrict__ land, const uint32_t* __restrict__ rowinfo_t,
                                                         const uint32_t* __restrict__ rowinfo_b, const uint4* __restrict__ ttab,  // [H][5]
                                                         const uint32_t* __restrict__ btab,    // [2 parities][33][17] packed robot stamp
                                                         const uint32_t* __restrict__ bspan,   // [33] klo | khi << 8
                                                         SceneDev P, PrunedDev Q, uint32_t* __restrict__ staging) {
  __shared__ __align__(16) uint32_t tile[kPrTileWords];
  __shared__ uint16_t list[kPrBand * 32];
  __shared__ uint16_t rows[kPrChunk];
  const int lane = threadIdx.x;
  const int strip = blockIdx.x / Q.nbands, band = blockIdx.x - strip * Q.nbands;
  const int f = blockIdx.y;
  const int x = strip * kStripW + lane;
  const int enc_lo = Q.r_min + band * kPrBand + P.py_bias;  // encoded landing row of the band's first row
  for (int i = lane; i < kPrTileWords / 4; i += 32) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  const uint16_t* L = land + int64_t(f) * P.W * P.H;
  const int64_t ri = (int64_t(f) * Q.nstrips + strip) * P.H;
  // ---- terrain (val = source row; row 0 stamps nothing: val = 0 makes every y_add NaN -> 0)
  {
    int cnt_n, cnt_d;
    pr_collect(L, rowinfo_t + ri, P, x, lane, kKindTerrain, 1, enc_lo, list, rows, cnt_n, cnt_d);
    // pass 0: the upper words (rows dy <= 0) of every entry; pass 1: the lower words of the entries not dominated below
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int cnt = pass == 0 ? cnt_n + cnt_d : cnt_n;
      const int maxcnt = __reduce_max_sync(0xffffffffu, unsigned(cnt));
      auto slot_of = [&](int i) { return i < cnt_n ? i : kPrBand - 1 - (i - cnt_n); };
      // the table row of the next entry is fetched while the current one is stamped
      uint32_t t[kPrRowWords];
      {
        const uint4* tr = ttab + int64_t(cnt > 0 ? (list[slot_of(0) * 32 + lane] & 1023u) : 0u) * (kPrRowWords / 4);
#pragma unroll
        for (int q = 0; q < kPrRowWords / 4; ++q) {
          const uint4 v = __ldg(tr + q);
          t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
        }
      }
#pragma unroll 1
      for (int i = 0; i < maxcnt; ++i) {
        const unsigned active = __ballot_sync(0xffffffffu, i < cnt);
        if (i < cnt) {
          const unsigned e = list[slot_of(i) * 32 + lane];
          const unsigned en = list[slot_of(i + 1 < cnt ? i + 1 : i) * 32 + lane];
          const int row0 = int(e >> 10) + (kPrHalo - kPrRadius);   // tile row of dy = -8
          uint32_t* base = tile + (row0 >> 1) * kPrTileW + lane + kPrHalo;
          const unsigned sel = (row0 & 1) ? 0x5432u : 0x7654u;
          const uint4* tr = ttab + int64_t(en & 1023u) * (kPrRowWords / 4);
          uint4 tn[kPrRowWords / 4];
#pragma unroll
          for (int q = 0; q < kPrRowWords / 4; ++q) tn[q] = __ldg(tr + q);
#define PR_PRMT(a, b, s) __byte_perm(a, b, s)
#define PR_SYNC() __syncwarp(active)
          if (pass == 0) {
#define PR_RMW_U(off, w) base[off] = __vmaxu2(base[off], w)
#define PR_RMW_L(off, w)
#include "stamp_pattern.inc"
#undef PR_RMW_U
#undef PR_RMW_L
          } else {
#define PR_RMW_U(off, w)
#define PR_RMW_L(off, w) base[off] = __vmaxu2(base[off], w)
#include "stamp_pattern.inc"
#undef PR_RMW_U
#undef PR_RMW_L
          }
#undef PR_PRMT
#undef PR_SYNC
#pragma unroll
          for (int q = 0; q < kPrRowWords / 4; ++q) {
            t[4 * q] = tn[q].x; t[4 * q + 1] = tn[q].y; t[4 * q + 2] = tn[q].z; t[4 * q + 3] = tn[q].w;
          }
        }
      }
      __syncwarp();
    }
  }
  // ---- robot (constant val: one stamp per (column, landing row) is enough)
  {
    __syncwarp();
    int cnt_n, cnt_d;
    pr_collect(L, rowinfo_b + ri, P, x, lane, kKindRobot, 0, enc_lo, list, rows, cnt_n, cnt_d);
    const int cnt = cnt_n + cnt_d;   // the constant robot table is applied whole to every entry
    const int maxcnt = __reduce_max_sync(0xffffffffu, unsigned(cnt));
#pragma unroll 1
    for (int ii = 0; ii < maxcnt; ++ii) {
      const unsigned active = __ballot_sync(0xffffffffu, ii < cnt);
      if (ii < cnt) {
        const int i = ii < cnt_n ? ii : kPrBand - 1 - (ii - cnt_n);
        const int row0 = int(list[i * 32 + lane] >> 10);   // tile row of dy = -16
        uint32_t* base = tile + (row0 >> 1) * kPrTileW + lane;   // column of dx = -16
        const uint32_t* tp = btab + (row0 & 1) * (kPrBotCols * kPrBotWords);
        for (int c = 0; c < kPrBotCols; ++c) {
          const unsigned sp = __ldg(bspan + c);
          const int klo = int(sp & 0xFFu), khi = int(sp >> 8);
          for (int k = klo; k < khi; ++k) {
            uint32_t* q = base + k * kPrTileW + c;
            *q = __vmaxu2(*q, __ldg(tp + c * kPrBotWords + k));
          }
          __syncwarp(active);
        }
      }
    }
  }
  __syncwarp();
  uint4* out = reinterpret_cast<uint4*>(staging + (int64_t(f) * gridDim.x + blockIdx.x) * kPrTileWords);
  for (int i = lane; i < kPrTileWords / 4; i += 32) out[i] = reinterpret_cast<const uint4*>(tile)[i];
}

// map cell = max over the tiles that cover it (two strips x one or two bands); pt_cloud.comp:67 border rule.
// S = this frame's tiles viewed as u16 (two map rows per word: even tile row in the low half).
__device__ __forceinline__ uint32_t merged_cell(const uint16_t* __restrict__ S, const SceneDev& P, const PrunedDev& Q, int x, int y) {
  uint32_t v = 0u;
  if (x > 0 && x < P.W - 1 && y > 0 && y < P.H - 1) {
    const int s1 = (x + kPrHalo) / kStripW;             // tile columns: x - 32 s + 16 in [0, 64)
    const int b1 = (y - Q.r_min + kPrHalo) / kPrBand;   // tile rows:    y - r_min - 64 b + 16 in [0, 96)
#pragma unroll
    for (int ds = 0; ds < 2; ++ds) {
      const int s = s1 - ds;
      if (s < 0 || s >= Q.nstrips) continue;
      const int c = x - s * kStripW + kPrHalo;
#pragma unroll
      for (int db = 0; db < 2; ++db) {
        const int b = b1 - db;
        const int r = y - Q.r_min - b * kPrBand + kPrHalo;
        if (b < 0 || b >= Q.nbands || r >= kPrBand + 2 * kPrHalo) continue;
        const uint32_t t = __ldg(S + size_t(s * Q.nbands + b) * (2 * kPrTileWords) + ((r >> 1) * kPrTileW + c) * 2 + (r & 1));
        v = t > v ? t : v;
      }
    }
  }
  return v;
}

// stand-alone merge: only when the caller wants the map without the weights (otherwise weights_kernel merges)
__global__ void __launch_bounds__(256) merge_kernel(const uint32_t* __restrict__ staging, SceneDev P, PrunedDev Q,
                                                   uint32_t* __restrict__ map, int frames) {
  const int64_t npx = int64_t(P.W) * P.H;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= npx * frames) return;
  const int f = int(idx / npx);
  const int p = int(idx - int64_t(f) * npx);
  const int y = p / P.W, x = p - y * P.W;
  const uint16_t* S = reinterpret_cast<const uint16_t*>(staging) + size_t(f) * Q.nstrips * Q.nbands * (2 * kPrTileWords);
  map[idx] = merged_cell(S, P, Q, x, y);
}

// ------------------------------------------------------------------ weights
// pt_cloud_weights.comp:49-123 as one streaming pass: every thread recomputes the (at most 8)
// distances it needs from the 3x3 map neighbourhood instead of exchanging them through images, which
// is what the shader's 3 stages + barriers do (and what SURVEY §9.6 orders globally).
__device__ __forceinline__ float dist3(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// One block = a 32 x 8 pixel tile.  In literal mode (pack() == 0, so every neighbour decodes to world[0,0])
// all eight link weights of a pixel are values of one field D[a] = |world[a] - world[0,0]| at the pixel and four of its
// neighbours: the tile (+ 1-pixel halo) of D is computed once into shared memory - 1.3 square roots per pixel instead of
// 5 - and the 48 output bytes per pixel leave as streaming 16-byte stores (nothing reads them back on the device).
constexpr int kWtX = 32;
static_assert(((1 - kPrHalo) & 1), "the fused merge pairs map rows (odd, odd + 1) with tile words");

// kFromTiles: the height map does not exist yet - the block merges its cells from the pruned stamp kernel's tiles
// (merged_cell) and also writes them to `map`, which saves the map's round trip through HBM and a launch.
template <bool kStream>
__device__ __forceinline__ void st16(float4* p, float4 v) {
  if (kStream) __stcs(p, v);
  else *p = v;
}

// kFromTiles: the height map does not exist yet - the block merges its cells from the pruned stamp kernel's tiles
// and also writes them to `map`, which saves the map's round trip through HBM and a launch.  A tile word holds map rows
// (y, y + 1) with y odd (r_min is odd), and a block's rows y0 - 1 .. y0 + kWtY (y0 even) are exactly kWtY / 2 + 1 such
// pairs: warp w merges pair w for the block's 34 columns, so the band arithmetic is warp-uniform and a lane's share is
// the strip choice and at most four 32-bit loads per two cells.
template <bool kFromTiles, int kWtY, bool kStream, bool kLiteral>
__global__ void __launch_bounds__(kWtX * kWtY, 1792 / (kWtX * kWtY)) weights_kernel(const uint32_t* __restrict__ src, SceneDev P, PrunedDev Q,
                                                             uint32_t* __restrict__ map_out, float4* __restrict__ world,
                                                             float4* __restrict__ conn0, float4* __restrict__ conn1) {
  __shared__ float fld[kWtY + 2][kWtX + 2];      // literal: D[a]; intent: h[a]
  __shared__ uint32_t cell[kWtY + 2][kWtX + 2];  // map values of the tile + halo
  static_assert(kWtY % 2 == 0 && kWtY / 2 + 1 <= kWtY, "one warp per row pair");
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int x0 = blockIdx.x * kWtX, y0 = blockIdx.y * kWtY;
  const int f = blockIdx.z;
  const size_t npx = size_t(P.W) * P.H;
  // world[0,0].y: cell (0,0) is on the border no stamp ever writes (pt_cloud.comp:67), so the merged value is 0
  const float h00 = kLiteral ? (kFromTiles ? 0.f : float(src[size_t(f) * npx])) : 0.f;
  auto put = [&](int ly, int lx, int gx, int gy, uint32_t m) {
    float v = 0.f;
    if (gx >= 0 && gx < P.W && gy >= 0 && gy < P.H) {
      const float h = float(m);
      v = kLiteral ? dist3(float(gx), h, float(gy), 0.f, h00, 0.f) : h;
    }
    fld[ly][lx] = v;
    cell[ly][lx] = m;
  };
  if (kFromTiles) {
    if (ty < kWtY / 2 + 1) {
      const uint32_t* S32 = src + size_t(f) * Q.nstrips * Q.nbands * kPrTileWords;
      const int gy = y0 - 1 + 2 * ty;                      // odd; rows gy, gy + 1 share the words of (at most) two bands
      const int rr = gy - Q.r_min + kPrHalo;               // >= 0
      const int b1 = rr / kPrBand, r1 = rr - b1 * kPrBand;  // band b1: tile row r1 in [0, 64); band b1 - 1: r1 + 64 (< 96 or absent)
      const bool rows_ok = gy < P.H - 1;                   // else both rows are border / outside
      const bool use_b1 = rows_ok && b1 < Q.nbands, use_b0 = rows_ok && b1 >= 1 && r1 + kPrBand < kPrBand + 2 * kPrHalo;
      const int s_blk = x0 / kStripW;                      // kWtX == kStripW: the block's columns are strip s_blk's
      for (int lx = tx; lx < kWtX + 2; lx += kWtX) {
        const int gx = x0 + lx - 1;
        uint32_t w = 0u;
        if (gx > 0 && gx < P.W - 1) {
          const int s1 = (gx + kPrHalo) / kStripW;         // s_blk - 1 .. s_blk + 1
          (void)s_blk;
#pragma unroll
          for (int ds = 0; ds < 2; ++ds) {
            const int sidx = s1 - ds;
            if (sidx < 0 || sidx >= Q.nstrips) continue;
            const int c = gx - sidx * kStripW + kPrHalo;
            const uint32_t* T0 = S32 + size_t(sidx * Q.nbands) * kPrTileWords + c;
            if (use_b1) w = __vmaxu2(w, __ldg(T0 + size_t(b1) * kPrTileWords + (r1 >> 1) * kPrTileW));
            if (use_b0) w = __vmaxu2(w, __ldg(T0 + size_t(b1 - 1) * kPrTileWords + ((r1 + kPrBand) >> 1) * kPrTileW));
          }
        }
        put(2 * ty, lx, gx, gy, gy > 0 ? (w & 0xFFFFu) : 0u);
        put(2 * ty + 1, lx, gx, gy + 1, (gy + 1 > 0 && gy + 1 < P.H - 1) ? (w >> 16) : 0u);
      }
    }
  } else {
    const uint32_t* M = src + size_t(f) * npx;
    for (int i = ty * kWtX + tx; i < (kWtY + 2) * (kWtX + 2); i += kWtX * kWtY) {
      const int ly = i / (kWtX + 2), lx = i - ly * (kWtX + 2);
      const int gx = x0 + lx - 1, gy = y0 + ly - 1;
      const bool in = gx >= 0 && gx < P.W && gy >= 0 && gy < P.H;
      put(ly, lx, gx, gy, in ? __ldg(M + size_t(gy) * P.W + gx) : 0u);
    }
  }
  __syncthreads();
  const int x = x0 + tx, y = y0 + ty;
  if (x >= P.W || y >= P.H) return;
  const size_t idx = size_t(f) * npx + size_t(y) * P.W + x;
  const uint32_t mine = cell[ty + 1][tx + 1];
  if (kFromTiles) map_out[idx] = mine;
  auto F = [&](int dx, int dy) { return fld[ty + 1 + dy][tx + 1 + dx]; };
  // weight of the link from pixel a = p + (adx, ady) (the "pos" of the shader invocation) to b = p + (bdx, bdy)
  auto link = [&](int adx, int ady, int bdx, int bdy) {
    if (kLiteral) return F(adx, ady);   // pack() == 0 -> unpack loads world[0,0]
    return dist3(float(x + adx), F(adx, ady), float(y + ady), float(x + bdx), F(bdx, bdy), float(y + bdy));
  };
  const bool nxmin = x > 0, nxmax = x < P.W - 1, nymin = y > 0, nymax = y < P.H - 1;
  if (world) st16<kStream>(world + idx, make_float4(float(x), float(mine), float(y), 0.f));  // :59-69
  if (conn1) {  // :86-107  (below, below-left, left, above-left)
    float4 c;
    c.x = nymax ? link(0, 0, 0, 1) : -1.f;
    c.y = (nxmin && nymax) ? link(0, 0, -1, 1) : -1.f;
    c.z = nxmin ? link(0, 0, -1, 0) : -1.f;
    c.w = (nxmin && nymin) ? link(0, 0, -1, -1) : -1.f;
    st16<kStream>(conn1 + idx, c);
  }
  if (conn0) {  // :112-122  conn1[up].r, conn1[up-right].g, conn1[right].b, conn1[down-right].a
    float4 c;
    c.x = nymin ? link(0, -1, 0, 0) : -1.f;
    c.y = (nxmax && nymin) ? link(1, -1, 0, 0) : -1.f;
    c.z = nxmax ? link(1, 0, 0, 0) : -1.f;
    c.w = (nxmax && nymax) ? link(1, 1, 0, 0) : -1.f;
    st16<kStream>(conn0 + idx, c);
  }
}

// ------------------------------------------------------------------ weights, literal mode, lean form (default)
// The general kernel above spends ~230 instructions per pixel and runs at half the rate a pure 52-byte-per-pixel writer
// could (measured: 3.5 TB/s against 7.5 TB/s for a memset) - it is instruction-bound.  In literal mode every link weight
// is a value of the one field D (see above), so this kernel keeps D for a 32 x 32 pixel block (+ halo) in shared memory
// and lets each warp walk four rows of one 32-column strip: per pixel three shared loads (D below, D below-right, the
// map cell), the boundary selects, and four 16-byte streaming stores with pointer increments.
constexpr int kWfRows = 32, kWfThreads = 256, kWfWarpRows = kWfRows / (kWfThreads / 32);

template <bool kFromTiles>
__global__ void __launch_bounds__(kWfThreads) weights_literal_kernel(const uint32_t* __restrict__ src, SceneDev P, PrunedDev Q,
                                                                     uint32_t* __restrict__ map_out, float4* __restrict__ world,
                                                                     float4* __restrict__ conn0, float4* __restrict__ conn1) {
  __shared__ float fld[kWfRows + 2][kWtX + 2];
  __shared__ uint32_t cell[kWfRows + 2][kWtX + 2];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * kWtX, y0 = blockIdx.y * kWfRows;
  const int f = blockIdx.z;
  const size_t npx = size_t(P.W) * P.H;
  const float h00 = kFromTiles ? 0.f : float(src[size_t(f) * npx]);   // cell (0,0) is never stamped (pt_cloud.comp:67)
  auto put = [&](int ly, int lx, int gx, int gy, uint32_t m) {
    float v = 0.f;
    if (gx >= 0 && gx < P.W && gy >= 0 && gy < P.H) v = dist3(float(gx), float(m), float(gy), 0.f, h00, 0.f);
    fld[ly][lx] = v;
    cell[ly][lx] = m;
  };
  if (kFromTiles) {
    // row pairs (y, y + 1), y odd, are the words of the stamp tiles: (kWfRows / 2 + 1) pairs x 34 columns
    const uint32_t* S32 = src + size_t(f) * Q.nstrips * Q.nbands * kPrTileWords;
    for (int i = tid; i < (kWfRows / 2 + 1) * (kWtX + 2); i += kWfThreads) {
      const int pr = i / (kWtX + 2), lx = i - pr * (kWtX + 2);
      const int gx = x0 + lx - 1, gy = y0 - 1 + 2 * pr;
      uint32_t w = 0u;
      if (gx > 0 && gx < P.W - 1 && gy < P.H - 1) {
        const int rr = gy - Q.r_min + kPrHalo;
        const int b1 = rr / kPrBand, r1 = rr - b1 * kPrBand;
        const int s1 = (gx + kPrHalo) / kStripW;
#pragma unroll
        for (int ds = 0; ds < 2; ++ds) {
          const int sidx = s1 - ds;
          if (sidx < 0 || sidx >= Q.nstrips) continue;
          const uint32_t* T0 = S32 + size_t(sidx * Q.nbands) * kPrTileWords + (gx - sidx * kStripW + kPrHalo);
          if (b1 < Q.nbands) w = __vmaxu2(w, __ldg(T0 + size_t(b1) * kPrTileWords + (r1 >> 1) * kPrTileW));
          if (b1 >= 1 && r1 < 2 * kPrHalo) w = __vmaxu2(w, __ldg(T0 + size_t(b1 - 1) * kPrTileWords + ((r1 + kPrBand) >> 1) * kPrTileW));
        }
      }
      put(2 * pr, lx, gx, gy, gy > 0 ? (w & 0xFFFFu) : 0u);
      put(2 * pr + 1, lx, gx, gy + 1, (gy + 1 > 0 && gy + 1 < P.H - 1) ? (w >> 16) : 0u);
    }
  } else {
    const uint32_t* M = src + size_t(f) * npx;
    for (int i = tid; i < (kWfRows + 2) * (kWtX + 2); i += kWfThreads) {
      const int ly = i / (kWtX + 2), lx = i - ly * (kWtX + 2);
      const int gx = x0 + lx - 1, gy = y0 + ly - 1;
      const bool in = gx >= 0 && gx < P.W && gy >= 0 && gy < P.H;
      put(ly, lx, gx, gy, in ? __ldg(M + size_t(gy) * P.W + gx) : 0u);
    }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int x = x0 + lane;
  if (x >= P.W) return;
  const bool nxmin = x > 0, nxmax = x < P.W - 1;
  const float fx = float(x);
  const int ly0 = warp * kWfWarpRows;   // first row of this warp inside the block
  float d_up = fld[ly0][lane + 1], dr_up = fld[ly0][lane + 2];             // row y - 1: D[x], D[x + 1]
  float d_me = fld[ly0 + 1][lane + 1], dr_me = fld[ly0 + 1][lane + 2];     // row y
  size_t idx = size_t(f) * npx + size_t(y0 + ly0) * P.W + x;
#pragma unroll
  for (int k = 0; k < kWfWarpRows; ++k) {
    const int y = y0 + ly0 + k;
    if (y >= P.H) break;   // warp-uniform
    const float d_dn = fld[ly0 + k + 2][lane + 1], dr_dn = fld[ly0 + k + 2][lane + 2];
    const uint32_t mine = cell[ly0 + k + 1][lane + 1];
    const bool nymin = y > 0, nymax = y < P.H - 1;
    if (kFromTiles) map_out[idx] = mine;
    if (world) __stcs(world + idx, make_float4(fx, float(mine), float(y), 0.f));  // :59-69
    if (conn1)   // :86-107 (below, below-left, left, above-left): every present link is D[p]
      __stcs(conn1 + idx, make_float4(nymax ? d_me : -1.f, (nxmin && nymax) ? d_me : -1.f, nxmin ? d_me : -1.f, (nxmin && nymin) ? d_me : -1.f));
    if (conn0)   // :112-122 conn1[up].r, conn1[up-right].g, conn1[right].b, conn1[down-right].a
      __stcs(conn0 + idx, make_float4(nymin ? d_up : -1.f, (nxmax && nymin) ? dr_up : -1.f, nxmax ? dr_me : -1.f, (nxmax && nymax) ? dr_dn : -1.f));
    d_up = d_me; dr_up = dr_me;
    d_me = d_dn; dr_me = dr_dn;
    idx += size_t(P.W);
  }
}

// ------------------------------------------------------------------ Scene materialisation (scene.rs:312-327)
__device__ __forceinline__ int f32_as_i32_sat(float f) {  // Rust `as i32`
  if (!(f == f)) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return -2147483647 - 1;
  return int(f);
}

__global__ void __launch_bounds__(256) materialize_kernel(const uint32_t* __restrict__ map, const float4* __restrict__ world,
                                                         const float4* __restrict__ conn0, const float4* __restrict__ conn1,
                                                         const float4* __restrict__ balls4, int npx,
                                                         float* __restrict__ height, float* __restrict__ pos3,
                                                         float4* __restrict__ conn8, int* __restrict__ balls2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kMaxBalls) {
    const float4 b = balls4[i];
    balls2[2 * i] = f32_as_i32_sat(b.x);
    balls2[2 * i + 1] = f32_as_i32_sat(b.y);
  }
  if (i >= npx) return;
  height[i] = float(map[i]);
  const float4 w = world[i];
  pos3[3 * i + 0] = w.x;
  pos3[3 * i + 1] = w.y;
  pos3[3 * i + 2] = w.z;
  conn8[2 * i] = conn0[i];
  conn8[2 * i + 1] = conn1[i];
}

// ------------------------------------------------------------------ host: tables
inline uint32_t float_to_uint_rz(float f) {  // uint(y_add); NaN -> 0
  if (!(f == f) || f <= 0.0f) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return uint32_t(f);
}

// pt_cloud.comp:59-72 for one (val, bump_size): out[ox][oy] = uint(y_add)
void bump_table(float val, int s, float bump_err, std::vector<uint32_t>* out) {
  out->assign(size_t(4) * s * s, 0u);
  const float c1 = val / bump_err - 1.0f;
  const float c2 = 2.0f / float(s);
  for (int ox = 0; ox < 2 * s; ++ox)
    for (int oy = 0; oy < 2 * s; ++oy) {
      const float dx = float(s - ox), dy = float(s - oy);  // pos - loc
      const float prox = std::sqrt(dx * dx + dy * dy);      // pow(a,2) := a*a
      const float e = c2 * prox - 1.0f;
      const float y_add = val / (1.0f + std::pow(c1, e));
      (*out)[size_t(ox) * 2 * s + oy] = float_to_uint_rz(y_add);
    }
}

void spans_of(const uint16_t* lut, int d, uint8_t* span) {  // per ox: [lo, hi) of non-zero entries
  for (int ox = 0; ox < d; ++ox) {
    int lo = d, hi = 0;
    for (int oy = 0; oy < d; ++oy)
      if (lut[ox * d + oy]) {
        lo = oy < lo ? oy : lo;
        hi = oy + 1;
      }
    if (hi == 0) lo = 0;
    span[2 * ox] = uint8_t(lo);
    span[2 * ox + 1] = uint8_t(hi);
  }
}

// ---- tables of the pruned stamp kernel; returns false when the bump tables do not have the structure it relies on
struct PrunedTables {
  std::vector<uint16_t> ttab;   // [H][40]: class k of source row y
  std::vector<uint32_t> btab;   // [2][33][17]
  std::vector<uint32_t> bspan;  // [33]
};

bool build_pruned_tables(const std::vector<uint16_t>& lut_t, int H, int st, const std::vector<uint16_t>& lut_b, int sb, PrunedTables* out) {
  if (H > 1024) return false;  // list entries hold the source row in 10 bits
  int cls_of[kPrD2Max + 1];
  int ncls = 0;
  for (int d2 = 0; d2 <= kPrD2Max; ++d2) {
    bool is_sum = false;
    for (int a = 0; a <= kPrRadius && !is_sum; ++a)
      for (int b = 0; b <= kPrRadius && !is_sum; ++b) is_sum = a * a + b * b == d2;
    cls_of[d2] = is_sum ? ncls++ : -1;
  }
  if (ncls != kPrClasses) return false;
  const int dt = 2 * st;
  out->ttab.assign(size_t(H) * 2 * kPrRowWords, 0);
  for (int y = 0; y < H; ++y) {
    std::vector<int> seen(kPrClasses, -1);
    for (int ox = 0; ox < dt; ++ox)
      for (int oy = 0; oy < dt; ++oy) {
        const int v = lut_t[(size_t(y) * dt + ox) * dt + oy];
        const int d2 = (st - ox) * (st - ox) + (st - oy) * (st - oy);
        if (d2 > kPrD2Max || d2 >= st * st) {   // outside the pattern, or a cell whose mirror image the shader never writes
          if (v) return false;
          continue;
        }
        const int k = cls_of[d2];
        if (seen[k] >= 0 && seen[k] != v) return false;  // not a function of d2 alone
        seen[k] = v;
      }
    for (int k = 0; k < kPrClasses; ++k) {
      const int v = seen[k] < 0 ? 0 : seen[k];
      if (y > 0 && v < out->ttab[size_t(y - 1) * 2 * kPrRowWords + k]) return false;  // dominance needs monotone in val ...
      if (k > 0 && v > out->ttab[size_t(y) * 2 * kPrRowWords + k - 1]) return false;    // ... and non-increasing in d2 (classes ascend in d2)
      out->ttab[size_t(y) * 2 * kPrRowWords + k] = uint16_t(v);
    }
  }
  // robot: column c <-> dx = c - 16, word k of parity p covers dy = 2k - 16 - p, 2k - 15 - p
  const int db = 2 * sb;
  auto bval = [&](int dx, int dy) -> uint32_t {
    const int ox = dx + sb, oy = dy + sb;
    return (ox >= 0 && ox < db && oy >= 0 && oy < db) ? lut_b[size_t(ox) * db + oy] : 0u;
  };
  for (int ox = 0; ox < db; ++ox)
    for (int oy = 0; oy < db; ++oy)
      if (lut_b[size_t(ox) * db + oy] && (std::abs(ox - sb) > kPrHalo || std::abs(oy - sb) > kPrHalo)) return false;
  out->btab.assign(size_t(2) * kPrBotCols * kPrBotWords, 0u);
  out->bspan.assign(kPrBotCols, 0u);
  for (int c = 0; c < kPrBotCols; ++c) {
    int klo = kPrBotWords, khi = 0;
    for (int p = 0; p < 2; ++p)
      for (int k = 0; k < kPrBotWords; ++k) {
        const int dy0 = 2 * k - kPrHalo - p;
        const uint32_t lo = (dy0 >= -kPrHalo && dy0 <= kPrHalo) ? bval(c - kPrHalo, dy0) : 0u;
        const uint32_t hi = (dy0 + 1 >= -kPrHalo && dy0 + 1 <= kPrHalo) ? bval(c - kPrHalo, dy0 + 1) : 0u;
        const uint32_t w = lo | (hi << 16);
        out->btab[(size_t(p) * kPrBotCols + c) * kPrBotWords + k] = w;
        if (w) {
          klo = k < klo ? k : klo;
          khi = k + 1 > khi ? k + 1 : khi;
        }
      }
    if (khi <= klo) klo = khi = 0;
    out->bspan[c] = uint32_t(klo) | (uint32_t(khi) << 8);
  }
  return true;
}

}  // namespace
}  // namespace tod

using namespace tod;

struct tod_scene {
  int device = 0;
  tod_scene_params prm{};
  SceneDev dev{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // tables
  float *cy = nullptr, *cx = nullptr;
  uint16_t *lut_t = nullptr, *lut_b = nullptr;
  uint8_t *span_t = nullptr, *span_b = nullptr;
  uint4* lut_pack = nullptr;         // packed terrain stamps, both row alignments (terrain_norm_const == 10 only)
  unsigned int* row_mask = nullptr;  // per source row: which stamp columns are non-empty
  size_t packed_smem = 0;
  size_t atomic_smem = 0;
  // pruned stamp kernel (default): class table, robot table, per (strip, row) landing ranges, tile staging
  uint4* ttab = nullptr;
  uint32_t *btab = nullptr, *bspan = nullptr, *rowinfo_t = nullptr, *rowinfo_b = nullptr, *staging = nullptr;
  PrunedDev pruned{};
  bool pruned_ok = false;
  int stamp_impl = 0;
  int weights_variant = 3;
  int chunk_frames = 0;      // frames per land -> stamp -> weights round (0 = the whole batch at once)
  int timed_frames = 0, timed_total = 0;
  bool own_results = false;   // the last append call left map / world / conn / balls in the handle's own buffers
  cudaEvent_t caller_done = nullptr;  // recorded on a caller's stream so that materialize can order itself behind it
  // per-batch buffers
  uint16_t *depth = nullptr, *target = nullptr, *land = nullptr;
  unsigned int* row_robot = nullptr;
  unsigned long long* ball_sums = nullptr;
  uint32_t* map = nullptr;
  float *world = nullptr, *conn0 = nullptr, *conn1 = nullptr, *balls = nullptr;
  // materialisation scratch (one frame)
  float *m_height = nullptr, *m_pos = nullptr, *m_conn = nullptr;
  int* m_balls = nullptr;
  size_t stamp_smem = 0;
  int last_n = 0;
  bool timed = false;
};

extern "C" {

void tod_scene_default_params(tod_scene_params* p) {
  if (!p) return;
  p->width = 640;                  // pt_cloud.comp:23
  p->height = 480;                 // :24
  p->max_depth_in = 4000.0f;       // :25
  p->y_fov = 1.01229096616f;       // :27
  p->x_fov = 1.51843644924f;       // :28
  p->bot_avoidance_const = 100.0f; // :32
  p->bot_norm_const = 20;          // :36
  p->terrain_norm_const = 10;      // :37
  p->bump_err = 0.1f;              // :39
  p->sample_shift = 0;
  p->weights_mode = 0;
  p->max_batch = 1;
}

void tod_scene_destroy(tod_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  void* ptrs[] = {s->ttab, s->btab, s->bspan, s->rowinfo_t, s->rowinfo_b, s->staging, s->lut_pack, s->row_mask, s->cy, s->cx, s->lut_t, s->lut_b, s->span_t, s->span_b, s->depth, s->target, s->land, s->row_robot,
                  s->ball_sums, s->map, s->world, s->conn0, s->conn1, s->balls, s->m_height, s->m_pos, s->m_conn, s->m_balls};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (cudaEvent_t e : s->ev)
    if (e) cudaEventDestroy(e);
  if (s->caller_done) cudaEventDestroy(s->caller_done);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int tod_scene_create(int device, const tod_scene_params* params, tod_scene** out) {
  if (!params || !out) return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: null argument");
  *out = nullptr;
  const tod_scene_params& p = *params;
  if (p.width < 3 || p.height < 3 || p.width > 8192 || p.height > 8192)
    return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: unsupported image size %dx%d", p.width, p.height);
  if (p.terrain_norm_const < 1 || p.terrain_norm_const > 64 || p.bot_norm_const < 1 || p.bot_norm_const > 64)
    return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: bump sizes must be in [1,64]");
  if (p.max_batch < 1) return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: max_batch must be >= 1");
  if (p.sample_shift != 0 && p.sample_shift != 1) return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: sample_shift must be 0 or 1");
  if (p.weights_mode != 0 && p.weights_mode != 1) return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: weights_mode must be 0 or 1");
  if (!(p.max_depth_in > 0.f) || !(p.bump_err > 0.f)) return fail(TOD_ERR_INVALID_ARG, "tod_scene_create: max_depth_in and bump_err must be positive");
  // stamp values are stored as u16 in shared memory: y_add <= val <= max(H-1, bot_avoidance_const)
  if (!(p.bot_avoidance_const >= 0.f && p.bot_avoidance_const < 65535.f))
    return fail(TOD_ERR_UNSUPPORTED, "tod_scene_create: bot_avoidance_const must be in [0, 65535)");
  TOD_TRY(select_device(device));

  tod_scene* s = new tod_scene();
  s->device = device;
  s->prm = p;
  s->stamp_impl = std::getenv("TOD_STAMP_IMPL") ? std::atoi(std::getenv("TOD_STAMP_IMPL")) : 0;
  s->weights_variant = std::getenv("TOD_WEIGHTS_VARIANT") ? std::atoi(std::getenv("TOD_WEIGHTS_VARIANT")) : 3;
  // default chunk: what keeps ~2.7 MB of inter-kernel data per 640x480 frame inside ~3/4 of the L2
  s->chunk_frames = std::getenv("TOD_SCENE_CHUNK") ? std::atoi(std::getenv("TOD_SCENE_CHUNK")) : 0;
  const int W = p.width, H = p.height, st = p.terrain_norm_const, sb = p.bot_norm_const;
  const int smax = st > sb ? st : sb;
  s->dev = SceneDev{W, H, st, sb, 2 * smax + 2, 2 * smax, p.max_depth_in, p.sample_shift, p.weights_mode};
  s->stamp_smem = size_t(H + 2 * s->dev.pad_rows) * kStripW * sizeof(uint16_t);

  // host tables: the shader's transcendentals as functions of small integers
  std::vector<float> cy(H), cx(W);
  const float ty = std::tan(p.y_fov / 2.0f), tx = std::tan(p.x_fov / 2.0f);
  for (int y = 0; y < H; ++y) cy[y] = std::cos(std::atan(ty * float(y) * 2.0f / float(H)));  // pt_cloud.comp:94
  for (int x = 0; x < W; ++x) cx[x] = std::cos(std::atan(tx * float(x) * 2.0f / float(W)));  // :95
  const int dt = 2 * st, db = 2 * sb;
  std::vector<uint16_t> lut_t(size_t(H) * dt * dt), lut_b(size_t(db) * db);
  std::vector<uint8_t> span_t(size_t(H) * dt * 2), span_b(size_t(db) * 2);
  std::vector<uint32_t> slab;
  for (int y = 0; y < H; ++y) {  // terrain: val = float(img_pos.y) (:116-117)
    bump_table(float(y), st, p.bump_err, &slab);
    for (size_t i = 0; i < slab.size(); ++i) lut_t[size_t(y) * dt * dt + i] = uint16_t(slab[i] > 65535u ? 65535u : slab[i]);
    spans_of(&lut_t[size_t(y) * dt * dt], dt, &span_t[size_t(y) * dt * 2]);
  }
  bump_table(p.bot_avoidance_const, sb, p.bump_err, &slab);  // robot (:121-122)
  for (size_t i = 0; i < slab.size(); ++i) lut_b[i] = uint16_t(slab[i] > 65535u ? 65535u : slab[i]);
  spans_of(lut_b.data(), db, span_b.data());

  const size_t npx = size_t(W) * H, nb = size_t(p.max_batch);
  auto cleanup_fail = [&](int rc) {
    tod_scene_destroy(s);
    return rc;
  };
#define SC_CUDA(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      return cleanup_fail(fail(TOD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__)); \
  } while (0)
  SC_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  for (cudaEvent_t& e : s->ev) SC_CUDA(cudaEventCreate(&e));
  SC_CUDA(cudaEventCreateWithFlags(&s->caller_done, cudaEventDisableTiming));
  SC_CUDA(cudaMalloc(&s->cy, H * sizeof(float)));
  SC_CUDA(cudaMalloc(&s->cx, W * sizeof(float)));
  SC_CUDA(cudaMalloc(&s->lut_t, lut_t.size() * 2));
  SC_CUDA(cudaMalloc(&s->lut_b, lut_b.size() * 2));
  SC_CUDA(cudaMalloc(&s->span_t, span_t.size()));
  SC_CUDA(cudaMalloc(&s->span_b, span_b.size()));
  SC_CUDA(cudaMemcpy(s->cy, cy.data(), H * sizeof(float), cudaMemcpyHostToDevice));
  SC_CUDA(cudaMemcpy(s->cx, cx.data(), W * sizeof(float), cudaMemcpyHostToDevice));
  SC_CUDA(cudaMemcpy(s->lut_t, lut_t.data(), lut_t.size() * 2, cudaMemcpyHostToDevice));
  SC_CUDA(cudaMemcpy(s->lut_b, lut_b.data(), lut_b.size() * 2, cudaMemcpyHostToDevice));
  SC_CUDA(cudaMemcpy(s->span_t, span_t.data(), span_t.size(), cudaMemcpyHostToDevice));
  SC_CUDA(cudaMemcpy(s->span_b, span_b.data(), span_b.size(), cudaMemcpyHostToDevice));
  SC_CUDA(cudaMalloc(&s->depth, nb * npx * 2));
  SC_CUDA(cudaMalloc(&s->target, nb * npx * 2));
  SC_CUDA(cudaMalloc(&s->land, nb * npx * 2));
  SC_CUDA(cudaMalloc(&s->row_robot, nb * H * sizeof(unsigned int)));
  SC_CUDA(cudaMalloc(&s->ball_sums, nb * kMaxBalls * 3 * sizeof(unsigned long long)));
  SC_CUDA(cudaMalloc(&s->map, nb * npx * 4));
  SC_CUDA(cudaMalloc(&s->world, nb * npx * 16));
  SC_CUDA(cudaMalloc(&s->conn0, nb * npx * 16));
  SC_CUDA(cudaMalloc(&s->conn1, nb * npx * 16));
  SC_CUDA(cudaMalloc(&s->balls, nb * kMaxBalls * 16));
  SC_CUDA(cudaMalloc(&s->m_height, npx * 4));
  SC_CUDA(cudaMalloc(&s->m_pos, npx * 12));
  SC_CUDA(cudaMalloc(&s->m_conn, npx * 32));
  SC_CUDA(cudaMalloc(&s->m_balls, kMaxBalls * 8));
  // Opt-in shared memory is a per-function, per-device attribute: every handle asks for the same fixed ceiling, so a
  // smaller handle created later cannot lower the limit under an existing one.
  constexpr size_t kSmemCeil = 227 * 1024;
  if (s->stamp_smem > kSmemCeil) return cleanup_fail(fail(TOD_ERR_UNSUPPORTED, "tod_scene_create: a %d-row strip does not fit shared memory", H));
  SC_CUDA(cudaFuncSetAttribute(stamp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemCeil)));
  // the eight-warp atomic kernel stages 64 source columns per warp: a strip needs 2 * s_t + 31 of them
  s->atomic_smem = size_t(H + 2 * s->dev.pad_rows) * kStripW * 4 + size_t(kAtomWarps) * size_t((dt * dt + 7) / 8 * 16);
  if (s->atomic_smem <= 110 * 1024 && 2 * st + 31 <= 64) SC_CUDA(cudaFuncSetAttribute(stamp_atomic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemCeil)));
  else s->atomic_smem = 0;
  {
    PrunedTables pt;
    if (build_pruned_tables(lut_t, H, st, lut_b, sb, &pt)) {
      const int nstrips = (W + kStripW - 1) / kStripW;
      const int r_min = 1 - kPrHalo;
      const int nbands = (H - 2 + 2 * kPrHalo + kPrBand - 1) / kPrBand;
      s->pruned = PrunedDev{r_min, nbands, nstrips};
      SC_CUDA(cudaMalloc(&s->ttab, pt.ttab.size() * 2));
      SC_CUDA(cudaMalloc(&s->btab, pt.btab.size() * 4));
      SC_CUDA(cudaMalloc(&s->bspan, pt.bspan.size() * 4));
      SC_CUDA(cudaMemcpy(s->ttab, pt.ttab.data(), pt.ttab.size() * 2, cudaMemcpyHostToDevice));
      SC_CUDA(cudaMemcpy(s->btab, pt.btab.data(), pt.btab.size() * 4, cudaMemcpyHostToDevice));
      SC_CUDA(cudaMemcpy(s->bspan, pt.bspan.data(), pt.bspan.size() * 4, cudaMemcpyHostToDevice));
      SC_CUDA(cudaMalloc(&s->staging, nb * size_t(nstrips) * nbands * kPrTileWords * 4));
      s->pruned_ok = true;
    }
  }
  {
    const size_t nstrips = size_t(W + kStripW - 1) / kStripW;
    SC_CUDA(cudaMalloc(&s->rowinfo_t, nb * nstrips * H * 4));
    SC_CUDA(cudaMalloc(&s->rowinfo_b, nb * nstrips * H * 4));
  }
  if (st == 10) {
    // word k of a stamp column covers tile rows (2k, 2k+1) counted from the even-aligned start row; per (y, ox)
    // the table holds the even-aligned and the odd-aligned packing, 12 words each
    std::vector<uint32_t> pk(size_t(H) * dt * 2 * kPackWords, 0u);
    std::vector<unsigned int> rmask(H, 0u);
    for (int y = 0; y < H; ++y)
      for (int ox = 0; ox < dt; ++ox) {
        const uint16_t* col = &lut_t[(size_t(y) * dt + ox) * dt];
        uint32_t* E = &pk[((size_t(y) * dt + ox) * 2 + 0) * kPackWords];
        uint32_t* O = &pk[((size_t(y) * dt + ox) * 2 + 1) * kPackWords];
        for (int k = 0; k < 10; ++k) E[k] = uint32_t(col[2 * k]) | (uint32_t(col[2 * k + 1]) << 16);
        O[0] = uint32_t(col[0]) << 16;
        for (int k = 1; k < 10; ++k) O[k] = uint32_t(col[2 * k - 1]) | (uint32_t(col[2 * k]) << 16);
        O[10] = uint32_t(col[19]);
        if (span_t[(size_t(y) * dt + ox) * 2 + 1] != 0) rmask[y] |= 1u << ox;
      }
    SC_CUDA(cudaMalloc(&s->lut_pack, pk.size() * 4));
    SC_CUDA(cudaMalloc(&s->row_mask, rmask.size() * 4));
    SC_CUDA(cudaMemcpy(s->lut_pack, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
    SC_CUDA(cudaMemcpy(s->row_mask, rmask.data(), rmask.size() * 4, cudaMemcpyHostToDevice));
    s->packed_smem = (size_t(H + 2 * s->dev.pad_rows + 1) / 2 + kPackWords) * kStripW * sizeof(uint32_t) + size_t(dt) * 6 * 16;
    SC_CUDA(cudaFuncSetAttribute(stamp_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemCeil)));
  }
#undef SC_CUDA
  *out = s;
  return TOD_OK;
}

// kernels only; all pointers on the device
// One chunk of frames through land -> stamp -> (merge +) weights -> balls.  The caller cuts a batch into chunks small enough
// that what one kernel leaves for the next (land + row ranges 0.7 MB, stamp tiles 2 MB per 640x480 frame) is still in the
// 126 MB L2 when it is read back; the 52 output bytes per pixel leave with streaming stores.
static int scene_chunk(tod_scene* s, const uint16_t* d_depth, const uint16_t* d_target, int n, uint32_t* d_map, float* d_world,
                       float* d_conn0, float* d_conn1, float* d_balls, cudaStream_t st, bool timed) {
  const SceneDev& P = s->dev;
  const int64_t npx = int64_t(P.W) * P.H;
  const unsigned blocks = unsigned((npx * n + 255) / 256);
  const int nstrips = (P.W + kStripW - 1) / kStripW;
  land_kernel<<<dim3(nstrips, (P.H + 8 * kLandRows - 1) / (8 * kLandRows), n), dim3(32, 8), 0, st>>>(d_depth, d_target, s->cy, s->cx, P, s->land, s->row_robot, s->ball_sums,
                                                                        s->rowinfo_t, s->rowinfo_b);
  if (timed) TOD_CUDA(cudaEventRecord(s->ev[0], st));
  dim3 grid(nstrips, n);
  // TOD_STAMP_IMPL: 0 = pruned (default: 1 entry per (column, landing row), no atomics), 1 = eight-warp shared-atomic kernel,
  // 2 = single-warp packed kernel (bump size 10), 3 = single-warp generic kernel.  A handle whose bump tables lack the
  // structure a kernel relies on falls through to the next one; all four produce the same bytes.
  const int impl_env = s->stamp_impl;   // read from the environment when the handle was created
  const bool want_weights = d_world || d_conn0 || d_conn1;
  bool merged_in_weights = false;
  if (impl_env <= 0 && s->pruned_ok) {
    stamp_pruned_kernel<<<dim3(nstrips * s->pruned.nbands, n), 32, 0, st>>>(s->land, s->rowinfo_t, s->rowinfo_b, s->ttab, s->btab, s->bspan, P,
                                                                              s->pruned, s->staging);
    if (want_weights) merged_in_weights = true;
    else merge_kernel<<<blocks, 256, 0, st>>>(s->staging, P, s->pruned, d_map, n);
  } else if (impl_env <= 1 && s->atomic_smem)
    stamp_atomic_kernel<<<grid, kAtomWarps * 32, s->atomic_smem, st>>>(s->land, s->row_robot, s->lut_t, s->span_t, s->lut_b, s->span_b, P, d_map);
  else if (impl_env <= 2 && s->lut_pack)
    stamp_packed_kernel<<<grid, 32, s->packed_smem, st>>>(s->land, s->row_robot, s->lut_pack, s->row_mask, s->lut_b, s->span_b, P, d_map);
  else
    stamp_kernel<<<grid, 32, s->stamp_smem, st>>>(s->land, s->row_robot, s->lut_t, s->span_t, s->lut_b, s->span_b, P, d_map);
  if (timed) TOD_CUDA(cudaEventRecord(s->ev[1], st));
  if (want_weights) {
    float4 *w4 = reinterpret_cast<float4*>(d_world), *c04 = reinterpret_cast<float4*>(d_conn0), *c14 = reinterpret_cast<float4*>(d_conn1);
    const uint32_t* src = merged_in_weights ? s->staging : d_map;
    uint32_t* mo = merged_in_weights ? d_map : nullptr;
#define TOD_WT(FT, WY, CS)                                                                                                          \
  do {                                                                                                                              \
    const dim3 g_((P.W + kWtX - 1) / kWtX, (P.H + WY - 1) / WY, n), b_(kWtX, WY);                                                   \
    if (P.weights_mode == 0) weights_kernel<FT, WY, CS, true><<<g_, b_, 0, st>>>(src, P, s->pruned, mo, w4, c04, c14);             \
    else weights_kernel<FT, WY, CS, false><<<g_, b_, 0, st>>>(src, P, s->pruned, mo, w4, c04, c14);                                \
  } while (0)
    const int v = s->weights_variant;   // TOD_WEIGHTS_VARIANT: 3 (default) = lean literal-mode kernel; 0 / 1 / 2 = general kernel with 4 / 8 / 16 rows per block
    if (P.weights_mode == 0 && v == 3) {
      const dim3 g_((P.W + kWtX - 1) / kWtX, (P.H + kWfRows - 1) / kWfRows, n);
      if (merged_in_weights) weights_literal_kernel<true><<<g_, kWfThreads, 0, st>>>(src, P, s->pruned, mo, w4, c04, c14);
      else weights_literal_kernel<false><<<g_, kWfThreads, 0, st>>>(src, P, s->pruned, mo, w4, c04, c14);
    } else if (merged_in_weights) {
      switch (v) {
        case 0: TOD_WT(true, 4, true); break;
        case 2: TOD_WT(true, 16, true); break;
        case 9: TOD_WT(true, 8, false); break;
        default: TOD_WT(true, 8, true); break;
      }
    } else {
      TOD_WT(false, 8, true);
    }
#undef TOD_WT
  }
  if (timed) TOD_CUDA(cudaEventRecord(s->ev[2], st));
  if (d_balls) balls_kernel<<<(n * kMaxBalls + 127) / 128, 128, 0, st>>>(s->ball_sums, d_balls, n * kMaxBalls);
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

// kernels only; all pointers on the device
static int scene_run(tod_scene* s, const uint16_t* d_depth, const uint16_t* d_target, int n, uint32_t* d_map,
                     float* d_world, float* d_conn0, float* d_conn1, float* d_balls, cudaStream_t st, bool timed) {
  const SceneDev& P = s->dev;
  const size_t npx = size_t(P.W) * P.H;
  TOD_CUDA(cudaMemsetAsync(s->row_robot, 0, size_t(n) * P.H * sizeof(unsigned int), st));
  TOD_CUDA(cudaMemsetAsync(s->ball_sums, 0, size_t(n) * kMaxBalls * 3 * sizeof(unsigned long long), st));
  const int chunk = s->chunk_frames > 0 ? s->chunk_frames : n;
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int cn = std::min(chunk, n - c0);
    const bool last = c0 + cn >= n;
    // per-chunk scratch (land, row ranges, tiles, ball sums, robot flags) is reused: the kernels of consecutive chunks
    // are ordered by the stream.  The timing events bracket the last chunk only; tod_scene_last_kernel_ms scales them.
    TOD_TRY(scene_chunk(s, d_depth + c0 * npx, d_target + c0 * npx, cn, d_map + c0 * npx, d_world ? d_world + c0 * npx * 4 : nullptr,
                        d_conn0 ? d_conn0 + c0 * npx * 4 : nullptr, d_conn1 ? d_conn1 + c0 * npx * 4 : nullptr,
                        d_balls ? d_balls + size_t(c0) * kMaxBalls * 4 : nullptr, st, timed && last));
    if (!last) {
      TOD_CUDA(cudaMemsetAsync(s->row_robot, 0, size_t(cn) * P.H * sizeof(unsigned int), st));
      TOD_CUDA(cudaMemsetAsync(s->ball_sums, 0, size_t(cn) * kMaxBalls * 3 * sizeof(unsigned long long), st));
    }
    s->timed_frames = cn;
  }
  s->timed = timed;
  s->timed_total = n;
  return TOD_OK;
}

int tod_scene_append_batch_device(tod_scene* s, const uint16_t* d_depth, const uint16_t* d_target, int n,
                                  uint32_t* d_map, float* d_world4, float* d_conn0, float* d_conn1, float* d_balls4,
                                  void* stream) {
  if (!s || !d_depth || !d_target) return fail(TOD_ERR_INVALID_ARG, "tod_scene_append_batch_device: null argument");
  if (n < 1 || n > s->prm.max_batch) return fail(TOD_ERR_CAPACITY, "tod_scene_append_batch_device: n=%d outside [1,%d]", n, s->prm.max_batch);
  TOD_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : s->stream;
  s->last_n = n;
  // tod_scene_materialize reads the handle's own images: it is only valid after a call that filled all of them
  s->own_results = !d_map && !d_world4 && !d_conn0 && !d_conn1 && !d_balls4;
  const bool own = s->own_results;
  TOD_TRY(scene_run(s, d_depth, d_target, n, d_map ? d_map : s->map, own ? s->world : d_world4, own ? s->conn0 : d_conn0,
                    own ? s->conn1 : d_conn1, own ? s->balls : d_balls4, st, stream == nullptr));
  if (stream) {  // later work on the handle's stream (materialize) must follow the caller's stream
    TOD_CUDA(cudaEventRecord(s->caller_done, st));
    TOD_CUDA(cudaStreamWaitEvent(s->stream, s->caller_done, 0));
  }
  return TOD_OK;
}

int tod_scene_append_batch(tod_scene* s, const uint16_t* depth, const uint16_t* target, int n, uint32_t* map,
                           float* world4, float* conn0, float* conn1, float* balls4) {
  if (!s || !depth || !target) return fail(TOD_ERR_INVALID_ARG, "tod_scene_append_batch: null argument");
  if (n < 1 || n > s->prm.max_batch) return fail(TOD_ERR_CAPACITY, "tod_scene_append_batch: n=%d outside [1,%d]", n, s->prm.max_batch);
  TOD_CUDA(cudaSetDevice(s->device));
  const size_t npx = size_t(s->dev.W) * s->dev.H;
  cudaStream_t st = s->stream;
  TOD_CUDA(cudaMemcpyAsync(s->depth, depth, n * npx * 2, cudaMemcpyHostToDevice, st));   // scene.rs:197
  TOD_CUDA(cudaMemcpyAsync(s->target, target, n * npx * 2, cudaMemcpyHostToDevice, st)); // scene.rs:198
  // Scene materialisation reads all four images, so they are always produced on the device
  TOD_TRY(scene_run(s, s->depth, s->target, n, s->map, s->world, s->conn0, s->conn1, s->balls, st, true));
  if (map) TOD_CUDA(cudaMemcpyAsync(map, s->map, n * npx * 4, cudaMemcpyDeviceToHost, st));          // scene.rs:246
  if (world4) TOD_CUDA(cudaMemcpyAsync(world4, s->world, n * npx * 16, cudaMemcpyDeviceToHost, st)); // :257
  if (conn0) TOD_CUDA(cudaMemcpyAsync(conn0, s->conn0, n * npx * 16, cudaMemcpyDeviceToHost, st));   // :258
  if (conn1) TOD_CUDA(cudaMemcpyAsync(conn1, s->conn1, n * npx * 16, cudaMemcpyDeviceToHost, st));   // :259
  if (balls4) TOD_CUDA(cudaMemcpyAsync(balls4, s->balls, size_t(n) * kMaxBalls * 16, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));  // scene.rs:282 future.wait
  s->last_n = n;
  s->own_results = true;
  return TOD_OK;
}

int tod_scene_materialize(tod_scene* s, int frame, float* height, float* pos3, int32_t* balls2, float* connections8) {
  if (!s) return fail(TOD_ERR_INVALID_ARG, "tod_scene_materialize: null handle");
  if (frame < 0 || frame >= s->last_n) return fail(TOD_ERR_INVALID_ARG, "tod_scene_materialize: frame %d not in the last batch of %d", frame, s->last_n);
  if (!s->own_results)
    return fail(TOD_ERR_INVALID_ARG, "tod_scene_materialize: the last tod_scene_append_batch_device call wrote to caller buffers; pass NULL outputs to keep the results in the handle");
  TOD_CUDA(cudaSetDevice(s->device));
  const size_t npx = size_t(s->dev.W) * s->dev.H;
  cudaStream_t st = s->stream;
  materialize_kernel<<<unsigned((npx + 255) / 256), 256, 0, st>>>(
      s->map + frame * npx, reinterpret_cast<const float4*>(s->world) + frame * npx,
      reinterpret_cast<const float4*>(s->conn0) + frame * npx, reinterpret_cast<const float4*>(s->conn1) + frame * npx,
      reinterpret_cast<const float4*>(s->balls) + size_t(frame) * kMaxBalls, int(npx), s->m_height, s->m_pos,
      reinterpret_cast<float4*>(s->m_conn), s->m_balls);
  TOD_CUDA(cudaGetLastError());
  if (height) TOD_CUDA(cudaMemcpyAsync(height, s->m_height, npx * 4, cudaMemcpyDeviceToHost, st));
  if (pos3) TOD_CUDA(cudaMemcpyAsync(pos3, s->m_pos, npx * 12, cudaMemcpyDeviceToHost, st));
  if (connections8) TOD_CUDA(cudaMemcpyAsync(connections8, s->m_conn, npx * 32, cudaMemcpyDeviceToHost, st));
  if (balls2) TOD_CUDA(cudaMemcpyAsync(balls2, s->m_balls, kMaxBalls * 8, cudaMemcpyDeviceToHost, st));
  TOD_CUDA(cudaStreamSynchronize(st));
  return TOD_OK;
}

int tod_scene_last_kernel_ms(tod_scene* s, float* stamp_ms, float* weights_ms) {
  if (!s) return fail(TOD_ERR_INVALID_ARG, "tod_scene_last_kernel_ms: null handle");
  if (!s->timed) return fail(TOD_ERR_INVALID_ARG, "tod_scene_last_kernel_ms: the last call ran on a caller stream and was not timed");
  TOD_CUDA(cudaSetDevice(s->device));
  TOD_CUDA(cudaEventSynchronize(s->ev[2]));
  // the events bracket the last chunk of the call: scale to the whole batch
  const float scale = s->timed_frames > 0 ? float(s->timed_total) / float(s->timed_frames) : 1.f;
  if (stamp_ms) {
    TOD_CUDA(cudaEventElapsedTime(stamp_ms, s->ev[0], s->ev[1]));
    *stamp_ms *= scale;
  }
  if (weights_ms) {
    TOD_CUDA(cudaEventElapsedTime(weights_ms, s->ev[1], s->ev[2]));
    *weights_ms *= scale;
  }
  return TOD_OK;
}

}  // extern "C"
