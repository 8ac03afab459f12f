// Kernels around the int8 graph.  See post.h.
//
// Float arithmetic that must match the CPU reference bit for bit is written with the explicit
// round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...) so that no FMA contraction can change a
// result (Rust never fuses a*b+c).
#include "post.h"

#include <algorithm>
#include <cstdlib>

#include "common.h"

namespace tod {
namespace {

// ================================================================= literal post-processing
// One CTA per tile, one thread per seg-grid cell.
//   yolact.rs:108-118  running-max classification over channels 0..3
//   yolact.rs:52-88    terrible_id   (literal: no visited set -> either every id stays -1 or the
//                                     reference never returns; intent: 4-connected components)
//   yolact.rs:127-128  pack + 8x nearest replicate
__global__ void __launch_bounds__(1024) seg_post_kernel(const uint8_t* __restrict__ seg, int64_t ts, SegPost P,
                                                       uint32_t* __restrict__ out, uint32_t* __restrict__ cells_out,
                                                       int* __restrict__ diverges) {
  __shared__ uint8_t s_cls[1024];
  __shared__ int s_label[1024];
  __shared__ uint32_t s_val[1024];
  __shared__ int s_warp_roots[32];
  const int t = blockIdx.x;
  const int tid = threadIdx.x;
  const int cells = P.gh * P.gw;
  const uint8_t* q = seg + int64_t(t) * ts;

  int cls = 0;
  if (tid < cells) {
    const uint8_t* chunk = q + int64_t(tid) * P.ch;
    float mx = 0.0f;  // :109
    bool f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool hit = false;
      if (i < P.ch) {
        const float v = __fmul_rn(P.scale, float(int(chunk[i]) - P.zp));  // yolact.rs:177
        hit = v > mx;                                                      // :110
        if (hit) mx = v;
      }
      f[i] = hit;
    }
    if (!f[0] && f[1] && !f[2] && !f[3]) cls = 1;  // :112-117
    else if (!f[0] && f[2] && !f[3]) cls = 2;
    else if (!f[0] && f[3]) cls = 3;
    else cls = 0;
  }
  s_cls[tid] = uint8_t(cls);
  __syncthreads();

  int id = -1;
  if (P.id_mode == 0) {
    // literal: neighbours by flat index +-1 / +-gw (wrapping across row ends)
    bool bad = false;
    if (tid < cells && cls == 3) {
      const int nb[4] = {tid - 1, tid + 1, tid - P.gw, tid + P.gw};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (nb[k] >= 0 && nb[k] < cells && s_cls[nb[k]] == 3) bad = true;
    }
    if (__syncthreads_or(bad) && tid == 0) diverges[t] = 1;
  } else {
    // intent: connected components by iterated min-label propagation + pointer jumping
    const int x = tid % P.gw, y = tid / P.gw;
    const bool ball = tid < cells && cls == 3;
    s_label[tid] = ball ? tid : -1;
    __syncthreads();
    bool changed = true;
    while (changed) {
      int l = s_label[tid];
      bool ch = false;
      if (ball) {
        int m = l;
        if (x > 0 && s_label[tid - 1] >= 0) m = min(m, s_label[tid - 1]);
        if (x < P.gw - 1 && s_label[tid + 1] >= 0) m = min(m, s_label[tid + 1]);
        if (y > 0 && s_label[tid - P.gw] >= 0) m = min(m, s_label[tid - P.gw]);
        if (y < P.gh - 1 && s_label[tid + P.gw] >= 0) m = min(m, s_label[tid + P.gw]);
        m = min(m, s_label[m]);  // jump
        ch = m < l;
        l = m;
      }
      __syncthreads();
      if (ball) s_label[tid] = l;
      changed = __syncthreads_or(ch);
    }
    // component id = raster-order rank of its first (minimum-index) cell
    const bool root = ball && s_label[tid] == tid;
    const unsigned bal = __ballot_sync(0xffffffffu, root);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) s_warp_roots[warp] = __popc(bal);
    __syncthreads();
    if (ball) {
      const int r = s_label[tid];
      int rank = 0;
      for (int w = 0; w < (r >> 5); ++w) rank += s_warp_roots[w];
      // roots below r inside r's warp: recount from labels (r's warp may differ from mine)
      const int base = r & ~31;
      for (int j = base; j < r; ++j) rank += (s_cls[j] == 3 && s_label[j] == j) ? 1 : 0;
      id = rank & 0x7F;
    }
  }

  uint32_t v = 0;
  if (tid < cells) {
    const uint32_t idu = uint32_t(int32_t(id));  // `id as u32` sign-extends
    if (P.id_mode == 0) v = (uint32_t(cls) << 24) & (idu << 16);  // literal `&`
    else v = (uint32_t(cls) << 24) | ((idu & 0xFFu) << 16);
  }
  s_val[tid] = v;
  if (tid < cells && cells_out) cells_out[int64_t(t) * cells + tid] = v;   // the grid before the 8x replication
  __syncthreads();
  const int OW = P.gw * P.up, OH = P.gh * P.up;
  uint32_t* o = out + int64_t(t) * OW * OH;
  for (int i = tid; i < OW * OH; i += blockDim.x) {
    const int ox = i % OW, oy = i / OW;
    o[i] = s_val[(oy / P.up) * P.gw + ox / P.up];
  }
}

// ================================================================= Triangle resampling (image 0.24.1)
// vertical_sample: u8 source -> f32 rows; horizontal_sample: f32 -> u8 with clamp + round.
// Source pixel fetchers differ between the pre (u32 frame) and post (two stitched u32 tiles) use.
struct FrameSrc {
  const uint32_t* frames;
  int W, H;
  __device__ __forceinline__ uint32_t px(int f, int x, int y) const { return frames[(int64_t(f) * H + y) * W + x]; }
};
struct TilePairSrc {
  const uint32_t* tiles;
  int tw, th;
  __device__ __forceinline__ uint32_t px(int f, int x, int y) const {  // yolact.rs:219-220 row interleave
    const int t = x / tw;
    return tiles[((int64_t(f) * 2 + t) * th + y) * tw + (x - t * tw)];
  }
};

template <class Src>
__global__ void __launch_bounds__(256) vertical_kernel(Src src, int n, int sw, ResampleAxis A, float* __restrict__ tmp) {
  const int64_t total = int64_t(n) * A.out_size * sw;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int x, oy, f;
  if (total < (int64_t(1) << 31)) {  // 32-bit index arithmetic whenever it fits: the 64-bit divisions outweighed the taps
    const unsigned i32 = unsigned(idx), r = i32 / unsigned(sw);
    x = int(i32 - r * unsigned(sw));
    f = int(r / unsigned(A.out_size));
    oy = int(r - unsigned(f) * unsigned(A.out_size));
  } else {
    x = int(idx % sw);
    oy = int((idx / sw) % A.out_size);
    f = int(idx / (int64_t(sw) * A.out_size));
  }
  const int left = A.left[oy], cnt = A.count[oy];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int i = 0; i < cnt; ++i) {
    const float w = A.weight[oy * kMaxTaps + i];
    const uint32_t p = src.px(f, x, left + i);
    a0 = __fadd_rn(a0, __fmul_rn(float(p >> 24), w));
    a1 = __fadd_rn(a1, __fmul_rn(float((p >> 16) & 0xFF), w));
    a2 = __fadd_rn(a2, __fmul_rn(float((p >> 8) & 0xFF), w));
  }
  float* o = tmp + idx * 3;
  o[0] = a0;
  o[1] = a1;
  o[2] = a2;
}

__device__ __forceinline__ uint32_t clamp_round_u8(float v) {
  v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
  return uint32_t(roundf(v));
}

// mode 0: write RGB8 tiles (yolact.rs:213-214 crops); mode 1: write u32 frame (+ target)
__global__ void __launch_bounds__(256) horizontal_kernel(const float* __restrict__ tmp, int n, int sw, int rows,
                                                        ResampleAxis A, int mode, uint8_t* __restrict__ tiles, int tw,
                                                        uint32_t* __restrict__ frames, uint16_t* __restrict__ target) {
  const int64_t total = int64_t(n) * rows * A.out_size;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int ox, y, f;
  if (total < (int64_t(1) << 31)) {
    const unsigned i32 = unsigned(idx), r = i32 / unsigned(A.out_size);
    ox = int(i32 - r * unsigned(A.out_size));
    f = int(r / unsigned(rows));
    y = int(r - unsigned(f) * unsigned(rows));
  } else {
    ox = int(idx % A.out_size);
    y = int((idx / A.out_size) % rows);
    f = int(idx / (int64_t(A.out_size) * rows));
  }
  const int left = A.left[ox], cnt = A.count[ox];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const float* row = tmp + (int64_t(f) * rows + y) * sw * 3;
  for (int i = 0; i < cnt; ++i) {
    const float w = A.weight[ox * kMaxTaps + i];
    const float* p = row + int64_t(left + i) * 3;
    a0 = __fadd_rn(a0, __fmul_rn(p[0], w));
    a1 = __fadd_rn(a1, __fmul_rn(p[1], w));
    a2 = __fadd_rn(a2, __fmul_rn(p[2], w));
  }
  const uint32_t r = clamp_round_u8(a0), g = clamp_round_u8(a1), b = clamp_round_u8(a2);
  if (mode == 0) {
    const int t = ox / tw;
    uint8_t* o = tiles + (((int64_t(f) * 2 + t) * rows + y) * tw + (ox - t * tw)) * 3;
    o[0] = uint8_t(r);
    o[1] = uint8_t(g);
    o[2] = uint8_t(b);
  } else {
    const uint32_t px = (r << 24) | (g << 16) | (b << 8);  // yolact.rs:231-232 from_be_bytes([r,g,b,0])
    frames[idx] = px;
    if (target) target[idx] = uint16_t(px & 0xFFFFu);  // scene.rs:93
  }
}

// ================================================================= detection
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* k, int n) {  // n = power of two
  for (int size = 2; size <= n; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = 2 * i - (i & (stride - 1));
        const unsigned long long a = k[pos], b = k[pos + stride];
        const bool desc = (pos & size) == 0;
        if ((a < b) == desc) {
          k[pos] = b;
          k[pos + stride] = a;
        }
      }
    }
  __syncthreads();
}

// decode + softmax + per-class candidate lists.  One thread per prior; the class bytes of the CTA's
// priors are staged through shared memory so the global read is coalesced.
constexpr int kDecodeThreads = 128;
__global__ void __launch_bounds__(kDecodeThreads, 12) decode_kernel(DetectCfg c, DetectBuffers b, const uint8_t* __restrict__ cls,
                                                               int64_t cls_ts, const uint8_t* __restrict__ box, int64_t box_ts) {
  extern __shared__ uint8_t s_q[];  // [kDecodeThreads][C]
  __shared__ float s_exp[256];
  __shared__ int s_cnt[256], s_base[256];
  for (int i = threadIdx.x; i < 256; i += kDecodeThreads) s_cnt[i] = 0;
  const int t = blockIdx.y;
  const int p0 = blockIdx.x * kDecodeThreads;
  const int np = min(kDecodeThreads, c.P - p0);
  for (int i = threadIdx.x; i < 256; i += kDecodeThreads) s_exp[i] = b.exp_diff[i];
  const uint8_t* src = cls + int64_t(t) * cls_ts + int64_t(p0) * c.C;
  const int nbytes = np * c.C;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(s_q)) & 15) == 0) {  // the CTA's class bytes are one contiguous block
    const int n16 = nbytes >> 4;
    for (int i = threadIdx.x; i < n16; i += kDecodeThreads) reinterpret_cast<uint4*>(s_q)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (int i = (n16 << 4) + threadIdx.x; i < nbytes; i += kDecodeThreads) s_q[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < nbytes; i += kDecodeThreads) s_q[i] = src[i];
  }
  __syncthreads();
  const bool active = threadIdx.x < np;
  const int p = min(p0 + int(threadIdx.x), c.P - 1);
  // box_utils.decode, variances 0.1 / 0.2
  const float4 pr = reinterpret_cast<const float4*>(b.priors)[p];
  const uint8_t* l = box + int64_t(t) * box_ts + int64_t(p) * 4;
  const float cx = __fadd_rn(pr.x, __fmul_rn(__fmul_rn(b.box_deq[l[0]], 0.1f), pr.z));
  const float cy = __fadd_rn(pr.y, __fmul_rn(__fmul_rn(b.box_deq[l[1]], 0.1f), pr.w));
  const float w = __fmul_rn(pr.z, b.box_exp[l[2]]);
  const float h = __fmul_rn(pr.w, b.box_exp[l[3]]);
  const float x1 = __fsub_rn(cx, __fdiv_rn(w, 2.0f)), y1 = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
  if (active) reinterpret_cast<float4*>(b.boxes)[int64_t(t) * c.P + p] = make_float4(x1, y1, __fadd_rn(w, x1), __fadd_rn(h, y1));
  // softmax over the u8 codes: exp(scale*(q - qmax)) from a 256-entry table, summed class 0..C-1 in order
  const uint8_t* q = s_q + threadIdx.x * c.C;
  int qmax = 0;
  for (int k = 0; k < c.C; ++k) qmax = max(qmax, int(q[k]));
  float sum = 0.f;
  for (int k = 0; k < c.C; ++k) sum = __fadd_rn(sum, s_exp[int(q[k]) - qmax + 255]);
  // fl(e / sum) > conf_thresh, deciding without the division wherever the margin allows (see iou_exceeds)
  const float tu = __fmul_rn(c.conf_thresh, sum);
  const float tu_hi = __fmul_rn(tu, 1.0000004f), tu_lo = __fmul_rn(tu, 0.9999996f);
  auto passes = [&](float e) { return c.conf_thresh > 0.f ? (e > tu_hi ? true : (e < tu_lo ? false : __fdiv_rn(e, sum) > c.conf_thresh)) : __fdiv_rn(e, sum) > c.conf_thresh; };
  // pass 1: per-class candidate counts of this CTA (shared-memory atomics); the thread remembers its candidates as a
  // bit set (C <= 256), so the emit pass below only revisits those few classes
  // `passes` is monotone in e and the table is monotone in its index, so the per-class test is a byte compare against the
  // smallest passing table index (binary search, 8 probes) instead of 80 table look-ups and float tests
  int lo = 0, hi = 256;  // smallest i in [0, 256) with passes(s_exp[i]); 256 if none
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (passes(s_exp[mid])) hi = mid;
    else lo = mid + 1;
  }
  const int q_thr = lo + qmax - 255;  // class k is a candidate  <=>  q[k] >= q_thr
  unsigned cbits[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
  if (active)
    for (int k = 1; k < c.C; ++k)
      if (int(q[k]) >= q_thr) {
        atomicAdd(&s_cnt[k - 1], 1);
        cbits[k >> 5] |= 1u << (k & 31);
      }
  __syncthreads();
  for (int k = threadIdx.x; k < c.C - 1; k += kDecodeThreads) {
    const int cnt = s_cnt[k];
    s_base[k] = cnt ? atomicAdd(b.cand_count + int64_t(t) * (c.C - 1) + k, cnt) : 0;
    s_cnt[k] = 0;
  }
  __syncthreads();
  if (!active) return;
#pragma unroll
  for (int wi = 0; wi < 8; ++wi) {
    unsigned bits = cbits[wi];
    while (bits) {
      const int k = wi * 32 + __ffs(int(bits)) - 1;
      bits &= bits - 1;
      const float e = s_exp[int(q[k]) - qmax + 255];
      const int slot = s_base[k - 1] + atomicAdd(&s_cnt[k - 1], 1);
      b.cand[(int64_t(t) * (c.C - 1) + (k - 1)) * c.P + slot] =
          (static_cast<unsigned long long>(__float_as_uint(__fdiv_rn(e, sum))) << 32) | (0xFFFFFFFFu - unsigned(p));
    }
  }
}

__device__ __forceinline__ float iou_rn(const float4 a, const float4 b) {
  const float ix = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float iy = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  const float inter = __fmul_rn(ix > 0.f ? ix : 0.f, iy > 0.f ? iy : 0.f);
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  return uni > 0.f ? __fdiv_rn(inter, uni) : 0.f;
}

// `iou_rn(a, b) > thresh` without the division in all but the borderline cases: fl(inter / uni) > t is implied by
// inter > t * uni * (1 + 4e-7) and excluded by inter < t * uni * (1 - 4e-7) (the margins cover the roundings of the
// two products and of the quotient: 2.8e-7 > 2 ulp); in between, the exact quotient decides.
__device__ __forceinline__ bool iou_exceeds(const float4 a, float area_a, const float4 b, float area_b, float thresh) {
  const float ix = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float iy = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  const float inter = __fmul_rn(ix > 0.f ? ix : 0.f, iy > 0.f ? iy : 0.f);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (!(uni > 0.f)) return 0.f > thresh;
  const float tu = __fmul_rn(thresh, uni);
  if (thresh > 0.f) {
    if (inter > __fmul_rn(tu, 1.0000004f)) return true;
    if (inter < __fmul_rn(tu, 0.9999996f)) return false;
  }
  return __fdiv_rn(inter, uni) > thresh;
}

// The scan's form of the same decision, ~15 instructions: with s = area_a + area_b (the rounded sum iou_exceeds uses),
// inter / (s - inter) > t  <=>  (1 + t) * inter > t * s in the reals; the two sides are compared with a 1e-6 margin (the roundings
// of either formulation are below 3e-7 relative) and only the sliver in between runs the exact iou_exceeds.
struct IouTest {
  float k1, k2_hi, k2_lo, thresh;
  __device__ __forceinline__ void init(float t) {
    k1 = __fadd_rn(1.0f, t);
    k2_hi = __fmul_rn(t, 1.000001f);
    k2_lo = __fmul_rn(t, 0.999999f);
    thresh = t;
  }
  __device__ __forceinline__ bool exceeds(const float4 a, float area_a, const float4 b, float area_b) const {
    const float ix = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float iy = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float inter = __fmul_rn(fmaxf(ix, 0.f), fmaxf(iy, 0.f));
    const float s = __fadd_rn(area_a, area_b);
    const float lhs = __fmul_rn(k1, inter);
    if (lhs < __fmul_rn(k2_lo, s)) return false;   // includes inter == 0 against any positive s
    if (thresh > 0.f && lhs > __fmul_rn(k2_hi, s) && s > 0.f) return true;
    return iou_exceeds(a, area_a, b, area_b, thresh);
  }
};

// Fast-NMS for one (class, tile): sort the class' candidates, keep top_k, then each warp lane owns a
// column j of the IoU matrix and scans the rows i < j (upper triangle); survivors are compacted with
// ballot/popc and appended to the tile's survivor list.
// Launched twice per batch: once sized for class lists of at most kNmsSmall candidates (8 KB of shared memory -> the whole
// grid is resident at once) and once sized for the worst case (every prior a candidate); a CTA whose list falls in the other
// launch's range returns immediately.  The common case no longer pays the worst case's 36 KB per CTA.
constexpr int kNmsThreads = 256;
constexpr int kNmsSmall = 512;
constexpr int kNmsRank = 256;   // lists up to this long are rank-sorted (needs top_k <= sort_cap / 2, checked per launch)
constexpr int kNmsBins = 1024;  // score histogram of the longer lists: 4 bins per thread
static_assert(kNmsBins == 4 * kNmsThreads, "one thread owns four histogram bins");

// monotone bin of a candidate key's score (float bits in the high word): exponents 2^-5 .. 2^-1 with 7 mantissa bits each, clamped
__device__ __forceinline__ int nms_bin(unsigned long long key) {
  const int v = int(unsigned(key >> 48)) - (122 << 7);
  return min(kNmsBins - 1, max(0, v));
}

// One (class, tile) list.  All threads of the CTA call it with the same arguments.
//  * sort: lists of <= kNmsRank candidates are rank-sorted (keys are unique - the prior index is part of the key - so a
//    key's position in the descending order is the number of larger keys: one pass over the list per thread and one
//    barrier, against 28 - 36 barrier-separated compare-exchange rounds of the bitonic network).  A thread issues the load
//    of its candidate's box before it counts, so the gather's latency hides behind the rank loop and the box lands at its
//    sorted position directly.
//  * Fast-NMS: keep[j] <=> no row i < j of the upper-triangular IoU matrix exceeds the threshold.  Column j costs j tests,
//    so one lane per column (round 1) left the warp of the last columns walking ~m rows alone (ncu: 21 - 39 % of the warps
//    active, the kernel pair 14 % of all instructions of the step).  Here columns are paired (q, m - 1 - q): every pair costs
//    m - 1 tests, and S = 256 / pairs threads share a pair's rows; a suppressed column is a flag in shared memory.
__device__ __forceinline__ void nms_one(const DetectCfg& c, const DetectBuffers& b, unsigned long long* s_keys, int sort_cap, int k, int t, int n) {
  float4* s_box = reinterpret_cast<float4*>(s_keys + sort_cap);
  float* s_area = reinterpret_cast<float*>(s_box + c.top_k);
  int* s_sup = reinterpret_cast<int*>(s_area + c.top_k);
  const unsigned long long* src = b.cand + (int64_t(t) * (c.C - 1) + k) * c.P;
  const int m = min(n, c.top_k);
  const float4* boxes = reinterpret_cast<const float4*>(b.boxes) + int64_t(t) * c.P;
  for (int i = threadIdx.x; i < m; i += kNmsThreads) s_sup[i] = 0;
  // Lists longer than the rank sort takes only need their top_k best keys: a 1024-bin histogram of the score (a monotone function
  // of its float bits between 2^-5 and 1), a suffix scan for the bin the top_k-th best key falls in, and a compaction of the keys
  // at or above that bin - usually a few more than top_k - which then take the rank sort below.  (The bitonic sort of all n keys
  // remains for the lists whose cut bin is too crowded: 45 - 66 barrier-separated rounds over 512 - 2048 keys.)
  int nn = n;
  bool compacted = false;
  if (n > kNmsRank && sort_cap / 2 >= kNmsRank) {
    __shared__ int s_hist[kNmsBins];
    __shared__ int s_warp[kNmsThreads / 32];
    __shared__ int s_cut, s_cnt;
    for (int i = threadIdx.x; i < kNmsBins; i += kNmsThreads) s_hist[i] = 0;
    if (threadIdx.x == 0) {
      s_cnt = 0;
      s_cut = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kNmsThreads) atomicAdd(&s_hist[nms_bin(src[i])], 1);
    __syncthreads();
    // suffix sums, 4 consecutive bins per thread, highest bins first (thread 0 owns the top four)
    const int b0 = kNmsBins - 4 * (int(threadIdx.x) + 1);
    const int h3 = s_hist[b0 + 3], h2 = s_hist[b0 + 2], h1 = s_hist[b0 + 1], h0 = s_hist[b0];
    const int mine = h0 + h1 + h2 + h3;
    int incl = mine;
    const int ln = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (ln >= o) incl += v;
    }
    if (ln == 31) s_warp[wid] = incl;
    __syncthreads();
    int above = incl - mine;
    for (int wq = 0; wq < wid; ++wq) above += s_warp[wq];
    if (above < m && above + mine >= m) {   // the m-th best key lies in one of my bins
      int cut = b0 + 3, acc = above + h3;
      if (acc < m) { cut = b0 + 2; acc += h2; }
      if (acc < m) { cut = b0 + 1; acc += h1; }
      if (acc < m) { cut = b0; }
      s_cut = cut;
    }
    __syncthreads();
    const int cut = s_cut;
    for (int i = threadIdx.x; i < n; i += kNmsThreads) {
      const unsigned long long key = src[i];
      if (nms_bin(key) >= cut) {
        const int slot = atomicAdd(&s_cnt, 1);
        if (slot < kNmsRank) s_keys[slot] = key;
      }
    }
    __syncthreads();
    if (s_cnt <= kNmsRank) {   // uniform
      nn = s_cnt;              // >= m by construction of the cut
      compacted = true;
    }
  }
  if (nn <= kNmsRank && nn <= sort_cap / 2) {
    unsigned long long key = 0ull;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (int(threadIdx.x) < nn) {   // nn <= kNmsRank == kNmsThreads: one candidate per thread
      if (compacted) {
        key = s_keys[threadIdx.x];
      } else {
        key = src[threadIdx.x];
        s_keys[threadIdx.x] = key;
      }
      bx = boxes[int(0xFFFFFFFFu - unsigned(key & 0xFFFFFFFFull))];
    }
    __syncthreads();
    if (int(threadIdx.x) < nn) {
      int rank = 0;
#pragma unroll 8
      for (int q = 0; q < nn; ++q) rank += s_keys[q] > key ? 1 : 0;
      if (rank < m) {
        s_keys[sort_cap / 2 + rank] = key;
        s_box[rank] = bx;
        s_area[rank] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
      }
    }
    __syncthreads();
    s_keys += sort_cap / 2;   // the sorted keys
  } else {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += kNmsThreads) s_keys[i] = i < n ? src[i] : 0ull;
    bitonic_sort_desc(s_keys, np2);
    for (int i = threadIdx.x; i < m; i += kNmsThreads) {
      const float4 bx = boxes[int(0xFFFFFFFFu - unsigned(s_keys[i] & 0xFFFFFFFFull))];
      s_box[i] = bx;
      s_area[i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    }
    __syncthreads();
  }
  // column pairs (q, m - 1 - q), S threads per pair taking interleaved rows
  const int pairs = (m + 1) >> 1;
  const int S = max(1, kNmsThreads / max(pairs, 1));
  IouTest iou;
  iou.init(c.nms_thresh);
  for (int u = threadIdx.x; u < pairs * S; u += kNmsThreads) {
    const int q = u % pairs, sub = u / pairs;
    const int ja = q, jb = m - 1 - q;
    {
      const float4 bj = s_box[jb];
      const float aj = s_area[jb];
      bool sup = false;
      for (int i = sub; i < jb && !sup; i += S) sup = iou.exceeds(s_box[i], s_area[i], bj, aj);
      if (sup) s_sup[jb] = 1;
    }
    if (ja != jb) {
      const float4 bj = s_box[ja];
      const float aj = s_area[ja];
      bool sup = false;
      for (int i = S - 1 - sub; i < ja && !sup; i += S) sup = iou.exceeds(s_box[i], s_area[i], bj, aj);
      if (sup) s_sup[ja] = 1;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int j0 = (threadIdx.x >> 5) * 32; j0 < m; j0 += kNmsThreads) {
    const int j = j0 + lane;
    const bool keep = j < m && !s_sup[j];
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(b.surv_count + t, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      const unsigned long long key = s_keys[j];
      const unsigned prior = 0xFFFFFFFFu - unsigned(key & 0xFFFFFFFFull);
      const unsigned order = 0xFFFFu - unsigned((k << 8) | j);  // ties: class asc, then rank asc
      b.surv[int64_t(t) * (c.C - 1) * c.top_k + base + __popc(bal & ((1u << lane) - 1))] =
          (key & 0xFFFFFFFF00000000ull) | (static_cast<unsigned long long>(order) << 16) | prior;
    }
  }
}

// Fast-NMS, one CTA per (class, tile) whose list has at most n_hi candidates (8 KB of shared memory: the whole grid is
// resident at once).  Longer lists belong to nms_long_kernel.
__global__ void __launch_bounds__(kNmsThreads, 6) nms_kernel(DetectCfg c, DetectBuffers b, int sort_cap, int n_hi) {
  extern __shared__ unsigned long long s_keys[];  // [sort_cap] then float4 boxes[top_k], float areas[top_k], int suppressed[top_k]
  const int k = blockIdx.x, t = blockIdx.y;
  const int n = b.cand_count[int64_t(t) * (c.C - 1) + k];
  if (n <= 0 || n > n_hi) return;
  nms_one(c, b, s_keys, sort_cap, k, t, n);
}

// The lists above n_lo candidates (worst case: every prior a candidate, 36 KB of sort buffer).  A few persistent CTAs instead
// of one CTA per list (5120 CTAs of 36 KB each, most of which only found out that they had nothing to do): every CTA reads
// the whole count array once (all loads in flight), builds the same ordered list of the long lists - a bitmap in shared
// memory, then a prefix over its words - and takes the entries blockIdx.x, blockIdx.x + gridDim.x, ...
constexpr int kNmsLongMaxLists = 16384;   // bitmap capacity: (C - 1) * tiles
__global__ void __launch_bounds__(kNmsThreads) nms_long_kernel(DetectCfg c, DetectBuffers b, int sort_cap, int n_lo, int lists) {
  extern __shared__ unsigned long long s_keys[];
  __shared__ unsigned s_bits[kNmsLongMaxLists / 32];
  __shared__ int s_pref[kNmsLongMaxLists / 32 + 1];
  const int words = (lists + 31) >> 5;
  for (int w = threadIdx.x; w < words; w += kNmsThreads) s_bits[w] = 0u;
  __syncthreads();
#pragma unroll 4
  for (int idx = threadIdx.x; idx < lists; idx += kNmsThreads)
    if (b.cand_count[idx] > n_lo) atomicOr(&s_bits[idx >> 5], 1u << (idx & 31));
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int w = 0; w < words; ++w) {
      s_pref[w] = acc;
      acc += __popc(s_bits[w]);
    }
    s_pref[words] = acc;
  }
  __syncthreads();
  const int total = s_pref[words];
  for (int e = blockIdx.x; e < total; e += gridDim.x) {
    // the e-th long list: the word whose prefix range holds e, then the (e - prefix)-th set bit of that word
    int w = 0;
    while (s_pref[w + 1] <= e) ++w;
    unsigned bits = s_bits[w];
    for (int r = e - s_pref[w]; r > 0; --r) bits &= bits - 1;
    const int idx = (w << 5) + __ffs(int(bits)) - 1;
    const int n = b.cand_count[idx];
    const int t = idx / (c.C - 1), k = idx - t * (c.C - 1);
    __syncthreads();           // the previous list's shared-memory readers are done
    nms_one(c, b, s_keys, sort_cap, k, t, n);
  }
}

// merge the classes of one tile: the max_dets best survivors in key order (score, then class / rank, then prior).
// A full sort of every survivor is wasted work when thousands survive and 100 are kept, so the kernel first bins the
// keys by the top bits of the score (a monotone 12-bit bin), finds by a suffix sum the lowest bin the top max_dets
// reach into, and sorts only the keys at or above that bin.  Same result as the full sort (which remains as the
// fallback for a pathological bin population); one CTA per tile.
constexpr int kSelThreads = 1024;
constexpr int kSelBins = 4096;
constexpr int kSelCap = 2048;

__device__ __forceinline__ int score_bin(unsigned long long key) {
  const int v = int(unsigned(key >> 47)) - (112 << 8);  // float bits >> 15, rebased at 2^-15: exponent low bits + 8 mantissa bits
  return min(kSelBins - 1, max(0, v));
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(DetectCfg c, DetectBuffers b, int full_cap) {
  extern __shared__ unsigned long long s_keys[];  // [full_cap] (fallback), then [kSelCap] short list, then int hist[kSelBins]
  unsigned long long* s_small = s_keys + full_cap;
  int* s_hist = reinterpret_cast<int*>(s_small + kSelCap);
  __shared__ int s_warp[32];
  __shared__ int s_cut, s_cnt;
  const int t = blockIdx.x;
  const int n = b.surv_count[t];
  const unsigned long long* src = b.surv + int64_t(t) * (c.C - 1) * c.top_k;
  const int nd = min(n, c.max_dets);
  unsigned long long* sorted = s_small;
  bool done = false;
  if (n <= kSelCap) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += kSelThreads) s_small[i] = i < n ? src[i] : 0ull;
    if (n > 1) bitonic_sort_desc(s_small, np2);
    else __syncthreads();
    done = true;
  } else {
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) s_hist[i] = 0;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kSelThreads) atomicAdd(&s_hist[score_bin(src[i])], 1);
    __syncthreads();
    // suffix sums over the bins, 4 consecutive bins per thread, highest bins first
    const int b0 = kSelBins - 4 * (threadIdx.x + 1);  // this thread owns bins b0 .. b0+3; thread 0 owns the top four
    const int h3 = s_hist[b0 + 3], h2 = s_hist[b0 + 2], h1 = s_hist[b0 + 1], h0 = s_hist[b0];
    const int mine = h0 + h1 + h2 + h3;
    int incl = mine;  // inclusive scan over threads 0..tid (= bins above and including mine)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int above = incl - mine + (wid ? s_warp[wid - 1] : 0);  // keys in bins strictly above this thread's four
    if (above < nd && above + mine >= nd) {  // the nd-th best key lies in one of my bins
      int cut = b0 + 3, acc = above + h3;
      if (acc < nd) { cut = b0 + 2; acc += h2; }
      if (acc < nd) { cut = b0 + 1; acc += h1; }
      if (acc < nd) { cut = b0; }
      s_cut = cut;
    }
    __syncthreads();
    const int cut = s_cut;
    for (int i = threadIdx.x; i < n; i += kSelThreads) {
      const unsigned long long key = src[i];
      if (score_bin(key) >= cut) {
        const int slot = atomicAdd(&s_cnt, 1);
        if (slot < kSelCap) s_small[slot] = key;
      }
    }
    __syncthreads();
    const int m = s_cnt;
    if (m <= kSelCap) {
      int np2 = 1;
      while (np2 < m) np2 <<= 1;
      for (int i = m + threadIdx.x; i < np2; i += kSelThreads) s_small[i] = 0ull;
      bitonic_sort_desc(s_small, np2);
      done = true;
    }
  }
  if (!done) {  // fallback: sort everything
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += kSelThreads) s_keys[i] = i < n ? src[i] : 0ull;
    bitonic_sort_desc(s_keys, np2);
    sorted = s_keys;
  }
  if (threadIdx.x == 0) b.det_count[t] = nd;
  for (int d = threadIdx.x; d < nd; d += kSelThreads) {
    const unsigned long long key = sorted[d];
    const int prior = int(key & 0xFFFFull);
    const unsigned order = 0xFFFFu - unsigned((key >> 16) & 0xFFFFull);
    const int64_t o = int64_t(t) * c.max_dets + d;
    b.det_score[o] = __uint_as_float(unsigned(key >> 32));
    b.det_class[o] = int(order >> 8);
    b.det_prior[o] = prior;
    reinterpret_cast<float4*>(b.det_box)[o] = reinterpret_cast<const float4*>(b.boxes)[int64_t(t) * c.P + prior];
  }
}

// prototype-mask assembly: logits = P[ph*pw, K] . C^T[K, nd] in exact int32 (u8 x u8 dp4a with the
// zero points folded algebraically), then sigmoid, crop to the box (+1 px), threshold.
// One thread per prototype pixel keeps its K bytes in registers and sweeps the tile's detections.
constexpr int kMaskThreads = 128;
template <int K>
__global__ void __launch_bounds__(kMaskThreads) mask_kernel(DetectCfg c, DetectBuffers b, const uint8_t* __restrict__ coef,
                                                           int64_t coef_ts, const uint8_t* __restrict__ proto,
                                                           int64_t proto_ts) {
  extern __shared__ float4 s_crop[];  // [max_dets] crop windows, then coefficients [max_dets][K/4], then sums
  unsigned int* s_coef = reinterpret_cast<unsigned int*>(s_crop + c.max_dets);
  int* s_csum = reinterpret_cast<int*>(s_coef + c.max_dets * (K / 4));
  const int t = blockIdx.y;
  const int nd = b.det_count[t];
  if (nd == 0) return;
  for (int i = threadIdx.x; i < nd * (K / 4); i += kMaskThreads) {
    const int d = i / (K / 4), w = i - d * (K / 4);
    const int prior = b.det_prior[int64_t(t) * c.max_dets + d];
    const uint8_t* cq = coef + int64_t(t) * coef_ts + int64_t(prior) * K + w * 4;
    s_coef[i] = unsigned(cq[0]) | (unsigned(cq[1]) << 8) | (unsigned(cq[2]) << 16) | (unsigned(cq[3]) << 24);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < nd; d += kMaskThreads) {
    unsigned int sum = 0;
    for (int w = 0; w < K / 4; ++w) sum = __dp4a(s_coef[d * (K / 4) + w], 0x01010101u, sum);
    s_csum[d] = int(sum);
    // output_utils.crop / sanitize_coordinates(padding = 1)
    const float4 bx = reinterpret_cast<const float4*>(b.det_box)[int64_t(t) * c.max_dets + d];
    const float x1 = __fmul_rn(bx.x, float(c.pw)), x2 = __fmul_rn(bx.z, float(c.pw));
    const float y1 = __fmul_rn(bx.y, float(c.ph)), y2 = __fmul_rn(bx.w, float(c.ph));
    float xa = __fsub_rn(fminf(x1, x2), 1.0f), xb = __fadd_rn(fmaxf(x1, x2), 1.0f);
    float ya = __fsub_rn(fminf(y1, y2), 1.0f), yb = __fadd_rn(fmaxf(y1, y2), 1.0f);
    xa = fmaxf(xa, 0.f);
    xb = fminf(xb, float(c.pw));
    ya = fmaxf(ya, 0.f);
    yb = fminf(yb, float(c.ph));
    s_crop[d] = make_float4(xa, xb, ya, yb);
  }
  __syncthreads();
  const int npx = c.ph * c.pw;
  const int px = blockIdx.x * kMaskThreads + threadIdx.x;
  const bool in_range = px < npx;
  const int pxc = in_range ? px : npx - 1;   // out-of-range lanes shadow the last pixel: the warp stays converged for the ballots
  unsigned int pw_[K / 4];
  const uint8_t* pq = proto + int64_t(t) * proto_ts + int64_t(pxc) * K;
  unsigned int upsum = 0;
#pragma unroll
  for (int w = 0; w < K / 4; ++w) {
    pw_[w] = *reinterpret_cast<const unsigned int*>(pq + 4 * w);
    upsum = __dp4a(pw_[w], 0x01010101u, upsum);
  }
  const int psum = int(upsum);
  const float fx = float(pxc % c.pw), fy = float(pxc / c.pw);
  const int kzz = K * c.proto_zp * c.coef_zp;
  const int words = (npx + 31) >> 5;
  // sigmoid(l) > 0.5  <=>  l > 0  <=>  idot > 0 for a positive mask scale (|l| >= scale >> 2^-24, so exp(-l) never rounds to 1):
  // when only the binary masks are wanted the transcendental is skipped; a pixel outside a detection's crop window is 0
  // whatever the logit, so it skips the contraction as well
  const bool sign_only = !b.masks && c.mask_scale >= 1e-6f;
  for (int d = 0; d < nd; ++d) {
    const float4 cr = s_crop[d];
    const bool inside = fx >= cr.x && fx < cr.y && fy >= cr.z && fy < cr.w;
    float mval = 0.f;
    bool on = false;
    if (inside) {
      unsigned int dot = 0;
#pragma unroll
      for (int w = 0; w < K / 4; ++w) dot = __dp4a(pw_[w], s_coef[d * (K / 4) + w], dot);
      const int idot = int(dot) - c.coef_zp * psum - c.proto_zp * s_csum[d] + kzz;
      if (sign_only) {
        on = in_range && idot > 0;
      } else {
        const float logit = __fmul_rn(float(idot), c.mask_scale);
        mval = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-logit)));
        on = in_range && mval > 0.5f;
      }
    }
    const int64_t o = (int64_t(t) * c.max_dets + d) * npx + px;
    if (in_range && b.masks) b.masks[o] = mval;
    if (in_range && b.masks_bin) b.masks_bin[o] = on ? 1 : 0;
    if (b.masks_bits) {
      const unsigned bits = __ballot_sync(0xffffffffu, on);
      if ((threadIdx.x & 31) == 0 && (px >> 5) < words) b.masks_bits[(int64_t(t) * c.max_dets + d) * words + (px >> 5)] = bits;
    }
  }
}

// YOLACT `postprocess` (layers/output_utils.py; north-star row 9): the cropped prototype-resolution masks resized to the tile
// with F.interpolate(mode='bilinear', align_corners=False) and thresholded at 0.5.  One thread per tile pixel, one ballot per 32
// pixels -> bit-packed [tile][det][th*tw/32].  PyTorch's arithmetic order, no contraction: src = max(scale*(dst+0.5)-0.5, 0).
__global__ void __launch_bounds__(256) mask_upsample_kernel(const float* __restrict__ masks, const int* __restrict__ det_count,
                                                           int max_dets, int ph, int pw, int th, int tw,
                                                           uint32_t* __restrict__ out_bits) {
  const int t = blockIdx.z, d = blockIdx.y;
  const int words = (th * tw + 31) >> 5;
  const int px = blockIdx.x * 256 + threadIdx.x;
  const bool in_range = px < th * tw;
  bool on = false;
  if (d < det_count[t] && in_range) {
    const int y = px / tw, x = px - y * tw;
    const float sy = __fdiv_rn(float(ph), float(th)), sx = __fdiv_rn(float(pw), float(tw));
    const float fy = fmaxf(__fsub_rn(__fmul_rn(sy, __fadd_rn(float(y), 0.5f)), 0.5f), 0.f);
    const float fx = fmaxf(__fsub_rn(__fmul_rn(sx, __fadd_rn(float(x), 0.5f)), 0.5f), 0.f);
    const int y0 = int(fy), x0 = int(fx);
    const int y1 = y0 + (y0 < ph - 1 ? 1 : 0), x1 = x0 + (x0 < pw - 1 ? 1 : 0);
    const float h1 = __fsub_rn(fy, float(y0)), h0 = __fsub_rn(1.f, h1);
    const float w1 = __fsub_rn(fx, float(x0)), w0 = __fsub_rn(1.f, w1);
    const float* m = masks + (int64_t(t) * max_dets + d) * ph * pw;
    const float top = __fadd_rn(__fmul_rn(w0, m[y0 * pw + x0]), __fmul_rn(w1, m[y0 * pw + x1]));
    const float bot = __fadd_rn(__fmul_rn(w0, m[y1 * pw + x0]), __fmul_rn(w1, m[y1 * pw + x1]));
    on = __fadd_rn(__fmul_rn(h0, top), __fmul_rn(h1, bot)) > 0.5f;
  }
  const unsigned bits = __ballot_sync(0xffffffffu, on);
  if ((threadIdx.x & 31) == 0 && (px >> 5) < words) out_bits[(int64_t(t) * max_dets + d) * words + (px >> 5)] = bits;
}

inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

void launch_seg_postprocess(const uint8_t* seg, int64_t ts, int tiles, const SegPost& p, uint32_t* out, uint32_t* cells_out,
                            int* diverges, cudaStream_t s) {
  cudaMemsetAsync(diverges, 0, sizeof(int) * tiles, s);
  seg_post_kernel<<<tiles, 1024, 0, s>>>(seg, ts, p, out, cells_out, diverges);
}

void launch_classify_pre(const uint32_t* frames, int n, int W, int H, const ResampleAxis& vert, const ResampleAxis& horz,
                         float* tmp, uint8_t* tiles, int tw, int th, cudaStream_t s) {
  // image 0.24.1 resize: vertical_sample first (W x H -> W x th), then horizontal_sample (-> 2tw x th)
  const int64_t tv = int64_t(n) * th * W;
  vertical_kernel<FrameSrc><<<unsigned((tv + 255) / 256), 256, 0, s>>>(FrameSrc{frames, W, H}, n, W, vert, tmp);
  const int64_t thz = int64_t(n) * th * 2 * tw;
  horizontal_kernel<<<unsigned((thz + 255) / 256), 256, 0, s>>>(tmp, n, W, th, horz, 0, tiles, tw, nullptr, nullptr);
}

void launch_classify_post(const uint32_t* tile_px, int n, int W, int H, const ResampleAxis& vert, const ResampleAxis& horz,
                          float* tmp, uint32_t* frames, uint16_t* target, int tw, int th, cudaStream_t s) {
  const int sw = 2 * tw;
  const int64_t tv = int64_t(n) * H * sw;
  vertical_kernel<TilePairSrc><<<unsigned((tv + 255) / 256), 256, 0, s>>>(TilePairSrc{tile_px, tw, th}, n, sw, vert, tmp);
  const int64_t thz = int64_t(n) * H * W;
  horizontal_kernel<<<unsigned((thz + 255) / 256), 256, 0, s>>>(tmp, n, sw, H, horz, 1, nullptr, tw, frames, target);
}

size_t detect_select_smem(const DetectCfg& c) { return size_t(next_pow2((c.C - 1) * c.top_k)) * 8 + size_t(kSelCap) * 8 + size_t(kSelBins) * 4; }
static size_t nms_smem_for(const DetectCfg& c, int cap) { return size_t(cap) * 8 + size_t(c.top_k) * 24; }
static size_t nms_smem(const DetectCfg& c) { return nms_smem_for(c, next_pow2(c.P)); }
static size_t mask_smem(const DetectCfg& c) { return size_t(c.max_dets) * (c.K + 4 + 16); }

int detect_setup_kernels(const DetectCfg& c) {
  if (c.P >= 65536 || (c.C - 1) > 255 || c.top_k > 256)
    return fail(TOD_ERR_UNSUPPORTED, "detection head too large (P=%d, C=%d, top_k=%d)", c.P, c.C, c.top_k);
  if (c.K != 32) return fail(TOD_ERR_UNSUPPORTED, "mask assembly is built for 32 coefficients, model has %d", c.K);
  if (detect_select_smem(c) > 200 * 1024 || nms_smem(c) > 200 * 1024)
    return fail(TOD_ERR_UNSUPPORTED, "detection head does not fit shared memory (P=%d, C=%d, top_k=%d)", c.P, c.C, c.top_k);
  // the opt-in limit is per function and per device, not per handle: always the fixed ceiling the check above enforces, so a
  // handle with a smaller head created later cannot lower it under an existing one
  TOD_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  TOD_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  TOD_CUDA(cudaFuncSetAttribute(nms_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return TOD_OK;
}

int launch_detect_boxes(const DetectCfg& c, const DetectBuffers& b, const uint8_t* cls, int64_t cls_ts, const uint8_t* box,
                        int64_t box_ts, int tiles, cudaStream_t s, cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join) {
  TOD_CUDA(cudaMemsetAsync(b.cand_count, 0, sizeof(int) * size_t(tiles) * (c.C - 1), s));
  TOD_CUDA(cudaMemsetAsync(b.surv_count, 0, sizeof(int) * size_t(tiles), s));
  static const int diag_skip = std::getenv("TOD_DIAG_SKIP") ? std::atoi(std::getenv("TOD_DIAG_SKIP")) : 0;  // timing attribution only
  dim3 g1((c.P + kDecodeThreads - 1) / kDecodeThreads, tiles);
  if (!(diag_skip & 128)) decode_kernel<<<g1, kDecodeThreads, size_t(kDecodeThreads) * c.C, s>>>(c, b, cls, cls_ts, box, box_ts);
  if (diag_skip & 64) {
    select_kernel<<<tiles, kSelThreads, detect_select_smem(c), s>>>(c, b, next_pow2((c.C - 1) * c.top_k));
    return TOD_OK;
  }
  dim3 g2(c.C - 1, tiles);
  const int small_cap = std::min(kNmsSmall, next_pow2(c.P));
  // the two NMS launches touch disjoint (class, tile) lists; the worst-case one is a handful of long CTAs (latency-bound)
  // while the small one fills the machine, so with a second stream they run side by side
  const bool large = next_pow2(c.P) > small_cap;
  const bool split = large && aux && ev_fork && ev_join;
  const int lists = (c.C - 1) * tiles;
  const int long_grid = std::min(lists, 4 * 148);
  if (lists > kNmsLongMaxLists) return fail(TOD_ERR_CAPACITY, "detection: %d (class, tile) lists exceed the Fast-NMS scan capacity %d", lists, kNmsLongMaxLists);
  if (split) {
    TOD_CUDA(cudaEventRecord(ev_fork, s));
    TOD_CUDA(cudaStreamWaitEvent(aux, ev_fork, 0));
    nms_long_kernel<<<long_grid, kNmsThreads, nms_smem(c), aux>>>(c, b, next_pow2(c.P), small_cap, lists);
    TOD_CUDA(cudaEventRecord(ev_join, aux));
  }
  nms_kernel<<<g2, kNmsThreads, nms_smem_for(c, small_cap), s>>>(c, b, small_cap, small_cap);
  if (large && !split) nms_long_kernel<<<long_grid, kNmsThreads, nms_smem(c), s>>>(c, b, next_pow2(c.P), small_cap, lists);
  if (split) TOD_CUDA(cudaStreamWaitEvent(s, ev_join, 0));
  select_kernel<<<tiles, kSelThreads, detect_select_smem(c), s>>>(c, b, next_pow2((c.C - 1) * c.top_k));
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

int launch_detect_masks(const DetectCfg& c, const DetectBuffers& b, const uint8_t* coef, int64_t coef_ts, const uint8_t* proto,
                        int64_t proto_ts, int tiles, cudaStream_t s) {
  static const int diag_skip = std::getenv("TOD_DIAG_SKIP") ? std::atoi(std::getenv("TOD_DIAG_SKIP")) : 0;
  if (diag_skip & 32) return TOD_OK;
  dim3 g4((c.ph * c.pw + kMaskThreads - 1) / kMaskThreads, tiles);
  mask_kernel<32><<<g4, kMaskThreads, mask_smem(c), s>>>(c, b, coef, coef_ts, proto, proto_ts);
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

namespace {
// yolact.rs:179-186: f32 = scale * ((u8 as i32 - zero_point) as f32), element by element (c_store > c: channel-padded storage)
__global__ void __launch_bounds__(256) dequant_u8_kernel(const uint8_t* __restrict__ in, int64_t tile_stride, int c, int c_store,
                                                        int64_t elems, float scale, int zp, float* __restrict__ out) {
  const uint8_t* tin = in + int64_t(blockIdx.y) * tile_stride;
  float* tout = out + int64_t(blockIdx.y) * elems;
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < elems; i += int64_t(gridDim.x) * 256) {
    const int64_t px = i / c;
    const int ch = int(i - px * c);
    tout[i] = __fmul_rn(scale, float(int(tin[px * c_store + ch]) - zp));
  }
}
}  // namespace

int launch_dequant_u8(const uint8_t* in, int64_t tile_stride, int c, int c_store, int64_t elems, int tiles, float scale, int zp,
                      float* out, cudaStream_t s) {
  dim3 g(unsigned(std::min<int64_t>((elems + 255) / 256, 1024)), tiles);
  dequant_u8_kernel<<<g, 256, 0, s>>>(in, tile_stride, c, c_store, elems, scale, zp, out);
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

int launch_mask_upsample(const DetectCfg& c, const DetectBuffers& b, int tiles, int th, int tw, uint32_t* out_bits, cudaStream_t s) {
  dim3 g((th * tw + 255) / 256, c.max_dets, tiles);
  mask_upsample_kernel<<<g, 256, 0, s>>>(b.masks, b.det_count, c.max_dets, c.ph, c.pw, th, tw, out_bits);
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

int launch_detect(const DetectCfg& c, const DetectBuffers& b, const uint8_t* cls, int64_t cls_ts, const uint8_t* box,
                  int64_t box_ts, const uint8_t* coef, int64_t coef_ts, const uint8_t* proto, int64_t proto_ts, int tiles,
                  bool want_masks, cudaStream_t s) {
  TOD_TRY(launch_detect_boxes(c, b, cls, cls_ts, box, box_ts, tiles, s, nullptr, nullptr, nullptr));
  if (want_masks) TOD_TRY(launch_detect_masks(c, b, coef, coef_ts, proto, proto_ts, tiles, s));
  return TOD_OK;
}

}  // namespace tod
