// The Scene's consumer: `path::modify_path` (/root/reference/src/path.rs:25-120) and `Path::serialize` (path.rs:17-21).
// SURVEY.md §8f-4, the last "next" row.
//
// The reference function cannot complete on any input (224*224-element arrays indexed with 640x480 node numbers,
// path.rs:29-30,38,99; cost[usize::MAX - 2] on the extraction's last step, path.rs:105) and is unreachable (main.rs:92),
// so there is no literal output to reproduce: `tod_path_modify` implements the evident intent, rule by rule as listed in
// oracle/path.cpp and DESIGN.md §5.5, and `tod_path_reference_panics()` reports the literal behaviour.
//
// Design: the LIFO relaxation of path.rs:57-96 converges (when it terminates) to the least fixed point of
//   cost[n] = min(cost[n], min over the neighbour list of n: (cost[nb] + connections[n][cn]) + |height[n] - height[nb]|)
// which does not depend on the visiting order (the update is monotone in cost[nb], float rounding included).  On the GPU
// that is a chaotic relaxation: a CTA keeps a band of image rows (+ one halo row above and below: the four neighbours of
// Scene::neighbors are the flat indices n-1, n+1, n-W, n+W) of `cost` in shared memory, sweeps it kInner times, writes
// it back and raises a flag if anything fell; launches repeat until no CTA raised it.  Predecessors are read off the
// converged costs in a last pass (first neighbour in list order that attains the minimum), and the (magnitude,
// rotation) list is walked on the host exactly as path.rs:99-117 does.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"

namespace tod {
namespace {

constexpr float kInf = 3.402823466e+38f;   // f32::MAX (path.rs:30)
constexpr int kRowsPerCta = 4, kInner = 24, kPathThreads = 256;

__host__ __device__ inline int neighbours(long long px, int W, int H, long long out[4]) {  // scene.rs:134-143 (680 read as W)
  int n = 0;
  if (px > 0) out[n++] = px - 1;
  if (px < (long long)W * H - 1) out[n++] = px + 1;
  if (px / W > 0) out[n++] = px - W;
  if (px / W < H - 1) out[n++] = px + W;
  return n;
}

__global__ void path_init_kernel(float* cost, int n, int t0, int t1, int t2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cost[i] = (i == t0 || i == t1 || i == t2) ? 0.0f : kInf;
}

__global__ void __launch_bounds__(kPathThreads) path_relax_kernel(float* __restrict__ cost, const float* __restrict__ height,
                                                                 const float* __restrict__ conn8, int W, int H, int t0, int t1, int t2,
                                                                 int* __restrict__ changed) {
  extern __shared__ float s_cost[];   // rows [r0 - 1, r0 + kRowsPerCta] as one flat run
  const int r0 = blockIdx.x * kRowsPerCta;
  const int rows = min(kRowsPerCta, H - r0);
  const long long N = (long long)W * H;
  const long long lo = (long long)(r0 - 1) * W;          // flat index of s_cost[0] (may be negative: row -1 does not exist)
  const int span = (rows + 2) * W;
  for (int i = threadIdx.x; i < span; i += kPathThreads) {
    const long long g = lo + i;
    s_cost[i] = (g >= 0 && g < N) ? cost[g] : kInf;
  }
  __syncthreads();
  bool fell = false;
  for (int it = 0; it < kInner; ++it) {
    bool any = false;
    for (int i = threadIdx.x; i < rows * W; i += kPathThreads) {
      const long long n = (long long)r0 * W + i;
      if (n == t0 || n == t1 || n == t2) continue;   // targets keep cost 0 (path.rs:41-42)
      long long nb[4];
      const int k = neighbours(n, W, H, nb);
      const float hn = __ldg(height + n);
      const float4 c = __ldg(reinterpret_cast<const float4*>(conn8 + 8 * n));
      float best = s_cost[W + i];
#pragma unroll
      for (int cn = 0; cn < 4; ++cn) {
        if (cn >= k) break;
        const float w = cn == 0 ? c.x : (cn == 1 ? c.y : (cn == 2 ? c.z : c.w));
        const float cb = s_cost[int(nb[cn] - lo)];
        if (w < 0.0f || cb == kInf) continue;        // the shaders' -1 marker: no such edge
        const float cand = __fadd_rn(__fadd_rn(cb, w), fabsf(__fsub_rn(hn, __ldg(height + nb[cn]))));   // path.rs:64, left to right
        if (cand < best) best = cand;
      }
      if (best < s_cost[W + i]) {
        s_cost[W + i] = best;
        any = true;
      }
    }
    fell |= any;
    if (!__syncthreads_or(any)) break;
  }
  if (__syncthreads_or(fell)) {
    for (int i = threadIdx.x; i < rows * W; i += kPathThreads) cost[(long long)r0 * W + i] = s_cost[W + i];
    if (threadIdx.x == 0) *changed = 1;
  }
}

__global__ void path_pred_kernel(const float* __restrict__ cost, const float* __restrict__ height, const float* __restrict__ conn8, int W, int H,
                                 int t0, int t1, int t2, int* __restrict__ pred) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)W * H) return;
  if (n == t0 || n == t1 || n == t2) {
    pred[n] = -2;
    return;
  }
  int p = -1;
  const float mine = cost[n];
  if (mine != kInf) {
    long long nb[4];
    const int k = neighbours(n, W, H, nb);
    const float hn = height[n];
    for (int cn = 0; cn < k && p < 0; ++cn) {
      const float w = conn8[8 * n + cn];
      const float cb = cost[nb[cn]];
      if (w < 0.0f || cb == kInf) continue;
      if (__fadd_rn(__fadd_rn(cb, w), fabsf(__fsub_rn(hn, height[nb[cn]]))) == mine) p = int(nb[cn]);
    }
  }
  pred[n] = p;
}

}  // namespace
}  // namespace tod

using namespace tod;

extern "C" {

int tod_path_reference_panics(void) { return 1; }

int tod_path_modify(int device, int width, int height_px, const float* height, const float* pos3, const int32_t* balls2,
                    const float* connections8, float* cost_out, int32_t* pred_out, float* directions, int cap, int32_t* n_directions) {
  if (!height || !pos3 || !balls2 || !connections8 || !n_directions) return fail(TOD_ERR_INVALID_ARG, "tod_path_modify: null argument");
  if (width < 2 || height_px < 2 || width > 8192 || height_px > 8192) return fail(TOD_ERR_INVALID_ARG, "tod_path_modify: unsupported size %dx%d", width, height_px);
  *n_directions = 0;
  TOD_TRY(select_device(device));
  const int W = width, H = height_px;
  const size_t N = size_t(W) * H;
  int t[3] = {-1, -1, -1};
  for (int b = 0; b < 3; ++b) {   // path.rs:36-37: the first three balls
    const long long x = balls2[2 * b], y = balls2[2 * b + 1];
    if (x >= 0 && x < W && y >= 0 && y < H) t[b] = int(x + y * W);
  }
  float *d_cost = nullptr, *d_h = nullptr, *d_conn = nullptr;
  int *d_pred = nullptr, *d_flag = nullptr;
  cudaStream_t st = nullptr;
  int rc = TOD_OK;
  auto done = [&](int code) {
    for (void* p : {(void*)d_cost, (void*)d_h, (void*)d_conn, (void*)d_pred, (void*)d_flag})
      if (p) cudaFree(p);
    if (st) cudaStreamDestroy(st);
    return code;
  };
#define P_CUDA(expr)                                                                                                            \
  do {                                                                                                                          \
    cudaError_t e_ = (expr);                                                                                                    \
    if (e_ != cudaSuccess) return done(fail(TOD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__)); \
  } while (0)
  P_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  P_CUDA(cudaMalloc(&d_cost, N * 4));
  P_CUDA(cudaMalloc(&d_h, N * 4));
  P_CUDA(cudaMalloc(&d_conn, N * 32));
  P_CUDA(cudaMalloc(&d_pred, N * 4));
  P_CUDA(cudaMalloc(&d_flag, 4));
  P_CUDA(cudaMemcpyAsync(d_h, height, N * 4, cudaMemcpyHostToDevice, st));
  P_CUDA(cudaMemcpyAsync(d_conn, connections8, N * 32, cudaMemcpyHostToDevice, st));
  path_init_kernel<<<unsigned((N + 255) / 256), 256, 0, st>>>(d_cost, int(N), t[0], t[1], t[2]);
  const size_t smem = size_t(kRowsPerCta + 2) * W * sizeof(float);
  if (smem > 200 * 1024) return done(fail(TOD_ERR_UNSUPPORTED, "tod_path_modify: rows of %d pixels do not fit shared memory", W));
  P_CUDA(cudaFuncSetAttribute(path_relax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const unsigned ctas = unsigned((H + kRowsPerCta - 1) / kRowsPerCta);
  // a shortest path crosses at most W*H nodes; every launch settles at least one more node of it, far more in practice
  const long long max_launches = (long long)N + 8;
  int flag = 1;
  for (long long l = 0; flag && l < max_launches; l += 8) {
    P_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
    for (int k = 0; k < 8; ++k) path_relax_kernel<<<ctas, kPathThreads, smem, st>>>(d_cost, d_h, d_conn, W, H, t[0], t[1], t[2], d_flag);
    P_CUDA(cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, st));
    P_CUDA(cudaStreamSynchronize(st));
  }
  path_pred_kernel<<<unsigned((N + 255) / 256), 256, 0, st>>>(d_cost, d_h, d_conn, W, H, t[0], t[1], t[2], d_pred);
  std::vector<float> cost(N);
  std::vector<int32_t> pred(N);
  P_CUDA(cudaMemcpyAsync(cost.data(), d_cost, N * 4, cudaMemcpyDeviceToHost, st));
  P_CUDA(cudaMemcpyAsync(pred.data(), d_pred, N * 4, cudaMemcpyDeviceToHost, st));
  P_CUDA(cudaStreamSynchronize(st));
  P_CUDA(cudaGetLastError());
#undef P_CUDA
  if (cost_out) std::memcpy(cost_out, cost.data(), N * 4);
  if (pred_out) std::memcpy(pred_out, pred.data(), N * 4);
  // extraction (path.rs:99-117): magnitude = cost[node] - cost[path[node]], rotation = angle at `node` in the (pos.0, pos.1) plane
  long long node = (long long)N - H / 2;
  int nd = 0;
  if (cost[node] == kInf) {
    *n_directions = -1;   // the start cannot reach a target
    return done(rc);
  }
  float rotation = 0.0f;
  while (pred[node] != -2) {
    const long long next = pred[node];
    if (next < 0) {
      *n_directions = -1;
      return done(rc);
    }
    if (directions && nd < cap) {
      directions[2 * nd] = cost[node] - cost[next];
      directions[2 * nd + 1] = rotation;
    }
    ++nd;
    const long long last = node;
    node = next;
    if (pred[node] == -2) break;
    const long long after = pred[node];
    const float ax = pos3[3 * last] - pos3[3 * node], ay = pos3[3 * last + 1] - pos3[3 * node + 1];
    const float bx = pos3[3 * after] - pos3[3 * node], by = pos3[3 * after + 1] - pos3[3 * node + 1];
    rotation = std::acos((ax * bx + ay * by) / (std::sqrt(ax * ax + ay * ay) * std::sqrt(bx * bx + by * by)));
  }
  *n_directions = nd;
  return done(rc);
}

// Path::serialize (path.rs:17-21): created.as_secs() big-endian, then big-endian (magnitude, rotation) pairs - the payload
// `handle_path_request` writes to the RoboRIO socket (path.rs:158-162).  Pure host code.
int tod_path_serialize(uint64_t created_secs, const float* directions, int n, uint8_t* out, size_t cap, size_t* bytes) {
  if ((!directions && n > 0) || !out || n < 0) return fail(TOD_ERR_INVALID_ARG, "tod_path_serialize: bad argument");
  const size_t need = 8 + size_t(n) * 8;
  if (bytes) *bytes = need;
  if (cap < need) return fail(TOD_ERR_CAPACITY, "tod_path_serialize: %zu bytes needed, buffer holds %zu", need, cap);
  for (int i = 0; i < 8; ++i) out[i] = uint8_t(created_secs >> (8 * (7 - i)));
  for (int i = 0; i < 2 * n; ++i) {
    uint32_t u;
    std::memcpy(&u, directions + i, 4);
    for (int k = 0; k < 4; ++k) out[8 + 4 * size_t(i) + k] = uint8_t(u >> (8 * (3 - k)));
  }
  return TOD_OK;
}

}  // extern "C"
