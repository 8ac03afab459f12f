// CUDA-core kernels of the int8 graph: the layers that have no tensor-core mapping (depthwise,
// element-wise, resampling) and a direct convolution used for the shapes the tcgen05 implicit-GEMM
// path does not take (Cin = 3 stem, strided 3x3) and as its on-device cross-check.
//
// Arithmetic follows TFLite's integer kernels; the reference runs them inside
// interpreter.invoke() (/root/reference/src/yolact.rs:163).
#include "ops.h"

#include "common.h"

#include <algorithm>
#include <cstdlib>

#include "fixedpoint.cuh"

namespace tod {
namespace {

__device__ __forceinline__ int8_t requant_store(int32_t acc, int32_t mult, int32_t shift, const Requant& rq) {
  int32_t v = mul_by_quant_mult_fast(acc, mult, shift) + rq.out_zp;  // q >= 0, |acc| < 2^30
  v = max(rq.act_min, min(rq.act_max, v));
  if (rq.post_lut) v = __ldg(rq.post_lut + (v & 0xFF));
  return int8_t(v);
}

// ------------------------------------------------------------------ direct convolution
// One thread: one output pixel x OCT output channels.  A warp covers 32 consecutive pixels of the
// same channel group, so weight loads are warp-uniform (one transaction) and each lane streams its
// own pixel's channels.  acc = sum in*w  +  (-zp) * sum_{valid taps} wsum[tap]  + bias.
template <int OCT, bool VEC4>
__global__ void __launch_bounds__(128) conv_direct_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                         const int8_t* __restrict__ w,
                                                         const int32_t* __restrict__ bias,
                                                         const int32_t* __restrict__ wsum, int32_t in_zp, ConvGeom g,
                                                         Requant rq, int8_t* __restrict__ out, int64_t out_ts,
                                                         int tiles) {
  const int64_t pix = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(tiles) * g.OH * g.OW;
  if (pix >= total) return;
  const int oc0 = blockIdx.y * OCT;
  const int ox = int(pix % g.OW);
  const int oy = int((pix / g.OW) % g.OH);
  const int t = int(pix / (int64_t(g.OW) * g.OH));
  const int8_t* tin = in + int64_t(t) * in_ts;
  const int taps = g.KH * g.KW;
  const int64_t wstride = int64_t(taps) * g.IC;

  int32_t acc[OCT];
#pragma unroll
  for (int j = 0; j < OCT; ++j) acc[j] = 0;

  for (int fy = 0; fy < g.KH; ++fy) {
    const int iy = oy * g.stride_h - g.pad_top + fy * g.dil_h;
    if (iy < 0 || iy >= g.IH) continue;
    for (int fx = 0; fx < g.KW; ++fx) {
      const int ix = ox * g.stride_w - g.pad_left + fx * g.dil_w;
      if (ix < 0 || ix >= g.IW) continue;
      const int tap = fy * g.KW + fx;
      const int8_t* ip = tin + (int64_t(iy) * g.IW + ix) * g.IC;
      const int8_t* wp = w + int64_t(oc0) * wstride + int64_t(tap) * g.IC;
#pragma unroll
      for (int j = 0; j < OCT; ++j)
        if (oc0 + j < g.OC) acc[j] -= in_zp * __ldg(wsum + int64_t(oc0 + j) * taps + tap);
      if (VEC4) {
        for (int ic = 0; ic < g.IC; ic += 4) {
          const int a = *reinterpret_cast<const int*>(ip + ic);
#pragma unroll
          for (int j = 0; j < OCT; ++j)
            if (oc0 + j < g.OC) acc[j] = __dp4a(a, __ldg(reinterpret_cast<const int*>(wp + j * wstride + ic)), acc[j]);
        }
      } else {
        for (int ic = 0; ic < g.IC; ++ic) {
          const int a = ip[ic];
#pragma unroll
          for (int j = 0; j < OCT; ++j)
            if (oc0 + j < g.OC) acc[j] += a * int(__ldg(wp + j * wstride + ic));
        }
      }
    }
  }
  int8_t* op = out + int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + oc0;
#pragma unroll
  for (int j = 0; j < OCT; ++j) {
    if (oc0 + j < g.OC) {
      const int32_t a = acc[j] + (bias ? __ldg(bias + oc0 + j) : 0);
      op[j] = requant_store(a, __ldg(rq.mult + oc0 + j), __ldg(rq.shift + oc0 + j), rq);
    }
  }
}

// ------------------------------------------------------------------ depthwise 3x3 (any KHxKW)
// One thread: one output pixel x 4 consecutive channels (int8x4 loads, coalesced along C).
__global__ void __launch_bounds__(256) depthwise_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                       const int8_t* __restrict__ w,
                                                       const int32_t* __restrict__ bias, int32_t in_zp, ConvGeom g,
                                                       Requant rq, int8_t* __restrict__ out, int64_t out_ts,
                                                       int tiles) {
  const int c4n = g.OC >> 2;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(tiles) * g.OH * g.OW * c4n;
  if (idx >= total) return;
  const int c = int(idx % c4n) * 4;
  int64_t p = idx / c4n;
  const int ox = int(p % g.OW);
  p /= g.OW;
  const int oy = int(p % g.OH);
  const int t = int(p / g.OH);
  const int8_t* tin = in + int64_t(t) * in_ts;
  int32_t acc[4] = {0, 0, 0, 0};
  for (int fy = 0; fy < g.KH; ++fy) {
    const int iy = oy * g.stride_h - g.pad_top + fy * g.dil_h;
    if (iy < 0 || iy >= g.IH) continue;
    for (int fx = 0; fx < g.KW; ++fx) {
      const int ix = ox * g.stride_w - g.pad_left + fx * g.dil_w;
      if (ix < 0 || ix >= g.IW) continue;
      const char4 a = *reinterpret_cast<const char4*>(tin + (int64_t(iy) * g.IW + ix) * g.IC + c);
      const char4 k = __ldg(reinterpret_cast<const char4*>(w + (int64_t(fy) * g.KW + fx) * g.OC + c));
      acc[0] += (int(a.x) - in_zp) * int(k.x);
      acc[1] += (int(a.y) - in_zp) * int(k.y);
      acc[2] += (int(a.z) - in_zp) * int(k.z);
      acc[3] += (int(a.w) - in_zp) * int(k.w);
    }
  }
  char4 o;
  int8_t* ob = reinterpret_cast<int8_t*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t a = acc[j] + (bias ? __ldg(bias + c + j) : 0);
    ob[j] = requant_store(a, __ldg(rq.mult + c + j), __ldg(rq.shift + c + j), rq);
  }
  *reinterpret_cast<char4*>(out + int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + c) = o;
}

// scalar-channel variant for C % 4 != 0
__global__ void __launch_bounds__(256) depthwise_scalar_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                              const int8_t* __restrict__ w,
                                                              const int32_t* __restrict__ bias, int32_t in_zp,
                                                              ConvGeom g, Requant rq, int8_t* __restrict__ out,
                                                              int64_t out_ts, int tiles) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(tiles) * g.OH * g.OW * g.OC;
  if (idx >= total) return;
  const int c = int(idx % g.OC);
  int64_t p = idx / g.OC;
  const int ox = int(p % g.OW);
  p /= g.OW;
  const int oy = int(p % g.OH);
  const int t = int(p / g.OH);
  const int8_t* tin = in + int64_t(t) * in_ts;
  int32_t acc = 0;
  for (int fy = 0; fy < g.KH; ++fy) {
    const int iy = oy * g.stride_h - g.pad_top + fy * g.dil_h;
    if (iy < 0 || iy >= g.IH) continue;
    for (int fx = 0; fx < g.KW; ++fx) {
      const int ix = ox * g.stride_w - g.pad_left + fx * g.dil_w;
      if (ix < 0 || ix >= g.IW) continue;
      acc += (int(tin[(int64_t(iy) * g.IW + ix) * g.IC + c]) - in_zp) * int(__ldg(w + (int64_t(fy) * g.KW + fx) * g.OC + c));
    }
  }
  acc += bias ? __ldg(bias + c) : 0;
  out[int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + c] = requant_store(acc, __ldg(rq.mult + c), __ldg(rq.shift + c), rq);
}

// ------------------------------------------------------------------ pixel-parallel convolution (CUDA cores)
// For the few CONV_2D shapes the tcgen05 path does not take: the Cin = 3 stem, Cin % 16 != 0 pointwise layers and
// the strided 3x3 FPN down-samplers.  One thread = one output pixel x OCT output channels; the CTA's weight slab
// ([K/4 words][OCT], K = taps * Cin padded to 4) sits in shared memory and is read as warp-wide broadcasts
// (LDS.128 = 4 channels), so the inner loop is one input word + 4 LDS + 16 dp4a.
constexpr int kPixOct = 16;
constexpr int kPixThreads = 128;

template <bool IC3>
__global__ void __launch_bounds__(kPixThreads) conv_pix_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                              const int8_t* __restrict__ w,
                                                              const int32_t* __restrict__ bias,
                                                              const int32_t* __restrict__ wsum, int32_t in_zp, ConvGeom g,
                                                              Requant rq, int8_t* __restrict__ out, int64_t out_ts,
                                                              int tiles) {
  extern __shared__ int s_w[];  // [taps * icw][OCT] weight words, then [taps][OCT] tap sums
  pdl_trigger();
  const int taps = g.KH * g.KW;
  const int icw = (g.IC + 3) >> 2;  // input words per tap
  int* s_ws = s_w + taps * icw * kPixOct;
  const int oc0 = blockIdx.y * kPixOct;
  for (int i = threadIdx.x; i < taps * icw * kPixOct; i += kPixThreads) {
    const int j = i % kPixOct, kw = i / kPixOct;
    const int tap = kw / icw, c4 = (kw - tap * icw) * 4;
    int word = 0;
    if (oc0 + j < g.OC) {
      const int8_t* src = w + (int64_t(oc0 + j) * taps + tap) * g.IC + c4;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (c4 + b < g.IC) word |= (int(src[b]) & 0xFF) << (8 * b);
    }
    s_w[i] = word;
  }
  for (int i = threadIdx.x; i < taps * kPixOct; i += kPixThreads) {
    const int j = i % kPixOct, tap = i / kPixOct;
    s_ws[i] = (oc0 + j < g.OC) ? wsum[int64_t(oc0 + j) * taps + tap] : 0;
  }
  __syncthreads();
  pdl_wait();
  const int pix = blockIdx.x * kPixThreads + threadIdx.x;
  const int total = tiles * g.OH * g.OW;
  if (pix >= total) return;
  const int ox = pix % g.OW;
  const int oy = (pix / g.OW) % g.OH;
  const int t = pix / (g.OW * g.OH);
  const int8_t* tin = in + int64_t(t) * in_ts;
  int acc[kPixOct];
#pragma unroll
  for (int j = 0; j < kPixOct; ++j) acc[j] = 0;
  for (int fy = 0; fy < g.KH; ++fy) {
    const int iy = oy * g.stride_h - g.pad_top + fy * g.dil_h;
    if (iy < 0 || iy >= g.IH) continue;
    for (int fx = 0; fx < g.KW; ++fx) {
      const int ix = ox * g.stride_w - g.pad_left + fx * g.dil_w;
      if (ix < 0 || ix >= g.IW) continue;
      const int tap = fy * g.KW + fx;
      const int8_t* ip = tin + (int64_t(iy) * g.IW + ix) * g.IC;
      const int4* wrow = reinterpret_cast<const int4*>(s_w + tap * icw * kPixOct);
      const int4* zs = reinterpret_cast<const int4*>(s_ws + tap * kPixOct);
#pragma unroll
      for (int q = 0; q < kPixOct / 4; ++q) {
        const int4 z = zs[q];
        acc[4 * q + 0] -= in_zp * z.x;
        acc[4 * q + 1] -= in_zp * z.y;
        acc[4 * q + 2] -= in_zp * z.z;
        acc[4 * q + 3] -= in_zp * z.w;
      }
      if (IC3) {
        const int a = (int(ip[0]) & 0xFF) | ((int(ip[1]) & 0xFF) << 8) | ((int(ip[2]) & 0xFF) << 16);
#pragma unroll
        for (int q = 0; q < kPixOct / 4; ++q) {
          const int4 wv = wrow[q];
          acc[4 * q + 0] = __dp4a(a, wv.x, acc[4 * q + 0]);
          acc[4 * q + 1] = __dp4a(a, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = __dp4a(a, wv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = __dp4a(a, wv.w, acc[4 * q + 3]);
        }
      } else {
        for (int k = 0; k < icw; ++k) {
          const int a = *reinterpret_cast<const int*>(ip + 4 * k);
#pragma unroll
          for (int q = 0; q < kPixOct / 4; ++q) {
            const int4 wv = wrow[k * (kPixOct / 4) + q];
            acc[4 * q + 0] = __dp4a(a, wv.x, acc[4 * q + 0]);
            acc[4 * q + 1] = __dp4a(a, wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = __dp4a(a, wv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = __dp4a(a, wv.w, acc[4 * q + 3]);
          }
        }
      }
    }
  }
  int8_t* op = out + int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + oc0;
  int8_t res[kPixOct];
#pragma unroll
  for (int j = 0; j < kPixOct; ++j) {
    const int oc = min(oc0 + j, g.OC - 1);
    const int32_t a = acc[j] + (bias ? __ldg(bias + oc) : 0);
    res[j] = requant_store(a, __ldg(rq.mult + oc), __ldg(rq.shift + oc), rq);
  }
  if (oc0 + kPixOct <= g.OC && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
    *reinterpret_cast<int4*>(op) = *reinterpret_cast<const int4*>(res);
  } else {
#pragma unroll
    for (int j = 0; j < kPixOct; ++j)
      if (oc0 + j < g.OC) op[j] = res[j];
  }
}

// ------------------------------------------------------------------ stem: 3x3, stride 2, Cin = 3
// The RGB stem is too thin for the tensor cores (K = 27) and was the slowest CUDA-core layer.  With Cin = 3 the three
// pixels under one filter row are nine contiguous bytes, so a filter row is three (unaligned, funnel-shifted) words and
// nine dp4a per output channel cover the whole 3x3x3 window (weights padded to 12 bytes per row with zeros).
// Thread = one output pixel x 16 channels; the packed weights sit in shared memory and are read as warp-wide
// broadcasts.  Out-of-image taps (TFLite SAME puts the single padding row / column at the bottom / right here) read as
// the zero point, with the bias pre-folded as bias - zp * sum(w); requantisation is the Requant::fast_tab form.
constexpr int kStemThreads = 128;
constexpr int kStemOct = 16;
constexpr int kStemIters = 8;

template <bool SAT, bool RELU>
__global__ void __launch_bounds__(kStemThreads) stem3x3s2_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                                const int8_t* __restrict__ w, const int32_t* __restrict__ bias,
                                                                int32_t in_zp, ConvGeom g, Requant rq, int8_t* __restrict__ out,
                                                                int64_t out_ts, int tiles) {
  __shared__ int4 s_w[9 * 8 * 4];   // [filter row][word 0..2][channel group of 4 (<= 8 groups)] -> 4 channels' weight words... see fill below
  __shared__ int s_bs[64];
  __shared__ int4 s_k[64];
  pdl_trigger();
  const int OCg = g.OC;  // <= 64 here (launcher checks)
  // s_wi[(fy*3 + k) * OC + oc] = weight bytes 4k .. 4k+3 of filter row fy (9 real bytes: 3 pixels x 3 channels), 0-padded
  int* s_wi = reinterpret_cast<int*>(s_w);
  for (int i = threadIdx.x; i < 9 * OCg; i += kStemThreads) {
    const int oc = i % OCg, fk = i / OCg, fy = fk / 3, k = fk - fy * 3;
    int word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = 4 * k + b;  // byte of the 9-byte row: pixel fx = j / 3, channel j % 3
      if (j < 9) word |= (int(w[(int64_t(oc) * 9 + fy * 3 + j / 3) * 3 + j % 3]) & 0xFF) << (8 * b);
    }
    s_wi[i] = word;
  }
  for (int oc = threadIdx.x; oc < OCg; oc += kStemThreads) {
    int sum = 0;
    for (int j = 0; j < 27; ++j) sum += int(w[int64_t(oc) * 27 + j]);
    s_bs[oc] = (bias ? bias[oc] : 0) - in_zp * sum;
    s_k[oc] = rq.fast_tab[oc];
  }
  __syncthreads();
  pdl_wait();
  const int groups = OCg / kStemOct;
  const int64_t total = int64_t(tiles) * g.OH * g.OW * groups;
  // kStemIters work items per thread: the packed-weight prologue above is amortised over 1024 outputs-groups per CTA
  for (int rep = 0; rep < kStemIters; ++rep) {
  const int64_t idx = (int64_t(blockIdx.x) * kStemIters + rep) * kStemThreads + threadIdx.x;
  if (idx >= total) return;
  // 32-bit index arithmetic (the launcher guarantees total < 2^31): four 64-bit divisions per work item were as many
  // instructions as the 144 dp4a they fed
  const unsigned idx32 = unsigned(idx);
  unsigned pix = idx32 / unsigned(groups);
  const int grp = int(idx32 - pix * unsigned(groups));
  unsigned row = pix / unsigned(g.OW);
  const int ox = int(pix - row * unsigned(g.OW));
  const int t = int(row / unsigned(g.OH));
  const int oy = int(row - unsigned(t) * unsigned(g.OH));
  const int oc0 = grp * kStemOct;
  const int zp4 = (in_zp & 0xFF) * 0x01010101;
  const int8_t* tin = in + int64_t(t) * in_ts;
  const int ix0 = ox * 2, iy0 = oy * 2;  // pad_top = pad_left = 0 (launcher checks)
  int acc[kStemOct];
#pragma unroll
  for (int j = 0; j < kStemOct; ++j) acc[j] = s_bs[oc0 + j];
#pragma unroll
  for (int fy = 0; fy < 3; ++fy) {
    int a0 = zp4, a1 = zp4, a2 = zp4;
    const int iy = iy0 + fy;
    if (iy < g.IH) {
      const int64_t off = (int64_t(iy) * g.IW + ix0) * 3;          // even: 0 or 2 mod 4
      const int* p = reinterpret_cast<const int*>(tin + (off & ~int64_t(3)));
      const int xr = g.in_xor * 0x01010101;  // folded uint8 -> int8 QUANTIZE (x ^ 0x80), 0 otherwise
      const int w0 = p[0] ^ xr, w1 = p[1] ^ xr, w2 = p[2] ^ xr;
      if (off & 2) {
        a0 = __funnelshift_r(w0, w1, 16);
        a1 = __funnelshift_r(w1, w2, 16);
        a2 = int(unsigned(w2) >> 16);
      } else {
        a0 = w0;
        a1 = w1;
        a2 = w2;
      }
      // bytes of pixels to the right of the image read as the zero point (only the last output column)
      const int valid_px = g.IW - ix0;  // 1, 2 or >= 3 pixels of this row exist
      if (valid_px < 3) {
        // byte j belongs to pixel j / 3: keep the bytes of pixels < valid_px (6 bytes or 3 bytes), the rest read as zp
        const int k0 = valid_px == 2 ? int(0xFFFFFFFFu) : 0x00FFFFFF;
        const int k1 = valid_px == 2 ? 0x0000FFFF : 0;
        a0 = (a0 & k0) | (zp4 & ~k0);
        a1 = (a1 & k1) | (zp4 & ~k1);
        a2 = zp4;
      }
    }
    const int4* wr0 = reinterpret_cast<const int4*>(s_wi + (fy * 3 + 0) * OCg + oc0);
    const int4* wr1 = reinterpret_cast<const int4*>(s_wi + (fy * 3 + 1) * OCg + oc0);
    const int4* wr2 = reinterpret_cast<const int4*>(s_wi + (fy * 3 + 2) * OCg + oc0);
#pragma unroll
    for (int q = 0; q < kStemOct / 4; ++q) {
      const int4 u0 = wr0[q], u1 = wr1[q], u2 = wr2[q];
      acc[4 * q + 0] = __dp4a(a2, u2.x, __dp4a(a1, u1.x, __dp4a(a0, u0.x, acc[4 * q + 0])));
      acc[4 * q + 1] = __dp4a(a2, u2.y, __dp4a(a1, u1.y, __dp4a(a0, u0.y, acc[4 * q + 1])));
      acc[4 * q + 2] = __dp4a(a2, u2.z, __dp4a(a1, u1.z, __dp4a(a0, u0.z, acc[4 * q + 2])));
      acc[4 * q + 3] = __dp4a(a2, u2.w, __dp4a(a1, u1.w, __dp4a(a0, u0.w, acc[4 * q + 3])));
    }
  }
  unsigned packed[kStemOct / 4];
#pragma unroll
  for (int q = 0; q < kStemOct / 4; ++q) {
    int o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int4 k = s_k[oc0 + 4 * q + j];
      if (RELU) o[j] = requant_relu(acc[4 * q + j], k.x, k.y, (int64_t(k.w) << 32) | int64_t(uint32_t(k.z)));
      else o[j] = requant_tab(acc[4 * q + j], k.x, k.y, k.w);
    }
    if (SAT) {
      unsigned hi;
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(o[3]), "r"(o[2]), "r"(0u));
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(packed[q]) : "r"(o[1]), "r"(o[0]), "r"(hi));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = max(rq.act_min, min(rq.act_max, o[j]));
      packed[q] = (unsigned(o[0]) & 0xFFu) | ((unsigned(o[1]) & 0xFFu) << 8) | ((unsigned(o[2]) & 0xFFu) << 16) | (unsigned(o[3]) << 24);
    }
  }
  *reinterpret_cast<uint4*>(out + int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + oc0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// ------------------------------------------------------------------ stem, two output pixels per thread (default)
// ncu on the kernel above: ~500 SASS instructions per (pixel, 16 channels) item for 144 dp4a - four 32-bit divisions to decode
// the item, 36 weight LDS.128, 16 constant LDS.128, the input words and their funnel shifts.  Here a thread owns two horizontally
// adjacent output pixels of one row (their windows share one of five input pixels: 15 contiguous bytes per filter row, and
// 12 * pair is a whole number of words), so every weight / constant load feeds two outputs, and the item is decoded once per
// thread: (pair, channel group) from the thread index, then kStemRows consecutive output rows by pointer increments.
constexpr int kStemRows = 8;

template <bool SAT, bool RELU>
__global__ void __launch_bounds__(kStemThreads) stem3x3s2_pair_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                                     const int8_t* __restrict__ w, const int32_t* __restrict__ bias,
                                                                     int32_t in_zp, ConvGeom g, Requant rq, int8_t* __restrict__ out,
                                                                     int64_t out_ts, int tiles, int rows_per_cta) {
  __shared__ int4 s_w[9 * 8 * 4];
  __shared__ int s_bs[64];
  __shared__ int4 s_k[64];
  pdl_trigger();
  const int OCg = g.OC;  // <= 64 (launcher checks)
  int* s_wi = reinterpret_cast<int*>(s_w);   // s_wi[(fy*3 + k) * OC + oc] = weight bytes 4k .. 4k+3 of filter row fy (9 real bytes), 0-padded
  for (int i = threadIdx.x; i < 9 * OCg; i += kStemThreads) {
    const int oc = i % OCg, fk = i / OCg, fy = fk / 3, k = fk - fy * 3;
    int word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = 4 * k + b;
      if (j < 9) word |= (int(w[(int64_t(oc) * 9 + fy * 3 + j / 3) * 3 + j % 3]) & 0xFF) << (8 * b);
    }
    s_wi[i] = word;
  }
  for (int oc = threadIdx.x; oc < OCg; oc += kStemThreads) {
    int sum = 0;
    for (int j = 0; j < 27; ++j) sum += int(w[int64_t(oc) * 27 + j]);
    s_bs[oc] = (bias ? bias[oc] : 0) - in_zp * sum;
    s_k[oc] = rq.fast_tab[oc];
  }
  __syncthreads();
  pdl_wait();
  const int groups = OCg / kStemOct;
  const int pairs = (g.OW + 1) >> 1;
  const int ig = blockIdx.x * kStemThreads + threadIdx.x;
  if (ig >= pairs * groups) return;
  const int pr = ig / groups, grp = ig - pr * groups;
  const int ox = 2 * pr, oc0 = grp * kStemOct;
  const bool has_b = ox + 1 < g.OW;
  const int ix0 = 2 * ox;                       // pad_left == 0; the pair's windows cover input columns ix0 .. ix0 + 4
  const int valid_a = g.IW - ix0, valid_b = g.IW - ix0 - 2;  // pixels of each window inside the image (>= 3: all)
  const int zp4 = (in_zp & 0xFF) * 0x01010101;
  const int xr = g.in_xor * 0x01010101;         // folded uint8 -> int8 QUANTIZE (x ^ 0x80), 0 otherwise
  const int rows_total = tiles * g.OH;
  int row = blockIdx.y * rows_per_cta;
  if (row >= rows_total) return;
  int t = row / g.OH, oy = row - t * g.OH;
  // bytes of pixels to the right of the image read as the zero point: keep the bytes of the first `valid` pixels of a window
  auto clip = [&](int valid, int& x0, int& x1, int& x2) {
    if (valid >= 3) return;
    const int k0 = valid == 2 ? int(0xFFFFFFFFu) : (valid == 1 ? 0x00FFFFFF : 0);
    const int k1 = valid == 2 ? 0x0000FFFF : 0;
    x0 = (x0 & k0) | (zp4 & ~k0);
    x1 = (x1 & k1) | (zp4 & ~k1);
    x2 = zp4;
  };
  for (int rep = 0; rep < rows_per_cta && row < rows_total; ++rep, ++row) {
    const int8_t* tin = in + int64_t(t) * in_ts;
    int acc_a[kStemOct], acc_b[kStemOct];
#pragma unroll
    for (int j = 0; j < kStemOct; ++j) acc_a[j] = acc_b[j] = s_bs[oc0 + j];
#pragma unroll
    for (int fy = 0; fy < 3; ++fy) {
      int a0 = zp4, a1 = zp4, a2 = zp4, b0 = zp4, b1 = zp4, b2 = zp4;
      const int iy = 2 * oy + fy;               // pad_top == 0
      if (iy < g.IH) {
        const int64_t off = (int64_t(iy) * g.IW + ix0) * 3;          // even: 0 or 2 mod 4
        const int* p = reinterpret_cast<const int*>(tin + (off & ~int64_t(3)));
        const int sh = int(off & 3) * 8;
        const int w0 = p[0] ^ xr, w1 = p[1] ^ xr, w2 = p[2] ^ xr, w3 = p[3] ^ xr, w4 = p[4] ^ xr;
        const int c0 = __funnelshift_r(w0, w1, sh), c1 = __funnelshift_r(w1, w2, sh), c2 = __funnelshift_r(w2, w3, sh), c3 = __funnelshift_r(w3, w4, sh);
        a0 = c0;                              // window A: bytes 0 .. 8 (the weight words are zero beyond byte 8)
        a1 = c1;
        a2 = c2;
        b0 = __funnelshift_r(c1, c2, 16);     // window B: bytes 6 .. 14
        b1 = __funnelshift_r(c2, c3, 16);
        b2 = int(unsigned(c3) >> 16);
        clip(valid_a, a0, a1, a2);
        clip(valid_b, b0, b1, b2);
      }
      const int4* wr0 = reinterpret_cast<const int4*>(s_wi + (fy * 3 + 0) * OCg + oc0);
      const int4* wr1 = reinterpret_cast<const int4*>(s_wi + (fy * 3 + 1) * OCg + oc0);
      const int4* wr2 = reinterpret_cast<const int4*>(s_wi + (fy * 3 + 2) * OCg + oc0);
#pragma unroll
      for (int q = 0; q < kStemOct / 4; ++q) {
        const int4 u0 = wr0[q], u1 = wr1[q], u2 = wr2[q];
        acc_a[4 * q + 0] = __dp4a(a2, u2.x, __dp4a(a1, u1.x, __dp4a(a0, u0.x, acc_a[4 * q + 0])));
        acc_a[4 * q + 1] = __dp4a(a2, u2.y, __dp4a(a1, u1.y, __dp4a(a0, u0.y, acc_a[4 * q + 1])));
        acc_a[4 * q + 2] = __dp4a(a2, u2.z, __dp4a(a1, u1.z, __dp4a(a0, u0.z, acc_a[4 * q + 2])));
        acc_a[4 * q + 3] = __dp4a(a2, u2.w, __dp4a(a1, u1.w, __dp4a(a0, u0.w, acc_a[4 * q + 3])));
        acc_b[4 * q + 0] = __dp4a(b2, u2.x, __dp4a(b1, u1.x, __dp4a(b0, u0.x, acc_b[4 * q + 0])));
        acc_b[4 * q + 1] = __dp4a(b2, u2.y, __dp4a(b1, u1.y, __dp4a(b0, u0.y, acc_b[4 * q + 1])));
        acc_b[4 * q + 2] = __dp4a(b2, u2.z, __dp4a(b1, u1.z, __dp4a(b0, u0.z, acc_b[4 * q + 2])));
        acc_b[4 * q + 3] = __dp4a(b2, u2.w, __dp4a(b1, u1.w, __dp4a(b0, u0.w, acc_b[4 * q + 3])));
      }
    }
    unsigned pk_a[kStemOct / 4], pk_b[kStemOct / 4];
#pragma unroll
    for (int q = 0; q < kStemOct / 4; ++q) {
      int oa[4], ob[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int4 k = s_k[oc0 + 4 * q + j];
        if (RELU) {
          const int64_t ad = (int64_t(k.w) << 32) | int64_t(uint32_t(k.z));
          oa[j] = requant_relu(acc_a[4 * q + j], k.x, k.y, ad);
          ob[j] = requant_relu(acc_b[4 * q + j], k.x, k.y, ad);
        } else {
          oa[j] = requant_tab(acc_a[4 * q + j], k.x, k.y, k.w);
          ob[j] = requant_tab(acc_b[4 * q + j], k.x, k.y, k.w);
        }
      }
      if (SAT) {
        unsigned hi;
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(oa[3]), "r"(oa[2]), "r"(0u));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(pk_a[q]) : "r"(oa[1]), "r"(oa[0]), "r"(hi));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(ob[3]), "r"(ob[2]), "r"(0u));
        asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(pk_b[q]) : "r"(ob[1]), "r"(ob[0]), "r"(hi));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          oa[j] = max(rq.act_min, min(rq.act_max, oa[j]));
          ob[j] = max(rq.act_min, min(rq.act_max, ob[j]));
        }
        pk_a[q] = (unsigned(oa[0]) & 0xFFu) | ((unsigned(oa[1]) & 0xFFu) << 8) | ((unsigned(oa[2]) & 0xFFu) << 16) | (unsigned(oa[3]) << 24);
        pk_b[q] = (unsigned(ob[0]) & 0xFFu) | ((unsigned(ob[1]) & 0xFFu) << 8) | ((unsigned(ob[2]) & 0xFFu) << 16) | (unsigned(ob[3]) << 24);
      }
    }
    int8_t* op = out + int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + oc0;
    *reinterpret_cast<uint4*>(op) = make_uint4(pk_a[0], pk_a[1], pk_a[2], pk_a[3]);
    if (has_b) *reinterpret_cast<uint4*>(op + g.OC) = make_uint4(pk_b[0], pk_b[1], pk_b[2], pk_b[3]);
    if (++oy == g.OH) {
      oy = 0;
      ++t;
    }
  }
}

// ------------------------------------------------------------------ depthwise, register-resident filters
// blockDim = (channel groups of 4, pixel lanes).  A thread keeps the 3x3 taps of its 4 channels (pre-masked so a
// dp4a isolates one channel), their bias and requantisation constants in registers and walks over output pixels;
// consecutive threads cover consecutive channels, so every global access of a warp is one contiguous segment.
template <int KK>
__global__ void __launch_bounds__(256, 2) depthwise_reg_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                              const int8_t* __restrict__ w,
                                                              const int32_t* __restrict__ bias, int32_t in_zp, ConvGeom g,
                                                              Requant rq, int8_t* __restrict__ out, int64_t out_ts,
                                                              int tiles, int pix_per_block) {
  const int cgi = blockIdx.x * blockDim.x + threadIdx.x;  // channel group
  if (cgi * 4 >= g.OC) return;
  const int c = cgi * 4;
  int wm[KK * KK][4];  // tap weights, masked to one byte lane each: dp4a(a, wm[tap][j]) == a_j * w_j
  int wall[4] = {0, 0, 0, 0};
#pragma unroll
  for (int tp = 0; tp < KK * KK; ++tp) {
    const int wv = *reinterpret_cast<const int*>(w + int64_t(tp) * g.OC + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wm[tp][j] = wv & (0xFF << (8 * j));
      wall[j] += (wv << (24 - 8 * j)) >> 24;
    }
  }
  int mult[4], shift[4], bs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mult[j] = rq.mult[c + j];
    shift[j] = rq.shift[c + j];
    bs[j] = (bias ? bias[c + j] : 0) - in_zp * wall[j];  // interior pixels: every tap is inside
  }
  const int total = tiles * g.OH * g.OW;
  const int p_end = min(total, (blockIdx.y + 1) * pix_per_block);
  int pix = blockIdx.y * pix_per_block + threadIdx.y;
  int ox = pix % g.OW;
  int oy = (pix / g.OW) % g.OH;
  int t = pix / (g.OW * g.OH);
  for (; pix < p_end; pix += blockDim.y) {
    const int8_t* tin = in + int64_t(t) * in_ts + c;
    const int iy0 = oy * g.stride_h - g.pad_top, ix0 = ox * g.stride_w - g.pad_left;
    int acc[4];
    if (iy0 >= 0 && ix0 >= 0 && iy0 + KK <= g.IH && ix0 + KK <= g.IW) {
      int a[KK * KK];
      const int8_t* p0 = tin + (int64_t(iy0) * g.IW + ix0) * g.IC;
#pragma unroll
      for (int fy = 0; fy < KK; ++fy)
#pragma unroll
        for (int fx = 0; fx < KK; ++fx) a[fy * KK + fx] = *reinterpret_cast<const int*>(p0 + (fy * g.IW + fx) * g.IC);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = bs[j];
#pragma unroll
      for (int tp = 0; tp < KK * KK; ++tp)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = __dp4a(a[tp], wm[tp][j], acc[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = bs[j] + in_zp * wall[j];
#pragma unroll
      for (int fy = 0; fy < KK; ++fy)
#pragma unroll
        for (int fx = 0; fx < KK; ++fx) {
          const int iy = iy0 + fy, ix = ix0 + fx;
          if (iy < 0 || iy >= g.IH || ix < 0 || ix >= g.IW) continue;
          const int av = *reinterpret_cast<const int*>(tin + (int64_t(iy) * g.IW + ix) * g.IC);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j] = __dp4a(av, wm[fy * KK + fx][j], acc[j]);
            acc[j] -= in_zp * ((wm[fy * KK + fx][j] << (24 - 8 * j)) >> 24);
          }
        }
    }
    unsigned packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int32_t v = mul_by_quant_mult_fast(acc[j], mult[j], shift[j]) + rq.out_zp;
      v = max(rq.act_min, min(rq.act_max, v));
      if (rq.post_lut) v = __ldg(rq.post_lut + (v & 0xFF));
      packed |= (unsigned(v) & 0xFFu) << (8 * j);
    }
    *reinterpret_cast<unsigned*>(out + int64_t(t) * out_ts + (int64_t(oy) * g.OW + ox) * g.OC + c) = packed;
    // next pixel of this thread (pix += blockDim.y) without divisions
    ox += blockDim.y;
    while (ox >= g.OW) {
      ox -= g.OW;
      if (++oy == g.OH) {
        oy = 0;
        ++t;
      }
    }
  }
}

// ------------------------------------------------------------------ depthwise 3x3, sliding window
// Thread = one 4-channel word of one output column, walking down a run of output rows.  The (column, channel word)
// pair is a flattened index along the NHWC row, so consecutive lanes read consecutive words whatever C is.
// A depthwise tap multiplies channel by channel, so a dp4a over a loaded word (4 channels of ONE tap) wastes three of
// its four products.  Each input row's three words (columns x-1, x, x+1) are therefore transposed with six PRMT into
// four words holding the three taps of ONE channel (fourth byte multiplied by a zero weight): a filter row then costs
// one dp4a per channel, 12 per output word instead of 36 (ncu on the 36-dp4a form: fmaheavy 54 %, issue 54 %, 95
// registers).  The transposed rows slide vertically: a stride-1 output row costs one new input row (3 loads, 6 PRMT),
// a stride-2 row two.  Out-of-image taps read as the input zero point, which makes their (in - zp) * w term vanish
// with the bias pre-folded as bias - zp * sum(w): no branches in the arithmetic.  Requantisation: Requant::fast_tab.
constexpr int kDwThreads = 128;

template <int STRIDE, bool SAT, bool RELU>
__global__ void __launch_bounds__(kDwThreads, 8) depthwise3x3_slide_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                                       const int8_t* __restrict__ w,
                                                                       const int32_t* __restrict__ bias, int32_t in_zp,
                                                                       ConvGeom g, Requant rq, int8_t* __restrict__ out,
                                                                       int64_t out_ts, int rows_per_block) {
  pdl_trigger();
  const int C = g.OC, CW = C >> 2;
  const int j = blockIdx.x * kDwThreads + threadIdx.x;
  if (j >= g.OW * CW) return;
  const int ox = j / CW, c = (j - ox * CW) * 4;
  // three input words (taps fx = 0, 1, 2; four channels each) -> four words (channel q; bytes = taps 0, 1, 2, don't care)
  auto transpose = [](unsigned a0, unsigned a1, unsigned a2, unsigned (&t)[4]) {
    const unsigned lo = __byte_perm(a0, a1, 0x5140);  // a0.b0 a1.b0 a0.b1 a1.b1
    const unsigned hi = __byte_perm(a0, a1, 0x7362);  // a0.b2 a1.b2 a0.b3 a1.b3
    t[0] = __byte_perm(lo, a2, 0x4410);
    t[1] = __byte_perm(lo, a2, 0x5532);
    t[2] = __byte_perm(hi, a2, 0x6610);
    t[3] = __byte_perm(hi, a2, 0x7732);
  };
  unsigned wt[3][4];  // wt[fy][q] = {w(fy,0,q), w(fy,1,q), w(fy,2,q), 0}
  int wall[4] = {0, 0, 0, 0};
#pragma unroll
  for (int fy = 0; fy < 3; ++fy) {
    unsigned wv[3];
#pragma unroll
    for (int fx = 0; fx < 3; ++fx) {
      wv[fx] = unsigned(__ldg(reinterpret_cast<const int*>(w + int64_t(fy * 3 + fx) * C + c)));
#pragma unroll
      for (int q = 0; q < 4; ++q) wall[q] += int(wv[fx] << (24 - 8 * q)) >> 24;
    }
    transpose(wv[0], wv[1], wv[2], wt[fy]);
#pragma unroll
    for (int q = 0; q < 4; ++q) wt[fy][q] &= 0x00FFFFFFu;
  }
  int4 k[4];
  int bs[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    k[q] = __ldg(rq.fast_tab + c + q);
    bs[q] = (bias ? __ldg(bias + c + q) : 0) - in_zp * wall[q];
  }
  const unsigned zp4 = unsigned(in_zp & 0xFF) * 0x01010101u;
  const int ix0 = ox * STRIDE - g.pad_left;
  const bool vx0 = ix0 >= 0 && ix0 < g.IW, vx1 = ix0 + 1 >= 0 && ix0 + 1 < g.IW, vx2 = ix0 + 2 >= 0 && ix0 + 2 < g.IW;
  const int8_t* base = in + int64_t(blockIdx.z) * in_ts + int64_t(ix0) * C + c;
  const int64_t row_pitch = int64_t(g.IW) * C;
  auto load_raw = [&](int iy, unsigned (&a)[3]) {
    a[0] = a[1] = a[2] = zp4;
    if (iy >= 0 && iy < g.IH) {
      const int8_t* p = base + int64_t(iy) * row_pitch;
      if (vx0) a[0] = *reinterpret_cast<const unsigned*>(p);
      if (vx1) a[1] = *reinterpret_cast<const unsigned*>(p + C);
      if (vx2) a[2] = *reinterpret_cast<const unsigned*>(p + 2 * C);
    }
  };
  auto load_row = [&](int iy, unsigned (&t)[4]) {
    unsigned a[3];
    load_raw(iy, a);
    transpose(a[0], a[1], a[2], t);
  };
  const int oy0 = blockIdx.y * rows_per_block, oy1 = min(g.OH, oy0 + rows_per_block);
  int iy = oy0 * STRIDE - g.pad_top;
  int8_t* op = out + int64_t(blockIdx.z) * out_ts + (int64_t(oy0) * g.OW + ox) * C + c;
  const int64_t ostep = int64_t(g.OW) * C;
  // one output word from three transposed window rows
  auto emit = [&](const unsigned (&ra)[4], const unsigned (&rb)[4], const unsigned (&rc)[4], int8_t* dst) {
    int o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int a = __dp4a(int(ra[q]), int(wt[0][q]), bs[q]);
      a = __dp4a(int(rb[q]), int(wt[1][q]), a);
      a = __dp4a(int(rc[q]), int(wt[2][q]), a);
      if (RELU) o[q] = requant_relu(a, k[q].x, k[q].y, (int64_t(k[q].w) << 32) | int64_t(uint32_t(k[q].z)));
      else o[q] = requant_tab(a, k[q].x, k[q].y, k[q].w);
    }
    unsigned packed;
    if (SAT) {
      unsigned hi;
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(o[3]), "r"(o[2]), "r"(0u));
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(packed) : "r"(o[1]), "r"(o[0]), "r"(hi));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = max(rq.act_min, min(rq.act_max, o[q]));
      packed = (unsigned(o[0]) & 0xFFu) | ((unsigned(o[1]) & 0xFFu) << 8) | ((unsigned(o[2]) & 0xFFu) << 16) | (unsigned(o[3]) << 24);
    }
    *reinterpret_cast<unsigned*>(dst) = packed;
  };
  pdl_wait();  // filters / requantisation constants are in registers; the activations come next
  if (STRIDE == 1) {
    // four register rows rotate through the roles (top, middle, bottom, incoming): the loop is unrolled by four so the
    // rotation is pure renaming; the next row's loads are issued before the current row's arithmetic and transposed after
    unsigned ra[4], rb[4], rc[4], rd[4], nx[3];
    load_row(iy, ra);
    load_row(iy + 1, rb);
    load_row(iy + 2, rc);
    for (int oy = oy0; oy < oy1; oy += 4, iy += 4, op += 4 * ostep) {
      if (oy + 1 < oy1) load_raw(iy + 3, nx);
      emit(ra, rb, rc, op);
      if (oy + 1 >= oy1) break;
      transpose(nx[0], nx[1], nx[2], rd);
      if (oy + 2 < oy1) load_raw(iy + 4, nx);
      emit(rb, rc, rd, op + ostep);
      if (oy + 2 >= oy1) break;
      transpose(nx[0], nx[1], nx[2], ra);
      if (oy + 3 < oy1) load_raw(iy + 5, nx);
      emit(rc, rd, ra, op + 2 * ostep);
      if (oy + 3 >= oy1) break;
      transpose(nx[0], nx[1], nx[2], rb);
      if (oy + 4 < oy1) load_raw(iy + 6, nx);
      emit(rd, ra, rb, op + 3 * ostep);
      if (oy + 4 < oy1) transpose(nx[0], nx[1], nx[2], rc);
    }
  } else {
    unsigned r0[4], r1[4], r2[4], n1[3], n2[3];
    load_row(iy, r0);
    load_row(iy + 1, r1);
    load_row(iy + 2, r2);
    for (int oy = oy0; oy < oy1; ++oy, iy += STRIDE, op += ostep) {
      const bool more = oy + 1 < oy1;
      if (more) {
        load_raw(iy + 3, n1);
        load_raw(iy + 4, n2);
      }
      emit(r0, r1, r2, op);
      if (more) {
#pragma unroll
        for (int q = 0; q < 4; ++q) r0[q] = r2[q];
        transpose(n1[0], n1[1], n1[2], r1);
        transpose(n2[0], n2[1], n2[2], r2);
      }
    }
  }
}

// ------------------------------------------------------------------ depthwise 3x3, lean form (default)
// ncu on the sliding kernel above (round 2, whole-step capture): 107 SASS instructions per output word of which ~32 are
// the arithmetic (3 loads, 6 PRMT, 12 dp4a, 8 requantisation, 2 pack, 1 store); issue 57 - 67 %, ALU pipe 45 - 54 %,
// "not selected" / "math pipe throttle" on top - the rest was 64-bit address arithmetic recomputed per load, per-load
// column / row predicates with their zero-point moves, and weight masking inside the loop.  Same dataflow here, with
//   * the channel count as a template parameter (CWT = channels / 4; 0 = any): the three taps of a row are
//     [p], [p + C], [p + 2C] off ONE pointer that advances by the row pitch;
//   * no column predicates: a thread on the left / right image border shifts its three columns into the image and
//     permutes its filter columns to match (the column that left the window gets zero weights, and the bias term
//     bias - zp * sum(w) only sums the taps that are inside), so every load is unconditional in x;
//   * rows outside the image read as the zero point behind a warp-uniform test (the row index depends on blockIdx.y and
//     the loop counter only).
// Bytes are identical to the sliding kernel (tests/test_gpu_graph.py, test_gpu_property.py run both against the oracle).
template <int STRIDE, bool SAT, bool RELU, int CWT>
__global__ void __launch_bounds__(kDwThreads, 6) depthwise3x3_lean_kernel(const int8_t* __restrict__ in, int64_t in_ts,
                                                                      const int8_t* __restrict__ w,
                                                                      const int32_t* __restrict__ bias, int32_t in_zp,
                                                                      ConvGeom g, Requant rq, int8_t* __restrict__ out,
                                                                      int64_t out_ts, int rows_per_block) {
  pdl_trigger();
  const int CW = CWT ? CWT : (g.OC >> 2);
  const int C = CW * 4;
  const int j = blockIdx.x * kDwThreads + threadIdx.x;
  if (j >= g.OW * CW) return;
  const int ox = j / CW, c = (j - ox * CW) * 4;
  auto transpose = [](unsigned a0, unsigned a1, unsigned a2, unsigned (&t)[4]) {
    const unsigned lo = __byte_perm(a0, a1, 0x5140);
    const unsigned hi = __byte_perm(a0, a1, 0x7362);
    t[0] = __byte_perm(lo, a2, 0x4410);
    t[1] = __byte_perm(lo, a2, 0x5532);
    t[2] = __byte_perm(hi, a2, 0x6610);
    t[3] = __byte_perm(hi, a2, 0x7732);
  };
  // window columns ix0 .. ix0 + 2; a border thread shifts them into the image (the host guarantees IW >= 3 and at most one
  // column outside on either side) and reads its filter columns through the same shift
  int ix0 = ox * STRIDE - g.pad_left;
  const int shift = ix0 < 0 ? 1 : (ix0 + 2 >= g.IW ? -1 : 0);
  ix0 += shift;
  unsigned wt[3][4];  // wt[fy][q] = {w(fy, col 0, q), w(fy, col 1, q), w(fy, col 2, q), 0} for the shifted columns
  int wall[4] = {0, 0, 0, 0};
#pragma unroll
  for (int fy = 0; fy < 3; ++fy) {
    unsigned wv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int fx = k + shift;
      wv[k] = (fx >= 0 && fx < 3) ? unsigned(__ldg(reinterpret_cast<const int*>(w + int64_t(fy * 3 + fx) * C + c))) : 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) wall[q] += int(wv[k] << (24 - 8 * q)) >> 24;
    }
    transpose(wv[0], wv[1], wv[2], wt[fy]);
#pragma unroll
    for (int q = 0; q < 4; ++q) wt[fy][q] &= 0x00FFFFFFu;
  }
  int4 k[4];
  int bs[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    k[q] = __ldg(rq.fast_tab + c + q);
    bs[q] = (bias ? __ldg(bias + c + q) : 0) - in_zp * wall[q];
  }
  const unsigned zp4 = unsigned(in_zp & 0xFF) * 0x01010101u;
  const int oy0 = blockIdx.y * rows_per_block;
  int n = min(g.OH, oy0 + rows_per_block) - oy0;   // output rows of this thread
  int iy = oy0 * STRIDE - g.pad_top;               // first input row of the window
  const int pitch = g.IW * CW;                     // input row pitch in words
  const unsigned* p = reinterpret_cast<const unsigned*>(in + int64_t(blockIdx.z) * in_ts + int64_t(ix0) * C + c) + int64_t(iy) * pitch;
  unsigned* op = reinterpret_cast<unsigned*>(out + int64_t(blockIdx.z) * out_ts + (int64_t(oy0) * g.OW + ox) * C + c);
  const int ostep = g.OW * CW;                     // output row pitch in words
  const int IH = g.IH;
  auto load_raw = [&](int row, const unsigned* q, unsigned (&a)[3]) {  // `row` is warp-uniform
    if (unsigned(row) < unsigned(IH)) {
      a[0] = q[0];
      a[1] = q[CW];
      a[2] = q[2 * CW];
    } else {
      a[0] = a[1] = a[2] = zp4;
    }
  };
  auto emit = [&](const unsigned (&ra)[4], const unsigned (&rb)[4], const unsigned (&rc)[4], unsigned* dst) {
    int o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int a = __dp4a(int(ra[q]), int(wt[0][q]), bs[q]);
      a = __dp4a(int(rb[q]), int(wt[1][q]), a);
      a = __dp4a(int(rc[q]), int(wt[2][q]), a);
      if (RELU) o[q] = requant_relu(a, k[q].x, k[q].y, (int64_t(k[q].w) << 32) | int64_t(uint32_t(k[q].z)));
      else o[q] = requant_tab(a, k[q].x, k[q].y, k[q].w);
    }
    unsigned packed;
    if (SAT) {
      unsigned hi;
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(o[3]), "r"(o[2]), "r"(0u));
      asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(packed) : "r"(o[1]), "r"(o[0]), "r"(hi));
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q] = max(rq.act_min, min(rq.act_max, o[q]));
      packed = (unsigned(o[0]) & 0xFFu) | ((unsigned(o[1]) & 0xFFu) << 8) | ((unsigned(o[2]) & 0xFFu) << 16) | (unsigned(o[3]) << 24);
    }
    *dst = packed;
  };
  pdl_wait();  // filters / requantisation constants are in registers; the activations come next
  unsigned nx[3];
  if (STRIDE == 1) {
    unsigned ra[4], rb[4], rc[4], rd[4];
    load_raw(iy, p, nx);
    transpose(nx[0], nx[1], nx[2], ra);
    load_raw(iy + 1, p + pitch, nx);
    transpose(nx[0], nx[1], nx[2], rb);
    load_raw(iy + 2, p + 2 * pitch, nx);
    transpose(nx[0], nx[1], nx[2], rc);
    p += 3 * int64_t(pitch);
    iy += 3;  // the next row to load
    // four register rows rotate through the roles (top, middle, bottom, incoming); the next row's loads are issued before
    // the current row's arithmetic and transposed after it.  (A row past the thread's last window is loaded and dropped.)
    for (; n >= 4; n -= 4) {
      load_raw(iy, p, nx);
      emit(ra, rb, rc, op);
      transpose(nx[0], nx[1], nx[2], rd);
      load_raw(iy + 1, p + pitch, nx);
      emit(rb, rc, rd, op + ostep);
      transpose(nx[0], nx[1], nx[2], ra);
      load_raw(iy + 2, p + 2 * pitch, nx);
      emit(rc, rd, ra, op + 2 * ostep);
      transpose(nx[0], nx[1], nx[2], rb);
      load_raw(iy + 3, p + 3 * pitch, nx);
      emit(rd, ra, rb, op + 3 * ostep);
      transpose(nx[0], nx[1], nx[2], rc);
      p += 4 * int64_t(pitch);
      op += 4 * int64_t(ostep);
      iy += 4;
    }
    for (; n > 0; --n) {
      load_raw(iy, p, nx);
      emit(ra, rb, rc, op);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ra[q] = rb[q];
        rb[q] = rc[q];
      }
      transpose(nx[0], nx[1], nx[2], rc);
      p += pitch;
      op += ostep;
      ++iy;
    }
  } else {
    unsigned r0[4], r1[4], r2[4], n2[3];
    load_raw(iy, p, nx);
    transpose(nx[0], nx[1], nx[2], r0);
    load_raw(iy + 1, p + pitch, nx);
    transpose(nx[0], nx[1], nx[2], r1);
    load_raw(iy + 2, p + 2 * pitch, nx);
    transpose(nx[0], nx[1], nx[2], r2);
    p += 3 * int64_t(pitch);
    iy += 3;
    for (; n > 0; --n) {
      load_raw(iy, p, nx);
      load_raw(iy + 1, p + pitch, n2);
      emit(r0, r1, r2, op);
#pragma unroll
      for (int q = 0; q < 4; ++q) r0[q] = r2[q];
      transpose(nx[0], nx[1], nx[2], r1);
      transpose(n2[0], n2[1], n2[2], r2);
      p += 2 * int64_t(pitch);
      op += ostep;
      iy += 2;
    }
  }
}

// ------------------------------------------------------------------ ADD (residual / FPN merge)
// The two input rescales depend on one byte each, so they are tabulated per CTA (256 entries each, built with
// the literal arithmetic); the output rescale uses the fast exact form (|ya + yb| < 2^29).
__device__ __forceinline__ int8_t add_one(int a, int b, const int* s_a, const int* s_b, const AddParams& p) {
  int32_t r = mul_by_quant_mult_fast(s_a[a & 0xFF] + s_b[b & 0xFF], p.mult_out, p.shift_out) + p.zp_out;
  r = max(p.act_min, min(p.act_max, r));
  return int8_t(r);
}

__global__ void __launch_bounds__(256) add_kernel(const int8_t* __restrict__ a, int64_t a_ts,
                                                 const int8_t* __restrict__ b, int64_t b_ts, int8_t* __restrict__ out,
                                                 int64_t out_ts, int64_t elems, int tiles, AddParams p, bool vec) {
  __shared__ int s_a[256], s_b[256];
  {
    const int v = int(int8_t(threadIdx.x));  // table index = the byte pattern
    s_a[threadIdx.x] = mul_by_quant_mult((v - p.zp_a) * (1 << 20), p.mult_a, p.shift_a);
    s_b[threadIdx.x] = mul_by_quant_mult((v - p.zp_b) * (1 << 20), p.mult_b, p.shift_b);
  }
  __syncthreads();
  const int t = blockIdx.y;
  const int8_t* pa = a + int64_t(t) * a_ts;
  const int8_t* pb = b + int64_t(t) * b_ts;
  int8_t* po = out + int64_t(t) * out_ts;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  if (vec) {  // 16 bytes per thread-iteration
    const int64_t n16 = elems >> 4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
      const int4 va = reinterpret_cast<const int4*>(pa)[i];
      const int4 vb = reinterpret_cast<const int4*>(pb)[i];
      int4 vo;
      const int8_t* ba = reinterpret_cast<const int8_t*>(&va);
      const int8_t* bb = reinterpret_cast<const int8_t*>(&vb);
      int8_t* bo = reinterpret_cast<int8_t*>(&vo);
#pragma unroll
      for (int j = 0; j < 16; ++j) bo[j] = add_one(ba[j], bb[j], s_a, s_b, p);
      reinterpret_cast<int4*>(po)[i] = vo;
    }
    for (int64_t i = (n16 << 4) + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < elems; i += stride)
      po[i] = add_one(pa[i], pb[i], s_a, s_b, p);
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < elems; i += stride) po[i] = add_one(pa[i], pb[i], s_a, s_b, p);
  }
}

// ------------------------------------------------------------------ byte LUT (QUANTIZE / RELU / TANH)
__global__ void __launch_bounds__(256) lut_kernel(const uint8_t* __restrict__ in, int64_t in_ts,
                                                 uint8_t* __restrict__ out, int64_t out_ts, int64_t elems,
                                                 const uint8_t* __restrict__ lut, bool vec) {
  __shared__ uint8_t s_lut[256];
  s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const int t = blockIdx.y;
  const uint8_t* pi = in + int64_t(t) * in_ts;
  uint8_t* po = out + int64_t(t) * out_ts;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  if (vec) {
    const int64_t n16 = elems >> 4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
      const uint4 v = reinterpret_cast<const uint4*>(pi)[i];
      uint4 o;
      const uint8_t* bv = reinterpret_cast<const uint8_t*>(&v);
      uint8_t* bo = reinterpret_cast<uint8_t*>(&o);
#pragma unroll
      for (int j = 0; j < 16; ++j) bo[j] = s_lut[bv[j]];
      reinterpret_cast<uint4*>(po)[i] = o;
    }
    for (int64_t i = (n16 << 4) + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < elems; i += stride)
      po[i] = s_lut[pi[i]];
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < elems; i += stride) po[i] = s_lut[pi[i]];
  }
}

// ------------------------------------------------------------------ PAD
__global__ void __launch_bounds__(256) pad_kernel(const int8_t* __restrict__ in, int64_t in_ts, int H, int W, int C,
                                                 int pt, int pl, int OH, int OW, int8_t fill, int8_t* __restrict__ out,
                                                 int64_t out_ts) {
  const int t = blockIdx.y;
  const int64_t total = int64_t(OH) * OW * C;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = int(i % C);
    const int x = int((i / C) % OW) - pl;
    const int y = int(i / (int64_t(C) * OW)) - pt;
    int8_t v = fill;
    if (x >= 0 && x < W && y >= 0 && y < H) v = in[int64_t(t) * in_ts + (int64_t(y) * W + x) * C + c];
    out[int64_t(t) * out_ts + i] = v;
  }
}

// ------------------------------------------------------------------ RESIZE_BILINEAR (integer, 10-bit weights)
__device__ __forceinline__ void resize_axis(int o, int scale10, int in_size, bool half_pixel, int* v, int* lo, int* hi) {
  int sv = o * scale10;
  if (half_pixel) sv += scale10 / 2 - (1 << 9);
  *v = sv;
  *lo = max(sv / (1 << 10), 0);
  *hi = min((sv + (1 << 10) - 1) / (1 << 10), in_size - 1);
}

__global__ void __launch_bounds__(256) resize_kernel(const int8_t* __restrict__ in, int64_t in_ts, int IH, int IW,
                                                    int C, int8_t* __restrict__ out, int64_t out_ts, int OH, int OW,
                                                    int hs, int ws, bool half_pixel) {
  const int t = blockIdx.y;
  const int c4n = C >> 2;
  const int64_t total = int64_t(OH) * OW * c4n;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int8_t* tin = in + int64_t(t) * in_ts;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = int(i % c4n) * 4;
    const int x = int((i / c4n) % OW);
    const int y = int(i / (int64_t(c4n) * OW));
    int iy, y0, y1, ix, x0, x1;
    resize_axis(y, hs, IH, half_pixel, &iy, &y0, &y1);
    resize_axis(x, ws, IW, half_pixel, &ix, &x0, &x1);
    const int64_t wy1 = iy - (1 << 10) * y0, wy0 = (1 << 10) - wy1;
    const int64_t wx1 = ix - (1 << 10) * x0, wx0 = (1 << 10) - wx1;
    const char4 p00 = *reinterpret_cast<const char4*>(tin + (int64_t(y0) * IW + x0) * C + c);
    const char4 p10 = *reinterpret_cast<const char4*>(tin + (int64_t(y1) * IW + x0) * C + c);
    const char4 p01 = *reinterpret_cast<const char4*>(tin + (int64_t(y0) * IW + x1) * C + c);
    const char4 p11 = *reinterpret_cast<const char4*>(tin + (int64_t(y1) * IW + x1) * C + c);
    const int8_t* a = reinterpret_cast<const int8_t*>(&p00);
    const int8_t* b = reinterpret_cast<const int8_t*>(&p10);
    const int8_t* d = reinterpret_cast<const int8_t*>(&p01);
    const int8_t* e = reinterpret_cast<const int8_t*>(&p11);
    char4 o;
    int8_t* ob = reinterpret_cast<int8_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t o20 = int64_t(a[j]) * wy0 * wx0 + int64_t(b[j]) * wy1 * wx0 + int64_t(d[j]) * wy0 * wx1 + int64_t(e[j]) * wy1 * wx1;
      const int64_t rnd = o20 > 0 ? (1 << 19) : -(1 << 19);
      ob[j] = int8_t((o20 + rnd) / (1 << 20));
    }
    *reinterpret_cast<char4*>(out + int64_t(t) * out_ts + (int64_t(y) * OW + x) * C + c) = o;
  }
}

// 16 channels per thread (one 16-byte access per corner); C % 16 == 0
__global__ void __launch_bounds__(256) resize16_kernel(const int8_t* __restrict__ in, int64_t in_ts, int IH, int IW,
                                                      int C, int8_t* __restrict__ out, int64_t out_ts, int OH, int OW,
                                                      int hs, int ws, bool half_pixel) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.y;
  const int c16n = C >> 4;
  const int total = OH * OW * c16n;
  const int8_t* tin = in + int64_t(t) * in_ts;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (i % c16n) * 16;
    const int x = (i / c16n) % OW;
    const int y = i / (c16n * OW);
    int iy, y0, y1, ix, x0, x1;
    resize_axis(y, hs, IH, half_pixel, &iy, &y0, &y1);
    resize_axis(x, ws, IW, half_pixel, &ix, &x0, &x1);
    // o20 = sum of the four corners x (wy * wx), weights < 2^20, |pixel| <= 128: fits int32.  Factored exactly as
    // wy0 * (wx0 a + wx1 d) + wy1 * (wx0 b + wx1 e); the inner sums are dp2a over (a, d) / (b, e) byte pairs interleaved
    // with one PRMT per two channels - 9 instructions per channel instead of 16 byte-extract / IMAD ones.
    const int wy1 = iy - (1 << 10) * y0, wy0 = (1 << 10) - wy1;
    const int wx1 = ix - (1 << 10) * x0, wx0 = (1 << 10) - wx1;
    const unsigned wx = (unsigned(wx1) << 16) | unsigned(wx0);  // two 16-bit weights (<= 1024)
    const uint4 p00 = *reinterpret_cast<const uint4*>(tin + (int64_t(y0) * IW + x0) * C + c);
    const uint4 p10 = *reinterpret_cast<const uint4*>(tin + (int64_t(y1) * IW + x0) * C + c);
    const uint4 p01 = *reinterpret_cast<const uint4*>(tin + (int64_t(y0) * IW + x1) * C + c);
    const uint4 p11 = *reinterpret_cast<const uint4*>(tin + (int64_t(y1) * IW + x1) * C + c);
    const unsigned ta[4] = {p00.x, p00.y, p00.z, p00.w}, td[4] = {p01.x, p01.y, p01.z, p01.w};
    const unsigned tb[4] = {p10.x, p10.y, p10.z, p10.w}, te[4] = {p11.x, p11.y, p11.z, p11.w};
    unsigned ow[4];
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      int q[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const unsigned sel = hf ? 0x7362u : 0x5140u;            // bytes {a_j, d_j, a_j+1, d_j+1} of channels 2 hf, 2 hf + 1
        const unsigned ad = __byte_perm(ta[wd], td[wd], sel), be = __byte_perm(tb[wd], te[wd], sel);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int ht = u ? __dp2a_hi(int(wx), int(ad), 0) : __dp2a_lo(int(wx), int(ad), 0);
          const int hb = u ? __dp2a_hi(int(wx), int(be), 0) : __dp2a_lo(int(wx), int(be), 0);
          const int o20 = wy0 * ht + wy1 * hb;
          // (o20 + (o20 > 0 ? 2^19 : -2^19)) / 2^20 truncating == floor((o20 + 2^19 + (o20 < 0 ? -1 : 0)) / 2^20)
          q[2 * hf + u] = (o20 + (o20 >> 31) + (1 << 19)) >> 20;
        }
      }
      ow[wd] = (unsigned(q[0]) & 0xFFu) | ((unsigned(q[1]) & 0xFFu) << 8) | ((unsigned(q[2]) & 0xFFu) << 16) | (unsigned(q[3]) << 24);
    }
    const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    *reinterpret_cast<uint4*>(out + int64_t(t) * out_ts + (int64_t(y) * OW + x) * C + c) = o;
  }
}

// 2 x 2 output pixels per thread.  With the 2x up-samplings of the FPN / protonet (half-pixel centres) the output rows 2k - 1 and 2k
// interpolate between the same two input rows, and likewise the columns, so the four outputs of such a block share their four
// corner loads and the byte interleave; only the dp2a against the two column weight pairs and the vertical blends are per output.
// Blocks are anchored at odd coordinates (rows 2k - 1, 2k; the -1 / OH positions of the border blocks do not exist).  A block
// whose rows or columns do not share their sources (other scales) computes its outputs one by one: same arithmetic, same bytes.
__device__ __forceinline__ void resize16_one(const int8_t* __restrict__ tin, int IW, int C, int c, int y0, int y1, int wy1, int x0, int x1, int wx1,
                                             int8_t* __restrict__ dst) {
  const int wy0 = (1 << 10) - wy1, wx0 = (1 << 10) - wx1;
  const unsigned wx = (unsigned(wx1) << 16) | (unsigned(wx0) & 0xFFFFu);
  const uint4 p00 = *reinterpret_cast<const uint4*>(tin + (int64_t(y0) * IW + x0) * C + c);
  const uint4 p10 = *reinterpret_cast<const uint4*>(tin + (int64_t(y1) * IW + x0) * C + c);
  const uint4 p01 = *reinterpret_cast<const uint4*>(tin + (int64_t(y0) * IW + x1) * C + c);
  const uint4 p11 = *reinterpret_cast<const uint4*>(tin + (int64_t(y1) * IW + x1) * C + c);
  const unsigned ta[4] = {p00.x, p00.y, p00.z, p00.w}, td[4] = {p01.x, p01.y, p01.z, p01.w};
  const unsigned tb[4] = {p10.x, p10.y, p10.z, p10.w}, te[4] = {p11.x, p11.y, p11.z, p11.w};
  unsigned ow[4];
#pragma unroll
  for (int wd = 0; wd < 4; ++wd) {
    int q[4];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const unsigned sel = hf ? 0x7362u : 0x5140u;
      const unsigned ad = __byte_perm(ta[wd], td[wd], sel), be = __byte_perm(tb[wd], te[wd], sel);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ht = u ? __dp2a_hi(int(wx), int(ad), 0) : __dp2a_lo(int(wx), int(ad), 0);
        const int hb = u ? __dp2a_hi(int(wx), int(be), 0) : __dp2a_lo(int(wx), int(be), 0);
        const int o20 = wy0 * ht + wy1 * hb;
        q[2 * hf + u] = (o20 + (o20 >> 31) + (1 << 19)) >> 20;
      }
    }
    ow[wd] = (unsigned(q[0]) & 0xFFu) | ((unsigned(q[1]) & 0xFFu) << 8) | ((unsigned(q[2]) & 0xFFu) << 16) | (unsigned(q[3]) << 24);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

__global__ void __launch_bounds__(256, 4) resize16_quad_kernel(const int8_t* __restrict__ in, int64_t in_ts, int IH, int IW,
                                                           int C, int8_t* __restrict__ out, int64_t out_ts, int OH, int OW,
                                                           int hs, int ws, bool half_pixel) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.y;
  const int c16n = C >> 4;
  const int bw = OW / 2 + 1, bh = OH / 2 + 1;        // blocks per row / column (anchored at -1: columns 2j - 1, 2j)
  const int total = bh * bw * c16n;
  const int8_t* tin = in + int64_t(t) * in_ts;
  int8_t* tout = out + int64_t(t) * out_ts;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (i % c16n) * 16;
    const int bx = (i / c16n) % bw;
    const int by = i / (c16n * bw);
    const int ya = 2 * by - 1, yb = 2 * by, xa = 2 * bx - 1, xb = 2 * bx;
    const bool va = ya >= 0, vb = yb < OH, ua = xa >= 0, ub = xb < OW;
    int iy[2], y0[2], y1[2], ix[2], x0[2], x1[2];
    resize_axis(va ? ya : yb, hs, IH, half_pixel, &iy[0], &y0[0], &y1[0]);
    resize_axis(vb ? yb : ya, hs, IH, half_pixel, &iy[1], &y0[1], &y1[1]);
    resize_axis(ua ? xa : xb, ws, IW, half_pixel, &ix[0], &x0[0], &x1[0]);
    resize_axis(ub ? xb : xa, ws, IW, half_pixel, &ix[1], &x0[1], &x1[1]);
    const int wy1[2] = {iy[0] - (1 << 10) * y0[0], iy[1] - (1 << 10) * y0[1]};
    const int wx1[2] = {ix[0] - (1 << 10) * x0[0], ix[1] - (1 << 10) * x0[1]};
    const bool shared = va && vb && ua && ub && y0[0] == y0[1] && y1[0] == y1[1] && x0[0] == x0[1] && x1[0] == x1[1];
    if (!shared) {
      if (va && ua) resize16_one(tin, IW, C, c, y0[0], y1[0], wy1[0], x0[0], x1[0], wx1[0], tout + (int64_t(ya) * OW + xa) * C + c);
      if (va && ub) resize16_one(tin, IW, C, c, y0[0], y1[0], wy1[0], x0[1], x1[1], wx1[1], tout + (int64_t(ya) * OW + xb) * C + c);
      if (vb && ua) resize16_one(tin, IW, C, c, y0[1], y1[1], wy1[1], x0[0], x1[0], wx1[0], tout + (int64_t(yb) * OW + xa) * C + c);
      if (vb && ub) resize16_one(tin, IW, C, c, y0[1], y1[1], wy1[1], x0[1], x1[1], wx1[1], tout + (int64_t(yb) * OW + xb) * C + c);
      continue;
    }
    const uint4 p00 = *reinterpret_cast<const uint4*>(tin + (int64_t(y0[0]) * IW + x0[0]) * C + c);
    const uint4 p10 = *reinterpret_cast<const uint4*>(tin + (int64_t(y1[0]) * IW + x0[0]) * C + c);
    const uint4 p01 = *reinterpret_cast<const uint4*>(tin + (int64_t(y0[0]) * IW + x1[0]) * C + c);
    const uint4 p11 = *reinterpret_cast<const uint4*>(tin + (int64_t(y1[0]) * IW + x1[0]) * C + c);
    const unsigned ta[4] = {p00.x, p00.y, p00.z, p00.w}, td[4] = {p01.x, p01.y, p01.z, p01.w};
    const unsigned tb[4] = {p10.x, p10.y, p10.z, p10.w}, te[4] = {p11.x, p11.y, p11.z, p11.w};
    const unsigned wxa = (unsigned(wx1[0]) << 16) | (unsigned((1 << 10) - wx1[0]) & 0xFFFFu);
    const unsigned wxb = (unsigned(wx1[1]) << 16) | (unsigned((1 << 10) - wx1[1]) & 0xFFFFu);
    const int wya0 = (1 << 10) - wy1[0], wya1 = wy1[0], wyb0 = (1 << 10) - wy1[1], wyb1 = wy1[1];
    unsigned o_aa[4], o_ab[4], o_ba[4], o_bb[4];   // (row a/b, column a/b)
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
      int qaa[4], qab[4], qba[4], qbb[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const unsigned sel = hf ? 0x7362u : 0x5140u;
        const unsigned ad = __byte_perm(ta[wd], td[wd], sel), be = __byte_perm(tb[wd], te[wd], sel);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int hta = u ? __dp2a_hi(int(wxa), int(ad), 0) : __dp2a_lo(int(wxa), int(ad), 0);
          const int hba = u ? __dp2a_hi(int(wxa), int(be), 0) : __dp2a_lo(int(wxa), int(be), 0);
          const int htb = u ? __dp2a_hi(int(wxb), int(ad), 0) : __dp2a_lo(int(wxb), int(ad), 0);
          const int hbb = u ? __dp2a_hi(int(wxb), int(be), 0) : __dp2a_lo(int(wxb), int(be), 0);
          int o20 = wya0 * hta + wya1 * hba;
          qaa[2 * hf + u] = (o20 + (o20 >> 31) + (1 << 19)) >> 20;
          o20 = wya0 * htb + wya1 * hbb;
          qab[2 * hf + u] = (o20 + (o20 >> 31) + (1 << 19)) >> 20;
          o20 = wyb0 * hta + wyb1 * hba;
          qba[2 * hf + u] = (o20 + (o20 >> 31) + (1 << 19)) >> 20;
          o20 = wyb0 * htb + wyb1 * hbb;
          qbb[2 * hf + u] = (o20 + (o20 >> 31) + (1 << 19)) >> 20;
        }
      }
      o_aa[wd] = (unsigned(qaa[0]) & 0xFFu) | ((unsigned(qaa[1]) & 0xFFu) << 8) | ((unsigned(qaa[2]) & 0xFFu) << 16) | (unsigned(qaa[3]) << 24);
      o_ab[wd] = (unsigned(qab[0]) & 0xFFu) | ((unsigned(qab[1]) & 0xFFu) << 8) | ((unsigned(qab[2]) & 0xFFu) << 16) | (unsigned(qab[3]) << 24);
      o_ba[wd] = (unsigned(qba[0]) & 0xFFu) | ((unsigned(qba[1]) & 0xFFu) << 8) | ((unsigned(qba[2]) & 0xFFu) << 16) | (unsigned(qba[3]) << 24);
      o_bb[wd] = (unsigned(qbb[0]) & 0xFFu) | ((unsigned(qbb[1]) & 0xFFu) << 8) | ((unsigned(qbb[2]) & 0xFFu) << 16) | (unsigned(qbb[3]) << 24);
    }
    int8_t* d0 = tout + (int64_t(ya) * OW + xa) * C + c;
    *reinterpret_cast<uint4*>(d0) = make_uint4(o_aa[0], o_aa[1], o_aa[2], o_aa[3]);
    *reinterpret_cast<uint4*>(d0 + C) = make_uint4(o_ab[0], o_ab[1], o_ab[2], o_ab[3]);
    *reinterpret_cast<uint4*>(d0 + int64_t(OW) * C) = make_uint4(o_ba[0], o_ba[1], o_ba[2], o_ba[3]);
    *reinterpret_cast<uint4*>(d0 + int64_t(OW) * C + C) = make_uint4(o_bb[0], o_bb[1], o_bb[2], o_bb[3]);
  }
}

__global__ void __launch_bounds__(256) copy_kernel(const uint8_t* __restrict__ in, int64_t in_ts,
                                                  uint8_t* __restrict__ out, int64_t out_ts, int64_t bytes, bool vec) {
  const int t = blockIdx.y;
  const uint8_t* pi = in + int64_t(t) * in_ts;
  uint8_t* po = out + int64_t(t) * out_ts;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  if (vec) {
    const int64_t n16 = bytes >> 4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
      reinterpret_cast<uint4*>(po)[i] = reinterpret_cast<const uint4*>(pi)[i];
    for (int64_t i = (n16 << 4) + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < bytes; i += stride) po[i] = pi[i];
  } else {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < bytes; i += stride) po[i] = pi[i];
  }
}

inline bool aligned16(const void* p, int64_t stride) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (stride & 15) == 0; }
inline int grid_for(int64_t work_items, int threads, int tiles) {
  // enough CTAs to fill 148 SMs a few times over, never more than the work needs
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (148 * 8 + tiles - 1) / tiles;
  if (blocks > cap) blocks = cap;
  return int(blocks < 1 ? 1 : blocks);
}

}  // namespace

bool stem_kernel_eligible(const ConvGeom& g, const Requant& rq, const void* in, int64_t in_ts, const void* out, int64_t out_ts) {
  return g.IC == 3 && g.KH == 3 && g.KW == 3 && g.stride_h == 2 && g.stride_w == 2 && g.dil_h == 1 && g.dil_w == 1 && g.pad_top == 0 &&
         g.pad_left == 0 && g.OC % 16 == 0 && g.OC <= 64 && rq.fast_tab && !rq.post_lut && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
         (in_ts & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (out_ts & 15) == 0 && (g.IW * 3) % 2 == 0;
}

void launch_conv_direct(const int8_t* in, int64_t in_ts, const int8_t* w, const int32_t* bias, const int32_t* wsum,
                        int32_t in_zp, const ConvGeom& g, const Requant& rq, int8_t* out, int64_t out_ts, int tiles,
                        cudaStream_t s) {
  // the RGB stem: 3x3, stride 2, no top / left padding, 16-byte-aligned 16-channel output groups
  if (stem_kernel_eligible(g, rq, in, in_ts, out, out_ts) && int64_t(tiles) * g.OH * g.OW * (g.OC / kStemOct) < (int64_t(1) << 31)) {
    static const int stem_impl = std::getenv("TOD_STEM_IMPL") ? std::atoi(std::getenv("TOD_STEM_IMPL")) : 0;  // 0 = two pixels per thread, 1 = one
    static const int stem_rows = std::getenv("TOD_STEM_ROWS") ? std::max(1, std::atoi(std::getenv("TOD_STEM_ROWS"))) : kStemRows;
    if (stem_impl == 0 && int64_t(tiles) * g.OH <= 65535 * int64_t(stem_rows)) {
      const int items = ((g.OW + 1) / 2) * (g.OC / kStemOct);
      dim3 grid2(unsigned((items + kStemThreads - 1) / kStemThreads), unsigned((int64_t(tiles) * g.OH + stem_rows - 1) / stem_rows));
      if (rq.act_min == -128 && rq.act_max == 127)
        rq.relu_tab ? launch_k(stem3x3s2_pair_kernel<true, true>, grid2, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles, stem_rows)
                    : launch_k(stem3x3s2_pair_kernel<true, false>, grid2, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles, stem_rows);
      else
        rq.relu_tab ? launch_k(stem3x3s2_pair_kernel<false, true>, grid2, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles, stem_rows)
                    : launch_k(stem3x3s2_pair_kernel<false, false>, grid2, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles, stem_rows);
      return;
    }
    const int64_t total = int64_t(tiles) * g.OH * g.OW * (g.OC / kStemOct);
    dim3 grid(unsigned((total + kStemThreads * kStemIters - 1) / (kStemThreads * kStemIters)));
    if (rq.act_min == -128 && rq.act_max == 127)
      rq.relu_tab ? launch_k(stem3x3s2_kernel<true, true>, grid, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles)
                  : launch_k(stem3x3s2_kernel<true, false>, grid, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles);
    else
      rq.relu_tab ? launch_k(stem3x3s2_kernel<false, true>, grid, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles)
                  : launch_k(stem3x3s2_kernel<false, false>, grid, dim3(kStemThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles);
    return;
  }
  // preferred CUDA-core path: shared-memory weight slab, one thread per pixel x 16 channels
  const int taps = g.KH * g.KW, icw = (g.IC + 3) / 4;
  const size_t pix_smem = size_t(taps) * icw * kPixOct * 4 + size_t(taps) * kPixOct * 4;
  const bool words_ok = g.IC == 3 || (g.IC % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (in_ts % 4) == 0);
  if (words_ok && pix_smem <= 48 * 1024 && int64_t(tiles) * g.OH * g.OW < (int64_t(1) << 31)) {
    const int total = tiles * g.OH * g.OW;
    dim3 grid(unsigned((total + kPixThreads - 1) / kPixThreads), unsigned((g.OC + kPixOct - 1) / kPixOct));
    if (g.IC == 3)
      launch_k(conv_pix_kernel<true>, grid, dim3(kPixThreads), pix_smem, s, in, in_ts, w, bias, wsum, in_zp, g, rq, out, out_ts, tiles);
    else
      launch_k(conv_pix_kernel<false>, grid, dim3(kPixThreads), pix_smem, s, in, in_ts, w, bias, wsum, in_zp, g, rq, out, out_ts, tiles);
    return;
  }
  constexpr int OCT = 8;
  const int64_t pixels = int64_t(tiles) * g.OH * g.OW;
  dim3 grid(unsigned((pixels + 127) / 128), unsigned((g.OC + OCT - 1) / OCT));
  const bool vec = (g.IC % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0) && (in_ts % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(w) & 3) == 0);
  if (vec)
    conv_direct_kernel<OCT, true><<<grid, 128, 0, s>>>(in, in_ts, w, bias, wsum, in_zp, g, rq, out, out_ts, tiles);
  else
    conv_direct_kernel<OCT, false><<<grid, 128, 0, s>>>(in, in_ts, w, bias, wsum, in_zp, g, rq, out, out_ts, tiles);
}

void launch_depthwise(const int8_t* in, int64_t in_ts, const int8_t* w, const int32_t* bias, int32_t in_zp,
                      const ConvGeom& g, const Requant& rq, int8_t* out, int64_t out_ts, int tiles, cudaStream_t s) {
  const bool vec = (g.OC % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0) && (in_ts % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 3) == 0) && (out_ts % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(w) & 3) == 0);
  if (vec && rq.fast_tab && !rq.post_lut && g.KH == 3 && g.KW == 3 && g.dil_h == 1 && g.dil_w == 1 && g.stride_h == g.stride_w &&
      (g.stride_h == 1 || g.stride_h == 2) && tiles <= 65535) {
    const int words = g.OW * (g.OC / 4);
    const int gx = (words + kDwThreads - 1) / kDwThreads;
    const bool sat = rq.act_min == -128 && rq.act_max == 127;
    // TOD_DW_IMPL: 0 (default) = lean kernel (sliding kernel where it does not apply), 2 = sliding kernel.  (A burst variant - all
    // of a thread's input rows in flight before any arithmetic - measured 2 % slower per step and was removed: DESIGN 5.2.)
    static const int dw_impl = std::getenv("TOD_DW_IMPL") ? std::atoi(std::getenv("TOD_DW_IMPL")) : 0;
    // lean kernel: ReLU-form table, at most one window column outside the image on either side, tile fits 32-bit word offsets
    const bool lean_ok = rq.relu_tab && g.IW >= 3 && g.pad_left <= 1 && (g.OW - 1) * g.stride_w - g.pad_left + 2 <= g.IW &&
                         int64_t(g.IH + 4) * g.IW * g.OC < (int64_t(1) << 31);
    if (dw_impl == 0 && lean_ok) {
      // rows per thread: every run re-loads and re-transposes two halo rows, so runs are as long as keeps ~4 CTAs per SM in the
      // grid (measured: 8 per SM / 16 rows -> 4 per SM / 28 rows = -2 % per step; longer runs or fewer CTAs no better)
      static const int dw_min_ctas = std::getenv("TOD_DW_MINCTAS") ? std::atoi(std::getenv("TOD_DW_MINCTAS")) : 148 * 4;
      static const int dw_max_rows = std::getenv("TOD_DW_MAXROWS") ? std::atoi(std::getenv("TOD_DW_MAXROWS")) : 28;
      int rpb = g.OH;
      while (rpb > 4 && int64_t(gx) * ((g.OH + rpb - 1) / rpb) * tiles < dw_min_ctas) rpb = (rpb + 1) / 2;
      rpb = std::min(rpb, dw_max_rows);
      dim3 grid(gx, (g.OH + rpb - 1) / rpb, tiles);
      const int cw = g.OC / 4;
#define TOD_DWL(ST, SA, CWT) launch_k(depthwise3x3_lean_kernel<ST, SA, true, CWT>, grid, dim3(kDwThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, rpb)
#define TOD_DWL_C(CWT)                                                              \
  do {                                                                              \
    if (g.stride_h == 1) { if (sat) TOD_DWL(1, true, CWT); else TOD_DWL(1, false, CWT); } \
    else { if (sat) TOD_DWL(2, true, CWT); else TOD_DWL(2, false, CWT); }           \
  } while (0)
      switch (cw) {  // MobileNetV2's depthwise widths (32 ... 960 channels); anything else takes the runtime-width build
        case 8: TOD_DWL_C(8); break;
        case 24: TOD_DWL_C(24); break;
        case 36: TOD_DWL_C(36); break;
        case 48: TOD_DWL_C(48); break;
        case 96: TOD_DWL_C(96); break;
        case 144: TOD_DWL_C(144); break;
        case 240: TOD_DWL_C(240); break;
        default: TOD_DWL_C(0); break;
      }
#undef TOD_DWL_C
#undef TOD_DWL
      return;
    }
    // rows per block: long enough to amortise the two warm-up rows, short enough for >= ~4 CTAs per SM
    int rpb = g.OH;
    while (rpb > 4 && int64_t(gx) * ((g.OH + rpb - 1) / rpb) * tiles < 148 * 8) rpb = (rpb + 1) / 2;
    rpb = std::min(rpb, 16);
    dim3 grid(gx, (g.OH + rpb - 1) / rpb, tiles);
#define TOD_DW(ST, SA, RE) launch_k(depthwise3x3_slide_kernel<ST, SA, RE>, grid, dim3(kDwThreads), 0, s, in, in_ts, w, bias, in_zp, g, rq, out, out_ts, rpb)
    if (g.stride_h == 1) {
      if (rq.relu_tab) { if (sat) TOD_DW(1, true, true); else TOD_DW(1, false, true); }
      else { if (sat) TOD_DW(1, true, false); else TOD_DW(1, false, false); }
    } else {
      if (rq.relu_tab) { if (sat) TOD_DW(2, true, true); else TOD_DW(2, false, true); }
      else { if (sat) TOD_DW(2, true, false); else TOD_DW(2, false, false); }
    }
#undef TOD_DW
    return;
  }
  if (vec && g.KH == 3 && g.KW == 3 && g.dil_h == 1 && g.dil_w == 1 && int64_t(tiles) * g.OH * g.OW < (int64_t(1) << 31)) {
    const int cg = g.OC / 4;
    const int bx = cg >= 32 ? 32 : cg;
    const int by = 256 / bx;
    const int total = tiles * g.OH * g.OW;
    const int gx = (cg + bx - 1) / bx;
    // enough CTAs for ~8 per SM, each walking a contiguous run of pixels
    int pix_per_block = (total + (148 * 8 / gx + 1) - 1) / (148 * 8 / gx + 1);
    pix_per_block = ((pix_per_block + by - 1) / by) * by;
    if (pix_per_block < by) pix_per_block = by;
    dim3 grid(gx, (total + pix_per_block - 1) / pix_per_block), block(bx, by);
    depthwise_reg_kernel<3><<<grid, block, 0, s>>>(in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles, pix_per_block);
    return;
  }
  if (vec) {
    const int64_t total = int64_t(tiles) * g.OH * g.OW * (g.OC / 4);
    depthwise_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles);
  } else {
    const int64_t total = int64_t(tiles) * g.OH * g.OW * g.OC;
    depthwise_scalar_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(in, in_ts, w, bias, in_zp, g, rq, out, out_ts, tiles);
  }
}

void launch_add(const int8_t* a, int64_t a_ts, const int8_t* b, int64_t b_ts, int8_t* out, int64_t out_ts,
                int64_t elems, int tiles, const AddParams& p, cudaStream_t s) {
  const bool vec = aligned16(a, a_ts) && aligned16(b, b_ts) && aligned16(out, out_ts);
  dim3 grid(grid_for(vec ? elems / 16 + 1 : elems, 256, tiles), tiles);
  add_kernel<<<grid, 256, 0, s>>>(a, a_ts, b, b_ts, out, out_ts, elems, tiles, p, vec);
}

void launch_lut(const uint8_t* in, int64_t in_ts, uint8_t* out, int64_t out_ts, int64_t elems, int tiles,
                const uint8_t* lut256, cudaStream_t s) {
  const bool vec = aligned16(in, in_ts) && aligned16(out, out_ts);
  dim3 grid(grid_for(vec ? elems / 16 + 1 : elems, 256, tiles), tiles);
  lut_kernel<<<grid, 256, 0, s>>>(in, in_ts, out, out_ts, elems, lut256, vec);
}

void launch_pad(const int8_t* in, int64_t in_ts, int H, int W, int C, int pt, int pl, int OH, int OW, int8_t fill,
                int8_t* out, int64_t out_ts, int tiles, cudaStream_t s) {
  dim3 grid(grid_for(int64_t(OH) * OW * C, 256, tiles), tiles);
  pad_kernel<<<grid, 256, 0, s>>>(in, in_ts, H, W, C, pt, pl, OH, OW, fill, out, out_ts);
}

void launch_resize_bilinear(const int8_t* in, int64_t in_ts, int IH, int IW, int C, int8_t* out, int64_t out_ts,
                            int OH, int OW, bool align_corners, bool half_pixel, int tiles, cudaStream_t s) {
  int hs = ((1 << 10) * IH + OH / 2) / OH, ws = ((1 << 10) * IW + OW / 2) / OW;
  if (align_corners && OH > 1) hs = ((1 << 10) * (IH - 1) + (OH - 1) / 2) / (OH - 1);
  if (align_corners && OW > 1) ws = ((1 << 10) * (IW - 1) + (OW - 1) / 2) / (OW - 1);
  if (C % 16 == 0 && aligned16(in, in_ts) && aligned16(out, out_ts) && int64_t(OH) * OW * C < (int64_t(1) << 31)) {
    static const int resize_impl = std::getenv("TOD_RESIZE_IMPL") ? std::atoi(std::getenv("TOD_RESIZE_IMPL")) : 0;  // 0 = 2 x 2 outputs per thread, 1 = one
    if (resize_impl == 0 && OH % 2 == 0 && OW % 2 == 0 && OH >= 2 && OW >= 2) {
      dim3 gridq(grid_for(int64_t(OH / 2 + 1) * (OW / 2 + 1) * (C / 16), 256, tiles), tiles);
      launch_k(resize16_quad_kernel, gridq, dim3(256), 0, s, in, in_ts, IH, IW, C, out, out_ts, OH, OW, hs, ws, half_pixel);
      return;
    }
    dim3 grid16(grid_for(int64_t(OH) * OW * (C / 16), 256, tiles), tiles);
    launch_k(resize16_kernel, grid16, dim3(256), 0, s, in, in_ts, IH, IW, C, out, out_ts, OH, OW, hs, ws, half_pixel);
    return;
  }
  dim3 grid(grid_for(int64_t(OH) * OW * (C / 4), 256, tiles), tiles);
  resize_kernel<<<grid, 256, 0, s>>>(in, in_ts, IH, IW, C, out, out_ts, OH, OW, hs, ws, half_pixel);
}

void launch_copy(const uint8_t* in, int64_t in_ts, uint8_t* out, int64_t out_ts, int64_t bytes, int tiles,
                 cudaStream_t s) {
  const bool vec = aligned16(in, in_ts) && aligned16(out, out_ts);
  dim3 grid(grid_for(vec ? bytes / 16 + 1 : bytes, 256, tiles), tiles);
  copy_kernel<<<grid, 256, 0, s>>>(in, in_ts, out, out_ts, bytes, vec);
}

}  // namespace tod
