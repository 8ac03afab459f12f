// Launchers of the CUDA-core graph kernels (ops.cu).  Every activation tensor is laid out
// [tile][H][W][C] with an explicit byte stride between tiles, so a tensor can live inside a larger
// buffer (a CONCATENATION output) and the concat itself costs nothing.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace tod {

struct ConvGeom {
  int IH, IW, IC;         // input, per tile
  int OH, OW, OC;         // output, per tile
  int KH, KW;
  int stride_h, stride_w, dil_h, dil_w;
  int pad_top, pad_left;  // taps that fall outside [0,IH)x[0,IW) are skipped (TFLite semantics)
  int in_xor;             // stem kernel only: every input byte is XORed with this first (0x80 = the folded uint8 -> int8 QUANTIZE)
};

struct Requant {          // per-output-channel fixed point requantisation + activation clamp
  const int32_t* mult;    // [OC] Q31
  const int32_t* shift;   // [OC] exponent (<= 0: right shift)
  int32_t out_zp, act_min, act_max;
  const uint8_t* post_lut;  // optional 256-entry byte map applied to the requantised byte (fused QUANTIZE / RELU / TANH chain)
  // optional [OC] {q, rs, 0x80000000, (1 << (rs-1)) + (out_zp << rs)}: present when every channel has shift in [-22,-1]
  // and q >= 0, so  out = (hi32(2*acc*q + 2^31 + (w << 32)) + (acc >> 31)) >> rs  equals the reference form
  const int4* fast_tab = nullptr;
  // fast_tab holds the ReLU form instead: {q, rs - 1, lo32(A), hi32(A)} with A = relu_addend(q, rs, out_zp), for
  // fixedpoint.cuh::requant_relu.  Only when act_min >= out_zp (every negative pre-activation clamps to act_min anyway).
  bool relu_tab = false;
};

// CONV_2D, any geometry.  w: [OC][KH][KW][IC] int8.  bias: [OC] int32 (may be null).
// wsum: [OC][KH*KW] int32 = sum over IC of the tap's weights (for the input zero point).
void launch_conv_direct(const int8_t* in, int64_t in_tile_stride, const int8_t* w, const int32_t* bias,
                        const int32_t* wsum, int32_t in_zp, const ConvGeom& g, const Requant& rq, int8_t* out,
                        int64_t out_tile_stride, int tiles, cudaStream_t s);

// true when launch_conv_direct will take the dedicated RGB-stem kernel (the only one that honours ConvGeom::in_xor)
bool stem_kernel_eligible(const ConvGeom& g, const Requant& rq, const void* in, int64_t in_tile_stride, const void* out,
                          int64_t out_tile_stride);

// DEPTHWISE_CONV_2D, depth multiplier 1.  w: [1][KH][KW][C] int8.
void launch_depthwise(const int8_t* in, int64_t in_tile_stride, const int8_t* w, const int32_t* bias, int32_t in_zp,
                      const ConvGeom& g, const Requant& rq, int8_t* out, int64_t out_tile_stride, int tiles,
                      cudaStream_t s);

struct AddParams {
  int32_t zp_a, zp_b, zp_out;
  int32_t mult_a, shift_a, mult_b, shift_b, mult_out, shift_out;
  int32_t act_min, act_max;
};
void launch_add(const int8_t* a, int64_t a_tile_stride, const int8_t* b, int64_t b_tile_stride, int8_t* out,
                int64_t out_tile_stride, int64_t elems_per_tile, int tiles, const AddParams& p, cudaStream_t s);

// 256-entry byte map (QUANTIZE int8/uint8 -> int8/uint8, RELU, TANH are all pure functions of the
// input byte, so the planner tabulates them with the exact integer / libm arithmetic).
void launch_lut(const uint8_t* in, int64_t in_tile_stride, uint8_t* out, int64_t out_tile_stride,
                int64_t elems_per_tile, int tiles, const uint8_t* lut256, cudaStream_t s);

// PAD (spatial only), fill value = the tensor's zero point.
void launch_pad(const int8_t* in, int64_t in_tile_stride, int H, int W, int C, int pad_top, int pad_left, int OH,
                int OW, int8_t fill, int8_t* out, int64_t out_tile_stride, int tiles, cudaStream_t s);

// RESIZE_BILINEAR, TFLite integer kernel (10-bit fixed point weights).
void launch_resize_bilinear(const int8_t* in, int64_t in_tile_stride, int IH, int IW, int C, int8_t* out,
                            int64_t out_tile_stride, int OH, int OW, bool align_corners, bool half_pixel, int tiles,
                            cudaStream_t s);

// strided tile copy (CONCATENATION / RESHAPE inputs that could not be aliased)
void launch_copy(const uint8_t* in, int64_t in_tile_stride, uint8_t* out, int64_t out_tile_stride,
                 int64_t bytes_per_tile, int tiles, cudaStream_t s);

}  // namespace tod
