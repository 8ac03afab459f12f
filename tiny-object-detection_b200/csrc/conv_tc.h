// tcgen05 int8 implicit-GEMM convolution (conv_tc.cu): CONV_2D layers whose shape maps onto the
// 5th-generation tensor cores.  D[pixels, OC] (s32, TMEM) = A[pixels, taps*IC] (s8, NHWC activations
// fetched tap by tap with TMA; out-of-image taps are zero-filled by the TMA unit) x B[OC, taps*IC]^T
// (s8, the TFLite OHWI weight tensor is already K-major), with TFLite's fixed-point requantisation,
// the input-zero-point correction and the activation clamp fused into the TMEM->register epilogue.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "ops.h"

namespace tod {

struct ConvTc;  // one planned layer (tensor maps, epilogue tables)

struct ConvTcArgs {
  ConvGeom g;
  const int8_t* in;      // device, [tile][IH][IW][IC]
  int64_t in_tile_stride;
  const int8_t* w;       // device, [OC][KH][KW][IC]
  int32_t in_zp;
  Requant rq;            // device pointers
  int8_t* out;           // device, [tile][OH][OW][OC]
  int64_t out_tile_stride;
  int max_tiles;
  const int32_t* h_bias; // host copies used to build the per-border-class bias tables (may be null)
  const int32_t* h_wsum; // host [OC][KH*KW]
  const int32_t* h_mult = nullptr;   // host copies of rq.mult / rq.shift: the fast epilogue is planned from them
  const int32_t* h_shift = nullptr;  // (null: general epilogue)
  int fast_epilogue = 1;             // 0 forces the general epilogue (cross-check)
};

// shape / alignment test only (no CUDA calls)
bool conv_tc_supported(const ConvGeom& g, int64_t in_tile_stride, const void* in, const void* w);
int conv_tc_create(const ConvTcArgs& a, ConvTc** plan);
int conv_tc_launch(ConvTc* plan, int tiles, cudaStream_t s);
void conv_tc_destroy(ConvTc* plan);

}  // namespace tod
