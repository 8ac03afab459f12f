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

// A residual / FPN ADD fused behind the convolution: out = clamp(rescale_out(tab[conv byte] + tab[256 + other byte]) + zp).
// The two 256-entry tables hold each input's term already rescaled to the common 2^20 fixed point with the literal
// TFLite arithmetic (index = the int8 byte pattern), exactly what ops.cu::add_kernel tabulates per CTA.
struct ConvTcAdd {
  const int8_t* resid;        // device, the ADD's other input, same geometry and tile stride rules as the output
  int64_t resid_tile_stride;
  int32_t tab[512];
  int32_t mult_out, shift_out, zp_out, act_min, act_max;
};

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
  const ConvTcAdd* add = nullptr;    // fused ADD (requires the fast epilogue + TMA-storable output)
  int min_rounds = 1;                // tiles per CTA a launch gives at least (tod_yolact_options::batches_in_flight >= 2: two)
  // ---- sibling output (a second, narrow convolution over the same input computed in the same launch: the box head riding the
  // coefficient head, see yolact.cu "sibling-head fusion").  g.OC counts BOTH layers' channels: the host layer's `out_oc`
  // (a multiple of 16) first, then `x_cols` sibling channels padded to a 16-column chunk; weights / bias / wsum / mult / shift
  // cover all g.OC rows.  The host layer's output tensor has `out_oc` channels; the sibling's `x_cols` bytes per pixel go to x_out.
  int out_oc = 0;                    // 0 = no sibling (the output has g.OC channels)
  int x_cols = 0;                    // sibling channels (multiple of 4, <= 16)
  int8_t* x_out = nullptr;           // device, [tile][OH][OW][x_cols]
  int64_t x_out_tile_stride = 0;
  int32_t x_out_zp = 0;              // the sibling's output zero point (rq.out_zp is the host layer's)
  const uint8_t* x_lut = nullptr;    // device: the sibling's fused byte map (null: none)
};

// shape / alignment test only (no CUDA calls)
bool conv_tc_supported(const ConvGeom& g, int64_t in_tile_stride, const void* in, const void* w);
int conv_tc_create(const ConvTcArgs& a, ConvTc** plan);
int conv_tc_launch(ConvTc* plan, int tiles, cudaStream_t s);
void conv_tc_destroy(ConvTc* plan);

}  // namespace tod
