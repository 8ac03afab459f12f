// TFLite full-integer fixed-point arithmetic for the device epilogues and the host planner.
//
// The reference reaches this arithmetic through interpreter.invoke() (/root/reference/src/yolact.rs:163;
// TensorFlow Lite C++ via crate tflite 0.9.0, Cargo.lock:1106-1108 — not vendored).  The rules are the
// published gemmlowp ones: a Q31 multiplier applied with a rounding doubling
// high multiply, then a rounding right shift (round half away from zero), both kept in integers so the
// result is bit-identical to the CPU interpreter.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define TOD_HD __host__ __device__ __forceinline__
#else
#define TOD_HD inline
#endif

namespace tod {

// (a*b*2 + 2^31 rounding) >> 32 with saturation of the single overflow case; the division by 2^31
// truncates toward zero, hence the sign-dependent nudge.
TOD_HD int32_t sat_round_doubling_high_mul(int32_t a, int32_t b) {
  if (a == b && a == INT32_MIN) return INT32_MAX;
  const int64_t prod = int64_t(a) * int64_t(b);
  const int64_t nudged = prod + (prod >= 0 ? (int64_t(1) << 30) : (int64_t(1) - (int64_t(1) << 30)));
  // truncating division by 2^31: add (2^31 - 1) to negatives before the arithmetic shift
  const int64_t bias = (nudged >> 63) & ((int64_t(1) << 31) - 1);
  return int32_t((nudged + bias) >> 31);
}

// x / 2^e rounded to nearest, ties away from zero
TOD_HD int32_t round_divide_pot(int32_t x, int e) {
  const int32_t mask = int32_t((int64_t(1) << e) - 1);
  const int32_t rem = x & mask;
  const int32_t thr = (mask >> 1) + (x < 0 ? 1 : 0);
  return (x >> e) + (rem > thr ? 1 : 0);
}

// MultiplyByQuantizedMultiplier, double-rounding form (the TFLite default build)
TOD_HD int32_t mul_by_quant_mult(int32_t x, int32_t q, int shift) {
  const int left = shift > 0 ? shift : 0;
  const int right = shift > 0 ? 0 : -shift;
  return round_divide_pot(sat_round_doubling_high_mul(int32_t(uint32_t(x) << left), q), right);
}

// Same result as mul_by_quant_mult for every (x, q, shift) with q >= 0 (TFLite multipliers are non-negative) and
// |x| < 2^30 (int8 accumulators stay below 2^26), in about half the instructions:
//   * SRDHM's sign-dependent nudge followed by a truncating division equals a plain floor:
//       trunc((p + (p >= 0 ? 2^30 : 1 - 2^30)) / 2^31) == (p + 2^30) >> 31          (arithmetic shift)
//   * the rounding right shift with ties away from zero equals
//       (v + 2^(e-1) + (v < 0 ? -1 : 0)) >> e                                        (e > 0)
// tests/test_fixedpoint.py checks the identity against the literal forms on the edge cases and 10^7 random draws.
TOD_HD int32_t mul_by_quant_mult_fast(int32_t x, int32_t q, int shift) {
  const int left = shift > 0 ? shift : 0;
  const int right = shift > 0 ? 0 : -shift;
  const int64_t p = int64_t(int32_t(uint32_t(x) << left)) * int64_t(q);
  const int32_t v = int32_t((p + (int64_t(1) << 30)) >> 31);
  const int32_t half = int32_t((uint32_t(1) << right) >> 1);
  const int32_t neg = right > 0 ? (v >> 31) : 0;
  return (v + half + neg) >> right;
}

// The tabulated form the conv / depthwise epilogues run (conv_tc.cu, ops.cu; table = Requant::fast_tab):
//   requant_tab(x, q, rs, (1 << (rs-1)) + (zp << rs))  ==  mul_by_quant_mult(x, q, -rs) + zp
// for q >= 0, 1 <= rs <= 22, |x| < 2^29, |zp| <= 128.  hi32(2x*q + 2^31) is SRDHM's result v; the rounding term
// and the output zero point ride in the high word of the 64-bit addend; sign(2x) stands in for sign(v): they differ
// only where v == 0 (x < 0 with x*q > -2^30), and there (half - 1) >> rs == half >> rs == 0 anyway.
TOD_HD int32_t requant_tab(int32_t x, int32_t q, int rs, int32_t halfp) {
  const int32_t x2 = int32_t(uint32_t(x) * 2u);
  const int64_t addend = int64_t((uint64_t(uint32_t(halfp)) << 32) | 0x80000000ull);
  const int32_t t = int32_t((int64_t(x2) * int64_t(q) + addend) >> 32);
  return (t + (x2 >> 31)) >> rs;
}

// The form for layers whose activation floor sits at or above the output zero point (ReLU / ReLU6: act_min is the
// quantised 0.0), after the clamp:
//   clamp(requant_relu(x, q, rs - 1, relu_addend(q, rs, zp)), act_min, act_max) == clamp(mul_by_quant_mult(x, q, -rs) + zp, ...)
// for q >= 0, 1 <= rs <= 22, |x| < 2^29, act_min >= zp.  With Y = x*q + 2^30 + halfp*2^31 (halfp as in requant_tab),
// requant_tab's value for x >= 0 is floor(2Y / 2^(32+rs)) = floor(Y / 2^(31+rs)) = hi32(Y) >> (rs - 1): exact.  For x < 0
// the exact result is <= zp (a negative value never rounds above zero) and so is this one (SRDHM's v <= 0 gives
// hi32-term <= half, and half >> rs == 0): both clamp to act_min.  One 64-bit multiply-add and one shift instead of five
// instructions; a per-channel bias folds into the addend as + bias * q.  tests/cpp/fixedpoint_check.cpp checks it.
TOD_HD int64_t relu_addend(int32_t q, int rs, int32_t zp, int64_t bias = 0) {
  const int64_t halfp = (int64_t(1) << (rs - 1)) + int64_t(zp) * (int64_t(1) << rs);
  return bias * int64_t(q) + (int64_t(1) << 30) + halfp * (int64_t(1) << 31);
}
TOD_HD int32_t requant_relu(int32_t x, int32_t q, int rs_minus_1, int64_t addend) {
  return int32_t((int64_t(x) * int64_t(q) + addend) >> 32) >> rs_minus_1;
}

// host only: real multiplier -> (Q31 mantissa, exponent)
inline void quantize_multiplier(double m, int32_t* q, int* shift) {
  if (m == 0.0) {
    *q = 0;
    *shift = 0;
    return;
  }
  const double frac = std::frexp(m, shift);
  int64_t fixed = int64_t(std::round(frac * double(int64_t(1) << 31)));
  if (fixed == (int64_t(1) << 31)) {
    fixed /= 2;
    ++*shift;
  }
  if (*shift < -31) {
    *shift = 0;
    fixed = 0;
  }
  *q = int32_t(fixed);
}

}  // namespace tod
