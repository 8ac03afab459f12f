// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
//
// TFLite full-integer fixed-point helpers, restated literally from the published
// gemmlowp / TensorFlow Lite rules recorded in SURVEY.md §10.1-10.2.  The reference
// reaches this arithmetic through the un-vendored crate `tflite 0.9.0`
// (git littletitan/tflite-rs@abcaeab, /root/reference/Cargo.lock:1106-1108) wrapped by
// `edgetpu 0.1.0` (Cargo.lock:314-316); call site /root/reference/src/yolact.rs:163
// (`interpreter.invoke()`).  PARITY UNPINNED against a real TFLite build: the
// reference tree holds no golden vector for it and no TFLite runtime exists here.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <algorithm>

namespace oracle {

// QuantizeMultiplier(double) -> (q31 multiplier, shift)
inline void QuantizeMultiplier(double m, int32_t* q, int* shift) {
  if (m == 0.0) { *q = 0; *shift = 0; return; }
  const double f = std::frexp(m, shift);
  int64_t q_fixed = static_cast<int64_t>(std::round(f * (1ll << 31)));
  if (q_fixed == (1ll << 31)) { q_fixed /= 2; ++*shift; }
  if (*shift < -31) { *shift = 0; q_fixed = 0; }
  *q = static_cast<int32_t>(q_fixed);
}

// SaturatingRoundingDoublingHighMul, literal form (nudge + truncating division).
inline int32_t SRDHM(int32_t a, int32_t b) {
  const bool overflow = a == b && a == std::numeric_limits<int32_t>::min();
  const int64_t ab = static_cast<int64_t>(a) * static_cast<int64_t>(b);
  const int32_t nudge = ab >= 0 ? (1 << 30) : (1 - (1 << 30));
  const int32_t r = static_cast<int32_t>((ab + nudge) / (1ll << 31));
  return overflow ? std::numeric_limits<int32_t>::max() : r;
}

// RoundingDivideByPOT, literal form (mask / remainder / threshold).
inline int32_t RDivPOT(int32_t x, int e) {
  const int32_t mask = static_cast<int32_t>((1ll << e) - 1);
  const int32_t rem = x & mask;
  const int32_t thr = (mask >> 1) + (x < 0 ? 1 : 0);
  return (x >> e) + (rem > thr ? 1 : 0);
}

// MultiplyByQuantizedMultiplier, double-rounding variant (TFLite default).
inline int32_t MBQM(int32_t x, int32_t q, int shift) {
  const int left = shift > 0 ? shift : 0;
  const int right = shift > 0 ? 0 : -shift;
  return RDivPOT(SRDHM(static_cast<int32_t>(static_cast<uint32_t>(x) << left), q), right);
}

// round-half-away-from-zero float->int (TfLiteRound == std::round)
inline int32_t RoundHalfAway(double x) { return static_cast<int32_t>(std::round(x)); }

}  // namespace oracle
