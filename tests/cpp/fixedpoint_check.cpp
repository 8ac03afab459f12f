// Compares the product's fixed-point helpers (csrc/fixedpoint.cuh) with the oracle's literal restatement
// (oracle/fixedpoint.h) on edge cases and random draws.  Built and run by tests/test_fixedpoint.py.
#include <cstdint>
#include <cstdio>
#include <random>
#include "../../oracle/fixedpoint.h"
#include "../../tiny-object-detection_b200/csrc/fixedpoint.cuh"

int main() {
  std::mt19937_64 rng(42);
  long long bad = 0, n = 0;
  const int32_t edge[] = {0, 1, -1, 2, -2, 127, -128, 255, (1 << 20), -(1 << 20), (1 << 26) - 1, -(1 << 26), (1 << 30) - 1, -(1 << 30) + 1,
                          1073741823, 1073741824 - 5, 123456789, -123456789};
  const int32_t qs[] = {0, 1, 1 << 30, (1 << 30) + 1, 0x7FFFFFFF, 1518500250, 1073741824 + 12345, 2000000000};
  auto check = [&](int32_t x, int32_t q, int sh) {
    const int32_t a = oracle::MBQM(x, q, sh), b = tod::mul_by_quant_mult(x, q, sh);
    ++n;
    if (a != b) { if (bad < 5) std::printf("generic mismatch x=%d q=%d sh=%d: %d vs %d\n", x, q, sh, a, b); ++bad; }
    const int64_t lim = int64_t(1) << 30;
    const int64_t xs = sh > 0 ? (int64_t(x) << sh) : x;
    if (q >= 0 && xs < lim && xs > -lim) {
      const int32_t c = tod::mul_by_quant_mult_fast(x, q, sh);
      if (a != c) { if (bad < 5) std::printf("fast mismatch x=%d q=%d sh=%d: %d vs %d\n", x, q, sh, a, c); ++bad; }
    }
  };
  for (int32_t x : edge)
    for (int32_t q : qs)
      for (int sh = -31; sh <= 4; ++sh) check(x, q, sh);
  // ties of the rounding shift: products landing exactly on .5
  for (int sh = -12; sh <= 0; ++sh)
    for (int k = -4000; k <= 4000; ++k) check(k, 1 << 30, sh), check(2 * k + 1, 1 << 30, sh), check(k, (1 << 30) + (1 << 29), sh);
  for (long long i = 0; i < 10000000; ++i) {
    const int32_t x = int32_t(rng() % (1u << 27)) - (1 << 26);
    const int32_t q = int32_t((1u << 30) + rng() % (1u << 30));
    const int sh = -int(rng() % 20) + (rng() % 50 == 0 ? 2 : 0);
    check(x, q, sh);
  }
  for (long long i = 0; i < 2000000; ++i) {  // the generic form over the full int32 range, either sign of q
    check(int32_t(rng()), int32_t(rng()), -int(rng() % 31));
  }
  // the tabulated epilogue form (zero point and rounding term folded into the 64-bit addend)
  auto check_tab = [&](int32_t x, int32_t qq, int rs, int zp) {
    const int32_t want = oracle::MBQM(x, qq, -rs) + zp;
    const int32_t got = tod::requant_tab(x, qq, rs, (1 << (rs - 1)) + zp * (1 << rs));
    ++n;
    if (want != got) { if (bad < 5) std::printf("tab mismatch x=%d q=%d rs=%d zp=%d: %d vs %d\n", x, qq, rs, zp, want, got); ++bad; }
  };
  const int32_t xe[] = {0, 1, -1, 2, -2, 3, -3, 127, -128, 1000, -1000, (1 << 20), -(1 << 20), (1 << 28), -(1 << 28), (1 << 29) - 1, -(1 << 29) + 1};
  const int32_t qe[] = {0, 1 << 30, (1 << 30) + 1, 0x7FFFFFFF, 1518500250, 2000000000};
  for (int32_t x : xe)
    for (int32_t qq : qe)
      for (int rs = 1; rs <= 22; ++rs)
        for (int zp : {-128, -3, 0, 5, 127, 128}) check_tab(x, qq, rs, zp);
  for (int rs = 1; rs <= 14; ++rs)  // exact .5 ties of the rounding shift, both signs
    for (int k = -5000; k <= 5000; ++k) check_tab(k, 1 << 30, rs, -128), check_tab(2 * k + 1, 1 << 30, rs, 7), check_tab(k, (1 << 30) + (1 << 29), rs, 0);
  for (long long i = 0; i < 10000000; ++i) {
    const int32_t x = int32_t(rng() % (1u << 30)) - (1 << 29);
    const int32_t qq = rng() % 64 == 0 ? int32_t(rng() % (1u << 30)) : int32_t((1u << 30) + rng() % (1u << 30));
    check_tab(x, qq, 1 + int(rng() % 22), int(rng() % 257) - 128);
  }
  // the ReLU form: equal to the literal result after the clamp whenever the activation floor is at or above the zero point
  auto check_relu = [&](int32_t x, int32_t qq, int rs, int zp, int lo, int hi, int32_t bias) {
    auto clamp = [&](int64_t v) { return int32_t(v < lo ? lo : (v > hi ? hi : v)); };
    const int32_t want = clamp(int64_t(oracle::MBQM(x + bias, qq, -rs)) + zp);
    const int32_t got = clamp(tod::requant_relu(x, qq, rs - 1, tod::relu_addend(qq, rs, zp, bias)));
    ++n;
    if (want != got) { if (bad < 5) std::printf("relu mismatch x=%d b=%d q=%d rs=%d zp=%d [%d,%d]: %d vs %d\n", x, bias, qq, rs, zp, lo, hi, want, got); ++bad; }
  };
  for (int32_t x : xe)
    for (int32_t qq : qe)
      for (int rs = 1; rs <= 22; ++rs)
        for (int zp : {-128, -100, -3, 0, 5}) {
          check_relu(x, qq, rs, zp, zp, 127, 0);                       // ReLU
          check_relu(x, qq, rs, zp, zp, zp + 100 > 127 ? 127 : zp + 100, 0);   // ReLU6-style ceiling
          check_relu(x, qq, rs, zp, zp + 3 > 127 ? 127 : zp + 3, 127, 17);     // floor above the zero point, with a bias
        }
  for (int rs = 1; rs <= 14; ++rs)  // exact .5 ties, both signs, around zero
    for (int k = -5000; k <= 5000; ++k)
      check_relu(k, 1 << 30, rs, -128, -128, 127, 0), check_relu(2 * k + 1, 1 << 30, rs, -7, -7, 127, -3), check_relu(k, (1 << 30) + (1 << 29), rs, 0, 0, 127, 1);
  for (long long i = 0; i < 10000000; ++i) {
    const int32_t x = int32_t(rng() % (1u << 29)) - (1 << 28);
    const int32_t b = int32_t(rng() % (1u << 24)) - (1 << 23);
    const int32_t qq = rng() % 64 == 0 ? int32_t(rng() % (1u << 30)) : int32_t((1u << 30) + rng() % (1u << 30));
    const int zp = int(rng() % 200) - 128;
    const int lo = zp + int(rng() % 3 == 0 ? rng() % 20 : 0);
    check_relu(x, qq, 1 + int(rng() % 22), zp, lo > 127 ? 127 : lo, 127, b);
  }
  int32_t q; int sh, q2, sh2;
  for (int i = 0; i < 200000; ++i) {
    const double m = std::ldexp(0.5 + (rng() % 1000000) / 2000000.0, -int(rng() % 40) + 3);
    oracle::QuantizeMultiplier(m, &q, &sh);
    tod::quantize_multiplier(m, &q2, &sh2);
    ++n;
    if (q != q2 || sh != sh2) ++bad;
  }
  std::printf("checked %lld cases, %lld mismatches\n", n, bad);
  return bad ? 1 : 0;
}
