// Host check of the generated terrain stamp (csrc/stamp_pattern.inc, used by scene.cu::stamp_pruned_kernel):
// the unrolled PRMT / packed-max sequence applied by one lane must equal the direct definition
//   tile[row0 + 8 + dy][col + dx] = max(., T[class(dx*dx + dy*dy)])   for dx*dx + dy*dy <= 80
// for both parities of the first tile row.  Build: g++ -O1 -o /tmp/spc tests/cpp/stamp_pattern_check.cpp && /tmp/spc
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

static uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t v = (uint64_t(b) << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const int n = (sel >> (4 * i)) & 7;  // selectors here never set the sign-replicate bit
    r |= uint32_t((v >> (8 * n)) & 0xFF) << (8 * i);
  }
  return r;
}
static uint32_t vmaxu2(uint32_t a, uint32_t b) {
  const uint32_t lo = (a & 0xFFFF) > (b & 0xFFFF) ? (a & 0xFFFF) : (b & 0xFFFF);
  const uint32_t hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
  return lo | (hi << 16);
}

static const int kClasses[36] = {0, 1, 2, 4, 5, 8, 9, 10, 13, 16, 17, 18, 20, 25, 26, 29, 32, 34, 36, 37, 40, 41, 45, 49, 50, 52, 53, 58, 61, 64, 65, 68, 72, 73, 74, 80};

int main() {
  std::mt19937 rng(7);
  const int pairs = 48, W = 64;
  int bad = 0;
  for (int iter = 0; iter < 2000; ++iter) {
    uint16_t T[36];
    for (int k = 0; k < 36; ++k) T[k] = uint16_t(rng() % 600);
    uint32_t t[18];
    for (int k = 0; k < 18; ++k) t[k] = uint32_t(T[2 * k]) | (uint32_t(T[2 * k + 1]) << 16);
    std::vector<uint32_t> tile(pairs * W), ref_rows(2 * pairs * W);
    for (auto& w : tile) w = (rng() % 300) | ((rng() % 300) << 16);
    for (int r = 0; r < 2 * pairs; ++r)
      for (int c = 0; c < W; ++c) ref_rows[r * W + c] = (tile[(r >> 1) * W + c] >> (16 * (r & 1))) & 0xFFFF;
    const int row0 = int(rng() % (2 * pairs - 17));  // tile row of dy = -8
    const int col = 8 + int(rng() % 48);
    for (int dx = -8; dx <= 8; ++dx)
      for (int dy = -8; dy <= 8; ++dy) {
        const int d2 = dx * dx + dy * dy;
        if (d2 > 80) continue;
        int k = -1;
        for (int i = 0; i < 36; ++i)
          if (kClasses[i] == d2) k = i;
        if (k < 0) { printf("d2 %d has no class\n", d2); return 1; }
        uint32_t& cell = ref_rows[(row0 + 8 + dy) * W + col + dx];
        if (T[k] > cell) cell = T[k];
      }
    uint32_t* base = tile.data() + (row0 >> 1) * W + col;
    const uint32_t sel = (row0 & 1) ? 0x5432u : 0x7654u;
#define PR_PRMT(a, b, s) prmt(a, b, s)
    // the kernel includes the pattern twice: upper words for every entry, lower words for the non-dominated ones
    {
#define PR_RMW_U(off, w) base[off] = vmaxu2(base[off], w)
#define PR_RMW_L(off, w)
#define PR_SYNC()
#include "../../tiny-object-detection_b200/csrc/stamp_pattern.inc"
#undef PR_RMW_U
#undef PR_RMW_L
    }
    {
#define PR_RMW_U(off, w)
#define PR_RMW_L(off, w) base[off] = vmaxu2(base[off], w)
#include "../../tiny-object-detection_b200/csrc/stamp_pattern.inc"
#undef PR_RMW_U
#undef PR_RMW_L
    }
    for (int r = 0; r < 2 * pairs; ++r)
      for (int c = 0; c < W; ++c) {
        const uint32_t got = (tile[(r >> 1) * W + c] >> (16 * (r & 1))) & 0xFFFF;
        if (got != ref_rows[r * W + c]) {
          if (bad < 5) printf("iter %d row0 %d col %d: cell (%d,%d) got %u want %u\n", iter, row0, col, r, c, got, ref_rows[r * W + c]);
          ++bad;
        }
      }
  }
  // the split the dominance argument needs: the lower words only touch rows dy >= 1
  for (int parity = 0; parity < 2; ++parity) {
    std::vector<uint32_t> tile(pairs * W, 0u);
    uint32_t t[18];
    for (int k = 0; k < 18; ++k) t[k] = 0x00010001u;
    const int row0 = 20 + parity, col = 30;
    uint32_t* base = tile.data() + (row0 >> 1) * W + col;
    const uint32_t sel = (row0 & 1) ? 0x5432u : 0x7654u;
#define PR_RMW_U(off, w)
#define PR_RMW_L(off, w) base[off] = vmaxu2(base[off], w)
#include "../../tiny-object-detection_b200/csrc/stamp_pattern.inc"
#undef PR_RMW_U
#undef PR_RMW_L
    for (int r = 0; r < 2 * pairs; ++r)
      for (int c = 0; c < W; ++c)
        if (((tile[(r >> 1) * W + c] >> (16 * (r & 1))) & 0xFFFF) && r - (row0 + 8) < 1) {
          printf("lower words touch row dy = %d (parity %d)\n", r - (row0 + 8), parity);
          ++bad;
        }
  }
  printf("stamp pattern: %s (%d mismatching cells)\n", bad ? "FAIL" : "ok", bad);
  return bad ? 1 : 0;
}
