// tcgen05 int8 implicit-GEMM convolution for sm_100a.  See conv_tc.h.
//
// Replaces the CONV_2D kernels TFLite runs inside interpreter.invoke() (/root/reference/src/yolact.rs:163)
// for stride-1 1x1 / 3x3 layers with Cin % 16 == 0 — >97 % of the graph's multiply-accumulates.
//
// Kernel anatomy (persistent, at most one CTA per SM - the launch asks for the CTAs that keep every round of tiles full, see
// conv_tc_launch - 12 warps):
//   warp 0   TMA producer   per (filter tap, K chunk): one 4-D activation box [BK ch x pw x ph x pn] fetched at the
//                           tap's shifted coordinates (out-of-image elements are zero-filled by the TMA unit = SAME
//                           padding) and one 3-D weight box [BK x 1 tap x BN], both 32/64/128B-swizzled, K-major
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=BN, K=32 per
//                           instruction) accumulating s32 in TMEM; tcgen05.commit releases smem stages / publishes
//                           the accumulator
//   warp 2   TMEM allocator 512 columns = two accumulator stages, so the epilogue of tile i overlaps the MMAs of i+1
//   warps 4-11 epilogue     tcgen05.ld 32 lanes x 16 columns -> registers; + bias with the input-zero-point
//                           correction of the taps that were inside the image; TFLite fixed-point requantisation
//                           (SRDHM + rounding shift, bit-exact); activation clamp; int8 store
//
// The input zero point: TFLite accumulates (in - zp) * w and *skips* out-of-image taps.  The tensor core multiplies
// raw int8, so  acc = sum in*w (zero-filled taps add 0)  and the epilogue adds  bias - zp * sum_{taps inside} wsum[tap],
// tabulated on the host per border class (which rows / columns of the filter are inside) x output channel.
#include "conv_tc.h"

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "common.h"
#include "fixedpoint.cuh"

namespace tod {
namespace {

constexpr int kBM = 128;
constexpr int kTcThreads = 384;  // 4 control warps + 8 epilogue warps
constexpr int kFastThreads = 384;  // conv_tc_fast_kernel: 4 control warps + kEpiWarps epilogue warps (16 measured 2 % slower per step)
constexpr int kEpiWarps = kFastThreads / 32 - 4;
constexpr int kAccStages = 2;     // accumulator stages in TMEM == epilogue groups (4 with 16 warps and BN <= 128 measured no faster)
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;  // TMEM columns between the two accumulator stages

struct TcParams {
  int Wd, Hd;                  // spatial extent the patch tiling runs over (flat 1x1: Wd = pixels of all tiles, Hd = 1)
  int OC, OCp;                 // real / padded (n_tiles * BN) output channels
  int pw, ph, pn, rows;        // M tile = pn images x ph rows x pw columns (rows = pw*ph*pn <= 128)
  int tiles_x, tiles_y;
  int KW, taps, kchunks, BK, BN, n_tiles;
  int KH, pad_top, pad_left, IH, IW;
  int stride;                  // 1 or 2 (both axes): the A box walks the input with TMA element strides
  int stages;
  int flat, HW;
  const int32_t* bias_eff;     // [classes][OCp]
  const int32_t* mult;         // [OCp]
  const int32_t* shift;        // [OCp]
  int32_t out_zp, act_min, act_max;
  int8_t* out;
  long long out_ts;
  uint32_t idesc;
  uint32_t tx_bytes;           // bytes landing per stage (A box + B box)
  uint32_t a_stage, b_stage;   // smem bytes reserved per stage for A / B (1024-aligned)
  uint32_t sbo;                // 8 rows * BK bytes
  uint32_t layout;             // UMMA swizzle code
  int vec_store;               // OC % 16 == 0 and 16B-aligned rows
  int sc, pitch;               // epilogue staging: columns per pass, bytes per staged row
  const uint8_t* post_lut;     // optional fused byte map (QUANTIZE / RELU / TANH chain), 256 entries
  // ---- fast epilogue (conv_tc_fast_kernel)
  const int4* qtab;            // [OCp] {Q31 multiplier, right shift, 2^31, rounding term with the output zero point folded in}
  const int32_t* b2tab;        // [ncls][OCp] 2 * (bias - zp * sum of in-image tap sums), compact border classes
  const long long* a64tab;     // kEpiRelu: [ncls][OCp] bias_eff * q + 2^30 + halfp * 2^31 (qtab then holds {q, rs - 1, -, -})
  int ncls, ncls_x;            // compact border classes: cls = ymap[ymask] * ncls_x + xmap[xmask]
  uint8_t ymap[8], xmap[8];
  int wo;                      // TMA-store staging row bytes (16 / 32 / 64 / 128), 0 = manual stores
  uint32_t stage_bytes;        // shared memory reserved for output staging
  // ---- cp.async A producer (flat 1x1 layers with short pixel rows: the TMA unit serves ~1 box row per 6-8 cycles
  // whatever its length, so 32-byte pixels starve it; 96 threads issuing 16-byte cp.async do not care)
  int dbg;                     // TOD_TC_DBG timing experiments: 1 = skip the output store
  int b_res;                   // fast kernel: the weights of the (single) N tile stay resident in shared memory, stages hold A only
  int wide;                    // epilogue: 1 = all eight warps convert every tile (two warps per TMEM lane quarter take alternate 16-column
                               // chunks; tiles alternate accumulator stages), 0 = two groups of four warps take alternate tiles.  Chosen per
                               // launch: with 1 - 3 tiles per CTA the last tile's epilogue is the CTA's tail, and eight warps halve it
  int OCm;                     // accumulator columns to convert (== OC, or OC + one 16-column chunk of a sibling layer's channels)
  int x_c0, x_cols;            // sibling output: its chunk's first column (== OC) and its real channel count (0 = none)
  int8_t* x_out;               // [tile][Hd][Wd][x_cols]
  long long x_out_ts;
  const uint8_t* x_lut;        // sibling byte map (kEpiLut; null = the conv's own map is not applied to the sibling either)
  int late_trig;               // 1 = griddepcontrol.launch_dependents when the CTA starts its LAST work item (the dependent grid's CTAs then
                               // wait one tile on their SMs, not the whole launch), 0 = at kernel start
  int lin;                     // fast kernel, TMA mode: tile rows are contiguous in global memory, staged linearly, one 1-D bulk store
  int run_w;                   // manual stores: pixels per staged run (pw, or pw * ph when the patch spans the image width)
  long long* trace;            // TOD_TC_TRACE: clock64 stamps of CTA 0 (epilogue warp 4 / MMA warp / producer), 16 per tile
  int a_cp;                    // 1 = A tiles are gathered with cp.async by warps 0, 2, 3
  const int8_t* in;            // flat [pixels][IC]
  int IC;
  // ---- fused residual ADD (kEpiAdd): out = requant_out(tab[conv byte] + tab[256 + residual byte])
  const int8_t* resid;         // same geometry as the output
  long long resid_ts;
  const int32_t* add_tab;      // device [512]: conv-side table, then residual-side table (terms rescaled to 2^20 fixed point)
  int32_t add_mult, add_shift, add_zp, add_min, add_max;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug becomes a trap (an error code on the host) instead of a hung GPU
// The suspend-time hint keeps a waiting warp parked in hardware instead of re-issuing try_wait every ~100 cycles: the
// spin loops of the producer / MMA warps were 11 % of all executed instructions on the epilogue-bound layers (ncu).
constexpr uint32_t kSuspendHintNs = 20000u;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(kSuspendHintNs)
        : "memory");
    if (!ok && spin > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {  // src_bytes = 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {  // arrives once this thread's earlier cp.async have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int kCpThreads = 96;

__device__ __forceinline__ bool elect_one() {  // true in exactly one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// d = sat_s8(hi) << 8 | sat_s8(lo) in the low half, `upper`'s low half in the high half
__device__ __forceinline__ uint32_t pack_sat_s8(int hi, int lo, uint32_t upper) {
  uint32_t d;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
  return d;
}

// K-major, swizzled shared-memory matrix descriptor (SM100 format: version 1 at bit 46)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;                      // leading byte offset: unused for swizzled K-major
  d |= uint64_t((sbo >> 4) & 0x3FFFu) << 32;   // stride between 8-row core-matrix groups
  d |= uint64_t(1) << 46;                      // descriptor version (Blackwell)
  d |= uint64_t(layout & 7u) << 61;
  return d;
}

struct SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint64_t b_full;   // resident weights have landed (TcParams::b_res)
  uint32_t tmem_base;
  uint8_t lut[256];
  uint8_t lut2[256];           // sibling output's byte map
  int32_t add_tab[512];
};

__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p, const int tiles) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible to the compiler
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + size_t(p.stages) * p.a_stage;
  uint8_t* stage_buf = smem_b + size_t(p.stages) * p.b_stage;                      // [128][pitch] requantised bytes
  long long* s_rowoff = reinterpret_cast<long long*>(stage_buf + size_t(kBM) * p.pitch);  // [128] global offset of each row
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(s_rowoff + kBM);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wd = p.flat ? tiles * p.HW : p.Wd;
  const int tiles_x = p.flat ? (Wd + kBM - 1) / kBM : p.tiles_x;
  const int groups = p.flat ? 1 : (tiles + p.pn - 1) / p.pn;
  const int m_tiles = groups * p.tiles_y * tiles_x;
  const int total_work = m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->acc_full[s], 1);
      mbar_init(&ctl->acc_empty[s], kEpiThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (p.post_lut && threadIdx.x >= 128) ctl->lut[threadIdx.x - 128] = p.post_lut[threadIdx.x - 128];
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  pdl_wait();  // everything above touched constants only; activations (and our output buffer) are safe from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int n_tile = work % p.n_tiles;
        int m = work / p.n_tiles;
        const int tx = m % tiles_x;
        m /= tiles_x;
        const int ty = m % p.tiles_y;
        const int g = m / p.tiles_y;
        const int x0 = tx * p.pw, y0 = ty * p.ph, n0 = g * p.pn;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int fy = tap / p.KW, fx = tap - fy * p.KW;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            mbar_expect_tx(&ctl->full[stage], p.tx_bytes);
            tma_load_4d(smem_a + size_t(stage) * p.a_stage, &map_a, &ctl->full[stage], kc * p.BK, x0 * p.stride + fx - p.pad_left, y0 * p.stride + fy - p.pad_top, n0);
            tma_load_3d(smem_b + size_t(stage) * p.b_stage, &map_b, &ctl->full[stage], kc * p.BK, tap, n_tile * p.BN);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint64_t adesc0 = make_desc(smem_u32(smem_a), p.sbo, p.layout), bdesc0 = make_desc(smem_u32(smem_b), p.sbo, p.layout);
    const uint32_t a_step = p.a_stage >> 4, b_step = p.b_stage >> 4;  // descriptor address field is in 16-byte units
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t use = uint32_t(it >> 1);
      mbar_wait(&ctl->acc_empty[as], (use & 1) ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kAccStride;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after();
        // the whole warp runs this loop converged with warp-uniform values, so the descriptors live in uniform registers;
        // only the tcgen05 instructions themselves are predicated on the elected lane
        const uint64_t ad0 = adesc0 + uint64_t(uint32_t(stage) * a_step), bd0 = bdesc0 + uint64_t(uint32_t(stage) * b_step);
        for (int kk = 0; kk < p.BK / 32; ++kk)
          if (leader) umma_i8(tmem_d, ad0 + uint64_t(2 * kk), bd0 + uint64_t(2 * kk), p.idesc, (k | kk) != 0 ? 1u : 0u);
        if (leader) {
          umma_commit(&ctl->empty[stage]);                       // frees the smem stage when these MMAs retire
          if (k == k_iters - 1) umma_commit(&ctl->acc_full[as]);  // accumulator complete
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // TMEM -> registers -> requantise -> shared staging tile -> coalesced 16-byte global stores.
    // Two warps share each TMEM lane quarter and split the columns (even / odd 16-column chunks).
    const int ew = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = (warp - 4) >> 2;        // which 16-column chunks of a pass this warp converts
    const int r = ew * 32 + lane;            // accumulator row == pixel of the tile
    const int et = threadIdx.x - 128;        // 0..255 among the epilogue threads
    int it = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t use = uint32_t(it >> 1);
      const int n_tile = work % p.n_tiles;
      int m = work / p.n_tiles;
      const int tx = m % tiles_x;
      m /= tiles_x;
      const int ty = m % p.tiles_y;
      const int g = m / p.tiles_y;
      // row -> pixel
      const int wx = r % p.pw;
      const int rest = r / p.pw;
      const int wy = rest % p.ph;
      const int wn = rest / p.ph;
      const int x = tx * p.pw + wx, yy = ty * p.ph + wy, n = g * p.pn + wn;
      const bool valid = r < p.rows && x < Wd && yy < p.Hd && (p.flat || n < tiles);
      // border class of this pixel: which filter rows / columns are inside the image
      int cls = 0;
      if (valid) {
        int px = x, py = yy;
        if (p.taps == 1) {  // 1x1: the single tap is always inside (x may run over a flattened pixel index)
          px = p.pad_left;
          py = p.pad_top;
        }
        int ymask = 0, xmask = 0;
        for (int f = 0; f < p.KH; ++f) {
          const int iy = py * p.stride + f - p.pad_top;
          ymask |= (iy >= 0 && iy < p.IH) ? (1 << f) : 0;
        }
        for (int f = 0; f < p.KW; ++f) {
          const int ix = px * p.stride + f - p.pad_left;
          xmask |= (ix >= 0 && ix < p.IW) ? (1 << f) : 0;
        }
        cls = ymask * (1 << p.KW) + xmask;
      }
      const int ocb = n_tile * p.BN;
      const int4* be = reinterpret_cast<const int4*>(p.bias_eff + size_t(cls) * p.OCp + ocb);
      const int4* mu = reinterpret_cast<const int4*>(p.mult + ocb);
      const int4* sh = reinterpret_cast<const int4*>(p.shift + ocb);
      s_rowoff[r] = valid ? ((p.flat ? 0ll : (long long)n * p.out_ts) + ((long long)yy * Wd + x) * p.OC) : -1ll;
      const int ncols_tile = min(p.BN, p.OC - ocb);

      mbar_wait(&ctl->acc_full[as], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * kAccStride;
      for (int pass0 = 0; pass0 < p.BN; pass0 += p.sc) {
        const int pass_cols = min(p.sc, p.BN - pass0);
        for (int c0 = half * 16; c0 < pass_cols; c0 += 32) {
          uint32_t v[16];
          tmem_ld16(taddr + pass0 + c0, v);
          tmem_wait_ld();
          uint32_t packed[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int4 b4 = __ldg(be + ((pass0 + c0) >> 2) + q4);
            const int4 m4 = __ldg(mu + ((pass0 + c0) >> 2) + q4);
            const int4 s4 = __ldg(sh + ((pass0 + c0) >> 2) + q4);
            int q0 = mul_by_quant_mult_fast(int32_t(v[4 * q4 + 0]) + b4.x, m4.x, s4.x) + p.out_zp;
            int q1 = mul_by_quant_mult_fast(int32_t(v[4 * q4 + 1]) + b4.y, m4.y, s4.y) + p.out_zp;
            int q2 = mul_by_quant_mult_fast(int32_t(v[4 * q4 + 2]) + b4.z, m4.z, s4.z) + p.out_zp;
            int q3 = mul_by_quant_mult_fast(int32_t(v[4 * q4 + 3]) + b4.w, m4.w, s4.w) + p.out_zp;
            q0 = max(p.act_min, min(p.act_max, q0));
            q1 = max(p.act_min, min(p.act_max, q1));
            q2 = max(p.act_min, min(p.act_max, q2));
            q3 = max(p.act_min, min(p.act_max, q3));
            if (p.post_lut) {
              q0 = ctl->lut[q0 & 0xFF];
              q1 = ctl->lut[q1 & 0xFF];
              q2 = ctl->lut[q2 & 0xFF];
              q3 = ctl->lut[q3 & 0xFF];
            }
            packed[q4] = (uint32_t(q0) & 0xFFu) | ((uint32_t(q1) & 0xFFu) << 8) | ((uint32_t(q2) & 0xFFu) << 16) | (uint32_t(q3) << 24);
          }
          *reinterpret_cast<uint4*>(stage_buf + size_t(r) * p.pitch + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        if (pass0 + p.sc >= p.BN) {  // accumulator fully read: hand the TMEM stage back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->acc_empty[as]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int ncols = min(pass_cols, ncols_tile - pass0);  // real output channels in this pass (<= 0: padding only)
        if (ncols > 0) {
          if (p.vec_store) {
            const int cpr = ncols >> 4;
            for (int idx = et; idx < kBM * cpr; idx += kEpiThreads) {
              const int rr = idx / cpr, ch = idx - rr * cpr;
              const long long off = s_rowoff[rr];
              if (off >= 0)
                *reinterpret_cast<uint4*>(p.out + off + ocb + pass0 + ch * 16) = *reinterpret_cast<const uint4*>(stage_buf + size_t(rr) * p.pitch + ch * 16);
            }
          } else {
            for (int idx = et; idx < kBM * ncols; idx += kEpiThreads) {
              const int rr = idx / ncols, bb = idx - rr * ncols;
              const long long off = s_rowoff[rr];
              if (off >= 0) p.out[off + ocb + pass0 + bb] = int8_t(stage_buf[size_t(rr) * p.pitch + bb]);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ fast-epilogue variant
// Same producer / MMA pipeline; the epilogue is rebuilt around what the first profile showed (it executed ~26 SASS
// instructions per output element and bounded every layer, large or small):
//   * per-channel constants live in shared memory ({q, rounding term, shift} and 2*bias per border class), read as
//     warp-wide broadcasts instead of three __ldg streams;
//   * TFLite's MultiplyByQuantizedMultiplier for shift <= -1 collapses to
//         x2 = 2*acc + 2*bias;   out = (hi32(x2 * q + 2^31 + ((half + (zp << rs)) << 32)) + (x2 >> 31)) >> rs
//     (identical to the double-rounding reference form: hi32(x2*q + 2^31) is SRDHM's result v, and sign(x2) differs
//     from sign(v) only where v == 0 and the rounding term cannot change the quotient; the host verifies the
//     preconditions, tests/test_fixedpoint.py checks the identity), 4 integer instructions + one LDS.128;
//   * a full-range clamp is the saturation of cvt.pack.sat.s8.s32;
//   * requantised bytes go to a swizzled staging tile and leave with ONE TMA store per 128-column pass (the TMA unit
//     clips rows / columns outside the tensor), double-buffered so a pass needs a single CTA-wide barrier.
// kEpiRelu: the layer's activation floor is at (or above) the output zero point (ReLU / ReLU6: act_min == quantised 0.0).
// Every negative pre-activation then clamps to act_min whatever its rounding, so the sign correction of the rounding
// right shift can go, and with it the doubling: with Y = acc*q + (bias*q + 2^30 + halfp*2^31) (one 64-bit
// multiply-add against a per-(border class, channel) addend), floor(Y / 2^(31+rs)) == hi32(Y) >> (rs - 1) is the exact
// result for acc + bias >= 0 and is <= zp (so clamps to act_min like the exact one) below: 2 ALU instructions per output
// instead of 5.  tests/cpp/fixedpoint_check.cpp checks the identity against the literal gemmlowp form.
enum : uint32_t { kEpiSat = 1, kEpiLut = 2, kEpiTma = 4, kEpiAdd = 8, kEpiRelu = 16 };

// n 16-byte words global -> shared, four loads per thread in flight per round
__device__ __forceinline__ void copy_table16(int4* dst, const int4* __restrict__ src, int n, int nthreads) {
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * nthreads) {
    int4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * nthreads < n) v[u] = __ldg(src + i0 + u * nthreads);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * nthreads < n) dst[i0 + u * nthreads] = v[u];
  }
}

struct WorkItem { int n_tile, tx, ty, g; };
__device__ __forceinline__ WorkItem decode_work(int work, int n_tiles, int tiles_x, int tiles_y) {
  WorkItem w;
  unsigned m = unsigned(work);
  if (n_tiles == 1 && tiles_y == 1 && m < unsigned(tiles_x)) {  // flat 1x1 layers and single-row tilings: no division at all
    w.n_tile = 0;
    w.tx = int(m);
    w.ty = 0;
    w.g = 0;
    return w;
  }
  if (n_tiles > 1) {
    w.n_tile = int(m % unsigned(n_tiles));
    m /= unsigned(n_tiles);
  } else {
    w.n_tile = 0;
  }
  w.tx = int(m % unsigned(tiles_x));
  m /= unsigned(tiles_x);
  if (tiles_y > 1) {
    w.ty = int(m % unsigned(tiles_y));
    w.g = int(m / unsigned(tiles_y));
  } else {
    w.ty = 0;
    w.g = int(m);
  }
  return w;
}

// One 16-column chunk of one accumulator row: requantise (+ fused ADD / byte map / clamp) and pack to 16 output bytes.
// kq / bq / aq point at this chunk's per-channel constants in shared memory (every load is base + immediate).
template <uint32_t MODE>
__device__ __forceinline__ void epi_chunk16(const uint32_t (&v)[16], const int4* kq, const int4* bq, const longlong2* aq, const uint4& rres,
                                            const SmemCtl* ctl, const TcParams& p, uint32_t (&packed)[4], const uint8_t* lut = nullptr) {
  if (!lut) lut = ctl->lut;
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    int o[4];
    if (MODE & kEpiRelu) {
      // kq is a table of int2 {q, rs - 1} here (two channels per 16-byte load): the shared-memory pipe, not the ALU, is the
      // busiest unit of the epilogue (ncu: 61 % of its wavefront rate on the 112x112 expand layer)
      const longlong2 a01 = aq[2 * q4], a23 = aq[2 * q4 + 1];
      const int4 k01 = reinterpret_cast<const int4*>(kq)[2 * q4], k23 = reinterpret_cast<const int4*>(kq)[2 * q4 + 1];
      o[0] = int((static_cast<long long>(int(v[4 * q4 + 0])) * k01.x + a01.x) >> 32) >> k01.y;
      o[1] = int((static_cast<long long>(int(v[4 * q4 + 1])) * k01.z + a01.y) >> 32) >> k01.w;
      o[2] = int((static_cast<long long>(int(v[4 * q4 + 2])) * k23.x + a23.x) >> 32) >> k23.y;
      o[3] = int((static_cast<long long>(int(v[4 * q4 + 3])) * k23.z + a23.y) >> 32) >> k23.w;
    } else {
      const int4 b4 = bq[q4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int4 k = kq[4 * q4 + j];  // {q, rs, 2^31, half + (zp << rs)}: z:w is the 64-bit addend, in place
        const int bj = j == 0 ? b4.x : (j == 1 ? b4.y : (j == 2 ? b4.z : b4.w));
        int x2;  // 2 * acc + 2 * bias on the FMA pipe (IMAD): the ALU pipe is the epilogue's bottleneck
        asm("mad.lo.s32 %0, %1, 2, %2;" : "=r"(x2) : "r"(int(v[4 * q4 + j])), "r"(bj));
        const long long addend = static_cast<long long>((static_cast<unsigned long long>(uint32_t(k.w)) << 32) | uint32_t(k.z));
        const int t = int((static_cast<long long>(x2) * k.x + addend) >> 32);
        o[j] = (t + (x2 >> 31)) >> k.y;  // sign(x2) == sign(v) wherever the rounding term can matter
      }
    }
    if (MODE & kEpiAdd) {
      const uint32_t rw = q4 == 0 ? rres.x : (q4 == 1 ? rres.y : (q4 == 2 ? rres.z : rres.w));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!(MODE & kEpiSat)) o[j] = max(p.act_min, min(p.act_max, o[j]));
        else o[j] = max(-128, min(127, o[j]));
        const int sum = ctl->add_tab[o[j] & 0xFF] + ctl->add_tab[256 + ((rw >> (8 * j)) & 0xFFu)];
        o[j] = max(p.add_min, min(p.add_max, mul_by_quant_mult_fast(sum, p.add_mult, p.add_shift) + p.add_zp));
      }
      packed[q4] = (uint32_t(o[0]) & 0xFFu) | ((uint32_t(o[1]) & 0xFFu) << 8) | ((uint32_t(o[2]) & 0xFFu) << 16) | (uint32_t(o[3]) << 24);
    } else if ((MODE & kEpiSat) && !(MODE & kEpiLut)) {
      packed[q4] = pack_sat_s8(o[1], o[0], pack_sat_s8(o[3], o[2], 0u));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[j] = max(p.act_min, min(p.act_max, o[j]));
        if (MODE & kEpiLut) o[j] = lut[o[j] & 0xFF];
      }
      packed[q4] = (uint32_t(o[0]) & 0xFFu) | ((uint32_t(o[1]) & 0xFFu) << 8) | ((uint32_t(o[2]) & 0xFFu) << 16) | (uint32_t(o[3]) << 24);
    }
  }
}

// DIAG = true compiles in the clock64 stamps (TOD_TC_TRACE) and the timing-experiment switches (TOD_TC_DBG); the production
// instantiations carry neither (predicated-off stamps inside the chunk loop alone were measurable).
// FL = true: the flat-linear specialisation (p.flat && p.lin: a 1x1 layer seen as one [pixels][OC] matrix, one N tile,
// linear staging, one 1-D bulk store per tile).  ncu on the 112x112 32 -> 96 layer showed the epilogue warps spending as
// many stall samples in the ~200 instructions of per-tile bookkeeping of the general path (work decode, border class,
// patch coordinates, run set-up, pass loop) as in the six 16-column chunks themselves; here a tile costs a handful.
// OCT > 0 (flat-linear ReLU layers only): the layer's output channel count at compile time and its per-channel requantisation
// constants as a kernel parameter (FlTab, constant bank).  With the chunk loop fully unrolled every constant is a
// constant-bank operand of the multiply-add / shift itself: no table loads on the shared-memory pipe, which - with the
// accumulator read - is what bounds these layers (~600 of the ~1100 wavefronts per 128 x 96 tile were
// the broadcast loads of the constants).
constexpr int kFlTabMax = 192;
struct FlTab {
  long long a64[kFlTabMax];  // bias * q + 2^30 + halfp * 2^31 (fixedpoint.cuh::relu_addend)
  int32_t q[kFlTabMax];
  uint8_t rs1[kFlTabMax];    // right shift - 1 (bytes: the whole parameter block stays below 4 KB)
};

template <uint32_t MODE, bool DIAG, bool FL, int OCT>
__device__ __forceinline__ void conv_tc_fast_body(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_o, const TcParams& p,
                                                  const int tiles, const FlTab* tab) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space visible to the compiler
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + size_t(p.stages) * p.a_stage;
  uint8_t* stage_buf = smem_b + size_t(p.b_res ? p.kchunks : p.stages) * p.b_stage;  // 1024-aligned (both stage sizes are)
  long long* s_rowoff = reinterpret_cast<long long*>(stage_buf + p.stage_bytes);
  int* s_runlen = reinterpret_cast<int*>(s_rowoff + 4 * kBM);   // [4 groups][128 runs]
  int4* s_qtab = reinterpret_cast<int4*>(s_runlen + 4 * kBM);
  int32_t* s_b2 = reinterpret_cast<int32_t*>(s_qtab + p.OCp);
  const long long* s_a64 = reinterpret_cast<const long long*>(s_b2);   // kEpiRelu: the same region holds 64-bit addends
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(s_b2 + size_t(p.ncls) * p.OCp * ((MODE & kEpiRelu) ? 2 : 1));

  if (!p.late_trig) pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wd = p.flat ? tiles * p.HW : p.Wd;
  const int tiles_x = p.flat ? (Wd + kBM - 1) / kBM : p.tiles_x;
  const int groups = p.flat ? 1 : (tiles + p.pn - 1) / p.pn;
  const int total_work = groups * p.tiles_y * tiles_x * p.n_tiles;
  const int k_iters = p.taps * p.kchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctl->full[s], p.a_cp ? kCpThreads + (p.b_res ? 0 : 1) : 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&ctl->acc_full[s], 1);
      mbar_init(&ctl->acc_empty[s], p.wide ? kEpiWarps : kEpiWarps / kAccStages);  // the warps that read each accumulator stage
    }
    mbar_init(&ctl->b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // epilogue tables -> shared memory with 16-byte loads, all of a thread's loads in flight at once (a 3x3 ReLU layer's
  // nine border classes are 18 KB: the word-by-word loop was twelve dependent round trips of prologue per CTA)
  if constexpr (OCT == 0) {
    copy_table16(s_qtab, p.qtab, p.OCp, kFastThreads);
    if (MODE & kEpiRelu) copy_table16(reinterpret_cast<int4*>(s_b2), reinterpret_cast<const int4*>(p.a64tab), p.ncls * p.OCp / 2, kFastThreads);
    else copy_table16(reinterpret_cast<int4*>(s_b2), reinterpret_cast<const int4*>(p.b2tab), p.ncls * p.OCp / 4, kFastThreads);
  }
  if ((MODE & kEpiLut) && threadIdx.x >= 128 && threadIdx.x < 384) {
    ctl->lut[threadIdx.x - 128] = p.post_lut[threadIdx.x - 128];
    if (p.x_cols) ctl->lut2[threadIdx.x - 128] = p.x_lut ? p.x_lut[threadIdx.x - 128] : uint8_t(threadIdx.x - 128);
  }
  if (MODE & kEpiAdd)
    for (int i = threadIdx.x; i < 512; i += kFastThreads) ctl->add_tab[i] = p.add_tab[i];
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (p.b_res && threadIdx.x == 0) {
    // resident weights (every tile of this CTA multiplies by the same N tile): the whole matrix once, and - weights being
    // constants - ahead of the wait for the previous kernel, so the load runs under its tail
    mbar_expect_tx(&ctl->b_full, uint32_t(p.kchunks * p.BN * p.BK));
    for (int kc = 0; kc < p.kchunks; ++kc) tma_load_3d(smem_b + size_t(kc) * p.b_stage, &map_b, &ctl->b_full, kc * p.BK, 0, 0);
  }
  pdl_wait();  // everything above touched constants only; activations (and our output buffer) are safe from here on

  if (p.a_cp && (warp == 0 || warp == 2 || warp == 3)) {
    // ===================== cp.async A producer (flat 1x1) + TMA for the weights =====================
    const int pt = (warp == 0 ? 0 : warp - 1) * 32 + lane;  // 0 .. 95
    const int cpr_shift = p.BK == 128 ? 3 : (p.BK == 64 ? 2 : 1);  // 16-byte chunks per row = 1 << cpr_shift
    const int cc = pt & ((1 << cpr_shift) - 1), m0 = pt >> cpr_shift, mstep = kCpThreads >> cpr_shift;
    const long long sstep = (long long)mstep * p.IC;
    const uint32_t ostep = uint32_t(mstep) * uint32_t(p.BK);
    const uint32_t swz_mask = p.BK == 128 ? 7u : (p.BK == 64 ? 3u : 1u);
    if (pt == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    int stage = 0;
    uint32_t phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      if (p.late_trig && pt == 0 && work + int(gridDim.x) >= total_work) pdl_trigger();
      const WorkItem w = decode_work(work, p.n_tiles, tiles_x, p.tiles_y);
      const long long row0 = (long long)w.tx * kBM;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        if (pt == 0 && !p.b_res) {
          mbar_expect_tx(&ctl->full[stage], uint32_t(p.BN * p.BK));
          tma_load_3d(smem_b + size_t(stage) * p.b_stage, &map_b, &ctl->full[stage], kc * p.BK, 0, w.n_tile * p.BN);
        }
        const uint32_t a_base = smem_u32(smem_a + size_t(stage) * p.a_stage);
        // 96 threads and 2 / 4 / 8 chunks per row: a thread keeps the same 16-byte chunk column `cc` and walks the rows
        // m0, m0 + mstep, ...; everything but the row offset is loop-invariant (the first version recomputed row, chunk,
        // 64-bit source address and bounds per copy: 49 instructions per 16 bytes, and the producer warps - not the
        // epilogue - bounded the K = 96 ... 192 project layers at 1600-1800 cycles per K chunk)
        const int kbyte = kc * p.BK + cc * 16;
        const bool kok = kbyte < p.IC;
        const int8_t* src = p.in + (row0 + m0) * p.IC + kbyte;
        const int rows_ok = int(min((long long)kBM, (long long)Wd - row0));  // rows of this tile inside the tensor
        uint32_t off0 = uint32_t(m0) * uint32_t(p.BK) + uint32_t(cc) * 16u;
#pragma unroll 4
        for (int m = m0; m < kBM; m += mstep) {
          const bool ok = kok && m < rows_ok;
          const uint32_t off = off0 ^ (((off0 >> 7) & swz_mask) << 4);
          cp_async16(a_base + off, ok ? src : p.in, ok ? 16u : 0u);
          src += sstep;
          off0 += ostep;
        }
        cp_async_arrive(&ctl->full[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        if (p.late_trig && work + int(gridDim.x) >= total_work) pdl_trigger();
        const WorkItem w = decode_work(work, p.n_tiles, tiles_x, p.tiles_y);
        const int x0 = w.tx * p.pw, y0 = w.ty * p.ph, n0 = w.g * p.pn;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int fy = tap / p.KW, fx = tap - fy * p.KW;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            if (DIAG && (p.dbg & 4)) {  // timing experiment: no operand traffic at all
              mbar_arrive(&ctl->full[stage]);
            } else {
            mbar_expect_tx(&ctl->full[stage], p.tx_bytes);
            tma_load_4d(smem_a + size_t(stage) * p.a_stage, &map_a, &ctl->full[stage], kc * p.BK, x0 * p.stride + fx - p.pad_left, y0 * p.stride + fy - p.pad_top, n0);
            if (!p.b_res) tma_load_3d(smem_b + size_t(stage) * p.b_stage, &map_b, &ctl->full[stage], kc * p.BK, tap, w.n_tile * p.BN);
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    const uint64_t adesc0 = make_desc(smem_u32(smem_a), p.sbo, p.layout), bdesc0 = make_desc(smem_u32(smem_b), p.sbo, p.layout);
    const uint32_t a_step = p.a_stage >> 4, b_step = p.b_stage >> 4;  // descriptor address field is in 16-byte units
    const uint32_t acc_stride = uint32_t(kTmemCols / kAccStages);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (p.b_res && blockIdx.x < total_work) {
      mbar_wait(&ctl->b_full, 0);
      tc_fence_after();
    }
    for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++it) {
      const int as = it & (kAccStages - 1);
      const uint32_t use = uint32_t(it) / uint32_t(kAccStages);
      mbar_wait(&ctl->acc_empty[as], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * acc_stride;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after();
        if (p.a_cp) fence_async_smem();  // cp.async wrote A through the generic proxy; the MMA reads it through the async one
        // the whole warp runs this loop converged with warp-uniform values, so the descriptors live in uniform registers;
        // only the tcgen05 instructions themselves are predicated on the elected lane
        const uint64_t ad0 = adesc0 + uint64_t(uint32_t(stage) * a_step), bd0 = bdesc0 + uint64_t(uint32_t(p.b_res ? k : stage) * b_step);
        for (int kk = 0; kk < p.BK / 32; ++kk)
          if (leader) umma_i8(tmem_d, ad0 + uint64_t(2 * kk), bd0 + uint64_t(2 * kk), p.idesc, (k | kk) != 0 ? 1u : 0u);
        if (leader) {
          umma_commit(&ctl->empty[stage]);                       // frees the smem stage when these MMAs retire
          if (k == k_iters - 1) umma_commit(&ctl->acc_full[as]);  // accumulator complete
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (FL && warp >= 4) {
    // ===================== epilogue, flat-linear layers =====================
    // group g owns accumulator stage g and takes this CTA's tiles g, g + 2, ...; tile = rows [128 work, 128 work + 128)
    // (p.wide: one group of eight warps takes every tile, the two warps of a lane quarter alternate 16-column chunks)
    const int wpg = p.wide ? kEpiWarps : kEpiWarps / kAccStages;
    const int gthreads = 32 * wpg;
    const int ew = warp & 3, grp = (warp - 4) / wpg;
    const uint32_t chunk0 = uint32_t(((warp - 4) % wpg) >> 2) * 16u, cstep = uint32_t(4 * wpg);
    const int r = ew * 32 + lane;
    const int eg = threadIdx.x - 128 - grp * gthreads;
    const uint32_t oc = OCT > 0 ? uint32_t(OCT) : uint32_t(p.OC);
    const uint32_t buf_bytes = uint32_t(kBM) * oc;
    uint8_t* grp_buf = stage_buf + size_t(grp) * (p.stage_bytes / uint32_t(kAccStages));
    const uint32_t taddr0 = tmem_base + (uint32_t(ew * 32) << 16);
    const int4* bq0 = reinterpret_cast<const int4*>(s_b2);
    const longlong2* aq0 = reinterpret_cast<const longlong2*>(s_a64);
    const int bar_id = 1 + grp;
    const int it_step = p.wide ? 1 : kAccStages;
    uint32_t use = 0;   // tiles this group has taken: staging buffer parity
    for (int it = grp; blockIdx.x + (long long)it * gridDim.x < total_work; it += it_step, ++use) {
      const int work = blockIdx.x + it * int(gridDim.x);
      const int as = it & (kAccStages - 1);
      const uint32_t taddr = taddr0 + uint32_t(as) * uint32_t(kTmemCols / kAccStages);
      const long long row0 = (long long)work * kBM;
      const int rows_ok = int(min((long long)kBM, (long long)Wd - row0));
      const int8_t* rrow = nullptr;
      if ((MODE & kEpiAdd) && r < rows_ok) rrow = p.resid + (row0 + r) * (long long)oc;
      uint8_t* sbuf = grp_buf + (use & 1u) * buf_bytes;
      uint8_t* srow = sbuf + uint32_t(r) * oc;
      mbar_wait(&ctl->acc_full[as], (uint32_t(it) / uint32_t(kAccStages)) & 1u);
      tc_fence_after();
      if constexpr (OCT > 0) {
        static_assert(OCT % 16 == 0 && OCT <= kFlTabMax, "flat-linear constant table");
        static_assert((MODE & kEpiRelu) && !(MODE & (kEpiAdd | kEpiLut)), "constant-table epilogue: ReLU form only");
#pragma unroll
        for (int ch = 0; ch < OCT / 16; ++ch) {
          if (cstep == 32u && uint32_t(ch & 1) * 16u != chunk0) continue;  // wide mode: the lane quarter's other warp takes this chunk
          uint32_t v[16];
          tmem_ld16(taddr + uint32_t(ch * 16), v);
          tmem_wait_ld();
          uint32_t packed[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            int o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = ch * 16 + q4 * 4 + j;
              o[j] = int((static_cast<long long>(int(v[4 * q4 + j])) * tab->q[col] + tab->a64[col]) >> 32) >> int(tab->rs1[col]);
            }
            if (MODE & kEpiSat) {
              packed[q4] = pack_sat_s8(o[1], o[0], pack_sat_s8(o[3], o[2], 0u));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = max(p.act_min, min(p.act_max, o[j]));
              packed[q4] = (uint32_t(o[0]) & 0xFFu) | ((uint32_t(o[1]) & 0xFFu) << 8) | ((uint32_t(o[2]) & 0xFFu) << 16) | (uint32_t(o[3]) << 24);
            }
          }
          *reinterpret_cast<uint4*>(srow + ch * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
      } else {
      for (uint32_t c0 = chunk0; c0 < oc; c0 += cstep) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        uint4 rres = make_uint4(0u, 0u, 0u, 0u);
        if ((MODE & kEpiAdd) && rrow) rres = *reinterpret_cast<const uint4*>(rrow + c0);
        tmem_wait_ld();
        uint32_t packed[4];
        const int4* kq = (MODE & kEpiRelu) ? reinterpret_cast<const int4*>(reinterpret_cast<const int2*>(s_qtab) + c0) : s_qtab + c0;
        epi_chunk16<MODE>(v, kq, bq0 + (c0 >> 2), aq0 + (c0 >> 1), rres, ctl, p, packed);
        *reinterpret_cast<uint4*>(srow + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      }
      }
      tc_fence_before();   // accumulator fully read: hand the TMEM stage back before the store
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->acc_empty[as]);
      if (eg == 0) tma_store_wait_read();   // this group's previous store has drained the other buffer
      fence_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
      if (eg == 0) {
        bulk_store_1d(p.out + row0 * (long long)oc, sbuf, uint32_t(rows_ok) * oc);
        tma_store_commit();
      }
    }
    if (eg == 0) tma_store_wait_read();  // staging must outlive the last store's reads
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // kEpiWarps warps in acc_stages groups take tiles round-robin: group g owns accumulator stage g, its own staging
    // buffers, named barrier and bulk-store queue, so the fixed per-tile latencies of one tile (accumulator wait, TMEM
    // loads, barrier, store issue) overlap the arithmetic of the other.  The code is written for any multiple of four
    // warps per group (two warps sharing a TMEM lane quarter take alternate 16-column chunks), but measured on the
    // MobileNet expand layers the step is bound by the integer instruction count of the requantisation (about 5100
    // warp instructions per 128 x 96 tile, IPC 0.65): 16 warps in 2 or 4 groups did not move the per-tile time.
    const int wpg = p.wide ? kEpiWarps : kEpiWarps / kAccStages;  // warps per group (p.wide: one group takes every tile)
    const int ew = warp & 3;                 // TMEM lane quarter this warp may touch
    const int grp = (warp - 4) / wpg;        // epilogue group (== accumulator stage with two groups)
    const int wg = (warp - 4) % wpg;         // warp inside the group
    const int half = wg >> 2;                // 8-warp groups: which 16-column chunks of a pass this warp takes (even / odd)
    const int cstep = 4 * wpg;               // columns between this warp's chunks (16 or 32)
    const int gthreads = 32 * wpg;
    const uint32_t acc_stride = uint32_t(kTmemCols / kAccStages);
    const int r = ew * 32 + lane;            // accumulator row == pixel of the tile
    const int eg = threadIdx.x - 128 - grp * gthreads;  // 0..gthreads-1 inside the group
    const int wx = r % p.pw;
    const int wy = (r / p.pw) % p.ph;
    const int wn = r / (p.pw * p.ph);
    // staging: TMA mode = dense rows of p.wo bytes in the tensor map's swizzle; manual mode = padded pitch
    const uint32_t swz_mask = p.lin ? 0u : (p.wo == 128 ? 7u : (p.wo == 64 ? 3u : (p.wo == 32 ? 1u : 0u)));
    const uint32_t row_base = (MODE & kEpiTma) ? uint32_t(r) * uint32_t(p.lin ? p.OC : p.wo) : uint32_t(r) * uint32_t(p.pitch);
    const uint32_t buf_bytes = uint32_t(kBM) * uint32_t(p.lin ? p.OC : p.wo);
    uint8_t* grp_buf = stage_buf + size_t(grp) * (p.stage_bytes / uint32_t(kAccStages));
    long long* g_rowoff = s_rowoff + grp * kBM;   // manual mode: global byte offset of each *run* (see below)
    int* g_runlen = s_runlen + grp * kBM;
    // Manual stores (rows that are not 16-byte multiples: OC = 243, 12, 81).  Pixels that are consecutive in the tile row
    // order and in memory form a run (one patch row, or the whole tile for flat layers) whose bytes are contiguous in
    // global memory.  A run is staged densely, shifted so that staging and global addresses are congruent mod 16, and
    // then leaves as aligned 16-byte stores with a byte head and tail.
    const int run_j = r / p.run_w;
    const uint32_t run_pitch = (uint32_t(p.run_w) * uint32_t(p.BN) + 15u) / 16u * 16u + 32u;
    const int nruns = p.rows / p.run_w;
    const int bar_id = 1 + grp;
    uint32_t pass_count = 0;
    int it = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++it) {
      if (!p.wide && (it & (kAccStages - 1)) != grp) continue;
      const int as = it & (kAccStages - 1);
      const uint32_t use = uint32_t(it) / uint32_t(kAccStages);
      const bool tr = DIAG && p.trace && blockIdx.x == 0 && threadIdx.x == 128 && use < 32;
#define TOD_TR(slot) do { if (DIAG && tr) p.trace[use * 16 + (slot)] = clock64(); } while (0)
      TOD_TR(0);
      const WorkItem w = decode_work(work, p.n_tiles, tiles_x, p.tiles_y);
      const int x = w.tx * p.pw + wx, yy = w.ty * p.ph + wy, n = w.g * p.pn + wn;
      int cls = 0;
      if (p.ncls > 1) {  // border class of this pixel: which filter rows / columns are inside the image
        int ymask = 0, xmask = 0;
        for (int f = 0; f < p.KH; ++f) {
          const int iy = yy * p.stride + f - p.pad_top;
          ymask |= (iy >= 0 && iy < p.IH) ? (1 << f) : 0;
        }
        for (int f = 0; f < p.KW; ++f) {
          const int ix = x * p.stride + f - p.pad_left;
          xmask |= (ix >= 0 && ix < p.IW) ? (1 << f) : 0;
        }
        cls = int(p.ymap[ymask & 7]) * p.ncls_x + int(p.xmap[xmask & 7]);
      }
      const int ocb = w.n_tile * p.BN;
      const int4* b2row = reinterpret_cast<const int4*>(s_b2 + size_t(cls) * p.OCp + ocb);
      const longlong2* a2row = reinterpret_cast<const longlong2*>(s_a64 + size_t(cls) * p.OCp + ocb);
      const int4* qrow = s_qtab + ocb;
      const int ncols_tile = min(p.BN, p.OCm - ocb);  // accumulator columns of this N tile that hold channels (a sibling's chunk included)
      uint32_t soff = 0;  // manual mode: this row's byte offset inside the group's staging buffer
      int8_t* xrow = nullptr;                         // this pixel's bytes in the sibling output (TMA mode only)
      if ((MODE & kEpiTma) && p.x_cols) {
        const bool valid = r < p.rows && x < Wd && yy < p.Hd && (p.flat || n < tiles);
        if (valid) xrow = p.x_out + (p.flat ? 0ll : (long long)n * p.x_out_ts) + ((long long)yy * Wd + x) * p.x_cols;
      }
      if (!(MODE & kEpiTma)) {
        // a run = one patch row; patches as wide as the image (run_w = pw * ph) are contiguous across their rows too
        const int x0 = w.tx * p.pw;
        const bool whole = p.run_w != p.pw;
        const int ry = whole ? w.ty * p.ph : yy;  // first image row of this pixel's run
        const bool valid_row = r < p.rows && ry < p.Hd && (p.flat || n < tiles);
        const int nvalid = !valid_row ? 0 : (whole ? min(p.ph, p.Hd - ry) * p.pw : max(0, min(p.pw, Wd - x0)));
        const long long g_run = (p.flat ? 0ll : (long long)n * p.out_ts) + ((long long)ry * Wd + x0) * p.OC + ocb;
        const uint32_t a = uint32_t(reinterpret_cast<uintptr_t>(p.out) + g_run) & 15u;
        const int pos = whole ? wy * p.pw + wx : wx;  // pixel index inside the run
        soff = uint32_t(run_j) * run_pitch + a + uint32_t(pos) * uint32_t(ncols_tile);
        if (pos == 0 && r < p.rows && half == 0) {
          g_rowoff[run_j] = g_run;
          g_runlen[run_j] = nvalid * ncols_tile;
        }
      }
      const int8_t* rrow = nullptr;                   // this pixel's residual bytes (kEpiAdd)
      if (MODE & kEpiAdd) {
        const bool valid = r < p.rows && x < Wd && yy < p.Hd && (p.flat || n < tiles);
        if (valid) rrow = p.resid + (p.flat ? 0ll : (long long)n * p.resid_ts) + ((long long)yy * Wd + x) * p.OC + ocb;
      }

      TOD_TR(1);
      mbar_wait(&ctl->acc_full[as], use & 1);
      tc_fence_after();
      TOD_TR(2);
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * acc_stride;
      if (DIAG && (p.dbg & 8)) {  // timing experiment: no epilogue work, hand the accumulator straight back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->acc_empty[as]);
        continue;
      }
      for (int pass0 = 0; pass0 < ncols_tile; pass0 += p.sc) {
        const int pass_cols = min(p.sc, ncols_tile - pass0);
        uint8_t* sbuf = grp_buf + ((MODE & kEpiTma) ? (pass_count & 1u) * buf_bytes : 0u);
        ++pass_count;
        for (int c0 = 16 * half; c0 < pass_cols; c0 += cstep) {
          uint32_t v[16];
          tmem_ld16(taddr + pass0 + c0, v);
          uint4 rres = make_uint4(0u, 0u, 0u, 0u);
          if ((MODE & kEpiAdd) && rrow) rres = *reinterpret_cast<const uint4*>(rrow + pass0 + c0);
          tmem_wait_ld();
          TOD_TR(3);
          uint32_t packed[4];
          const int4* bq = b2row + ((pass0 + c0) >> 2);  // chunk bases: every table load below is base + immediate
          const int4* kq = (MODE & kEpiRelu) ? reinterpret_cast<const int4*>(reinterpret_cast<const int2*>(s_qtab) + ocb + pass0 + c0) : qrow + (pass0 + c0);
          const longlong2* aq = a2row + ((pass0 + c0) >> 1);
          const bool sibling = (MODE & kEpiTma) && p.x_cols && pass0 + c0 == p.x_c0;
          epi_chunk16<MODE>(v, kq, bq, aq, rres, ctl, p, packed, sibling ? ctl->lut2 : nullptr);
          if (sibling) {   // the sibling layer's bytes of this pixel leave straight from the registers (x_cols <= 16, 4-byte rows)
            if (xrow) {
#pragma unroll
              for (int wq = 0; wq < 4; ++wq)
                if (4 * wq < p.x_cols) reinterpret_cast<uint32_t*>(xrow)[wq] = packed[wq];
            }
          }
          if (MODE & kEpiTma) {
            uint32_t off = row_base + uint32_t(c0);
            off ^= ((off >> 7) & swz_mask) << 4;
            *reinterpret_cast<uint4*>(sbuf + off) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          } else if (r < p.rows) {  // dense run staging: byte stores, never past this row's last real channel
            uint8_t* dst = sbuf + soff + uint32_t(pass0 + c0);
            const int nb = min(16, ncols_tile - pass0 - c0);
#pragma unroll
            for (int bi = 0; bi < 16; ++bi)
              if (bi < nb) dst[bi] = uint8_t(packed[bi >> 2] >> (8 * (bi & 3)));
          }
        }
        TOD_TR(4);
        if (pass0 + p.sc >= ncols_tile) {  // accumulator fully read: hand the TMEM stage back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->acc_empty[as]);
        }
        if (MODE & kEpiTma) {
          TOD_TR(5);
          if (eg == 0) tma_store_wait_read();   // this group's previous store has drained the *other* buffer
          TOD_TR(6);
          fence_async_smem();
          TOD_TR(7);
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
          TOD_TR(8);
          if (eg == 0 && !(DIAG && (p.dbg & 1))) {
            if (p.lin) {  // the tile's rows are one contiguous block: a single 1-D bulk store instead of 128 box rows
              const long long row0 = (long long)w.tx * kBM;
              bulk_store_1d(p.out + row0 * p.OC, sbuf, uint32_t(min((long long)kBM, (long long)Wd - row0)) * uint32_t(p.OC));
            } else if (p.flat) tma_store_4d(&map_o, sbuf, ocb + pass0, w.tx * kBM, 0, 0);
            else tma_store_4d(&map_o, sbuf, ocb + pass0, w.tx * p.pw, w.ty * p.ph, w.g * p.pn);
            tma_store_commit();
          }
        } else {
          TOD_TR(5);
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
          TOD_TR(6);
          // one warp per run: the per-run address set-up is a dependent chain of shared loads, so the eight warps walk
          // different runs; lanes stride the run's aligned 16-byte chunks and the first lanes carry the head / tail bytes
          for (int j = wg; j < nruns; j += wpg) {
            const int len = g_runlen[j];
            if (len <= 0) continue;
            int8_t* gdst = p.out + g_rowoff[j];
            const uint32_t a = uint32_t(reinterpret_cast<uintptr_t>(gdst)) & 15u;
            const uint8_t* ssrc = sbuf + uint32_t(j) * run_pitch + a;
            const int head = min(len, int((16u - a) & 15u));
            const int body = (len - head) >> 4;
            const int tail = len - head - (body << 4);
            for (int c = lane; c < body; c += 32)
              *reinterpret_cast<uint4*>(gdst + head + 16 * c) = *reinterpret_cast<const uint4*>(ssrc + head + 16 * c);
            if (lane < head) gdst[lane] = int8_t(ssrc[lane]);
            if (lane < tail) gdst[head + 16 * body + lane] = int8_t(ssrc[head + 16 * body + lane]);
          }
          TOD_TR(7);
          asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
          TOD_TR(8);
        }
      }
    }
    if ((MODE & kEpiTma) && eg == 0) tma_store_wait_read();  // staging must outlive the last store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <uint32_t MODE, bool DIAG = false, bool FL = false>
__global__ void __launch_bounds__(kFastThreads, 1)
conv_tc_fast_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_o, const TcParams p, const int tiles) {
  conv_tc_fast_body<MODE, DIAG, FL, 0>(map_a, map_b, map_o, p, tiles, nullptr);
}

template <uint32_t MODE, int OCT>
__global__ void __launch_bounds__(kFastThreads, 1)
conv_tc_flc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_o, const TcParams p, const int tiles, const __grid_constant__ FlTab tab) {
  conv_tc_fast_body<MODE, false, true, OCT>(map_a, map_b, map_o, p, tiles, &tab);
}

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// The large 3x3 layers are limited by how many operand bytes each SM can keep in flight from L2 (three 48 KB stages),
// not by the tensor pipe (ncu: tensor 44 %, L2 31 %).  A CTA pair on one TPC shares the B operand: each CTA stages its
// own 128 pixel rows of A and *half* of the weight rows; one tcgen05.mma.cta_group::2 (M = 256) issued by the even CTA
// feeds both tensor cores, each accumulating its 128 x BN tile in its own TMEM.  Per SM and tile that is 295 + 295 KB
// instead of 295 + 590 KB, and a stage shrinks to 32 KB so five fit.
//   * both CTAs' TMA loads signal the even CTA's `full` barrier (cta_group::2 loads, peer bit cleared);
//   * tcgen05.commit.cta_group::2 multicasts the `empty` / `acc_full` arrivals to both CTAs;
//   * the epilogue warps of both CTAs arrive on the even CTA's `acc_empty` (count 8) through its cluster address.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared::cluster address: "the even CTA's copy"

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(uint16_t(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_even_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}

template <uint32_t MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bh,
                    const __grid_constant__ CUtensorMap map_o, const TcParams p, const int tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + size_t(p.stages) * p.a_stage;
  uint8_t* stage_buf = smem_b + size_t(p.stages) * p.b_stage;
  int4* s_qtab = reinterpret_cast<int4*>(stage_buf + p.stage_bytes);
  int32_t* s_b2 = reinterpret_cast<int32_t*>(s_qtab + p.OCp);
  const long long* s_a64 = reinterpret_cast<const long long*>(s_b2);   // kEpiRelu: 64-bit addends
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(s_b2 + size_t(p.ncls) * p.OCp * ((MODE & kEpiRelu) ? 2 : 1));

  if (!p.late_trig) pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int Wd = p.flat ? tiles * p.HW : p.Wd;
  const int tiles_x = p.flat ? (Wd + kBM - 1) / kBM : p.tiles_x;
  const int groups = p.flat ? 1 : (tiles + p.pn - 1) / p.pn;
  const int m_tiles = groups * p.tiles_y * tiles_x;
  const int pair_items = ((m_tiles + 1) >> 1) * p.n_tiles;   // one item = two consecutive M tiles x one N tile
  const int k_iters = p.taps * p.kchunks;
  const uint32_t half_bn = uint32_t(p.BN) >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctl->full[s], 1);    // used in the even CTA only
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->acc_full[s], 1);
      mbar_init(&ctl->acc_empty[s], p.wide ? 16 : 8);  // the epilogue warps of both CTAs that read the stage (used in the even CTA only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  copy_table16(s_qtab, p.qtab, p.OCp, kTcThreads);
  if (MODE & kEpiRelu) copy_table16(reinterpret_cast<int4*>(s_b2), reinterpret_cast<const int4*>(p.a64tab), p.ncls * p.OCp / 2, kTcThreads);
  else copy_table16(reinterpret_cast<int4*>(s_b2), reinterpret_cast<const int4*>(p.b2tab), p.ncls * p.OCp / 4, kTcThreads);
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anything signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  pdl_wait();

  // this CTA's M tile of pair item `item`
  auto my_tile = [&](int item, int* n_tile, int* tx, int* ty, int* g) -> bool {
    *n_tile = p.n_tiles > 1 ? item % p.n_tiles : 0;
    const int pm = p.n_tiles > 1 ? item / p.n_tiles : item;
    unsigned mt = unsigned(2 * pm) + rank;   // one past the end for the odd CTA of the last pair when the tile count is odd:
    const bool real = mt < unsigned(m_tiles);  // it still runs the whole protocol on a recomputed tile, but stores nothing
    if (!real) mt = 0;
    *tx = int(mt % unsigned(tiles_x));
    mt /= unsigned(tiles_x);
    *ty = int(mt % unsigned(p.tiles_y));
    *g = int(mt / unsigned(p.tiles_y));
    return real;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bh)) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < pair_items; item += npairs) {
        if (p.late_trig && item + npairs >= pair_items) pdl_trigger();
        int n_tile, tx, ty, g;
        my_tile(item, &n_tile, &tx, &ty, &g);
        const int x0 = p.flat ? tx * kBM : tx * p.pw, y0 = ty * p.ph, n0 = g * p.pn;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int fy = tap / p.KW, fx = tap - fy * p.KW;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            const bool skip_b = (p.dbg & 64) != 0, skip_a = (p.dbg & 128) != 0;  // timing experiments: operand traffic of one side only
            if (rank == 0) mbar_expect_tx(&ctl->full[stage], 2u * (p.tx_bytes - (skip_b ? half_bn * uint32_t(p.BK) : 0u) - (skip_a ? uint32_t(p.rows * p.BK) : 0u)));  // both CTAs' A tile + weight half
            if (!skip_a)
            tma_load_4d_pair(smem_a + size_t(stage) * p.a_stage, &map_a, &ctl->full[stage], kc * p.BK, x0 * p.stride + fx - p.pad_left,
                             y0 * p.stride + fy - p.pad_top, n0);
            if (!skip_b)
            tma_load_3d_pair(smem_b + size_t(stage) * p.b_stage, &map_bh, &ctl->full[stage], kc * p.BK, tap, n_tile * p.BN + int(rank * half_bn));
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (even CTA only) =====================
    const bool leader = elect_one();
    const uint64_t adesc0 = make_desc(smem_u32(smem_a), p.sbo, p.layout), bdesc0 = make_desc(smem_u32(smem_b), p.sbo, p.layout);
    const uint32_t a_step = p.a_stage >> 4, b_step = p.b_stage >> 4;  // descriptor address field is in 16-byte units
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = pair; item < pair_items; item += npairs, ++it) {
      const int as = it & 1;
      const uint32_t use = uint32_t(it >> 1);
      mbar_wait(&ctl->acc_empty[as], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kAccStride;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after();
        // the whole warp runs this loop converged with warp-uniform values, so the descriptors live in uniform registers;
        // only the tcgen05 instructions themselves are predicated on the elected lane
        const uint64_t ad0 = adesc0 + uint64_t(uint32_t(stage) * a_step), bd0 = bdesc0 + uint64_t(uint32_t(stage) * b_step);
        for (int kk = 0; kk < p.BK / 32; ++kk)
          if (leader) umma_i8_pair(tmem_d, ad0 + uint64_t(2 * kk), bd0 + uint64_t(2 * kk), p.idesc, (k | kk) != 0 ? 1u : 0u);
        if (leader) {
          umma_commit_pair(&ctl->empty[stage]);                       // frees the smem stage when these MMAs retire
          if (k == k_iters - 1) umma_commit_pair(&ctl->acc_full[as]);  // accumulator complete
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs; same scheme as conv_tc_fast_kernel) =====================
    const int ew = warp & 3;
    const int grp = p.wide ? 0 : (warp - 4) >> 2;
    const int c_first = p.wide ? ((warp - 4) >> 2) * 16 : 0, cstep = p.wide ? 32 : 16;
    const int gthreads = p.wide ? 256 : 128;
    const int r = ew * 32 + lane;
    const int eg = threadIdx.x - 128 - grp * 128;
    const int wx = r % p.pw;
    const int wy = (r / p.pw) % p.ph;
    const uint32_t swz_mask = p.wo == 128 ? 7u : (p.wo == 64 ? 3u : (p.wo == 32 ? 1u : 0u));
    const uint32_t row_base = uint32_t(r) * uint32_t(p.wo);
    const uint32_t buf_bytes = uint32_t(kBM) * uint32_t(p.wo);
    uint8_t* grp_buf = stage_buf + size_t(grp) * (p.stage_bytes >> 1);
    const int bar_id = 1 + grp;
    uint32_t pass_count = 0;
    int it = 0;
    for (int item = pair; item < pair_items; item += npairs, ++it) {
      if (!p.wide && (it & 1) != grp) continue;
      const int as = it & 1;
      const uint32_t use = uint32_t(it >> 1);
      int n_tile, tx, ty, g;
      const bool real_tile = my_tile(item, &n_tile, &tx, &ty, &g);
      const int x = tx * p.pw + wx, yy = ty * p.ph + wy;
      int cls = 0;
      if (p.ncls > 1) {
        int ymask = 0, xmask = 0;
        for (int f = 0; f < p.KH; ++f) {
          const int iy = yy * p.stride + f - p.pad_top;
          ymask |= (iy >= 0 && iy < p.IH) ? (1 << f) : 0;
        }
        for (int f = 0; f < p.KW; ++f) {
          const int ix = x * p.stride + f - p.pad_left;
          xmask |= (ix >= 0 && ix < p.IW) ? (1 << f) : 0;
        }
        cls = int(p.ymap[ymask & 7]) * p.ncls_x + int(p.xmap[xmask & 7]);
      }
      const int ocb = n_tile * p.BN;
      const int4* b2row = reinterpret_cast<const int4*>(s_b2 + size_t(cls) * p.OCp + ocb);
      const longlong2* a2row = reinterpret_cast<const longlong2*>(s_a64 + size_t(cls) * p.OCp + ocb);
      const int4* qrow = s_qtab + ocb;
      const int ncols_tile = min(p.BN, p.OC - ocb);
      mbar_wait(&ctl->acc_full[as], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * kAccStride;
      for (int pass0 = 0; pass0 < ncols_tile; pass0 += p.sc) {
        const int pass_cols = min(p.sc, ncols_tile - pass0);
        uint8_t* sbuf = grp_buf + (pass_count & 1u) * buf_bytes;
        ++pass_count;
        for (int c0 = c_first; c0 < pass_cols; c0 += cstep) {
          uint32_t v[16];
          tmem_ld16(taddr + pass0 + c0, v);
          tmem_wait_ld();
          uint32_t packed[4];
          const int4* bq = b2row + ((pass0 + c0) >> 2);  // chunk bases: every table load below is base + immediate
          const int4* kq = qrow + (pass0 + c0);
          const longlong2* aq = a2row + ((pass0 + c0) >> 1);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            int o[4];
            if (MODE & kEpiRelu) {
              const longlong2 a01 = aq[2 * q4], a23 = aq[2 * q4 + 1];
              const int4* k2 = reinterpret_cast<const int4*>(reinterpret_cast<const int2*>(s_qtab) + ocb + pass0 + c0);   // int2 {q, rs - 1} per channel
              const int4 k01 = k2[2 * q4], k23 = k2[2 * q4 + 1];
              o[0] = int((static_cast<long long>(int(v[4 * q4 + 0])) * k01.x + a01.x) >> 32) >> k01.y;
              o[1] = int((static_cast<long long>(int(v[4 * q4 + 1])) * k01.z + a01.y) >> 32) >> k01.w;
              o[2] = int((static_cast<long long>(int(v[4 * q4 + 2])) * k23.x + a23.x) >> 32) >> k23.y;
              o[3] = int((static_cast<long long>(int(v[4 * q4 + 3])) * k23.z + a23.y) >> 32) >> k23.w;
            } else {
            const int4 b4 = bq[q4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int4 k = kq[4 * q4 + j];
              const int bj = j == 0 ? b4.x : (j == 1 ? b4.y : (j == 2 ? b4.z : b4.w));
              int x2;
              asm("mad.lo.s32 %0, %1, 2, %2;" : "=r"(x2) : "r"(int(v[4 * q4 + j])), "r"(bj));
              const long long addend = static_cast<long long>((static_cast<unsigned long long>(uint32_t(k.w)) << 32) | uint32_t(k.z));
              const int t = int((static_cast<long long>(x2) * k.x + addend) >> 32);
              o[j] = (t + (x2 >> 31)) >> k.y;
            }
            }
            if (MODE & kEpiSat) {
              packed[q4] = pack_sat_s8(o[1], o[0], pack_sat_s8(o[3], o[2], 0u));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = max(p.act_min, min(p.act_max, o[j]));
              packed[q4] = (uint32_t(o[0]) & 0xFFu) | ((uint32_t(o[1]) & 0xFFu) << 8) | ((uint32_t(o[2]) & 0xFFu) << 16) | (uint32_t(o[3]) << 24);
            }
          }
          uint32_t off = row_base + uint32_t(c0);
          off ^= ((off >> 7) & swz_mask) << 4;
          *reinterpret_cast<uint4*>(sbuf + off) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        if (pass0 + p.sc >= ncols_tile) {  // accumulator fully read: tell the even CTA's MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_even_cta(&ctl->acc_empty[as]);
        }
        if (eg == 0) tma_store_wait_read();
        fence_async_smem();
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
        if (eg == 0 && real_tile) {
          if (p.flat) tma_store_4d(&map_o, sbuf, ocb + pass0, tx * kBM, 0, 0);
          else tma_store_4d(&map_o, sbuf, ocb + pass0, tx * p.pw, ty * p.ph, g * p.pn);
          tma_store_commit();
        }
      }
    }
    if (eg == 0) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while its partner can still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode(CUtensorMap* map, const void* base_c, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk,
           const uint32_t* elem_strides = nullptr) {
  EncodeTiledFn fn = encode_fn();
  void* base = const_cast<void*>(base_c);
  if (!fn) return fail(TOD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  static const int promo_env = std::getenv("TOD_TMA_PROMO") ? std::atoi(std::getenv("TOD_TMA_PROMO")) : -1;
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (promo_env >= 0) promo = promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (promo_env == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, cuuint32_t(rank), base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TOD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu/%llu, box %u/%u)", int(r), rank,
                                     (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return TOD_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

struct ConvTc {
  CUtensorMap map_a, map_b, map_o, map_bh;
  int pair = 0;        // conv_tc_pair_kernel (cta_group::2): map_bh loads half of the weight rows per CTA
  TcParams p{};
  int32_t *d_bias_eff = nullptr, *d_mult = nullptr, *d_shift = nullptr, *d_b2 = nullptr;
  long long* d_a64 = nullptr;
  int4* d_qtab = nullptr;
  int32_t* d_add_tab = nullptr;
  size_t smem_bytes = 0;
  int max_tiles = 0;
  int fast = 0;        // conv_tc_fast_kernel is eligible
  uint32_t mode = 0;   // kEpi* bits
  int min_rounds = 1;  // ConvTcArgs::min_rounds
  int fl_oct = 0;      // conv_tc_flc_kernel<mode, fl_oct>: per-channel constants as a kernel parameter (0 = not eligible)
  FlTab fl_tab{};
};

bool conv_tc_supported(const ConvGeom& g, int64_t in_ts, const void* in, const void* w) {
  if (g.stride_h != g.stride_w || (g.stride_h != 1 && g.stride_h != 2) || g.dil_h != 1 || g.dil_w != 1) return false;
  if (g.stride_h == 2 && !(g.KH == 3 && g.KW == 3)) return false;
  if (!((g.KH == 1 && g.KW == 1) || (g.KH == 3 && g.KW == 3))) return false;
  if (g.IC % 16 != 0 || g.IC < 16) return false;
  if (g.OH > g.IH || g.OW > g.IW) return false;  // input coordinates are derived per tap from the output pixel
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (in_ts & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return false;
  if (g.OW > 100000 || g.OH > 100000) return false;
  return true;
}

void conv_tc_destroy(ConvTc* c) {
  if (!c) return;
  cudaFree(c->d_bias_eff);
  cudaFree(c->d_mult);
  cudaFree(c->d_shift);
  cudaFree(c->d_b2);
  cudaFree(c->d_a64);
  cudaFree(c->d_qtab);
  cudaFree(c->d_add_tab);
  delete c;
}

int conv_tc_create(const ConvTcArgs& a, ConvTc** out) {
  *out = nullptr;
  const ConvGeom& g = a.g;
  ConvTc* c = new ConvTc();
  auto bail = [&](int rc) {
    conv_tc_destroy(c);
    return rc;
  };
  TcParams& p = c->p;
  c->max_tiles = a.max_tiles;
  c->min_rounds = a.min_rounds;
  const int out_oc = a.out_oc > 0 ? a.out_oc : g.OC;   // channels of the output tensor (a sibling layer's chunk follows them in the accumulator)
  if (a.out_oc > 0 && (a.out_oc % 16 != 0 || a.x_cols < 4 || a.x_cols % 4 != 0 || a.x_cols > 16 || g.OC != a.out_oc + 16 || g.OC > 256 || !a.x_out ||
                       (reinterpret_cast<uintptr_t>(a.x_out) & 3) || (a.x_out_tile_stride & 3) || a.add))
    return bail(fail(TOD_ERR_INVALID_ARG, "conv_tc: malformed sibling output (out_oc=%d, x_cols=%d, OC=%d)", a.out_oc, a.x_cols, g.OC));
  p.OC = out_oc;
  p.OCm = g.OC;
  p.x_c0 = a.out_oc > 0 ? a.out_oc : 0;
  p.x_cols = a.out_oc > 0 ? a.x_cols : 0;
  p.x_out = a.x_out;
  p.x_out_ts = a.x_out_tile_stride;
  p.x_lut = a.x_lut;
  p.KH = g.KH;
  p.KW = g.KW;
  p.taps = g.KH * g.KW;
  p.pad_top = g.pad_top;
  p.pad_left = g.pad_left;
  p.IH = g.IH;
  p.IW = g.IW;
  p.stride = g.stride_h;
  p.BK = (g.IC % 128 == 0) ? 128 : ((g.IC % 64 == 0) ? 64 : 32);
  p.kchunks = (g.IC + p.BK - 1) / p.BK;
  // flat (dense 1x1) layers whose pixel rows are shorter than a 128-byte TMA row: cp.async gathers A instead, which also
  // frees the K chunk from having to divide IC (the tail is zero-filled chunk by chunk)
  static const int cp_env = std::getenv("TOD_TC_CP") ? std::atoi(std::getenv("TOD_TC_CP")) : -1;
  const bool want_cp = a.fast_epilogue && a.h_mult && a.h_shift && g.KH == 1 && g.KW == 1 && g.stride_h == 1 &&
                       a.in_tile_stride == int64_t(g.IH) * g.IW * g.IC && a.out_tile_stride == int64_t(g.OH) * g.OW * g.OC && g.IH == g.OH &&
                       g.IW == g.OW && (cp_env >= 0 ? cp_env != 0 : g.IC % 128 != 0);
  if (want_cp) {
    p.BK = g.IC <= 32 ? 32 : (g.IC <= 64 ? 64 : 128);
    p.kchunks = (g.IC + p.BK - 1) / p.BK;
  }
  p.a_cp = want_cp ? 1 : 0;
  p.dbg = std::getenv("TOD_TC_DBG") ? std::atoi(std::getenv("TOD_TC_DBG")) : 0;
  p.trace = nullptr;
  p.in = a.in;
  p.IC = g.IC;
  p.n_tiles = (g.OC + 255) / 256;
  // several N tiles: full 256-column tiles, so a 128-column TMA-store pass never overhangs into the next tile's
  // columns (the last tile's overhang is clipped at the tensor edge; its weight rows past OC are TMA zero fill)
  p.BN = p.n_tiles > 1 ? 256 : (g.OC + 15) / 16 * 16;
  p.OCp = p.n_tiles * p.BN;
  p.HW = g.OH * g.OW;
  const bool one = g.KH == 1 && g.KW == 1;
  const bool dense = a.in_tile_stride == int64_t(g.IH) * g.IW * g.IC && a.out_tile_stride == int64_t(g.OH) * g.OW * g.OC && g.IH == g.OH && g.IW == g.OW;
  p.flat = (one && dense && g.stride_h == 1) ? 1 : 0;
  if (p.flat) {
    p.Wd = 0;  // runtime: tiles * HW
    p.Hd = 1;
    p.pw = kBM;
    p.ph = 1;
    p.pn = 1;
    p.tiles_x = 0;
    p.tiles_y = 1;
  } else {
    // pixels of one image as a (Wd x Hd) grid: the real grid for 3x3, one long row for a strided 1x1
    p.Wd = one ? g.OH * g.OW : g.OW;
    p.Hd = one ? 1 : g.OH;
    // choose the patch (pw, ph, pn): maximise useful rows per 128-row tile, then minimise the halo
    double best = -1.0;
    for (int pw = 1; pw <= std::min(p.Wd, kBM); ++pw)
      for (int ph = 1; ph <= std::min(p.Hd, kBM / pw); ++ph) {
        const int pn = std::max(1, std::min(kBM / (pw * ph), a.max_tiles));
        if (pw > 256 || ph > 256 || pn > 256) continue;
        const int txs = (p.Wd + pw - 1) / pw, tys = (p.Hd + ph - 1) / ph;
        const int groups = (a.max_tiles + pn - 1) / pn;
        const double util = double(p.Wd) * p.Hd * a.max_tiles / (double(txs) * tys * groups * kBM);
        const double halo = one ? 1.0 : double(pw + 2) * (ph + 2) / (double(pw) * ph);
        // every (patch row, image) pair is one TMA box row per tap and K chunk and one output run: long rows are cheaper
        // than many short ones (pw = 1 patches measured 4x slower main loops on the 7 x 7 and 14 x 14 maps)
        const double score = util - 0.02 * halo - 0.002 * double(ph * pn);
        if (score > best + 1e-12) {
          best = score;
          p.pw = pw;
          p.ph = ph;
          p.pn = pn;
        }
      }
    p.tiles_x = (p.Wd + p.pw - 1) / p.pw;
    p.tiles_y = (p.Hd + p.ph - 1) / p.ph;
  }
  p.rows = p.pw * p.ph * p.pn;
  p.run_w = p.pw;
  p.sbo = 8u * uint32_t(p.BK);
  p.layout = p.BK == 128 ? 2u : (p.BK == 64 ? 4u : 6u);
  p.a_stage = uint32_t((kBM * p.BK + 1023) / 1024 * 1024);
  p.b_stage = uint32_t((p.BN * p.BK + 1023) / 1024 * 1024);
  p.tx_bytes = uint32_t(p.rows * p.BK + p.BN * p.BK);
  p.vec_store = (out_oc % 16 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && (a.out_tile_stride & 15) == 0) ? 1 : 0;
  // ---- fast epilogue eligibility: every channel requantises with a right shift in [1, 22], a non-negative multiplier,
  // and 2 * (accumulator + bias) stays inside int32
  const int ncls_full = (1 << g.KH) * (1 << g.KW);
  std::vector<int32_t> be(size_t(ncls_full) * p.OCp, 0);
  int64_t max_bias = 0;
  for (int ym = 0; ym < (1 << g.KH); ++ym)
    for (int xm = 0; xm < (1 << g.KW); ++xm) {
      int32_t* row = &be[size_t(ym * (1 << g.KW) + xm) * p.OCp];
      for (int oc = 0; oc < g.OC; ++oc) {
        int32_t s = 0;
        for (int fy = 0; fy < g.KH; ++fy)
          for (int fx = 0; fx < g.KW; ++fx)
            if (((ym >> fy) & 1) && ((xm >> fx) & 1)) s += a.h_wsum[size_t(oc) * p.taps + fy * g.KW + fx];
        row[oc] = int32_t(uint32_t(a.h_bias ? a.h_bias[oc] : 0) - uint32_t(a.in_zp) * uint32_t(s));
        max_bias = std::max<int64_t>(max_bias, std::llabs(int64_t(a.h_bias ? a.h_bias[oc] : 0) - int64_t(a.in_zp) * s));
      }
    }
  c->fast = (a.h_mult && a.h_shift && a.fast_epilogue && g.KH <= 3 && g.KW <= 3) ? 1 : 0;
  if (c->fast) {
    if (int64_t(p.taps) * g.IC * 128 * 128 + max_bias >= (int64_t(1) << 29)) c->fast = 0;
    for (int oc = 0; oc < g.OC && c->fast; ++oc)
      if (a.h_shift[oc] > -1 || a.h_shift[oc] < -22 || a.h_mult[oc] < 0) c->fast = 0;
    if (std::abs(a.rq.out_zp) > 128) c->fast = 0;
  }
  // compact border classes: the (row mask, column mask) pairs that occur for some output pixel
  std::vector<int> ymasks, xmasks;
  std::memset(p.ymap, 0, sizeof(p.ymap));
  std::memset(p.xmap, 0, sizeof(p.xmap));
  if (p.taps == 1) {
    ymasks.push_back(1);
    xmasks.push_back(1);
  } else {
    for (int oy = 0; oy < g.OH; ++oy) {
      int m = 0;
      for (int f = 0; f < g.KH; ++f) m |= (oy * g.stride_h + f - g.pad_top >= 0 && oy * g.stride_h + f - g.pad_top < g.IH) ? (1 << f) : 0;
      if (std::find(ymasks.begin(), ymasks.end(), m) == ymasks.end()) ymasks.push_back(m);
    }
    for (int ox = 0; ox < g.OW; ++ox) {
      int m = 0;
      for (int f = 0; f < g.KW; ++f) m |= (ox * g.stride_w + f - g.pad_left >= 0 && ox * g.stride_w + f - g.pad_left < g.IW) ? (1 << f) : 0;
      if (std::find(xmasks.begin(), xmasks.end(), m) == xmasks.end()) xmasks.push_back(m);
    }
  }
  p.ncls_x = int(xmasks.size());
  p.ncls = int(ymasks.size() * xmasks.size());
  if (c->fast)
    for (size_t i = 0; i < ymasks.size(); ++i) p.ymap[ymasks[i] & 7] = uint8_t(i);
  if (c->fast)
    for (size_t i = 0; i < xmasks.size(); ++i) p.xmap[xmasks[i] & 7] = uint8_t(i);
  // ReLU-type activation (floor at or above the output zero point): the two-instruction requantisation (kEpiRelu)
  static const int relu_env = std::getenv("TOD_TC_RELU") ? std::atoi(std::getenv("TOD_TC_RELU")) : -1;
  const bool relu = c->fast && !a.add && !a.rq.post_lut && !p.x_cols && a.rq.act_min >= a.rq.out_zp && p.vec_store && relu_env != 0;
  const size_t table_bytes = c->fast ? size_t(p.OCp) * 16 + size_t(p.ncls) * p.OCp * (relu ? 8 : 4) : 0;
  if (table_bytes > 40 * 1024) c->fast = 0;
  if (!p.vec_store && p.n_tiles > 1) c->fast = 0;  // run staging needs whole rows (every output channel) in one tile
  if (p.a_cp && !c->fast) {  // the general-epilogue kernel only has the TMA producer: back to a K chunk that TMA can serve
    p.a_cp = 0;
    p.BK = (g.IC % 128 == 0) ? 128 : ((g.IC % 64 == 0) ? 64 : 32);
    p.kchunks = (g.IC + p.BK - 1) / p.BK;
    p.sbo = 8u * uint32_t(p.BK);
    p.layout = p.BK == 128 ? 2u : (p.BK == 64 ? 4u : 6u);
    p.a_stage = uint32_t((kBM * p.BK + 1023) / 1024 * 1024);
    p.b_stage = uint32_t((p.BN * p.BK + 1023) / 1024 * 1024);
    p.tx_bytes = uint32_t(p.rows * p.BK + p.BN * p.BK);
  }
  c->mode = 0;
  if (c->fast) {
    if (a.rq.act_min == -128 && a.rq.act_max == 127) c->mode |= kEpiSat;
    if (a.rq.post_lut) c->mode |= kEpiLut;
    if (p.vec_store) c->mode |= kEpiTma;
    if (a.add) c->mode |= kEpiAdd;
    if (relu) c->mode |= kEpiRelu;
  }
  if (p.x_cols && (!c->fast || !(c->mode & kEpiTma) || p.n_tiles != 1))
    return bail(fail(TOD_ERR_UNSUPPORTED, "conv_tc: a sibling output needs the fast epilogue, a 16-byte-row host output and one N tile"));
  if (a.add && (!c->fast || (c->mode & kEpiLut) || !p.vec_store))
    return bail(fail(TOD_ERR_UNSUPPORTED, "conv_tc: a fused residual ADD needs the fast epilogue, 16-byte rows and no byte map"));
  // CTA pairs (cta_group::2) for the MMA-heavy layers: see conv_tc_pair_kernel
  static const int pair_env = std::getenv("TOD_TC_PAIR") ? std::atoi(std::getenv("TOD_TC_PAIR")) : -1;
  {
    const int groups_max = p.flat ? 1 : (a.max_tiles + p.pn - 1) / p.pn;
    const long long m_tiles_max = p.flat ? ((long long)a.max_tiles * p.HW + kBM - 1) / kBM : (long long)groups_max * p.tiles_y * p.tiles_x;
    c->pair = (c->fast && !p.x_cols && (c->mode & kEpiTma) && !(c->mode & (kEpiLut | kEpiAdd)) && !p.a_cp && p.BN % 32 == 0 && p.BN >= 64 &&
               p.taps * p.kchunks >= 8 && m_tiles_max >= 64 && pair_env != 0) ? 1 : 0;
  }
  if (c->pair) {
    p.b_stage = uint32_t(((p.BN / 2) * p.BK + 1023) / 1024 * 1024);
    p.tx_bytes = uint32_t(p.rows * p.BK + (p.BN / 2) * p.BK);
  }
  static const int lin_env = std::getenv("TOD_TC_LIN") ? std::atoi(std::getenv("TOD_TC_LIN")) : -1;
  static const int bres_env = std::getenv("TOD_TC_BRES") ? std::atoi(std::getenv("TOD_TC_BRES")) : -1;
  p.lin = 0;
  p.b_res = 0;
  {
    const long long m_tiles_flat = p.flat ? ((long long)a.max_tiles * p.HW + kBM - 1) / kBM : 0;
    // weights resident: 1x1, one N tile, several tiles per CTA to amortise over, at most 64 KB
    if (c->fast && !c->pair && p.taps == 1 && p.n_tiles == 1 && size_t(p.kchunks) * p.b_stage <= 64 * 1024 &&
        (p.flat ? m_tiles_flat : (long long)((a.max_tiles + p.pn - 1) / p.pn) * p.tiles_y * p.tiles_x) > sm_count() && bres_env != 0)
      p.b_res = 1;
    // linear staging + one bulk store per tile: flat layers whose rows are 16-byte multiples; 16-byte row-strided shared
    // stores conflict gcd(OC / 16, 8) ways, so rows of 128 / 256 bytes keep the swizzled tensor store
    const int u = g.OC / 16;
    const int conflict = (u % 8 == 0) ? 8 : ((u % 4 == 0) ? 4 : ((u % 2 == 0) ? 2 : 1));
    if (c->fast && !c->pair && !p.x_cols && (c->mode & kEpiTma) && p.flat && p.n_tiles == 1 && p.BN == g.OC && conflict <= 4 && lin_env != 0) p.lin = 1;
  }
  if (p.lin) {
    p.wo = 16;  // unused by the linear path (the tensor map is still encoded, with its smallest box)
    p.sc = p.BN;
    p.pitch = g.OC;
    p.stage_bytes = uint32_t((2 * kAccStages * kBM * g.OC + 1023) / 1024 * 1024);  // epilogue groups x double buffer
  } else if (c->mode & kEpiTma) {
    const int width = c->pair ? std::min(p.BN, 64) : std::min(p.BN, 128);  // pairs: narrower passes buy a fifth operand stage
    p.wo = width <= 16 ? 16 : (width <= 32 ? 32 : (width <= 64 ? 64 : 128));
    p.sc = p.wo;
    p.pitch = p.wo;
    p.stage_bytes = uint32_t(2 * (c->pair ? 2 : kAccStages) * kBM * p.wo);  // epilogue groups x double buffer
  } else {
    p.wo = 0;
    if (c->fast) {
      // conv_tc_fast_kernel, manual stores: the whole tile (every column) is staged as dense runs, one buffer per group
      p.sc = p.BN;
      p.pitch = p.BN;
      p.run_w = (!p.flat && p.pw == p.Wd) ? p.pw * p.ph : p.pw;
      const uint32_t run_pitch = (uint32_t(p.run_w) * uint32_t(p.BN) + 15u) / 16u * 16u + 32u;
      p.stage_bytes = uint32_t((uint32_t(p.rows / p.run_w) * run_pitch + 1023u) / 1024u * 1024u) * uint32_t(kAccStages);
    } else {
      p.sc = std::min(p.BN, 128);
      int pitch16 = p.sc / 16 + 1;
      if (pitch16 % 2 == 0) ++pitch16;  // odd number of 16-byte units per row: conflict-free 16-byte row-strided stores
      p.pitch = pitch16 * 16;
      p.stage_bytes = uint32_t((kBM * p.pitch + 1023) / 1024 * 1024) * 2u;
    }
  }
  const size_t fixed = size_t(p.stage_bytes) + size_t(kBM) * 48 + (c->fast ? table_bytes : 0) + sizeof(SmemCtl) + 1024;
  const size_t budget = 227 * 1024 - fixed;
  if (p.b_res) {
    p.tx_bytes = uint32_t(p.rows * p.BK);
    p.stages = int(std::min<size_t>(kMaxStages, (budget - size_t(p.kchunks) * p.b_stage) / p.a_stage));
  } else {
    p.stages = int(std::min<size_t>(kMaxStages, budget / (p.a_stage + p.b_stage)));
  }
  if (p.stages < 2) return bail(fail(TOD_ERR_UNSUPPORTED, "conv_tc: stage does not fit shared memory"));
  c->smem_bytes = size_t(p.stages) * p.a_stage + size_t(p.b_res ? p.kchunks : p.stages) * p.b_stage + fixed;
  // instruction descriptor: D = s32, A = B = s8, both K-major, N, M = 128
  p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(p.BN >> 3) << 17) | (uint32_t((c->pair ? 2 * kBM : kBM) >> 4) << 24);
  p.out_zp = a.rq.out_zp;
  p.act_min = a.rq.act_min;
  p.act_max = a.rq.act_max;
  p.post_lut = a.rq.post_lut;
  p.out = a.out;
  p.out_ts = a.out_tile_stride;

  // ---- epilogue tables
  std::vector<int32_t> mu(p.OCp, 0), sh(p.OCp, 0);
  cudaError_t ce;
  if (c->fast) {
    std::vector<int32_t> qt(size_t(p.OCp) * 4, 0), b2(size_t(p.ncls) * p.OCp, 0);
    const bool relu_mode = (c->mode & kEpiRelu) != 0;
    for (int oc = 0; oc < g.OC; ++oc) {
      const int rs = -a.h_shift[oc];
      if (relu_mode) {   // int2 {q, rs - 1} per channel, packed (the kernel still copies OCp int4 entries: the tail is padding)
        qt[size_t(oc) * 2 + 0] = a.h_mult[oc];
        qt[size_t(oc) * 2 + 1] = rs - 1;
        continue;
      }
      qt[size_t(oc) * 4 + 0] = a.h_mult[oc];
      qt[size_t(oc) * 4 + 1] = rs;
      qt[size_t(oc) * 4 + 2] = int32_t(0x80000000u);
      qt[size_t(oc) * 4 + 3] = (1 << (rs - 1)) + (oc >= out_oc ? a.x_out_zp : a.rq.out_zp) * (1 << rs);
    }
    std::vector<long long> a64(relu_mode ? size_t(p.ncls) * p.OCp : 0, 0);
    for (size_t yi = 0; yi < ymasks.size(); ++yi)
      for (size_t xi = 0; xi < xmasks.size(); ++xi) {
        const int code = p.taps == 1 ? ((1 << g.KW) + 1) : (ymasks[yi] * (1 << g.KW) + xmasks[xi]);
        const int32_t* src = &be[size_t(code) * p.OCp];
        int32_t* dst = &b2[(yi * xmasks.size() + xi) * p.OCp];
        for (int oc = 0; oc < g.OC; ++oc) dst[oc] = 2 * src[oc];
        if (relu_mode)
          for (int oc = 0; oc < g.OC; ++oc) {
            const int rs = -a.h_shift[oc];
            const long long halfp = (1ll << (rs - 1)) + (long long)a.rq.out_zp * (1ll << rs);
            a64[(yi * xmasks.size() + xi) * p.OCp + oc] = (long long)src[oc] * a.h_mult[oc] + (1ll << 30) + halfp * (1ll << 31);
          }
      }
    static const int flc_env = std::getenv("TOD_TC_FLC") ? std::atoi(std::getenv("TOD_TC_FLC")) : 1;
    if (relu_mode && p.flat && p.lin && p.n_tiles == 1 && p.ncls == 1 && !(c->mode & (kEpiAdd | kEpiLut)) && flc_env &&
        (g.OC == 96 || g.OC == 144 || g.OC == 192)) {
      c->fl_oct = g.OC;
      for (int oc = 0; oc < g.OC; ++oc) {
        c->fl_tab.a64[oc] = a64[oc];
        c->fl_tab.q[oc] = a.h_mult[oc];
        c->fl_tab.rs1[oc] = uint8_t(-a.h_shift[oc] - 1);
      }
    }
    if (relu_mode) {
      if ((ce = cudaMalloc(&c->d_a64, a64.size() * 8)) != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "conv_tc: cudaMalloc: %s", cudaGetErrorString(ce)));
      cudaMemcpy(c->d_a64, a64.data(), a64.size() * 8, cudaMemcpyHostToDevice);
      p.a64tab = c->d_a64;
    }
    if ((ce = cudaMalloc(&c->d_qtab, qt.size() * 4)) != cudaSuccess || (ce = cudaMalloc(&c->d_b2, b2.size() * 4)) != cudaSuccess)
      return bail(fail(TOD_ERR_CUDA, "conv_tc: cudaMalloc: %s", cudaGetErrorString(ce)));
    cudaMemcpy(c->d_qtab, qt.data(), qt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_b2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice);
    p.qtab = c->d_qtab;
    p.b2tab = c->d_b2;
  }
  if (a.add) {
    if ((ce = cudaMalloc(&c->d_add_tab, 512 * 4)) != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "conv_tc: cudaMalloc: %s", cudaGetErrorString(ce)));
    cudaMemcpy(c->d_add_tab, a.add->tab, 512 * 4, cudaMemcpyHostToDevice);
    p.add_tab = c->d_add_tab;
    p.resid = a.add->resid;
    p.resid_ts = a.add->resid_tile_stride;
    p.add_mult = a.add->mult_out;
    p.add_shift = a.add->shift_out;
    p.add_zp = a.add->zp_out;
    p.add_min = a.add->act_min;
    p.add_max = a.add->act_max;
  }
  if ((ce = cudaMalloc(&c->d_bias_eff, be.size() * 4)) != cudaSuccess || (ce = cudaMalloc(&c->d_mult, mu.size() * 4)) != cudaSuccess ||
      (ce = cudaMalloc(&c->d_shift, sh.size() * 4)) != cudaSuccess)
    return bail(fail(TOD_ERR_CUDA, "conv_tc: cudaMalloc: %s", cudaGetErrorString(ce)));
  cudaMemcpy(c->d_bias_eff, be.data(), be.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(c->d_mult, 0, mu.size() * 4);
  cudaMemset(c->d_shift, 0, sh.size() * 4);
  cudaMemcpy(c->d_mult, a.rq.mult, size_t(g.OC) * 4, cudaMemcpyDeviceToDevice);
  cudaMemcpy(c->d_shift, a.rq.shift, size_t(g.OC) * 4, cudaMemcpyDeviceToDevice);
  p.bias_eff = c->d_bias_eff;
  p.mult = c->d_mult;
  p.shift = c->d_shift;

  // ---- tensor maps
  int rc;
  if (p.flat) {
    const uint64_t dims[4] = {uint64_t(g.IC), uint64_t(a.max_tiles) * p.HW, 1, 1};
    const uint64_t str[3] = {uint64_t(g.IC), uint64_t(a.max_tiles) * p.HW * g.IC, uint64_t(a.max_tiles) * p.HW * g.IC};
    const uint32_t box[4] = {uint32_t(p.BK), uint32_t(kBM), 1, 1};
    rc = encode(&c->map_a, const_cast<int8_t*>(a.in), 4, dims, str, box, p.BK);
  } else if (one) {
    const uint64_t dims[4] = {uint64_t(g.IC), uint64_t(p.Wd), 1, uint64_t(a.max_tiles)};
    const uint64_t str[3] = {uint64_t(g.IC), uint64_t(a.in_tile_stride), uint64_t(a.in_tile_stride)};
    const uint32_t box[4] = {uint32_t(p.BK), uint32_t(p.pw), 1, uint32_t(p.pn)};
    rc = encode(&c->map_a, const_cast<int8_t*>(a.in), 4, dims, str, box, p.BK);
  } else {
    const uint64_t dims[4] = {uint64_t(g.IC), uint64_t(g.IW), uint64_t(g.IH), uint64_t(a.max_tiles)};
    const uint64_t str[3] = {uint64_t(g.IC), uint64_t(g.IW) * g.IC, uint64_t(a.in_tile_stride)};
    const uint32_t sd = uint32_t(p.stride);
    const uint32_t box[4] = {uint32_t(p.BK), uint32_t(p.pw) * sd, uint32_t(p.ph) * sd, uint32_t(p.pn)};
    const uint32_t es[4] = {1, sd, sd, 1};
    if (box[1] > 256 || box[2] > 256) return bail(fail(TOD_ERR_UNSUPPORTED, "conv_tc: strided patch exceeds the TMA box limit"));
    rc = encode(&c->map_a, const_cast<int8_t*>(a.in), 4, dims, str, box, p.BK, es);
  }
  if (rc < 0) return bail(rc);
  {
    const uint64_t dims[3] = {uint64_t(g.IC), uint64_t(p.taps), uint64_t(g.OC)};
    const uint64_t str[2] = {uint64_t(g.IC), uint64_t(p.taps) * g.IC};
    const uint32_t box[3] = {uint32_t(p.BK), 1, uint32_t(p.BN)};
    rc = encode(&c->map_b, const_cast<int8_t*>(a.w), 3, dims, str, box, p.BK);
  }
  if (rc < 0) return bail(rc);
  std::memset(&c->map_bh, 0, sizeof(c->map_bh));
  if (c->pair) {
    const uint64_t dims[3] = {uint64_t(g.IC), uint64_t(p.taps), uint64_t(g.OC)};
    const uint64_t str[2] = {uint64_t(g.IC), uint64_t(p.taps) * g.IC};
    const uint32_t box[3] = {uint32_t(p.BK), 1, uint32_t(p.BN / 2)};
    rc = encode(&c->map_bh, a.w, 3, dims, str, box, p.BK);
    if (rc < 0) return bail(rc);
  }
  std::memset(&c->map_o, 0, sizeof(c->map_o));
  if (c->mode & kEpiTma) {
    // the output seen through the same patch tiling as the A operand; the box is [wo channels x patch]
    if (p.flat) {
      const uint64_t dims[4] = {uint64_t(out_oc), uint64_t(a.max_tiles) * p.HW, 1, 1};
      const uint64_t str[3] = {uint64_t(out_oc), uint64_t(a.max_tiles) * p.HW * out_oc, uint64_t(a.max_tiles) * p.HW * out_oc};
      const uint32_t box[4] = {uint32_t(p.wo), uint32_t(kBM), 1, 1};
      rc = encode(&c->map_o, a.out, 4, dims, str, box, p.wo);
    } else if (one) {
      const uint64_t dims[4] = {uint64_t(out_oc), uint64_t(p.Wd), 1, uint64_t(a.max_tiles)};
      const uint64_t str[3] = {uint64_t(out_oc), uint64_t(a.out_tile_stride), uint64_t(a.out_tile_stride)};
      const uint32_t box[4] = {uint32_t(p.wo), uint32_t(p.pw), 1, uint32_t(p.pn)};
      rc = encode(&c->map_o, a.out, 4, dims, str, box, p.wo);
    } else {
      const uint64_t dims[4] = {uint64_t(out_oc), uint64_t(g.OW), uint64_t(g.OH), uint64_t(a.max_tiles)};
      const uint64_t str[3] = {uint64_t(out_oc), uint64_t(g.OW) * out_oc, uint64_t(a.out_tile_stride)};
      const uint32_t box[4] = {uint32_t(p.wo), uint32_t(p.pw), uint32_t(p.ph), uint32_t(p.pn)};
      rc = encode(&c->map_o, a.out, 4, dims, str, box, p.wo);
    }
    if (rc < 0) return bail(rc);
  }
  static bool attr_set = false;
  if (!attr_set) {
    const void* kernels[] = {(const void*)conv_tc_kernel,
                             (const void*)conv_tc_fast_kernel<0>, (const void*)conv_tc_fast_kernel<1>, (const void*)conv_tc_fast_kernel<2>,
                             (const void*)conv_tc_fast_kernel<3>, (const void*)conv_tc_fast_kernel<4>, (const void*)conv_tc_fast_kernel<5>,
                             (const void*)conv_tc_fast_kernel<6>, (const void*)conv_tc_fast_kernel<7>,
                             (const void*)conv_tc_fast_kernel<12>, (const void*)conv_tc_fast_kernel<13>,
                             (const void*)conv_tc_fast_kernel<20>, (const void*)conv_tc_fast_kernel<21>,
                             (const void*)conv_tc_fast_kernel<4, false, true>, (const void*)conv_tc_fast_kernel<5, false, true>,
                             (const void*)conv_tc_fast_kernel<12, false, true>, (const void*)conv_tc_fast_kernel<13, false, true>,
                             (const void*)conv_tc_fast_kernel<20, false, true>, (const void*)conv_tc_fast_kernel<21, false, true>,
                             (const void*)conv_tc_flc_kernel<20, 96>, (const void*)conv_tc_flc_kernel<20, 144>, (const void*)conv_tc_flc_kernel<20, 192>,
                             (const void*)conv_tc_flc_kernel<21, 96>, (const void*)conv_tc_flc_kernel<21, 144>, (const void*)conv_tc_flc_kernel<21, 192>,
                             (const void*)conv_tc_pair_kernel<20>, (const void*)conv_tc_pair_kernel<21>,
                             (const void*)conv_tc_fast_kernel<3, true>, (const void*)conv_tc_fast_kernel<5, true>,
                             (const void*)conv_tc_pair_kernel<4>, (const void*)conv_tc_pair_kernel<5>};
    for (const void* k : kernels) {
      ce = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (ce != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(ce)));
    }
    attr_set = true;
  }
  if (std::getenv("TOD_TC_PLAN")) {  // one line per planned layer, for tools/layer_table.py
    const int groups_max = p.flat ? 1 : (a.max_tiles + p.pn - 1) / p.pn;
    const long long m_tiles = p.flat ? ((long long)a.max_tiles * p.HW + kBM - 1) / kBM : (long long)groups_max * p.tiles_y * p.tiles_x;
    std::fprintf(stderr, "TC_PLAN in=%dx%dx%d out=%dx%dx%d k=%d s=%d flat=%d fast=%d mode=%u pair=%d cp=%d BK=%d kchunks=%d BN=%d n_tiles=%d m_tiles=%lld patch=%dx%dx%d stages=%d wo=%d ncls=%d bres=%d lin=%d\n",
                 g.IH, g.IW, g.IC, g.OH, g.OW, g.OC, g.KH, g.stride_h, p.flat, c->fast, c->mode, c->pair, p.a_cp, p.BK, p.kchunks, p.BN, p.n_tiles,
                 m_tiles, p.pw, p.ph, p.pn, p.stages, p.wo, p.ncls, p.b_res, p.lin);
  }
  *out = c;
  return TOD_OK;
}

int conv_tc_launch(ConvTc* c, int tiles, cudaStream_t s) {
  if (tiles < 1 || tiles > c->max_tiles) return fail(TOD_ERR_CAPACITY, "conv_tc_launch: tiles=%d outside [1,%d]", tiles, c->max_tiles);
  TcParams p = c->p;
  const int tiles_x = p.flat ? (tiles * p.HW + kBM - 1) / kBM : p.tiles_x;
  const int groups = p.flat ? 1 : (tiles + p.pn - 1) / p.pn;
  const long long work = (long long)groups * p.tiles_y * tiles_x * p.n_tiles;
  // balanced grid: the launch takes ceil(work / SMs) rounds whatever the grid, so it only asks for the CTAs that keep every
  // round full (392 tiles: 131 CTAs x 3 instead of 148 CTAs of which 96 do 3 and 52 do 2) and leaves the other SMs to the
  // batches in flight beside this one (TOD_TC_BALANCE=0: one CTA per SM)
  static const int balance_env = std::getenv("TOD_TC_BALANCE") ? std::atoi(std::getenv("TOD_TC_BALANCE")) : 1;
  // Tiles per CTA at least (ConvTcArgs::min_rounds <- tod_yolact_options::batches_in_flight; TOD_TC_MINROUNDS overrides): a launch
  // whose work fits one round still gives every CTA this many tiles.  A CTA's fixed cost (prologue, pipeline fill, last tile's
  // epilogue, drain: ~5 us of an SM that no other convolution CTA can share) is then paid by half as many CTAs; with batches in
  // flight the freed SMs run the other batches' kernels (3 in flight: 1.015 -> 0.987 ms per step; one handle alone: 1.153 ->
  // 1.169 ms, which is why a handle that has the GPU to itself keeps one tile per CTA)
  static const int min_rounds_env = std::getenv("TOD_TC_MINROUNDS") ? std::max(1, std::atoi(std::getenv("TOD_TC_MINROUNDS"))) : 0;
  const int min_rounds = min_rounds_env ? min_rounds_env : std::max(1, c->min_rounds);
  auto balanced = [min_rounds](long long items, int max_ctas) {
    const long long full = std::min<long long>(items, max_ctas);
    if (!balance_env || (items <= max_ctas && min_rounds <= 1)) return int(full);
    const long long rounds = std::max<long long>((items + max_ctas - 1) / max_ctas, min_rounds);
    return int((items + rounds - 1) / rounds);
  };
  const int grid = balanced(work, sm_count());
  // eight-warp epilogue when a CTA has only a few tiles (TOD_TC_WIDE: 0 = never, 1 = always, N > 1 = up to N tiles per CTA)
  static const int late_env = std::getenv("TOD_TC_LATE") ? std::atoi(std::getenv("TOD_TC_LATE")) : 0;  // measured: +0.4 % per pipelined step, -0.6 % single stream
  p.late_trig = late_env ? 1 : 0;
  static const int wide_env = std::getenv("TOD_TC_WIDE") ? std::atoi(std::getenv("TOD_TC_WIDE")) : -1;
  const int wide_max = wide_env < 0 ? 5 : (wide_env == 1 ? (1 << 30) : wide_env);
  if (c->pair) {
    const long long m_tiles = (long long)groups * p.tiles_y * tiles_x;
    const long long pair_items = ((m_tiles + 1) / 2) * p.n_tiles;
    const int pairs = balanced(pair_items, sm_count() / 2);
    p.wide = (pair_items + pairs - 1) / pairs <= wide_max ? 1 : 0;
    if (c->mode == 21)
      TOD_CUDA(launch_k(conv_tc_pair_kernel<21>, dim3(2 * pairs), dim3(kTcThreads), c->smem_bytes, s, c->map_a, c->map_bh, c->map_o, p, tiles));
    else if (c->mode == 20)
      TOD_CUDA(launch_k(conv_tc_pair_kernel<20>, dim3(2 * pairs), dim3(kTcThreads), c->smem_bytes, s, c->map_a, c->map_bh, c->map_o, p, tiles));
    else if (c->mode & kEpiSat)
      TOD_CUDA(launch_k(conv_tc_pair_kernel<5>, dim3(2 * pairs), dim3(kTcThreads), c->smem_bytes, s, c->map_a, c->map_bh, c->map_o, p, tiles));
    else
      TOD_CUDA(launch_k(conv_tc_pair_kernel<4>, dim3(2 * pairs), dim3(kTcThreads), c->smem_bytes, s, c->map_a, c->map_bh, c->map_o, p, tiles));
    return TOD_OK;
  }
  p.wide = (c->fast && (work + grid - 1) / grid <= wide_max) ? 1 : 0;
  if (!c->fast) {
    TOD_CUDA(launch_k(conv_tc_kernel, dim3(grid), dim3(kTcThreads), c->smem_bytes, s, c->map_a, c->map_b, p, tiles));
  } else {
    if ((p.trace || p.dbg) && (c->mode == 3 || c->mode == 5)) {  // diagnostics builds of the two most common epilogues
      if (c->mode == 3) TOD_CUDA(launch_k(conv_tc_fast_kernel<3, true>, dim3(grid), dim3(kFastThreads), c->smem_bytes, s, c->map_a, c->map_b, c->map_o, p, tiles));
      else TOD_CUDA(launch_k(conv_tc_fast_kernel<5, true>, dim3(grid), dim3(kFastThreads), c->smem_bytes, s, c->map_a, c->map_b, c->map_o, p, tiles));
      TOD_CUDA(cudaGetLastError());
      return TOD_OK;
    }
    static const int fl_env = std::getenv("TOD_TC_FL") ? std::atoi(std::getenv("TOD_TC_FL")) : -1;
    if (c->fl_oct && fl_env != 0) {
#define TOD_TC_FLC(M, O) if (c->mode == M && c->fl_oct == O) { TOD_CUDA(launch_k(conv_tc_flc_kernel<M, O>, dim3(grid), dim3(kFastThreads), c->smem_bytes, s, c->map_a, c->map_b, c->map_o, p, tiles, c->fl_tab)); TOD_CUDA(cudaGetLastError()); return TOD_OK; }
      TOD_TC_FLC(20, 96) TOD_TC_FLC(20, 144) TOD_TC_FLC(20, 192) TOD_TC_FLC(21, 96) TOD_TC_FLC(21, 144) TOD_TC_FLC(21, 192)
#undef TOD_TC_FLC
    }
    if (p.flat && p.lin && p.n_tiles == 1 && fl_env != 0) {
      switch (c->mode) {
#define TOD_TC_FL(M) case M: TOD_CUDA(launch_k(conv_tc_fast_kernel<M, false, true>, dim3(grid), dim3(kFastThreads), c->smem_bytes, s, c->map_a, c->map_b, c->map_o, p, tiles)); TOD_CUDA(cudaGetLastError()); return TOD_OK;
        TOD_TC_FL(4) TOD_TC_FL(5) TOD_TC_FL(12) TOD_TC_FL(13) TOD_TC_FL(20) TOD_TC_FL(21)
#undef TOD_TC_FL
        default: break;
      }
    }
    switch (c->mode) {
#define TOD_TC_CASE(M) case M: TOD_CUDA(launch_k(conv_tc_fast_kernel<M>, dim3(grid), dim3(kFastThreads), c->smem_bytes, s, c->map_a, c->map_b, c->map_o, p, tiles)); break;
      TOD_TC_CASE(0) TOD_TC_CASE(1) TOD_TC_CASE(2) TOD_TC_CASE(3) TOD_TC_CASE(4) TOD_TC_CASE(5) TOD_TC_CASE(6) TOD_TC_CASE(7)
      TOD_TC_CASE(12) TOD_TC_CASE(13) TOD_TC_CASE(20) TOD_TC_CASE(21)
      default: return fail(TOD_ERR_UNSUPPORTED, "conv_tc: no kernel for epilogue mode %u", c->mode);
#undef TOD_TC_CASE
    }
  }
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

}  // namespace tod

// ------------------------------------------------------------------ tensor-pipe peak for kind::i8
// MEASURED_PEAKS.json records HBM and bf16 peaks only; the conv roofline needs the int8 one.  One CTA per SM keeps a
// 128 x 128 B A tile and a 256 x 128 B B tile resident in shared memory and one thread issues `n_mma`
// tcgen05.mma.kind::i8 (M = 128, N = 256, K = 32) back to back into one TMEM accumulator: no loads, no epilogue.
namespace tod {
namespace {
__global__ void __launch_bounds__(128, 1) i8_mma_peak_kernel(int n_mma, uint32_t idesc, int group) {
  extern __shared__ uint8_t peak_raw[];
  uint8_t* smem = peak_raw + ((1024u - (smem_u32(peak_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, ready, freed;   // `ready` / `freed` emulate the conv kernel's per-stage full / empty barriers
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * uint32_t(i & 3);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&ready, 1);
    mbar_init(&freed, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {  // warp 0, converged: uniform descriptors, the elected lane issues (same scheme as the conv kernels)
    const bool leader = elect_one();
    const uint64_t ad0 = make_desc(smem_u32(smem), 1024u, 2u), bd0 = make_desc(smem_u32(smem + 128 * 128), 1024u, 2u);
    uint32_t rphase = 0;
    if (group <= 0) {
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (leader) umma_i8(tmem, ad0 + uint64_t(2 * kk), bd0 + uint64_t(2 * kk), idesc, (i | kk) ? 1u : 0u);
      }
    } else {
      for (int i = 0; i < n_mma; i += group) {   // the conv kernel's per-stage protocol: wait `full`, fence, MMAs, commit `empty`
        if (leader) mbar_arrive(&ready);
        mbar_wait(&ready, rphase);
        rphase ^= 1;
        tc_fence_after();
        for (int kk = 0; kk < group; ++kk)
          if (leader) umma_i8(tmem, ad0 + uint64_t(2 * (kk & 3)), bd0 + uint64_t(2 * (kk & 3)), idesc, (i | kk) ? 1u : 0u);
        if (leader) umma_commit(&freed);
      }
    }
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}
}  // namespace
}  // namespace tod

extern "C" int tod_i8_mma_peak(int device, int n_mma, int iters, double* tops) {
  using namespace tod;
  if (n_mma < 1 || iters < 1 || !tops) return fail(TOD_ERR_INVALID_ARG, "tod_i8_mma_peak: bad argument");
  TOD_TRY(select_device(device));
  const size_t smem = (128 + 256) * 128 + 1024;
  TOD_CUDA(cudaFuncSetAttribute(i8_mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(256 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  const int sms = sm_count();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int group = std::getenv("TOD_PEAK_GROUP") ? std::atoi(std::getenv("TOD_PEAK_GROUP")) : 0;
  i8_mma_peak_kernel<<<sms, 128, smem>>>(n_mma, idesc, group);
  TOD_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    cudaEventRecord(e0);
    i8_mma_peak_kernel<<<sms, 128, smem>>>(n_mma, idesc, group);
    cudaEventRecord(e1);
    TOD_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tops = 2.0 * 128 * 256 * 32 * double(n_mma) * sms / (double(best) * 1e-3) / 1e12;
  return TOD_OK;
}

// ------------------------------------------------------------------ self tests / micro-benchmark (C ABI)
using namespace tod;

namespace {
// runs one convolution through both the tcgen05 path and the CUDA-core direct kernel on random data
int conv_selftest_impl(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, int flags, float* ms_tc, float* ms_direct, long long* mismatches) {
  TOD_TRY(select_device(device));
  ConvGeom g{};
  const int sd = (flags & 32) ? 2 : 1;
  g.IH = H;
  g.IW = W;
  g.OH = (H + sd - 1) / sd;
  g.OW = (W + sd - 1) / sd;
  g.IC = IC;
  g.OC = OC;
  g.KH = g.KW = K;
  g.stride_h = g.stride_w = sd;
  g.dil_h = g.dil_w = 1;
  g.pad_top = std::max(0, (g.OH - 1) * sd + K - H) / 2;   // TFLite SAME
  g.pad_left = std::max(0, (g.OW - 1) * sd + K - W) / 2;
  const size_t in_elems = size_t(tiles) * H * W * IC, out_elems = size_t(tiles) * g.OH * g.OW * OC, w_elems = size_t(OC) * K * K * IC;
  std::mt19937 rng(1234);
  std::vector<int8_t> h_in(in_elems), h_w(w_elems);
  for (auto& v : h_in) v = int8_t(int(rng() % 255) - 127);
  for (auto& v : h_w) v = int8_t(int(rng() % 255) - 127);
  std::vector<int32_t> h_bias(OC), h_mult(OC), h_shift(OC), h_wsum(size_t(OC) * K * K);
  const double eff = 1.0 / (40.0 * std::sqrt(double(K * K * IC)) * 127.0 / 30.0);
  for (int oc = 0; oc < OC; ++oc) {
    h_bias[oc] = int32_t(rng() % 20001) - 10000;
    int sh;
    quantize_multiplier(((flags & 8) ? 0.75 : eff) * (0.5 + (rng() % 1000) / 1000.0), &h_mult[oc], &sh);
    h_shift[oc] = sh;
    for (int t = 0; t < K * K; ++t) {
      int32_t s = 0;
      for (int ic = 0; ic < IC; ++ic) s += h_w[(size_t(oc) * K * K + t) * IC + ic];
      h_wsum[size_t(oc) * K * K + t] = s;
    }
  }
  int8_t *d_in, *d_w, *d_o1, *d_o2;
  int32_t *d_bias, *d_mult, *d_shift, *d_wsum;
  TOD_CUDA(cudaMalloc(&d_in, in_elems));
  TOD_CUDA(cudaMalloc(&d_w, w_elems));
  TOD_CUDA(cudaMalloc(&d_o1, out_elems));
  TOD_CUDA(cudaMalloc(&d_o2, out_elems));
  TOD_CUDA(cudaMalloc(&d_bias, OC * 4));
  TOD_CUDA(cudaMalloc(&d_mult, OC * 4));
  TOD_CUDA(cudaMalloc(&d_shift, OC * 4));
  TOD_CUDA(cudaMalloc(&d_wsum, h_wsum.size() * 4));
  TOD_CUDA(cudaMemcpy(d_in, h_in.data(), in_elems, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(d_w, h_w.data(), w_elems, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(d_bias, h_bias.data(), OC * 4, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(d_mult, h_mult.data(), OC * 4, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(d_shift, h_shift.data(), OC * 4, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemcpy(d_wsum, h_wsum.data(), h_wsum.size() * 4, cudaMemcpyHostToDevice));
  TOD_CUDA(cudaMemset(d_o1, 0x55, out_elems));
  TOD_CUDA(cudaMemset(d_o2, 0x33, out_elems));
  const int32_t in_zp = -3;
  uint8_t* d_lut = nullptr;
  if (flags & 2) {  // a fused byte map (any permutation-free table will do)
    uint8_t h_lut[256];
    for (int i = 0; i < 256; ++i) h_lut[i] = uint8_t((i * 7 + 13) ^ (i >> 3));
    TOD_CUDA(cudaMalloc(&d_lut, 256));
    TOD_CUDA(cudaMemcpy(d_lut, h_lut, 256, cudaMemcpyHostToDevice));
  }
  Requant rq{d_mult, d_shift, (flags & 1) ? -128 : 5, (flags & 1) ? -128 : -128, (flags & 1) ? 90 : 127, d_lut};
  if (flags & 16) rq.act_min = -77;
  const int64_t in_ts = int64_t(H) * W * IC, out_ts = int64_t(g.OH) * g.OW * OC;
  int rc = TOD_OK;
  ConvTc* plan = nullptr;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (!conv_tc_supported(g, in_ts, d_in, d_w)) rc = fail(TOD_ERR_UNSUPPORTED, "selftest shape is not eligible for the tcgen05 path");
  if (rc == TOD_OK) {
    ConvTcArgs a{g, d_in, in_ts, d_w, in_zp, rq, d_o1, out_ts, tiles, h_bias.data(), h_wsum.data()};
    a.h_mult = h_mult.data();
    a.h_shift = h_shift.data();
    a.fast_epilogue = (flags & 4) ? 0 : 1;
    rc = conv_tc_create(a, &plan);
  }
  long long* d_trace = nullptr;
  if (rc == TOD_OK && std::getenv("TOD_TC_TRACE")) {
    cudaMalloc(&d_trace, 32 * 16 * 8);
    cudaMemset(d_trace, 0, 32 * 16 * 8);
    plan->p.trace = d_trace;
  }
  if (rc == TOD_OK) rc = conv_tc_launch(plan, tiles, nullptr);
  if (rc == TOD_OK && d_trace) {
    cudaDeviceSynchronize();
    rc = conv_tc_launch(plan, tiles, nullptr);  // warm second run is the one reported
    cudaDeviceSynchronize();
    long long h[32 * 16];
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    const int trace_rows = std::min(32, std::max(1, std::atoi(std::getenv("TOD_TC_TRACE"))));
    for (int t = 0; t < trace_rows; ++t) {
      std::printf("tile %2d t0=%8lld :", t, h[t * 16] - h[0]);
      for (int k = 1; k <= 8; ++k) std::printf(" %6lld", h[t * 16 + k] - h[t * 16 + k - 1]);
      std::printf("\n");
    }
    plan->p.trace = nullptr;
    cudaFree(d_trace);
  }
  if (rc == TOD_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(TOD_ERR_CUDA, "tcgen05 conv kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc == TOD_OK) {
    cudaEventRecord(e0);
    for (int i = 0; i < iters && rc == TOD_OK; ++i) rc = conv_tc_launch(plan, tiles, nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms_tc) *ms_tc = t / std::max(iters, 1);
  }
  if (rc == TOD_OK) {
    launch_conv_direct(d_in, in_ts, d_w, d_bias, d_wsum, in_zp, g, rq, d_o2, out_ts, tiles, nullptr);
    cudaEventRecord(e0);
    launch_conv_direct(d_in, in_ts, d_w, d_bias, d_wsum, in_zp, g, rq, d_o2, out_ts, tiles, nullptr);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = fail(TOD_ERR_CUDA, "direct conv kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms_direct) *ms_direct = t;
  }
  if (rc == TOD_OK) {
    std::vector<int8_t> o1(out_elems), o2(out_elems);
    cudaMemcpy(o1.data(), d_o1, out_elems, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2.data(), d_o2, out_elems, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (size_t i = 0; i < out_elems; ++i) bad += o1[i] != o2[i];
    if (mismatches) *mismatches = bad;
  }
  conv_tc_destroy(plan);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_in); cudaFree(d_w); cudaFree(d_o1); cudaFree(d_o2);
  cudaFree(d_bias); cudaFree(d_mult); cudaFree(d_shift); cudaFree(d_wsum);
  cudaFree(d_lut);
  return rc;
}
}  // namespace

extern "C" {

int tod_conv_selftest(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, float* ms_tc, float* ms_direct, long long* mismatches) {
  if (tiles < 1 || H < 1 || W < 1 || IC < 1 || OC < 1 || (K != 1 && K != 3)) return fail(TOD_ERR_INVALID_ARG, "tod_conv_selftest: bad shape");
  return conv_selftest_impl(device, tiles, H, W, IC, OC, K, iters, 0, ms_tc, ms_direct, mismatches);
}

// flags: 1 = ReLU6-style clamp (not the full int8 range), 2 = fused byte map, 4 = force the general epilogue,
//        8 = multipliers >= 0.5 (shift 0: the planner must fall back to the general epilogue), 16 = act_min above -128
int tod_conv_selftest_ex(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, int flags, float* ms_tc, float* ms_direct,
                         long long* mismatches) {
  if (tiles < 1 || H < 1 || W < 1 || IC < 1 || OC < 1 || (K != 1 && K != 3)) return fail(TOD_ERR_INVALID_ARG, "tod_conv_selftest_ex: bad shape");
  return conv_selftest_impl(device, tiles, H, W, IC, OC, K, iters, flags, ms_tc, ms_direct, mismatches);
}

// plain GEMM C[M,N] = A[M,K] * B[N,K]^T as a 1x1 convolution over M "pixels": the int8 roofline denominator
int tod_i8_gemm_selftest(int device, int M, int N, int K, int iters, float* ms_per_iter, double* max_abs_err) {
  if (M < 1 || N < 1 || K < 16) return fail(TOD_ERR_INVALID_ARG, "tod_i8_gemm_selftest: bad shape");
  long long bad = 0;
  float ms_direct = 0;
  const int rc = conv_selftest_impl(device, 1, 1, M, K, N, 1, iters, 0, ms_per_iter, &ms_direct, &bad);
  if (max_abs_err) *max_abs_err = double(bad);
  return rc;
}

}  // extern "C"
