// tcgen05 int8 implicit-GEMM convolution for sm_100a.  See conv_tc.h.
//
// Replaces the CONV_2D kernels TFLite runs inside interpreter.invoke() (/root/reference/src/yolact.rs:163)
// for stride-1 1x1 / 3x3 layers with Cin % 16 == 0 — >97 % of the graph's multiply-accumulates.
//
// Kernel anatomy (persistent, one CTA per SM, 8 warps):
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
#include <cstring>
#include <random>
#include <vector>

#include "common.h"
#include "fixedpoint.cuh"

namespace tod {
namespace {

constexpr int kBM = 128;
constexpr int kTcThreads = 384;  // 4 control warps + 8 epilogue warps
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
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
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint8_t lut[256];
};

__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p, const int tiles) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + size_t(p.stages) * p.a_stage;
  uint8_t* stage_buf = smem_b + size_t(p.stages) * p.b_stage;                      // [128][pitch] requantised bytes
  long long* s_rowoff = reinterpret_cast<long long*>(stage_buf + size_t(kBM) * p.pitch);  // [128] global offset of each row
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(s_rowoff + kBM);

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
            tma_load_4d(smem_a + size_t(stage) * p.a_stage, &map_a, &ctl->full[stage], kc * p.BK, x0 + fx - p.pad_left, y0 + fy - p.pad_top, n0);
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
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem_a + size_t(stage) * p.a_stage);
          const uint32_t b_addr = smem_u32(smem_b + size_t(stage) * p.b_stage);
          for (int kk = 0; kk < p.BK / 32; ++kk) {
            const uint64_t ad = make_desc(a_addr + kk * 32, p.sbo, p.layout);
            const uint64_t bd = make_desc(b_addr + kk * 32, p.sbo, p.layout);
            umma_i8(tmem_d, ad, bd, p.idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&ctl->empty[stage]);                     // frees the smem stage when these MMAs retire
          if (k == k_iters - 1) umma_commit(&ctl->acc_full[as]);  // accumulator complete
        }
        __syncwarp();
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
          const int iy = py + f - p.pad_top;
          ymask |= (iy >= 0 && iy < p.IH) ? (1 << f) : 0;
        }
        for (int f = 0; f < p.KW; ++f) {
          const int ix = px + f - p.pad_left;
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

int encode(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int bk) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(TOD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = bk == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, cuuint32_t(rank), base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  CUtensorMap map_a, map_b;
  TcParams p{};
  int32_t *d_bias_eff = nullptr, *d_mult = nullptr, *d_shift = nullptr;
  size_t smem_bytes = 0;
  int max_tiles = 0;
};

bool conv_tc_supported(const ConvGeom& g, int64_t in_ts, const void* in, const void* w) {
  if (g.stride_h != 1 || g.stride_w != 1 || g.dil_h != 1 || g.dil_w != 1) return false;
  if (!((g.KH == 1 && g.KW == 1) || (g.KH == 3 && g.KW == 3))) return false;
  if (g.IC % 16 != 0 || g.IC < 16) return false;
  if (g.OH != g.IH || g.OW != g.IW) {
    // stride-1 VALID (no padding) shrinks the output; the patch tiling assumes output == input extent only for
    // addressing the *output*, and input coordinates are derived per tap, so this is fine as long as pads >= 0
    if (g.OH > g.IH || g.OW > g.IW) return false;
  }
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (in_ts & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return false;
  if (g.OW > 100000 || g.OH > 100000) return false;
  return true;
}

void conv_tc_destroy(ConvTc* c) {
  if (!c) return;
  cudaFree(c->d_bias_eff);
  cudaFree(c->d_mult);
  cudaFree(c->d_shift);
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
  p.OC = g.OC;
  p.KH = g.KH;
  p.KW = g.KW;
  p.taps = g.KH * g.KW;
  p.pad_top = g.pad_top;
  p.pad_left = g.pad_left;
  p.IH = g.IH;
  p.IW = g.IW;
  p.BK = (g.IC % 128 == 0) ? 128 : ((g.IC % 64 == 0) ? 64 : 32);
  p.kchunks = (g.IC + p.BK - 1) / p.BK;
  p.n_tiles = (g.OC + 255) / 256;
  p.BN = ((g.OC + p.n_tiles - 1) / p.n_tiles + 15) / 16 * 16;
  p.OCp = p.n_tiles * p.BN;
  p.HW = g.OH * g.OW;
  const bool one = g.KH == 1 && g.KW == 1;
  const bool dense = a.in_tile_stride == int64_t(g.IH) * g.IW * g.IC && a.out_tile_stride == int64_t(g.OH) * g.OW * g.OC && g.IH == g.OH && g.IW == g.OW;
  p.flat = (one && dense) ? 1 : 0;
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
        const double score = util - 0.02 * halo;
        if (score > best) {
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
  p.sbo = 8u * uint32_t(p.BK);
  p.layout = p.BK == 128 ? 2u : (p.BK == 64 ? 4u : 6u);
  p.a_stage = uint32_t((kBM * p.BK + 1023) / 1024 * 1024);
  p.b_stage = uint32_t((p.BN * p.BK + 1023) / 1024 * 1024);
  p.tx_bytes = uint32_t(p.rows * p.BK + p.BN * p.BK);
  p.sc = std::min(p.BN, 128);
  int pitch16 = p.sc / 16 + 1;
  if (pitch16 % 2 == 0) ++pitch16;  // odd number of 16-byte units per row: conflict-free 16-byte row-strided stores
  p.pitch = pitch16 * 16;
  const size_t fixed = size_t(kBM) * p.pitch + size_t(kBM) * 8 + sizeof(SmemCtl) + 1024;
  const size_t budget = 227 * 1024 - fixed;
  p.stages = int(std::min<size_t>(kMaxStages, budget / (p.a_stage + p.b_stage)));
  if (p.stages < 2) return bail(fail(TOD_ERR_UNSUPPORTED, "conv_tc: stage does not fit shared memory"));
  c->smem_bytes = size_t(p.stages) * (p.a_stage + p.b_stage) + fixed;
  // instruction descriptor: D = s32, A = B = s8, both K-major, N, M = 128
  p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(p.BN >> 3) << 17) | (uint32_t(kBM >> 4) << 24);
  p.out_zp = a.rq.out_zp;
  p.act_min = a.rq.act_min;
  p.act_max = a.rq.act_max;
  p.post_lut = a.rq.post_lut;
  p.out = a.out;
  p.out_ts = a.out_tile_stride;
  p.vec_store = (g.OC % 16 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && (a.out_tile_stride & 15) == 0) ? 1 : 0;

  // ---- epilogue tables
  const int ncls = (1 << g.KH) * (1 << g.KW);
  std::vector<int32_t> be(size_t(ncls) * p.OCp, 0), mu(p.OCp, 0), sh(p.OCp, 0);
  for (int ym = 0; ym < (1 << g.KH); ++ym)
    for (int xm = 0; xm < (1 << g.KW); ++xm) {
      int32_t* row = &be[size_t(ym * (1 << g.KW) + xm) * p.OCp];
      for (int oc = 0; oc < g.OC; ++oc) {
        int32_t s = 0;
        for (int fy = 0; fy < g.KH; ++fy)
          for (int fx = 0; fx < g.KW; ++fx)
            if (((ym >> fy) & 1) && ((xm >> fx) & 1)) s += a.h_wsum[size_t(oc) * p.taps + fy * g.KW + fx];
        row[oc] = int32_t(uint32_t(a.h_bias ? a.h_bias[oc] : 0) - uint32_t(a.in_zp) * uint32_t(s));
      }
    }
  cudaError_t ce;
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
    const uint32_t box[4] = {uint32_t(p.BK), uint32_t(p.pw), uint32_t(p.ph), uint32_t(p.pn)};
    rc = encode(&c->map_a, const_cast<int8_t*>(a.in), 4, dims, str, box, p.BK);
  }
  if (rc < 0) return bail(rc);
  {
    const uint64_t dims[3] = {uint64_t(g.IC), uint64_t(p.taps), uint64_t(g.OC)};
    const uint64_t str[2] = {uint64_t(g.IC), uint64_t(p.taps) * g.IC};
    const uint32_t box[3] = {uint32_t(p.BK), 1, uint32_t(p.BN)};
    rc = encode(&c->map_b, const_cast<int8_t*>(a.w), 3, dims, str, box, p.BK);
  }
  if (rc < 0) return bail(rc);
  static bool attr_set = false;
  if (!attr_set) {
    ce = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (ce != cudaSuccess) return bail(fail(TOD_ERR_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(ce)));
    attr_set = true;
  }
  *out = c;
  return TOD_OK;
}

int conv_tc_launch(ConvTc* c, int tiles, cudaStream_t s) {
  if (tiles < 1 || tiles > c->max_tiles) return fail(TOD_ERR_CAPACITY, "conv_tc_launch: tiles=%d outside [1,%d]", tiles, c->max_tiles);
  const TcParams& p = c->p;
  const int tiles_x = p.flat ? (tiles * p.HW + kBM - 1) / kBM : p.tiles_x;
  const int groups = p.flat ? 1 : (tiles + p.pn - 1) / p.pn;
  const long long work = (long long)groups * p.tiles_y * tiles_x * p.n_tiles;
  const int grid = int(std::min<long long>(work, sm_count()));
  conv_tc_kernel<<<grid, kTcThreads, c->smem_bytes, s>>>(c->map_a, c->map_b, p, tiles);
  TOD_CUDA(cudaGetLastError());
  return TOD_OK;
}

}  // namespace tod

// ------------------------------------------------------------------ self tests / micro-benchmark (C ABI)
using namespace tod;

namespace {
// runs one convolution through both the tcgen05 path and the CUDA-core direct kernel on random data
int conv_selftest_impl(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, float* ms_tc, float* ms_direct, long long* mismatches) {
  TOD_TRY(select_device(device));
  ConvGeom g{};
  g.IH = g.OH = H;
  g.IW = g.OW = W;
  g.IC = IC;
  g.OC = OC;
  g.KH = g.KW = K;
  g.stride_h = g.stride_w = g.dil_h = g.dil_w = 1;
  g.pad_top = g.pad_left = K / 2;
  const size_t in_elems = size_t(tiles) * H * W * IC, out_elems = size_t(tiles) * H * W * OC, w_elems = size_t(OC) * K * K * IC;
  std::mt19937 rng(1234);
  std::vector<int8_t> h_in(in_elems), h_w(w_elems);
  for (auto& v : h_in) v = int8_t(int(rng() % 255) - 127);
  for (auto& v : h_w) v = int8_t(int(rng() % 255) - 127);
  std::vector<int32_t> h_bias(OC), h_mult(OC), h_shift(OC), h_wsum(size_t(OC) * K * K);
  const double eff = 1.0 / (40.0 * std::sqrt(double(K * K * IC)) * 127.0 / 30.0);
  for (int oc = 0; oc < OC; ++oc) {
    h_bias[oc] = int32_t(rng() % 20001) - 10000;
    int sh;
    quantize_multiplier(eff * (0.5 + (rng() % 1000) / 1000.0), &h_mult[oc], &sh);
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
  Requant rq{d_mult, d_shift, 5, -128, 127, nullptr};
  const int64_t in_ts = int64_t(H) * W * IC, out_ts = int64_t(H) * W * OC;
  int rc = TOD_OK;
  ConvTc* plan = nullptr;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  if (!conv_tc_supported(g, in_ts, d_in, d_w)) rc = fail(TOD_ERR_UNSUPPORTED, "selftest shape is not eligible for the tcgen05 path");
  if (rc == TOD_OK) {
    ConvTcArgs a{g, d_in, in_ts, d_w, in_zp, rq, d_o1, out_ts, tiles, h_bias.data(), h_wsum.data()};
    rc = conv_tc_create(a, &plan);
  }
  if (rc == TOD_OK) rc = conv_tc_launch(plan, tiles, nullptr);
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
  return rc;
}
}  // namespace

extern "C" {

int tod_conv_selftest(int device, int tiles, int H, int W, int IC, int OC, int K, int iters, float* ms_tc, float* ms_direct, long long* mismatches) {
  if (tiles < 1 || H < 1 || W < 1 || IC < 1 || OC < 1 || (K != 1 && K != 3)) return fail(TOD_ERR_INVALID_ARG, "tod_conv_selftest: bad shape");
  return conv_selftest_impl(device, tiles, H, W, IC, OC, K, iters, ms_tc, ms_direct, mismatches);
}

// plain GEMM C[M,N] = A[M,K] * B[N,K]^T as a 1x1 convolution over M "pixels": the int8 roofline denominator
int tod_i8_gemm_selftest(int device, int M, int N, int K, int iters, float* ms_per_iter, double* max_abs_err) {
  if (M < 1 || N < 1 || K < 16) return fail(TOD_ERR_INVALID_ARG, "tod_i8_gemm_selftest: bad shape");
  long long bad = 0;
  float ms_direct = 0;
  const int rc = conv_selftest_impl(device, 1, 1, M, K, N, 1, iters, ms_per_iter, &ms_direct, &bad);
  if (max_abs_err) *max_abs_err = double(bad);
  return rc;
}

}  // extern "C"
