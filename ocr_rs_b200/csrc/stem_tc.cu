// Detector stem on the tensor cores (BF16 mode):
//   conv 7x7 s2 p3 (1 -> 64, model.rs:68,109) + batch-norm + ReLU (:69,110-111) + max_pool2d
//   3x3 s2 p1 (:112), fused; u8 or f32 grey levels in, NHWC bf16 [B][H/4][W/4][64] out.
//
// The 1-channel 7x7 convolution is an implicit GEMM with K = 7 rows x 8 columns (the 8th
// column carries a zero weight) = 56, padded to 64: for conv pixel (cy, cx) the 16-byte K
// chunk j is the 8 consecutive input pixels (2cy + j, 2cx .. 2cx + 7), so the A tile is a pure
// 16-byte gather from a bf16 copy of the input patch — built by the CTA in shared memory in
// the 128B-swizzled K-major layout UMMA expects (no im2col matrix ever touches HBM).
// One CTA unit = 8 x 14 pooled pixels <- 17 x 29 conv pixels (493 GEMM rows = 4 MMA tiles of
// 128, accumulators in TMEM) <- 39 x 64 input patch.  After the MMAs the A region is reused
// as the bf16 conv tile from which the 3x3/s2 max-pool is taken.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

#ifndef OCRB_STEM_PW
#define OCRB_STEM_PW 14
#endif
constexpr int SK_PH = 8, SK_PW = OCRB_STEM_PW;             // pooled tile: 14 wide = 4 M tiles, 2 CTAs per SM (7 wide / 4 CTAs measured 17% slower: the kernel is bound by shared-memory wavefronts, not latency)
constexpr int SK_CH = 2 * SK_PH + 1, SK_CW = 2 * SK_PW + 1;  // conv tile 17 x 29
constexpr int SK_ROWS = SK_CH * SK_CW;                     // 493 valid GEMM rows
constexpr int SK_MT = (SK_ROWS + 127) / 128;               // M tiles of 128 rows
#ifndef OCRB_STEM_OCC
#define OCRB_STEM_OCC (SK_MT <= 2 ? 4 : 2)
#endif
constexpr int SK_OCC = OCRB_STEM_OCC;                      // CTAs per SM (shared memory and 512 TMEM columns)
constexpr int SK_IH = 2 * SK_CH + 5, SK_IW = 2 * SK_CW + 6;  // input patch 39 x 64 (63 used + the zero-weight column)
constexpr int SK_THREADS = 256;
constexpr int SK_A_BYTES = SK_MT * 128 * 128;              // M tiles x 128 rows x 128 B
constexpr int SK_OFF_B = SK_A_BYTES;                       // 64 x 128 B
constexpr int SK_OFF_PATCH = SK_OFF_B + 64 * 128;          // bf16 [39][SK_IW]
constexpr int SK_RAW_WORDS = (SK_IW + 6) / 4;              // raw bytes per patch row: 3 lead-in + SK_IW (+ pad to a word)
constexpr int SK_RAW_BYTES = SK_IH * SK_RAW_WORDS * 4;     // one raw u8 patch (cp.async prefetch target)
constexpr int SK_OFF_RAW = SK_OFF_PATCH + SK_IH * SK_IW * 2;
constexpr int SK_OFF_MISC = SK_OFF_RAW + 2 * SK_RAW_BYTES;
constexpr int SK_SMEM = SK_OFF_MISC + 64 + 1024;           // barrier + TMEM slot + alignment slack

__device__ __forceinline__ void cp_async_4_zfill(uint32_t dst, const void *src, bool ok) {
  const int n = ok ? 4 : 0;  // src-size 0: the 4 destination bytes are zero-filled (= conv / image padding)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// folded bn1 scale / shift, passed by value: after unrolling they are constant-bank operands
struct StemConsts { float scale[64], shift[64]; };

template <class TIn>
__global__ void __launch_bounds__(SK_THREADS, SK_OCC)
stem_tc_kernel(const TIn *__restrict__ in, int B, int H, int W, const float *__restrict__ w /*[49][64]*/,
               const __grid_constant__ StemConsts sc, __nv_bfloat16 *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;
  uint8_t *sB = smem + SK_OFF_B;
  __nv_bfloat16 *s_patch = reinterpret_cast<__nv_bfloat16 *>(smem + SK_OFF_PATCH);
  uint8_t *s_raw = smem + SK_OFF_RAW;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SK_OFF_MISC);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
  const uint32_t sA_u32 = smem_u32(sA), patch_u32 = smem_u32(s_patch), raw_u32 = smem_u32(s_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Hc = H / 2, Wc = W / 2, Hp = H / 4, Wp = W / 4;
  const int tiles_x = (Wp + SK_PW - 1) / SK_PW, tiles_y = (Hp + SK_PH - 1) / SK_PH;
  const int units = tiles_x * tiles_y * B;

  // ---- one-time setup: zero A (no NaN bit patterns under the zero weights), weights -> B, TMEM
  for (int i = tid; i < SK_A_BYTES / 16; i += SK_THREADS) reinterpret_cast<uint4 *>(sA)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 64 * 8; i += SK_THREADS) {  // (co, chunk j): k = 8j + s <-> tap (r = j, s)
    const int co = i >> 3, j = i & 7;
    uint32_t pk[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float a = 0.f, b = 0.f;
      if (j < 7) {
        a = w[(j * 7 + 2 * h) * 64 + co];
        if (2 * h + 1 < 7) b = w[(j * 7 + 2 * h + 1) * 64 + co];
      }
      pk[h] = pack_bf16(a, b);
    }
    *reinterpret_cast<uint4 *>(sB + co * 128 + ((j ^ (co & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, SK_MT * 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = make_idesc(64);
  uint32_t parity = 0;

  // u8 input: the raw patch of the NEXT unit is prefetched with cp.async while this one is
  // computed.  Patch column j is raw byte 3 + j: ix0 = 4*px0 - 5 is always 3 (mod 4), so the
  // 4-byte-aligned row start is ix0 - 3.
  auto prefetch_raw = [&](int unit, int buf) {
    const int b = unit / (tiles_x * tiles_y), t = unit - b * (tiles_x * tiles_y);
    const int py0 = (t / tiles_x) * SK_PH, px0 = (t % tiles_x) * SK_PW;
    const int iy0 = 4 * py0 - 5, xa = 4 * px0 - 8;
    const uint8_t *img = reinterpret_cast<const uint8_t *>(in) + (int64_t)b * H * W;
    for (int i = tid; i < SK_IH * SK_RAW_WORDS; i += SK_THREADS) {
      const int row = i / SK_RAW_WORDS, w = i - row * SK_RAW_WORDS;
      const int yy = iy0 + row, xx = xa + 4 * w;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      cp_async_4_zfill(raw_u32 + buf * SK_RAW_BYTES + i * 4, ok ? img + (int64_t)yy * W + xx : img, ok);
    }
    cp_async_commit();
  };
  int buf = 0;
  if (sizeof(TIn) == 1 && (int)blockIdx.x < units) prefetch_raw(blockIdx.x, 0);

  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int b = unit / (tiles_x * tiles_y), t = unit - b * (tiles_x * tiles_y);
    const int py0 = (t / tiles_x) * SK_PH, px0 = (t % tiles_x) * SK_PW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // conv-grid origin (pool pad 1)
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;  // input origin (conv pad 3)
    // ---- (a) input patch -> bf16 (u8 grey levels are exact in bf16)
    if (sizeof(TIn) == 1) {
      cp_async_wait_all();
      __syncthreads();
      if (unit + (int)gridDim.x < units) prefetch_raw(unit + gridDim.x, buf ^ 1);
      const uint32_t *raw = reinterpret_cast<const uint32_t *>(s_raw + buf * SK_RAW_BYTES);
      for (int i = tid; i < SK_IH * SK_RAW_WORDS; i += SK_THREADS) {
        const int row = i / SK_RAW_WORDS, w = i - row * SK_RAW_WORDS;
        const uint32_t word = raw[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = 4 * w + k - 3;
          if (col >= 0 && col < SK_IW) s_patch[row * SK_IW + col] = __float2bfloat16((float)((word >> (8 * k)) & 0xffu));
        }
      }
      buf ^= 1;
    } else {
      const TIn *img = in + (int64_t)b * H * W;
      for (int i = tid; i < SK_IH * SK_IW; i += SK_THREADS) {
        const int yy = iy0 + i / SK_IW, xx = ix0 + i % SK_IW;
        const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (float)img[(int64_t)yy * W + xx] : 0.0f;
        s_patch[i] = __float2bfloat16(v);
      }
    }
    __syncthreads();
    // ---- (b) A rows: chunk j of row (cy, cx) = patch[2cy + j][2cx .. 2cx + 7]
    for (int m = tid; m < SK_ROWS; m += SK_THREADS) {
      const int cy = m / SK_CW, cx = m - cy * SK_CW;
      const uint32_t row = sA_u32 + m * 128;
      const uint32_t src = patch_u32 + ((2 * cy) * SK_IW + 2 * cx) * 2;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        uint32_t w0, w1, w2, w3;
        asm volatile("ld.shared.u32 %0, [%4];\n\tld.shared.u32 %1, [%4+4];\n\tld.shared.u32 %2, [%4+8];\n\tld.shared.u32 %3, [%4+12];"
                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(src + j * SK_IW * 2));
        sts_16(row + ((j ^ (m & 7)) << 4), make_uint4(w0, w1, w2, w3));
      }
      sts_16(row + ((7 ^ (m & 7)) << 4), make_uint4(0, 0, 0, 0));
    }
    fence_proxy_async();
    __syncthreads();
    // ---- (c) 4 M tiles x 4 K steps of UMMA 128 x 64 x 16
    if (tid == 0) {
      tc_fence_after();
      const uint64_t bdesc = make_smem_desc(sB);
#pragma unroll
      for (int mt = 0; mt < SK_MT; ++mt) {
        const uint64_t adesc = make_smem_desc(sA + mt * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + mt * 64, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, parity, err, 21);
    parity ^= 1;
    tc_fence_after();
    // ---- (d) epilogue: BN + ReLU -> bf16 conv tile over the A region (same swizzle)
    {
      const int q = warp & 3;
#pragma unroll 1
      for (int mt = warp >> 2; mt < SK_MT; mt += 2) {
        const int m = mt * 128 + q * 32 + lane;
        const int cy = m / SK_CW, cx = m - cy * SK_CW;
        const bool in_grid = (cy0 + cy) >= 0 && (cy0 + cy) < Hc && (cx0 + cx) >= 0 && (cx0 + cx) < Wc;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 64);
        const uint32_t row = sA_u32 + m * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tmem_ld32(taddr + half * 32, v);
          if (m < SK_ROWS) {
            if (in_grid) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const int c = half * 32 + j4 * 8;
                uint32_t pk[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  const float2 y = ffma2(make_float2(v[j4 * 8 + 2 * h], v[j4 * 8 + 2 * h + 1]), make_float2(sc.scale[c + 2 * h], sc.scale[c + 2 * h + 1]),
                                         make_float2(sc.shift[c + 2 * h], sc.shift[c + 2 * h + 1]));
                  pk[h] = pack_bf16(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f));
                }
                sts_16(row + (((half * 4 + j4) ^ (m & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
              }
            } else {
              // outside the conv grid = max-pool padding (-inf): a large negative finite bf16 (0xFF7F)
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) sts_16(row + (((half * 4 + j4) ^ (m & 7)) << 4), make_uint4(0xFF7FFF7Fu, 0xFF7FFF7Fu, 0xFF7FFF7Fu, 0xFF7FFF7Fu));
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- (e) 3x3 / s2 max-pool from the conv tile -> global.  One thread = one pooled column
    // (px, 8-channel chunk) over 4 pooled rows: the horizontal 3-max of each of the 9 conv rows
    // is computed once and shared by the two pooled rows that overlap it (27 loads / 4 outputs).
    if (tid < SK_PW * 8 * 2) {
      const int chunk = tid & 7, px = (tid >> 3) % SK_PW, hrow = tid / (8 * SK_PW);  // hrow: pooled rows 4*hrow .. 4*hrow+3
      if (px0 + px < Wp) {
        __nv_bfloat162 prev[4];  // horizontal max of the conv row shared with the previous pooled row
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const int cyr = 8 * hrow + k;  // conv row within the tile
          __nv_bfloat162 hm[4];
#pragma unroll
          for (int s3 = 0; s3 < 3; ++s3) {
            const int m = cyr * SK_CW + 2 * px + s3;
            const uint4 u = lds_16(sA_u32 + m * 128 + ((chunk ^ (m & 7)) << 4));
            const __nv_bfloat162 *hv = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
            for (int h = 0; h < 4; ++h) hm[h] = s3 == 0 ? hv[h] : __hmax2(hm[h], hv[h]);
          }
          if (k == 0) {
#pragma unroll
            for (int h = 0; h < 4; ++h) prev[h] = hm[h];
          } else if (k & 1) {  // middle conv row of pooled row (k-1)/2: start accumulating
#pragma unroll
            for (int h = 0; h < 4; ++h) prev[h] = __hmax2(prev[h], hm[h]);
          } else {             // last conv row of pooled row k/2 - 1 = first conv row of pooled row k/2
            const int py = 4 * hrow + k / 2 - 1;
            __nv_bfloat162 o2[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) { o2[h] = __hmax2(prev[h], hm[h]); prev[h] = hm[h]; }
            if (py0 + py < Hp) {
              uint4 o;
              o.x = *reinterpret_cast<uint32_t *>(&o2[0]); o.y = *reinterpret_cast<uint32_t *>(&o2[1]);
              o.z = *reinterpret_cast<uint32_t *>(&o2[2]); o.w = *reinterpret_cast<uint32_t *>(&o2[3]);
              st_16(out + (((int64_t)b * Hp + py0 + py) * Wp + px0 + px) * 64 + chunk * 8, o);
            }
          }
        }
      }
    }
    __syncthreads();  // the conv tile / patch are rewritten by the next unit
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, SK_MT * 64);
  }
}

int launch_stem_tc(ocrb_ctx *ctx, const void *in, int is_u8, int B, int H, int W, const float *w, const float *scale_host,
                   const float *shift_host, __nv_bfloat16 *out, int *err) {
  StemConsts sc;
  memcpy(sc.scale, scale_host, sizeof(sc.scale));
  memcpy(sc.shift, shift_host, sizeof(sc.shift));
  OCRB_TRY(ensure_dyn_smem(ctx, stem_tc_kernel<uint8_t>, SK_SMEM));
  OCRB_TRY(ensure_dyn_smem(ctx, stem_tc_kernel<float>, SK_SMEM));
  const int Hp = H / 4, Wp = W / 4;
  const int64_t units = cdiv(Wp, SK_PW) * cdiv(Hp, SK_PH) * B;
  const int grid = (int)(units < SK_OCC * ctx->sm_count ? units : SK_OCC * ctx->sm_count);
  if (is_u8)
    stem_tc_kernel<uint8_t><<<grid, SK_THREADS, SK_SMEM, ctx->stream>>>((const uint8_t *)in, B, H, W, w, sc, out, err);
  else
    stem_tc_kernel<float><<<grid, SK_THREADS, SK_SMEM, ctx->stream>>>((const float *)in, B, H, W, w, sc, out, err);
  return check_launch(ctx, "tc:stem");
}

}  // namespace ocrb
