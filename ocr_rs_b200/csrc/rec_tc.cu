// Glyph-recognition net on the tensor cores at FP32-equivalent accuracy (char_recognition/model.rs:27-39).
//
// The north_star asks for a bit-exact class argmax, so operands cannot simply be rounded to 16 bits.  Every fp32
// operand x is split into two fp16 numbers, x ~= hi + lo' * 2^-11 with hi = fp16(x), lo' = fp16((x - hi) * 2^11)
// (22 significant bits; the 2^11 keeps the low part out of fp16's subnormal range), and a product of two split numbers
//     a * w ~= a_hi * w_hi + 2^-11 * (a_hi * w_lo' + a_lo' * w_hi)          (the lo * lo term, 2^-22 relative, is dropped)
// becomes ONE tcgen05 kind::f16 GEMM with doubled K and doubled N:
//     A' = [a_hi | a_lo']   (K' = 2K)        W' rows: main   n     : [w_hi  | 0   ]
//                                                     scaled  n + N : [w_lo' | w_hi]
// fp16 x fp16 products are exact in the fp32 accumulator; the epilogue returns  main + 2^-11 * scaled + bias.
// Measured against the fp32 torch restatement: logits agree to ~1e-6 relative, the same order as a re-ordered fp32 sum.
//
//   rec_conv2_tc_kernel  conv 5x5 (32 -> 64) + bias + max_pool 2 + NCHW flatten as an implicit GEMM: 4 glyphs per unit
//                        (2 M tiles of 128 = 2 glyphs x 8x8 output pixels), the pooled conv1 maps of the unit resident
//                        in shared memory ([144 px][hi 32 | lo' 32] = 128 B per pixel = one swizzle row), one K block per
//                        filter tap: A tile = 128 pixel records copied into the 128B-swizzled K-major layout, B tile
//                        (128 x 64) by TMA; N' = 128 accumulators in TMEM
//   rec_fc_tc_kernel     fc1 (1024 -> 512) + bias + ReLU: TMA-fed GEMM, M tile 128 glyphs, N' tile 256 = 128 outputs x
//                        {main, scaled}; the zero block of W' is never multiplied (K blocks of the lo' half update the
//                        scaled columns only)
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

constexpr float SPLIT_SCALE = 2048.0f, SPLIT_INV = 1.0f / 2048.0f;

// kind::f16 instruction descriptor with fp16 operands: D = f32, A = B = f16 (format 0), both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int n, int m = 128) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void split_f16(float x, __half &hi, __half &lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * SPLIT_SCALE);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// conv2 + pool2 + flatten
// ---------------------------------------------------------------------------------------------------------------
constexpr int RC2_GLYPHS = 4;                        // per unit: 2 M tiles x 2 glyphs
constexpr int RC2_BUILDERS = 256;                    // warps 0-7: one A row per thread and tap
constexpr int RC2_EPI = 128;                         // warps 9-12: epilogue (TMEM -> pool -> global)
constexpr int RC2_THREADS = RC2_BUILDERS + 32 + RC2_EPI;  // warp 8: TMA / MMA issuer
constexpr int RC2_ACT_BYTES = RC2_GLYPHS * 144 * 128;  // 73,728
constexpr int RC2_A_STAGE = 2 * 128 * 128;           // two M tiles
constexpr int RC2_A_STAGES = 3;
constexpr int RC2_B_STAGE = 128 * 128;
constexpr int RC2_B_STAGES = 3;                      // weight tiles are requested two taps ahead
constexpr int RC2_OFF_A = RC2_ACT_BYTES;
constexpr int RC2_OFF_B = RC2_OFF_A + RC2_A_STAGES * RC2_A_STAGE;
constexpr int RC2_OFF_BAR = RC2_OFF_B + RC2_B_STAGES * RC2_B_STAGE;
constexpr int RC2_SMEM = RC2_OFF_BAR + 256 + 1024;

// act: [B][144][64] half (hi 32 | lo' 32 per pooled conv1 pixel); tmW: [25 taps x 128 rows][64] half
// out: [B][2048] half = [hi(c * 16 + p) | lo'(c * 16 + p)]
__global__ void __launch_bounds__(RC2_THREADS, 1)
rec_conv2_tc_kernel(const __half *__restrict__ act, const __grid_constant__ CUtensorMap tmW, const float *__restrict__ bias, int B,
                    __half *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[64];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sAct = smem;
  uint8_t *sA = smem + RC2_OFF_A;
  uint8_t *sB = smem + RC2_OFF_B;
  uint64_t *fullA = reinterpret_cast<uint64_t *>(smem + RC2_OFF_BAR);  // [2] builders -> MMA
  uint64_t *emptyA = fullA + RC2_A_STAGES;                             // [2] MMA done with the A stage
  uint64_t *fullB = emptyA + RC2_A_STAGES;                             // [4] TMA landed
  uint64_t *emptyB = fullB + RC2_B_STAGES;                             // [4]
  uint64_t *tfull = emptyB + RC2_B_STAGES, *tempty = tfull + 2;        // [2] accumulator sets
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int units = (B + RC2_GLYPHS - 1) / RC2_GLYPHS;
  const int my_units = (int)blockIdx.x < units ? (units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (tid < 64) s_bias[tid] = bias[tid];
  if (tid == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < RC2_A_STAGES; ++s) { mbar_init(&fullA[s], RC2_BUILDERS); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < RC2_B_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], RC2_EPI / 32); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ================= TMA (weights) + MMA issuer =================
    constexpr uint32_t idesc = make_idesc_f16(128);
    const uint32_t total = (uint32_t)my_units * 25u;
    auto request_weights = [&](uint32_t it) {
      const int s = it % RC2_B_STAGES;
      mbar_wait(&emptyB[s], ((it / RC2_B_STAGES) & 1) ^ 1, err, 32);
      if (elect_one()) {
        mbar_expect_tx(&fullB[s], RC2_B_STAGE);
        tma_load_2d(sB + s * RC2_B_STAGE, &tmW, &fullB[s], 0, (int)(it % 25u) * 128);
      }
      __syncwarp();
    };
    for (uint32_t it = 0; it < (uint32_t)(RC2_B_STAGES - 1) && it < total; ++it) request_weights(it);
    for (uint32_t it = 0; it < total; ++it) {
      const int sa = it % RC2_A_STAGES, sb = it % RC2_B_STAGES;
      const uint32_t tap = it % 25u, unit_no = it / 25u, acc = unit_no & 1;
      if (it + RC2_B_STAGES - 1 < total) request_weights(it + RC2_B_STAGES - 1);
      if (tap == 0) {
        // this accumulator set is free once the epilogue of two units ago has read it
        mbar_wait(&tempty[acc], ((unit_no >> 1) & 1) ^ 1, err, 31);
        tc_fence_after();
      }
      mbar_wait(&fullB[sb], (it / RC2_B_STAGES) & 1, err, 33);
      mbar_wait(&fullA[sa], (it / RC2_A_STAGES) & 1, err, 36);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = make_smem_desc(sB + sb * RC2_B_STAGE);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t adesc = make_smem_desc(sA + sa * RC2_A_STAGE + mt * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + acc * 256 + mt * 128, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | (uint32_t)k) != 0 ? 1u : 0u);
        }
        umma_commit(&emptyA[sa]);
        umma_commit(&emptyB[sb]);
        if (tap == 24) umma_commit(&tfull[acc]);
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ================= builders: the unit's activations -> shared memory, then one A row per thread and tap =================
    const int mt = tid >> 7, r = tid & 127;
    const int g_local = 2 * mt + (r >> 6), oy = (r >> 3) & 7, ox = r & 7;
    const uint32_t act_u32 = smem_u32(sAct), sA_u32 = smem_u32(sA);
    uint32_t it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int g0 = unit * RC2_GLYPHS;
      // every builder has finished reading the previous unit's activations (its last A row is written)
      named_bar_sync(1, RC2_BUILDERS);
      {
        const uint4 *src = reinterpret_cast<const uint4 *>(act + (int64_t)g0 * 144 * 64);
        const int valid16 = (B - g0 < RC2_GLYPHS ? B - g0 : RC2_GLYPHS) * (144 * 128 / 16);
        for (int i = tid; i < RC2_ACT_BYTES / 16; i += RC2_BUILDERS)
          reinterpret_cast<uint4 *>(sAct)[i] = i < valid16 ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
      }
      named_bar_sync(1, RC2_BUILDERS);
      for (int tap = 0; tap < 25; ++tap, ++it) {
        const int s = it % RC2_A_STAGES;
        const int dy = tap / 5, dx = tap - dy * 5;
        const uint32_t src = act_u32 + (uint32_t)((g_local * 144 + (oy + dy) * 12 + ox + dx) * 128);
        const uint32_t dst = sA_u32 + (uint32_t)(s * RC2_A_STAGE + mt * 16384 + r * 128);
        mbar_wait(&emptyA[s], ((it / RC2_A_STAGES) & 1) ^ 1, err, 34);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = (j + r) & 7;  // rotated start: the eight rows of a group hit different banks
          sts_16(dst + (uint32_t)((c ^ (r & 7)) << 4), lds_16(src + (uint32_t)(c << 4)));
        }
        fence_proxy_async();
        mbar_arrive(&fullA[s]);
      }
    }
  } else {
    // ================= epilogue: main + 2^-11 * scaled + bias -> 2x2 max-pool by warp shuffles -> split halves =================
    // A warp reads TMEM lane quarter q = 32 GEMM rows = 4 output rows x 8 columns of one glyph: the pool partners of a
    // pixel are lane ^ 1 (x) and lane ^ 8 (y), so no shared memory is needed.
    const int q = warp & 3;
    uint32_t unit_no = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++unit_no) {
      const int g0 = unit * RC2_GLYPHS;
      const uint32_t acc = unit_no & 1;
      mbar_wait(&tfull[acc], (unit_no >> 1) & 1, err, 35);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int row = q * 32 + lane;                 // GEMM row within the M tile
        const int g = g0 + 2 * mt + (row >> 6);
        const int oy = (row >> 3) & 7, ox = row & 7;
        const bool writer = ((lane & 1) | (lane & 8)) == 0;  // even x, even y
        const int p = (oy >> 1) * 4 + (ox >> 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + (uint32_t)(mt * 128);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float m[32], sc[32];
          tmem_ld32(taddr + h * 32, m);
          tmem_ld32(taddr + 64 + h * 32, sc);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float v = (m[c] + sc[c] * SPLIT_INV) + s_bias[h * 32 + c];
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
            if (writer && g < B) {
              __half hi, lo;
              split_f16(v, hi, lo);
              __half *o = out + (int64_t)g * 2048 + (h * 32 + c) * 16 + p;
              o[0] = hi;
              o[1024] = lo;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// conv2, second form (default): the A tile of a tap is ONE tiled TMA box {64 halves, 8, 8, 4 glyphs} of the activation
// tensor [B][12][12][64] at (x, y) = (dx, dy) — rows land in (glyph, oy, ox) order, 128B-swizzled, exactly the two M tiles
// the builder warps of rec_conv2_tc_kernel used to copy out of a shared-memory image of the unit (those copies, 64 KB of
// shared-memory traffic per tap, were what the first form waited for: ncu showed the builders stalled on the load-store
// unit and the MMA warp on their barrier).  The 25-fold re-read of a unit's 72 KB comes out of L2.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RC2T_A_STAGES = 4, RC2T_B_STAGES = 4;
constexpr int RC2T_OFF_B = RC2T_A_STAGES * RC2_A_STAGE;
constexpr int RC2T_OFF_BAR = RC2T_OFF_B + RC2T_B_STAGES * RC2_B_STAGE;
constexpr int RC2T_SMEM = RC2T_OFF_BAR + 256 + 1024;
constexpr int RC2T_THREADS = 6 * 32;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

__global__ void __launch_bounds__(RC2T_THREADS, 1)
rec_conv2_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const float *__restrict__ bias, int B,
                     __half *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_bias[64];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem, *sB = smem + RC2T_OFF_B;
  uint64_t *fullA = reinterpret_cast<uint64_t *>(smem + RC2T_OFF_BAR), *emptyA = fullA + RC2T_A_STAGES;
  uint64_t *fullB = emptyA + RC2T_A_STAGES, *emptyB = fullB + RC2T_B_STAGES;
  uint64_t *tfull = emptyB + RC2T_B_STAGES, *tempty = tfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int units = (B + RC2_GLYPHS - 1) / RC2_GLYPHS;
  const int my_units = (int)blockIdx.x < units ? (units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total = (uint32_t)my_units * 25u;

  if (tid < 64) s_bias[tid] = bias[tid];
  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < RC2T_A_STAGES; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < RC2T_B_STAGES; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: per (unit, tap) one activation box and one weight tile =================
    for (uint32_t it = 0; it < total; ++it) {
      const int sa = it % RC2T_A_STAGES, sb = it % RC2T_B_STAGES;
      const uint32_t tap = it % 25u, unit_no = it / 25u;
      const int g0 = ((int)blockIdx.x + (int)unit_no * (int)gridDim.x) * RC2_GLYPHS;
      const int dy = (int)tap / 5, dx = (int)tap - dy * 5;
      mbar_wait(&emptyA[sa], ((it / RC2T_A_STAGES) & 1) ^ 1, err, 34);
      mbar_wait(&emptyB[sb], ((it / RC2T_B_STAGES) & 1) ^ 1, err, 32);
      if (elect_one()) {
        mbar_expect_tx(&fullA[sa], RC2_A_STAGE);
        tma_load_4d(sA + sa * RC2_A_STAGE, &tmA, &fullA[sa], 0, dx, dy, g0);  // glyphs past the end: zero fill
        mbar_expect_tx(&fullB[sb], RC2_B_STAGE);
        tma_load_2d(sB + sb * RC2_B_STAGE, &tmW, &fullB[sb], 0, (int)tap * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc_f16(128);
    for (uint32_t it = 0; it < total; ++it) {
      const int sa = it % RC2T_A_STAGES, sb = it % RC2T_B_STAGES;
      const uint32_t tap = it % 25u, unit_no = it / 25u, acc = unit_no & 1;
      if (tap == 0) {
        mbar_wait(&tempty[acc], ((unit_no >> 1) & 1) ^ 1, err, 31);
        tc_fence_after();
      }
      mbar_wait(&fullB[sb], (it / RC2T_B_STAGES) & 1, err, 33);
      mbar_wait(&fullA[sa], (it / RC2T_A_STAGES) & 1, err, 36);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = make_smem_desc(sB + sb * RC2_B_STAGE);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint64_t adesc = make_smem_desc(sA + sa * RC2_A_STAGE + mt * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + acc * 256 + mt * 128, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | (uint32_t)k) != 0 ? 1u : 0u);
        }
        umma_commit(&emptyA[sa]);
        umma_commit(&emptyB[sb]);
        if (tap == 24) umma_commit(&tfull[acc]);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue (as in rec_conv2_tc_kernel): main + 2^-11 * scaled + bias -> 2x2 max-pool -> split halves =================
    const int q = warp & 3;
    uint32_t unit_no = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++unit_no) {
      const int g0 = unit * RC2_GLYPHS;
      const uint32_t acc = unit_no & 1;
      mbar_wait(&tfull[acc], (unit_no >> 1) & 1, err, 35);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int row = q * 32 + lane;
        const int g = g0 + 2 * mt + (row >> 6);
        const int oy = (row >> 3) & 7, ox = row & 7;
        const bool writer = ((lane & 1) | (lane & 8)) == 0;
        const int p = (oy >> 1) * 4 + (ox >> 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + (uint32_t)(mt * 128);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float m[32], sc[32];
          tmem_ld32(taddr + h * 32, m);
          tmem_ld32(taddr + 64 + h * 32, sc);
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float v = (m[c] + sc[c] * SPLIT_INV) + s_bias[h * 32 + c];
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
            if (writer && g < B) {
              __half hi, lo;
              split_f16(v, hi, lo);
              __half *o = out + (int64_t)g * 2048 + (h * 32 + c) * 16 + p;
              o[0] = hi;
              o[1024] = lo;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fc1 + bias + ReLU as a split-fp16 GEMM
// ---------------------------------------------------------------------------------------------------------------
constexpr int RFC_STAGES = 4;
constexpr int RFC_A_STAGE = 128 * 128, RFC_B_STAGE = 256 * 128;
constexpr int RFC_OFF_B = RFC_STAGES * RFC_A_STAGE;
constexpr int RFC_OFF_BAR = RFC_OFF_B + RFC_STAGES * RFC_B_STAGE;
constexpr int RFC_SMEM = RFC_OFF_BAR + 256 + 1024;
constexpr int RFC_THREADS = 6 * 32;  // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

// tmA: [M][2 * K] half (hi K | lo' K); tmB: [(N / 128) x (K / 64) x 256 rows][64] half (rows 0-127 w_hi, 128-255 w_lo' of the
// block's 128 outputs for that K block); out: [M][N] float = relu(A W^T + bias)
__global__ void __launch_bounds__(RFC_THREADS, 1)
rec_fc_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float *__restrict__ bias, int M, int K,
                 int N, int relu, float *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem, *sB = smem + RFC_OFF_B;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + RFC_OFF_BAR);
  uint64_t *empty = full + RFC_STAGES, *tfull = empty + RFC_STAGES, *tempty = tfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = K / 64, nblocks = N / 128, mtiles = (M + 127) / 128;
  const int units = mtiles * nblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < RFC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int nb = unit % nblocks, mtile = unit / nblocks;  // the N blocks of one M tile run back to back: A stays in L2
      for (int kc = 0; kc < 2 * kblocks; ++kc) {
        mbar_wait(&empty[stage], phase ^ 1, err, 41);
        if (elect_one()) {
          const bool first_half = kc < kblocks;  // A chunk from the hi half: both the main and the scaled rows take part
          mbar_expect_tx(&full[stage], RFC_A_STAGE + (first_half ? RFC_B_STAGE : RFC_B_STAGE / 2));
          tma_load_2d(sA + stage * RFC_A_STAGE, &tmA, &full[stage], kc * 64, mtile * 128);
          const int brow = (nb * kblocks + (first_half ? kc : kc - kblocks)) * 256;
          tma_load_2d(sB + stage * RFC_B_STAGE, &tmB, &full[stage], 0, brow);
          if (first_half) tma_load_2d(sB + stage * RFC_B_STAGE + RFC_B_STAGE / 2, &tmB, &full[stage], 0, brow + 128);
        }
        __syncwarp();
        if (++stage == RFC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc256 = make_idesc_f16(256), idesc128 = make_idesc_f16(128);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1, err, 42);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kc = 0; kc < 2 * kblocks; ++kc) {
        mbar_wait(&full[stage], phase, err, 43);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc(sA + stage * RFC_A_STAGE), bdesc = make_smem_desc(sB + stage * RFC_B_STAGE);
          if (kc < kblocks) {  // a_hi x [w_hi | w_lo'] -> main and scaled columns
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc256, (kc | k) != 0 ? 1u : 0u);
          } else {             // a_lo' x w_hi -> scaled columns only
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem + 128, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc128, 1u);
          }
          umma_commit(&empty[stage]);
          if (kc == 2 * kblocks - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == RFC_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int nb = unit % nblocks, mtile = unit / nblocks;
      mbar_wait(&tfull[acc], acc_phase, err, 44);
      tc_fence_after();
      const int row = mtile * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      float *orow = out + (int64_t)row * N + nb * 128;
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        float m[32], sc[32];
        tmem_ld32(taddr + h * 32, m);
        tmem_ld32(taddr + 128 + h * 32, sc);
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v;
            const float4 bz = __ldg(reinterpret_cast<const float4 *>(bias + nb * 128 + h * 32) + j);
            v.x = (m[4 * j + 0] + sc[4 * j + 0] * SPLIT_INV) + bz.x;
            v.y = (m[4 * j + 1] + sc[4 * j + 1] * SPLIT_INV) + bz.y;
            v.z = (m[4 * j + 2] + sc[4 * j + 2] * SPLIT_INV) + bz.z;
            v.w = (m[4 * j + 3] + sc[4 * j + 3] * SPLIT_INV) + bz.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            reinterpret_cast<float4 *>(orow + h * 32)[j] = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
static inline void split_host(float x, uint16_t &hi, uint16_t &lo) {
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn((x - __half2float(h)) * SPLIT_SCALE);
  hi = *reinterpret_cast<const uint16_t *>(&h);
  lo = *reinterpret_cast<const uint16_t *>(&l);
}

// conv2 weights OIHW [64][32][5][5] -> [25 taps][128 rows][64] halves (file header)
void rec_tc_pack_conv2(const float *w, std::vector<uint16_t> &out) {
  out.assign((size_t)25 * 128 * 64, 0);
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < 32; ++ci)
      for (int tp = 0; tp < 25; ++tp) {
        uint16_t hi, lo;
        split_host(w[((size_t)co * 32 + ci) * 25 + tp], hi, lo);
        out[((size_t)tp * 128 + co) * 64 + ci] = hi;             // main:   a_hi * w_hi
        out[((size_t)tp * 128 + 64 + co) * 64 + ci] = lo;        // scaled: a_hi * w_lo'
        out[((size_t)tp * 128 + 64 + co) * 64 + 32 + ci] = hi;   //         a_lo' * w_hi
      }
}

// fc weights [N][K] -> [(N / 128) x (K / 64) blocks][256 rows][64] halves
void rec_tc_pack_fc(const float *w, int N, int K, std::vector<uint16_t> &out) {
  const int kb = K / 64;
  out.assign((size_t)(N / 128) * kb * 256 * 64, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      uint16_t hi, lo;
      split_host(w[(size_t)n * K + k], hi, lo);
      const size_t blk = ((size_t)(n / 128) * kb + k / 64) * 256;
      out[(blk + n % 128) * 64 + k % 64] = hi;
      out[(blk + 128 + n % 128) * 64 + k % 64] = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// conv1 + pool2 (u8 glyphs): conv 5x5 (1 -> 32) + bias + max_pool 2 (char_recognition/model.rs:30-32)
// ---------------------------------------------------------------------------------------------------------------
// Transposed implicit GEMM without im2col, D[channel][conv pixel] (the detector stem's scheme, stem_tc3.cu):
//   * the grey levels are integers <= 255, exact in fp16, so only the WEIGHTS are split: w * 2^e = w_hi + w_lo (two fp16
//     terms; the power of two puts max |w| near 2^14, which keeps w_lo of every weight that matters out of the subnormal
//     range) and the two terms sit side by side in K: A = [w_hi (6 chunks of 8) | w_lo (6 chunks)], both multiplied with the
//     same pixels.  u8 x fp16 products are exact in the fp32 accumulator;  conv = acc / (255 * 2^e);
//   * B = the glyph itself through NO-SWIZZLE K-major descriptors: T[x][y] = the 16 bytes fp16(img[y][x .. x + 7]) (5 used,
//     the weights of the other 3 are zero), x-columns of 32 rows (28 + 4 zero rows).  GEMM row n = x * 32 + y: the eight rows
//     of a core matrix are 8 consecutive y (128 contiguous bytes), the K-adjacent core matrix (filter row r + 1) is the SAME
//     array one row further (LBO = 16 bytes), the next row group is + 128 bytes (SBO, uniform across x-columns);
//   * M = 128 rows = the 32 channels four times, so each TMEM lane quarter holds all channels and the four epilogue warps
//     split the pixels: warp q pools the x-column pair q of an N tile (8 x-columns x 32 y = 256 TMEM columns, two
//     accumulator stages) in registers and writes the fp16-split [hi 32 | lo' 32] record conv2 reads.
constexpr int RC1T_T_BYTES = 24 * 512;                  // T of one glyph: [24 x][32 y][16 B]
constexpr int RC1T_RAW_BYTES = 800;                     // 784 bytes + slack for the 8-byte reads at the end
constexpr int RC1T_OFF_A = 2 * RC1T_T_BYTES;            // two A tiles [128][64] fp16, 128B-swizzled: w_hi, w_lo
constexpr int RC1T_OFF_RAW = RC1T_OFF_A + 2 * 128 * 128;
constexpr int RC1T_OFF_BAR = RC1T_OFF_RAW + 3 * RC1T_RAW_BYTES;
constexpr int RC1T_SMEM = RC1T_OFF_BAR + 256 + 1024;
constexpr int RC1T_THREADS = 9 * 32;                    // warps 0-3 epilogue, 4-7 producers, 8 MMA
static_assert(RC1T_OFF_A % 1024 == 0 && RC1T_OFF_RAW % 16 == 0 && RC1T_OFF_BAR % 8 == 0, "rec conv1 shared-memory layout");

__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ void u8x4_to_f16x4_r(uint32_t x, uint32_t &lo, uint32_t &hi) {
  const uint32_t a = __byte_perm(x, 0x64646464u, 0x4140), b = __byte_perm(x, 0x64646464u, 0x4342);  // 1024 + byte
  asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(0x64006400u));
  asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(b), "r"(0x64006400u));
}
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait32_raw(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// in: [B][784] u8 (16-byte aligned), w: [25][32] fp32 (tap-major), out: [B][144][64] half (hi 32 | lo' 32)
__global__ void __launch_bounds__(RC1T_THREADS, 1)
rec_conv1_tc_kernel(const uint8_t *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias, int B,
                    __half *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_red[RC1T_THREADS / 32];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem + RC1T_OFF_A;
  uint64_t *tfullT = reinterpret_cast<uint64_t *>(smem + RC1T_OFF_BAR);  // [2] T built
  uint64_t *temptyT = tfullT + 2;                                        // [2] MMAs done reading it
  uint64_t *accfull = temptyT + 2, *accempty = accfull + 2;              // [2] accumulator stages
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accempty + 2);
  const uint32_t t_u32 = smem_u32(smem), raw_u32 = smem_u32(smem + RC1T_OFF_RAW);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int step = gridDim.x;

  // ---- one-time setup: weight scale 2^e, A tiles, zero rows of T, barriers, TMEM
  float m = 0.f;
  for (int i = tid; i < 800; i += RC1T_THREADS) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfullT[s], 128);
      mbar_init(&temptyT[s], 1);
      mbar_init(&accfull[s], 1);
      mbar_init(&accempty[s], 4);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * RC1T_T_BYTES / 16; i += RC1T_THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 3 * RC1T_RAW_BYTES / 16; i += RC1T_THREADS) reinterpret_cast<uint4 *>(smem + RC1T_OFF_RAW)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  m = 0.f;
#pragma unroll
  for (int i = 0; i < RC1T_THREADS / 32; ++i) m = fmaxf(m, s_red[i]);
  int e2 = 0;
  if (m > 0.f && m < 3.0e38f) frexpf(m, &e2);  // m = f * 2^e2, f in [0.5, 1)
  int ex = 14 - e2;
  ex = ex < -100 ? -100 : ex > 100 ? 100 : ex;
  const float wscale = ldexpf(1.0f, ex), inv = ldexpf(1.0f, -ex) / 255.0f;
  for (int i = tid; i < 2 * 128 * 8; i += RC1T_THREADS) {
    const int part = i >> 10, row = (i >> 3) & 127, j = i & 7, co = row & 31;  // part 0: w_hi, 1: w_lo; chunk j = filter row
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (j < 5) {
      __half h[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        float v = 0.f;
        if (x < 5) {
          const float ws = w[(j * 5 + x) * 32 + co] * wscale;
          const __half hi = __float2half_rn(ws);
          v = part == 0 ? __half2float(hi) : ws - __half2float(hi);
        }
        h[x] = __float2half_rn(v);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) pk[x] = (uint32_t)__half_as_ushort(h[2 * x]) | ((uint32_t)__half_as_ushort(h[2 * x + 1]) << 16);
    }
    *reinterpret_cast<uint4 *>(sA + part * 16384 + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================= epilogue: lane = channel, warp q = pooled column q of the N tile; 2x2 max-pool on registers =================
    const float bs = bias[lane];
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(warp * 64);
    uint32_t it = 0;
    for (int b = blockIdx.x; b < B; b += step) {
      unsigned short *ob = reinterpret_cast<unsigned short *>(out) + (int64_t)b * 144 * 64 + lane;
#pragma unroll 1
      for (int t = 0; t < 3; ++t, ++it) {
        const uint32_t st = it & 1;
        mbar_wait(&accfull[st], (it >> 1) & 1, err, 41);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld32_raw(tlane + st * 256u, v0);
        tmem_ld32_raw(tlane + st * 256u + 32u, v1);
        tmem_ld_wait32_raw(v0);
        tmem_ld_wait32_raw(v1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&accempty[st]);
        const int px = t * 4 + warp;
#pragma unroll
        for (int py = 0; py < 12; ++py) {
          const float mx = fmaxf(fmaxf(__uint_as_float(v0[2 * py]), __uint_as_float(v0[2 * py + 1])),
                                 fmaxf(__uint_as_float(v1[2 * py]), __uint_as_float(v1[2 * py + 1])));
          const float val = fmaf(mx, inv, bs);  // max(conv) + bias == max(conv + bias)
          const __half hi = __float2half_rn(val);
          const __half lo = __float2half_rn((val - __half2float(hi)) * SPLIT_SCALE);
          unsigned short *op = ob + (py * 12 + px) * 64;
          op[0] = __half_as_ushort(hi);
          op[32] = __half_as_ushort(lo);
        }
      }
    }
  } else if (warp < 8) {
    // ================= producers: glyph bytes (cp.async two glyphs ahead) -> T =================
    const int pt = (warp - 4) * 32 + lane;  // 0..127
    auto prefetch_raw = [&](int b, int rb) {
      if (pt < 49) cp_async_16_zfill(raw_u32 + rb * RC1T_RAW_BYTES + pt * 16, in + (int64_t)b * 784 + pt * 16, true);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if ((int)blockIdx.x < B) prefetch_raw(blockIdx.x, 0); else asm volatile("cp.async.commit_group;" ::: "memory");
    if ((int)blockIdx.x + step < B) prefetch_raw(blockIdx.x + step, 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t n = 0;
    for (int b = blockIdx.x; b < B; b += step, ++n) {
      const uint32_t tb = n & 1;
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      named_bar_sync(9, 128);  // glyph n has landed for everyone; everyone is done with the raw buffer of glyph n - 1
      if (b + 2 * step < B) prefetch_raw(b + 2 * step, (n + 2) % 3); else asm volatile("cp.async.commit_group;" ::: "memory");
      mbar_wait(&temptyT[tb], ((n >> 1) & 1) ^ 1, err, 42);
      const uint32_t rawb = raw_u32 + (n % 3) * RC1T_RAW_BYTES, dst = t_u32 + tb * RC1T_T_BYTES;
      // item (x, y): the 8 bytes img[y][x .. x + 7] (runs past column 27 into the next row: those weights are zero)
      for (int k = pt; k < 24 * 28; k += 128) {
        const int y = k / 24, x = k - y * 24;
        const uint32_t o = (uint32_t)(y * 28 + x);
        const uint32_t wa = rawb + (o & ~3u);
        uint32_t w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];" : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(wa));
        const uint32_t sh = (o & 3u) * 8u;
        uint4 v;
        u8x4_to_f16x4_r(__funnelshift_r(w0, w1, sh), v.x, v.y);
        u8x4_to_f16x4_r(__funnelshift_r(w1, w2, sh), v.z, v.w);
        sts_16(dst + (uint32_t)(x * 512 + y * 16), v);
      }
      fence_proxy_async();
      mbar_arrive(&tfullT[tb]);
    }
  } else {
    // ================= MMA issuer: per glyph 3 N tiles x 6 K steps of UMMA 128 x 256 x 16 =================
    constexpr uint32_t idesc = make_idesc_f16(256);
    const uint64_t adesc = make_smem_desc(sA);
    uint32_t n = 0, it = 0;
    for (int b = blockIdx.x; b < B; b += step, ++n) {
      const uint32_t tb = n & 1;
      mbar_wait(&tfullT[tb], (n >> 1) & 1, err, 43);
#pragma unroll 1
      for (int t = 0; t < 3; ++t, ++it) {
        const uint32_t st = it & 1;
        mbar_wait(&accempty[st], ((it >> 1) & 1) ^ 1, err, 44);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 6; ++k)
            umma_bf16(tmem_base + st * 256u, adesc + (uint64_t)((k / 3) * (16384 >> 4) + 2 * (k % 3)),
                      make_smem_desc_nosw(t_u32 + tb * RC1T_T_BYTES + (uint32_t)(t * 8 * 512 + 2 * (k % 3) * 16), 16, 128), idesc, k != 0 ? 1u : 0u);
          umma_commit(&accfull[st]);
          if (t == 2) umma_commit(&temptyT[tb]);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_rec_conv1_tc(ocrb_ctx *ctx, const uint8_t *glyphs, const float *w, const float *bias, int B, __half *out, int *err) {
  OCRB_TRY(ensure_dyn_smem(ctx, rec_conv1_tc_kernel, RC1T_SMEM));
  const int grid = B < ctx->sm_count ? B : ctx->sm_count;
  rec_conv1_tc_kernel<<<grid, RC1T_THREADS, RC1T_SMEM, ctx->stream>>>(glyphs, w, bias, B, out, err);
  return check_launch(ctx, "rec_tc:conv1");
}

int make_act_tensor_map_box_b(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h, int box_b);

int launch_rec_conv2_tc(ocrb_ctx *ctx, const __half *act, const uint16_t *w_packed, const float *bias, int B, __half *out, int *err) {
  CUtensorMap tmW;
  OCRB_TRY(make_weight_tensor_map(&tmW, w_packed, 25 * 128, 64, 128));
  static const bool builders = getenv("OCRB_REC_CONV2") && strcmp(getenv("OCRB_REC_CONV2"), "smem") == 0;  // first form (knob)
  if (!builders && B >= RC2_GLYPHS) {  // (a box may not be larger than the tensor)
    CUtensorMap tmA;
    OCRB_TRY(make_act_tensor_map_box_b(&tmA, act, B, 12, 12, 64, 8, 8, RC2_GLYPHS));
    OCRB_TRY(ensure_dyn_smem(ctx, rec_conv2_tma_kernel, RC2T_SMEM));
    const int units_t = (B + RC2_GLYPHS - 1) / RC2_GLYPHS;
    const int grid_t = units_t < ctx->sm_count ? units_t : ctx->sm_count;
    rec_conv2_tma_kernel<<<grid_t, RC2T_THREADS, RC2T_SMEM, ctx->stream>>>(tmA, tmW, bias, B, out, err);
    return check_launch(ctx, "rec_tc:conv2");
  }
  OCRB_TRY(ensure_dyn_smem(ctx, rec_conv2_tc_kernel, RC2_SMEM));
  const int units = (B + RC2_GLYPHS - 1) / RC2_GLYPHS;
  const int grid = units < ctx->sm_count ? units : ctx->sm_count;
  rec_conv2_tc_kernel<<<grid, RC2_THREADS, RC2_SMEM, ctx->stream>>>(act, tmW, bias, B, out, err);
  return check_launch(ctx, "rec_tc:conv2");
}

int launch_rec_fc_tc(ocrb_ctx *ctx, const __half *a_split, const uint16_t *w_packed, const float *bias, int M, int K, int N, int relu,
                     float *out, int *err) {
  CUtensorMap tmA, tmB;
  OCRB_TRY(make_weight_tensor_map(&tmA, a_split, M, 2 * K, 128));
  OCRB_TRY(make_weight_tensor_map(&tmB, w_packed, (N / 128) * (K / 64) * 256, 64, 128));
  OCRB_TRY(ensure_dyn_smem(ctx, rec_fc_tc_kernel, RFC_SMEM));
  const int units = ((M + 127) / 128) * (N / 128);
  const int grid = units < ctx->sm_count ? units : ctx->sm_count;
  rec_fc_tc_kernel<<<grid, RFC_THREADS, RFC_SMEM, ctx->stream>>>(tmA, tmB, bias, M, K, N, relu, out, err);
  return check_launch(ctx, "rec_tc:fc1");
}

}  // namespace ocrb
