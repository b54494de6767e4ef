// Detector stem on the tensor cores, third generation (BF16 mode) — the TRANSPOSED implicit GEMM:
//   conv 7x7 s2 p3 (1 -> 64, model.rs:68,109) + batch-norm + ReLU (:69,110-111) + max_pool2d 3x3 s2 p1 (:112), fused;
//   u8 or f32 grey levels in, NHWC bf16 [B][H/4][W/4][64] out.
//
// stem_tc.cu builds an im2col tile and pools from a conv tile in shared memory (shared-memory wavefront bound);
// stem_tc2.cu has no im2col but keeps conv PIXELS in the TMEM lanes, so the 3x3 max-pool crosses threads (shuffles,
// packing, a named barrier: 13 k warp instructions per unit, issue bound).  Here the GEMM is D[channel][pixel]:
//   * A = the weights, [64 channels, stored TWICE (rows m and m + 64: M = 128 costs the tensor pipe the same as M = 64, and the
//     copy puts every channel into two TMEM lane quarters, so epilogue warps on all four schedulers can read it)][K = 64], fp16,
//     batch-norm scale folded in and normalised per channel by a power of two (undone in the epilogue's FFMA) so that any
//     scale stays inside the fp16 range; 128B-swizzled, resident in shared memory for the whole kernel;
//   * B = the input patch itself, read through NO-SWIZZLE K-major descriptors (no im2col): K chunk j of conv pixel (g, cx)
//     is the 16 bytes patch[2g + j][2cx .. 2cx + 7]; conv columns of equal phase (cx mod 4) are 16 bytes apart, so with four
//     phase-shifted fp16 copies of the patch, interleaved per PAIR of patch rows as [pair][phase][row parity][8 x 16 B], the
//     eight rows of a core matrix are one 128-byte run, the K-adjacent core matrix is the other row of the pair (LBO = 128),
//     the next core-matrix group is the next phase / the next conv row (SBO = 256, uniform), and K step k starts one pair
//     further (+ 1024 B): ONE UMMA 128 x 128 x 16 covers 4 conv rows x 32 conv columns.  Grey levels 0..255 are exact in
//     fp16, and u8 -> fp16 needs no conversion instruction: byte b interleaved with 0x64 is the fp16 number 1024 + b, one
//     packed HSUB2 removes the 1024 (PRMT + HADD2 instead of the quarter-rate I2F + F2FP);
//   * D: TMEM lane = channel, column = conv pixel ((g * 4 + phase) * 8 + u, cx = 4u + phase).  An epilogue thread owns one
//     channel and reads whole conv rows: the 3x3 / s2 max-pool is 3-input max instructions on registers — no shuffles, no
//     shared memory, no packing before the pool; scale-back + shift + ReLU + bf16 rounding are applied to the pooled values
//     only (max commutes with the monotonic affine / ReLU / rounding).
// Warp-specialised, one CTA per SM, 21 warps: 12 epilogue warps (three per scheduler), 8 producer warps (raw patch by
// cp.async in aligned 16-byte chunks two units ahead -> one 8-pixel group of all four phase copies per thread,
// double-buffered), 1 MMA warp.  A unit's 13 conv rows sit in FOUR accumulator stages of TMEM columns (4 conv rows = 128
// columns each; the last one a single row, N = 32); epilogue warp group W_q owns stage q = pooled rows 2q, 2q + 1 (conv rows
// 4q .. 4q + 4: the fifth is the first row of stage q + 1, read by both neighbours) and is split once more into two column
// halves (pooled columns 0..7 / 8..14) and two channel halves, so a stage is handed back to the MMA warp after a few row
// reads and the next unit's MMAs run under the current unit's epilogue.  Vertical maximum first (17 three-input maxima per
// pooled row and thread), then the horizontal one.  Unit = 6 x 15 pooled pixels <- 13 x 32 conv pixels <- 32 x 70 input
// pixels.  Measured (profiles/r2_stem3_ncu.md): 3.6 ms per 1024 images of 800 x 800 against 8.2 ms for stem_tc.cu; issue
// slots 62 % busy, ALU pipe 52 % — the kernel is bound by instruction issue (maxima, conversions, 2-byte stores), not by the
// tensor pipe (41 % busy) or shared memory (46 %).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

constexpr int S3_PH = 6, S3_PW = 15;                  // pooled pixels per unit
constexpr int S3_ROWS = 32;                           // patch rows: conv rows 0..12 of the unit read rows 2r .. 2r + 7
constexpr int S3_RAW_ROW = 96;                        // raw bytes per patch row: six 16-byte chunks from a 16-byte-aligned x
                                                      // (patch column p = raw byte delta + 3 + p, delta in {0, 4, 8, 12})
constexpr int S3_RAW_BYTES = S3_ROWS * S3_RAW_ROW;
constexpr int S3_PATCH_BYTES = (S3_ROWS / 2) * 1024;  // [16 pairs][4 phases][2 rows][8 x 16 B] = 16,384
constexpr int S3_EPI_WARPS = 12, S3_PROD_WARPS = 8;
constexpr int S3_THREADS = (S3_EPI_WARPS + S3_PROD_WARPS + 1) * 32;  // warps 0-11 epilogue, 12-19 producers, 20 MMA
constexpr int S3_OFF_A = 2 * S3_PATCH_BYTES;          // weights [128][64] fp16, 128B-swizzled (rows 64.. = rows 0..63 again)
constexpr int S3_OFF_RAW = S3_OFF_A + 128 * 128;      // 3 raw buffers (u8 path)
constexpr int S3_OFF_BAR = S3_OFF_RAW + 3 * S3_RAW_BYTES;
constexpr int S3_OFF_MUL = S3_OFF_BAR + 256;          // per-channel 2^e (fp32) that the weights were divided by
constexpr int S3_SMEM = S3_OFF_MUL + 256 + 1024;
static_assert(S3_OFF_A % 1024 == 0 && S3_OFF_RAW % 16 == 0 && S3_OFF_BAR % 8 == 0, "stem_tc3 shared-memory layout");
static_assert(S3_ROWS * 8 == S3_PROD_WARPS * 32 && S3_ROWS * (S3_RAW_ROW / 16) <= S3_PROD_WARPS * 32, "one build item per producer thread");

struct Stem3Consts { float scale[64], shift[64]; };

// kind::f16 instruction descriptor with fp16 operands: D = f32, A = B = f16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_h(int n, int m = 128) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_smem_desc_nosw3(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = SWIZZLE_NONE
}
// n / d for n, d < 2^20 by multiply-shift (m = ceil(2^40 / d): exact while n * d < 2^40)
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint64_t m) { return (uint32_t)(((uint64_t)n * m) >> 40); }
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
// four grey levels (the bytes of x) -> four fp16 numbers: byte b next to 0x64 reads as 1024 + b, HSUB2 takes the 1024 away
__device__ __forceinline__ void u8x4_to_f16x4(uint32_t x, uint32_t &lo, uint32_t &hi) {
  const uint32_t a = __byte_perm(x, 0x64646464u, 0x4140), b = __byte_perm(x, 0x64646464u, 0x4342);
  asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(0x64006400u));
  asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(b), "r"(0x64006400u));
}
// relu(x) rounded to bf16 (F2FP with the relu modifier: the ReLU costs no instruction)
__device__ __forceinline__ unsigned short relu_bf16(float x) {
  unsigned short h;
  asm("cvt.rn.relu.bf16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
// one conv row of a column half in registers: [0..3] phase 0 (u0 .. u0+3), [4] phase 0 (u0+4), [5..8] phase 1, [9..12] phase 2,
// [13..16] phase 3
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1_nowait(uint32_t taddr, uint32_t &a) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_row17(uint32_t taddr /* column of (phase 0, u0) */, uint32_t (&r)[17]) {
  tmem_ld4_nowait(taddr, r[0], r[1], r[2], r[3]);
  tmem_ld1_nowait(taddr + 4u, r[4]);
  tmem_ld4_nowait(taddr + 8u, r[5], r[6], r[7], r[8]);
  tmem_ld4_nowait(taddr + 16u, r[9], r[10], r[11], r[12]);
  tmem_ld4_nowait(taddr + 24u, r[13], r[14], r[15], r[16]);
}
// the registers of `r` are valid after this (the "+r" operands order every later use behind the wait)
__device__ __forceinline__ void tmem_ld_wait17(uint32_t (&r)[17]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16])
               :
               : "memory");
}

template <class TIn>
__global__ void __launch_bounds__(S3_THREADS, 1)
stem_tc3_kernel(const TIn *__restrict__ in, int B, int H, int W, const float *__restrict__ w /*[49][64]*/,
                const __grid_constant__ Stem3Consts sc, __nv_bfloat16 *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem + S3_OFF_A;
  uint64_t *pfull = reinterpret_cast<uint64_t *>(smem + S3_OFF_BAR);  // [2] patch copies ready
  uint64_t *pempty = pfull + 2;                                       // [2] MMAs done reading them
  uint64_t *tfull = pempty + 2;                                       // [4] accumulator stage ready
  uint64_t *tempty = tfull + 4;                                       // [4] accumulator stage drained
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 4);
  float *s_mul = reinterpret_cast<float *>(smem + S3_OFF_MUL);
  const uint32_t patch_u32 = smem_u32(smem), raw_u32 = smem_u32(smem + S3_OFF_RAW);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for the compiler */, lane = tid & 31;
  const int Hp = H / 4, Wp = W / 4;
  const int tiles_x = (Wp + S3_PW - 1) / S3_PW, tiles_y = (Hp + S3_PH - 1) / S3_PH;
  const int units = tiles_x * tiles_y * B;

  // ---- one-time setup: per-channel power of two, weights -> A (k = 8 j + s <-> tap (r = j, s); row 7 / column 7 zero;
  //      rows 64.. repeat rows 0..63), barriers, TMEM
  if (tid < 64) {
    float m = 0.f;
    for (int t = 0; t < 49; ++t) m = fmaxf(m, fabsf(w[t * 64 + tid] * sc.scale[tid]));
    int e = 0;
    if (m > 0.f && m < 3.0e38f) frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1): the scaled weights lie in (-1, 1)
    e = e < -100 ? -100 : e > 100 ? 100 : e;
    s_mul[tid] = ldexpf(1.0f, e);
  }
  __syncthreads();
  for (int i = tid; i < 128 * 8; i += S3_THREADS) {
    const int row = i >> 3, j = i & 7, co = row & 63;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (j < 7) {
      const float f = sc.scale[co] / s_mul[co];  // (division by a power of two: exact)
#pragma unroll
      for (int h = 0; h < 4; ++h) {  // batch-norm scale folded into the weights (in fp32, before the fp16 rounding)
        const float a = w[(j * 7 + 2 * h) * 64 + co] * f;
        const float b = 2 * h + 1 < 7 ? w[(j * 7 + 2 * h + 1) * 64 + co] * f : 0.f;
        pk[h] = pack_f16(a, b);
      }
    }
    *reinterpret_cast<uint4 *>(sA + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pfull[s], S3_PROD_WARPS * 32);
      mbar_init(&pempty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], s == 0 || s == 3 ? 4 : 8);  // the stage's own four warps + the four above it (their fifth row)
    }
    fence_barrier_init();
  }
  if (warp == S3_EPI_WARPS + S3_PROD_WARPS) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t tiles_xy = (uint32_t)(tiles_x * tiles_y);
  const uint64_t magic_xy = ((1ull << 40) + tiles_xy - 1) / tiles_xy, magic_x = ((1ull << 40) + tiles_x - 1) / (uint32_t)tiles_x;
  auto unit_origin = [&](int unit, int &b, int &py0, int &px0) {
    b = (int)div_magic((uint32_t)unit, magic_xy);
    const uint32_t t = (uint32_t)unit - (uint32_t)b * tiles_xy, ty = div_magic(t, magic_x);
    py0 = (int)ty * S3_PH;
    px0 = (int)(t - ty * (uint32_t)tiles_x) * S3_PW;
  };
  const int step = gridDim.x;

  if (warp < S3_EPI_WARPS) {
    // ================= epilogue: one thread = one channel; 3x3 / s2 max-pool on registers =================
    // warp -> TMEM lane quarter warp & 3 (hardware rule), channel half warp & 1 (quarters 2, 3 hold the second copy), pooled
    // column half c (columns 0..7 / 8..14), stage = pooled-row pair q.  Vertical maximum first (three conv rows of the column
    // half in registers, 17 three-input maxima), then the horizontal one, then FFMA + convert-with-ReLU + store per pixel.
    const int quarter = warp & 3, c = (warp >> 1) & 1, q = warp >> 2, ch = (warp & 1) * 32 + lane;
    const float mul = s_mul[ch], shift = sc.shift[ch];
    const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * 128 + 4 * c);  // (stage q, row 0, phase 0, u0)
    const uint32_t NEG = 0xff800000u;  // -inf = max-pool padding
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < units; unit += step, ++n) {
      int b, py0, px0;
      unit_origin(unit, b, py0, px0);
      const bool top = py0 == 0 && q == 0, left = px0 == 0 && c == 0;
      const uint32_t par = n & 1;
      const int rows_ok = Hp - (py0 + 2 * q), cols_ok = Wp - (px0 + 8 * c);  // this warp's pooled rows / columns inside the map
      unsigned short *obase = reinterpret_cast<unsigned short *>(out) + (((int64_t)b * Hp + py0 + 2 * q) * Wp + px0 + 8 * c) * 64 + ch;
      uint32_t r0[17], r1[17], r2[17];
      unsigned short o[8];
      auto vmax = [&](uint32_t (&d)[17], const uint32_t (&x)[17], const uint32_t (&y)[17], const uint32_t (&z)[17]) {
#pragma unroll
        for (int k = 0; k < 17; ++k) d[k] = __float_as_uint(max3f(__uint_as_float(x[k]), __uint_as_float(y[k]), __uint_as_float(z[k])));
      };
      auto pool_row = [&](const uint32_t (&m)[17]) {  // conv column 4u + phase of the unit; this half holds u = 4c .. 4c + 4
        const float m0 = left ? __uint_as_float(NEG) : __uint_as_float(m[0]);  // conv column -1
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float e = max3f(v == 0 ? m0 : __uint_as_float(m[v]), __uint_as_float(m[5 + v]), __uint_as_float(m[9 + v]));    // 4u .. 4u+2
          const float f = max3f(__uint_as_float(m[9 + v]), __uint_as_float(m[13 + v]), __uint_as_float(m[v + 1]));            // 4u+2 .. 4u+4
          o[2 * v] = relu_bf16(fmaf(e, mul, shift));
          o[2 * v + 1] = relu_bf16(fmaf(f, mul, shift));  // (column half 1: o[7] is pooled column 15 — not stored)
        }
      };
      auto store_row = [&](int t) {
        if (t < rows_ok) {
          unsigned short *orow = obase + (int64_t)t * Wp * 64;
          if (cols_ok >= 8) {
#pragma unroll
            for (int j = 0; j < 7; ++j) orow[j * 64] = o[j];
            if (c == 0) orow[7 * 64] = o[7];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < cols_ok && j < 8 - c) orow[j * 64] = o[j];
          }
        }
      };
      auto release = [&](int s) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[s]);
      };
      mbar_wait(&tfull[q], par, err, 61);
      tc_fence_after();
      tmem_ld_row17(tcol, r0);
      tmem_ld_row17(tcol + 32u, r1);
      tmem_ld_row17(tcol + 64u, r2);
      tmem_ld_wait17(r0);
      tmem_ld_wait17(r1);
      tmem_ld_wait17(r2);
      if (top) {
#pragma unroll
        for (int k = 0; k < 17; ++k) r0[k] = NEG;  // conv row -1
      }
      vmax(r0, r0, r1, r2);
      tmem_ld_row17(tcol + 96u, r1);  // conv row 4q + 3, in flight under the first pooled row
      pool_row(r0);
      tmem_ld_wait17(r1);
      release(q);  // stage q is in registers
      mbar_wait(&tfull[q + 1], par, err, 62);
      tc_fence_after();
      tmem_ld_row17(tcol + 128u, r0);  // conv row 4q + 4 = first row of stage q + 1
      store_row(0);
      tmem_ld_wait17(r0);
      release(q + 1);
      vmax(r0, r2, r1, r0);  // (loading four rows at once and releasing the stage earlier measured 10 % slower: registers)
      pool_row(r0);
      store_row(1);
    }
  } else if (warp < S3_EPI_WARPS + S3_PROD_WARPS) {
    // ================= producers: input patch -> four phase-shifted fp16 copies, interleaved per row pair =================
    const int pt = (warp - S3_EPI_WARPS) * 32 + lane;  // 0..127
    constexpr int PT = S3_PROD_WARPS * 32;
    // per-thread constants: this thread's 16-byte chunk of the raw prefetch (row prow, chunk pc) and its build item (row by,
    // 8-pixel group bu)
    const int prow = pt / (S3_RAW_ROW / 16), pc = pt - prow * (S3_RAW_ROW / 16);
    const bool pf = pt < S3_ROWS * (S3_RAW_ROW / 16);
    const int bu = pt & 7, by = pt >> 3;
    const uint32_t b_src = (uint32_t)(by * S3_RAW_ROW + 8 * bu), b_dst = (uint32_t)((by >> 1) * 1024 + (by & 1) * 128 + bu * 16);
    uint32_t deltas = 0;  // byte offset (0, 4, 8, 12) of patch column -3 inside raw buffer rb: bits [4 rb, 4 rb + 4)
    auto prefetch_raw = [&](int unit, int rb) {  // u8 path: raw patch rows as 16-byte chunks, zero-filled outside the image
      int b, py0, px0;
      unit_origin(unit, b, py0, px0);
      const int x8 = 4 * px0 - 8;
      deltas = (deltas & ~(0xFu << (4 * rb))) | ((uint32_t)(x8 & 15) << (4 * rb));
      if (pf) {
        const int yy = 4 * py0 - 5 + prow, xx = (x8 & ~15) + 16 * pc;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;  // (W % 16 == 0: a chunk is inside or outside as a whole)
        const uint8_t *img = reinterpret_cast<const uint8_t *>(in);
        cp_async_16_zfill(raw_u32 + (uint32_t)(rb * S3_RAW_BYTES + pt * 16), ok ? img + ((int64_t)b * H + yy) * W + xx : img, ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (sizeof(TIn) == 1) {
      if ((int)blockIdx.x < units) prefetch_raw(blockIdx.x, 0); else asm volatile("cp.async.commit_group;" ::: "memory");
      if ((int)blockIdx.x + step < units) prefetch_raw(blockIdx.x + step, 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t n = 0;
    int rb_cur = 0;  // n % 3
    for (int unit = blockIdx.x; unit < units; unit += step, ++n) {
      const uint32_t pb = n & 1;
      const uint32_t dst0 = patch_u32 + pb * S3_PATCH_BYTES;
      const int rb_next = rb_cur == 0 ? 2 : rb_cur - 1;  // (n + 2) % 3
      if (sizeof(TIn) == 1) {
        // raw buffer n % 3 holds this unit.  After the barrier every producer's share of it has landed AND every producer
        // has finished building unit n - 1, whose raw buffer the prefetch of unit n + 2 reuses.
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        asm volatile("bar.sync 8, %0;" ::"r"(PT) : "memory");
        if (unit + 2 * step < units) prefetch_raw(unit + 2 * step, rb_next); else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      mbar_wait(&pempty[pb], ((n >> 1) & 1) ^ 1, err, 52);
      if (sizeof(TIn) == 1) {
        // one item = 8 patch columns group u of patch row y, all four phases: raw bytes delta + 8u + 3 + 2 phase .. + 7
        {
          {
            const uint32_t ra = raw_u32 + (uint32_t)(rb_cur * S3_RAW_BYTES) + ((deltas >> (4 * rb_cur)) & 15u) + b_src;
            uint32_t w0, w1, w2, w3, w4;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1+12];" : "=r"(w3) : "r"(ra));
            asm volatile("ld.shared.u32 %0, [%1+16];" : "=r"(w4) : "r"(ra));
            // 4-byte groups starting at byte 3, 7, 11 (shift 24) and 5, 9, 13 (shift 8) of the 20-byte window
            uint32_t g[6][2];
            u8x4_to_f16x4(__funnelshift_r(w0, w1, 24), g[0][0], g[0][1]);
            u8x4_to_f16x4(__funnelshift_r(w1, w2, 24), g[1][0], g[1][1]);
            u8x4_to_f16x4(__funnelshift_r(w2, w3, 24), g[2][0], g[2][1]);
            u8x4_to_f16x4(__funnelshift_r(w1, w2, 8), g[3][0], g[3][1]);
            u8x4_to_f16x4(__funnelshift_r(w2, w3, 8), g[4][0], g[4][1]);
            u8x4_to_f16x4(__funnelshift_r(w3, w4, 8), g[5][0], g[5][1]);
            const uint32_t d = dst0 + b_dst;
            sts_16(d, make_uint4(g[0][0], g[0][1], g[1][0], g[1][1]));        // phase 0: bytes 3..10
            sts_16(d + 256, make_uint4(g[3][0], g[3][1], g[4][0], g[4][1]));  // phase 1: bytes 5..12
            sts_16(d + 512, make_uint4(g[1][0], g[1][1], g[2][0], g[2][1]));  // phase 2: bytes 7..14
            sts_16(d + 768, make_uint4(g[4][0], g[4][1], g[5][0], g[5][1]));  // phase 3: bytes 9..16
          }
        }
      } else {
        int b, py0, px0;
        unit_origin(unit, b, py0, px0);
        const int iy0 = 4 * py0 - 5, ix0 = 4 * px0 - 5;
        const TIn *img = in + (int64_t)b * H * W;
        for (int k = pt; k < S3_ROWS * 32; k += PT) {
          const int u = k & 7, phi = (k >> 3) & 3, y = k >> 5;
          const int yy = iy0 + y, x0 = ix0 + 8 * u + 2 * phi;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (yy >= 0 && yy < H && x0 + e >= 0 && x0 + e < W) ? (float)img[(int64_t)yy * W + x0 + e] : 0.0f;
          sts_16(dst0 + (uint32_t)((y >> 1) * 1024 + phi * 256 + (y & 1) * 128 + u * 16),
                 make_uint4(pack_f16(f[0], f[1]), pack_f16(f[2], f[3]), pack_f16(f[4], f[5]), pack_f16(f[6], f[7])));
        }
      }
      fence_proxy_async();
      mbar_arrive(&pfull[pb]);
      rb_cur = rb_cur == 2 ? 0 : rb_cur + 1;
    }
  } else {
    // ================= MMA issuer: per unit 3 stages x 4 K steps of UMMA 128 x 128 x 16 + 1 stage of 128 x 32 x 16 =================
    constexpr uint32_t idesc128 = make_idesc_h(128), idesc32 = make_idesc_h(32);
    const uint64_t adesc = make_smem_desc(sA);
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < units; unit += step, ++n) {
      const uint32_t pb = n & 1;
      mbar_wait(&pfull[pb], (n >> 1) & 1, err, 54);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (n > 0) mbar_wait(&tempty[s], (n - 1) & 1, err, 55 + s);  // stage s drained of unit n - 1
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + (uint32_t)(s * 128), adesc + (uint64_t)(2 * k),
                      make_smem_desc_nosw3(patch_u32 + pb * S3_PATCH_BYTES + (uint32_t)((4 * s + k) * 1024), 128, 256),
                      s < 3 ? idesc128 : idesc32, k != 0 ? 1u : 0u);
          umma_commit(&tfull[s]);
          if (s == 3) umma_commit(&pempty[pb]);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == S3_EPI_WARPS + S3_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_stem_tc3(ocrb_ctx *ctx, const void *in, int is_u8, int B, int H, int W, const float *w, const float *scale_host,
                    const float *shift_host, __nv_bfloat16 *out, int *err) {
  Stem3Consts sc;
  memcpy(sc.scale, scale_host, sizeof(sc.scale));
  memcpy(sc.shift, shift_host, sizeof(sc.shift));
  const int Hp = H / 4, Wp = W / 4;
  const int64_t units = cdiv(Wp, S3_PW) * cdiv(Hp, S3_PH) * B;
  const int sms = ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count;
  const int grid = (int)(units < sms ? units : sms);
  OCRB_TRY(ensure_dyn_smem(ctx, stem_tc3_kernel<uint8_t>, S3_SMEM));
  OCRB_TRY(ensure_dyn_smem(ctx, stem_tc3_kernel<float>, S3_SMEM));
  if (is_u8)
    stem_tc3_kernel<uint8_t><<<grid, S3_THREADS, S3_SMEM, ctx->stream>>>((const uint8_t *)in, B, H, W, w, sc, out, err);
  else
    stem_tc3_kernel<float><<<grid, S3_THREADS, S3_SMEM, ctx->stream>>>((const float *)in, B, H, W, w, sc, out, err);
  return check_launch(ctx, "tc:stem");
}

}  // namespace ocrb
