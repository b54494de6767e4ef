// Detector stem on the tensor cores, second generation (BF16 mode):
//   conv 7x7 s2 p3 (1 -> 64, model.rs:68,109) + batch-norm + ReLU (:69,110-111) + max_pool2d 3x3 s2 p1 (:112), fused;
//   u8 or f32 grey levels in, NHWC bf16 [B][H/4][W/4][64] out.
//
// The first version (stem_tc.cu) built an im2col tile in shared memory (128 B per conv pixel) and pooled from a bf16
// conv tile written back to shared memory: it ran at the shared-memory wavefront limit, 10 % of either roofline.
// This version moves almost nothing through shared memory:
//   * NO im2col.  The 1-channel 7x7 convolution is a GEMM with K = 8 input rows x 8 input columns (row 8 and column 8 carry
//     zero weights).  For conv pixel (cy, cx) K chunk j is the 16 bytes patch[2cy + j][2cx .. 2cx + 7] (bf16).  Conv columns
//     of equal phase (cx mod 4 = phi) are 8 pixels = 16 bytes apart, so with FOUR copies of the bf16 input patch, copy phi
//     shifted by 2 phi pixels, the eight rows of a UMMA core matrix (cx = 4 i + phi, i = 0..7) are one contiguous 128-byte
//     patch row, the K-adjacent core matrix is the next patch row (LBO = 128 B) and the M-adjacent one (cy + 1) is two patch
//     rows further (SBO = 256 B): the A operand is a NO-SWIZZLE K-major descriptor straight onto the patch copy.
//     One M tile = 16 conv rows x 8 columns of one phase; four tiles (phases) = a 16 x 32 conv tile from 19 KB of patch.
//   * NO conv tile in shared memory.  A thread owns TMEM lane (cy, i) of all four phase tiles = conv columns 4i .. 4i+3 of
//     one row: the horizontal 3-max of the pool is in-thread (+ one shuffle from lane i + 1), the vertical one is two
//     shuffles (rows of a warp) plus ONE row handed from the next warp through 4 KB of shared memory.
//   * warp-specialised: 4 producer warps (raw patch by cp.async two units ahead, 4 phase copies), 1 MMA warp, 16 epilogue
//     warps (lane quarter x 16-channel chunk: the epilogue is a latency chain — TMEM load, shuffles, a named barrier — so it
//     wants many warps, not many instructions per warp); patch copies and accumulators double-buffered, one CTA per SM.
// Unit = 7 x 15 pooled pixels <- 16 x 32 conv pixels <- 38 x 70 input pixels.
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

constexpr int S2_PH = 7, S2_PW = 15;                  // pooled pixels per unit
constexpr int S2_ROWS = 38;                           // patch rows: 2 * 15 + 8
constexpr int S2_RAW_WORDS = 20;                      // raw bytes per patch row: 3 lead-in + 70, rounded up to words (76 -> 80)
constexpr int S2_RAW_BYTES = S2_ROWS * S2_RAW_WORDS * 4;
constexpr int S2_PHASE_BYTES = S2_ROWS * 128;         // one phase copy: [38][8 units][16 B]
constexpr int S2_PATCH_BYTES = 4 * S2_PHASE_BYTES;    // 19,456
constexpr int S2_EPI_WARPS = 16, S2_PROD_WARPS = 4;      // epilogue warp = (TMEM lane quarter, 16-channel chunk)
constexpr int S2_THREADS = (S2_EPI_WARPS + S2_PROD_WARPS + 1) * 32;
constexpr int S2_OFF_B = 2 * S2_PATCH_BYTES;          // weights [64][64] bf16, 128B-swizzled
constexpr int S2_OFF_RAW = S2_OFF_B + 64 * 128;       // 3 raw buffers (u8 path)
constexpr int S2_OFF_EXCH = S2_OFF_RAW + 3 * S2_RAW_BYTES;   // [2 unit parities][4 chunks][4 quarters][8 lanes][16 words]
constexpr int S2_OFF_BAR = S2_OFF_EXCH + 2 * 4 * 4 * 8 * 64;
constexpr int S2_SMEM = S2_OFF_BAR + 128 + 1024;

struct Stem2Consts { float scale[64], shift[64]; };

// K-major, NO swizzle: core matrix = 8 rows x 16 B contiguous; lbo = bytes to the K-adjacent core matrix, sbo = bytes to the
// M-adjacent one (cute::UMMA canonical layout INTERLEAVE: ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void cp_async_4_zfill2(uint32_t dst, const void *src, bool ok) {
  const int n = ok ? 4 : 0;  // src-size 0: zero fill (= conv / image padding)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&a), *reinterpret_cast<__nv_bfloat162 *>(&b));
  return *reinterpret_cast<uint32_t *>(&r);
}

template <class TIn>
__global__ void __launch_bounds__(S2_THREADS, 1)
stem_tc2_kernel(const TIn *__restrict__ in, int B, int H, int W, const float *__restrict__ w /*[49][64]*/,
                const __grid_constant__ Stem2Consts sc, __nv_bfloat16 *__restrict__ out, int *err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sB = smem + S2_OFF_B;
  uint64_t *pfull = reinterpret_cast<uint64_t *>(smem + S2_OFF_BAR);  // [2] patch copies ready
  uint64_t *pempty = pfull + 2;                                       // [2] MMAs done reading them
  uint64_t *tfull = pempty + 2, *tempty = tfull + 2;                  // [2] accumulator sets
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
  const uint32_t patch_u32 = smem_u32(smem), raw_u32 = smem_u32(smem + S2_OFF_RAW), exch_u32 = smem_u32(smem + S2_OFF_EXCH);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Hc = H / 2, Wc = W / 2, Hp = H / 4, Wp = W / 4;
  const int tiles_x = (Wp + S2_PW - 1) / S2_PW, tiles_y = (Hp + S2_PH - 1) / S2_PH;
  const int units = tiles_x * tiles_y * B;

  // ---- one-time setup: weights -> B (k = 8 j + s <-> tap (r = j, s); row 7 / column 7 zero), barriers, TMEM
  for (int i = tid; i < 64 * 8; i += S2_THREADS) {
    const int co = i >> 3, j = i & 7;
    uint32_t pk[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float a = 0.f, b = 0.f;
      if (j < 7) {  // batch-norm scale folded into the weights (in fp32, before the bf16 rounding)
        a = w[(j * 7 + 2 * h) * 64 + co] * sc.scale[co];
        if (2 * h + 1 < 7) b = w[(j * 7 + 2 * h + 1) * 64 + co] * sc.scale[co];
      }
      pk[h] = pack_bf16(a, b);
    }
    *reinterpret_cast<uint4 *>(sB + co * 128 + ((j ^ (co & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&pfull[s], S2_PROD_WARPS * 32);
      mbar_init(&pempty[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], S2_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == S2_EPI_WARPS + S2_PROD_WARPS) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto unit_origin = [&](int unit, int &b, int &py0, int &px0) {
    b = unit / (tiles_x * tiles_y);
    const int t = unit - b * (tiles_x * tiles_y);
    py0 = (t / tiles_x) * S2_PH;
    px0 = (t % tiles_x) * S2_PW;
  };

  if (warp < S2_EPI_WARPS) {
    // ================= epilogue: BN + ReLU -> bf16 -> 3x3 / s2 max-pool in registers -> global =================
    const int q = warp & 3, chunk = warp >> 2;         // TMEM lane quarter; 16-channel chunk
    const int g = 4 * q + (lane >> 3), i = lane & 7;   // conv row within the tile; unit (conv columns 4 i .. 4 i + 3)
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++n) {
      int b, py0, px0;
      unit_origin(unit, b, py0, px0);
      const uint32_t acc = n & 1;
      const int cy = 2 * py0 - 1 + g, cxb = 2 * px0 - 1 + 4 * i;
      const bool row_ok = cy >= 0 && cy < Hc;
      mbar_wait_relaxed(&tfull[acc], (n >> 1) & 1, err, 51);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      // units on the image border mask conv pixels outside the conv grid (= pool padding); interior units skip the tests
      const bool edge = py0 == 0 || px0 == 0 || 2 * (py0 + S2_PH) >= Hc || 2 * (px0 + S2_PW) >= Wc;
      auto body = [&](auto CHUNK) {
        constexpr int c = decltype(CHUNK)::value;  // compile-time chunk: the batch-norm shifts become immediates
        uint32_t raw[4][16], P[4][8];
        // all four phase tiles of the chunk are requested at once
#pragma unroll
        for (int phi = 0; phi < 4; ++phi) tmem_ld16_nowait(tbase + (uint32_t)(phi * 64 + c * 16), raw[phi]);
#pragma unroll
        for (int phi = 0; phi < 4; ++phi) {
          tmem_ld_wait16(raw[phi]);
          // the scale sits in the weights; ReLU commutes with max and with the bf16 rounding and is applied to the pooled values
#pragma unroll
          for (int k = 0; k < 8; ++k)
            P[phi][k] = pack_bf16(__uint_as_float(raw[phi][2 * k]) + sc.shift[c * 16 + 2 * k], __uint_as_float(raw[phi][2 * k + 1]) + sc.shift[c * 16 + 2 * k + 1]);
          if (edge) {  // outside the conv grid = pool padding: the most negative finite bf16 stands in for -inf
            const bool ok = row_ok && (cxb + phi) >= 0 && (cxb + phi) < Wc;
#pragma unroll
            for (int k = 0; k < 8; ++k) P[phi][k] = ok ? P[phi][k] : 0xFF7FFF7Fu;
          }
        }
        // horizontal: pooled column 2 i <- conv columns 4i, 4i+1, 4i+2; pooled column 2 i + 1 <- 4i+2, 4i+3, 4(i+1)
        uint32_t He[8], Ho[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t nxt = __shfl_down_sync(0xffffffffu, P[0][k], 1);
          He[k] = hmax2_u32(hmax2_u32(P[0][k], P[1][k]), P[2][k]);
          Ho[k] = hmax2_u32(hmax2_u32(P[2][k], P[3][k]), nxt);
        }
        // the first row of this warp is the third row of the previous quarter's second pooled row (slots alternate with the
        // unit so that a fast warp's next write cannot overtake a slow neighbour's read)
        const uint32_t slot = exch_u32 + (uint32_t)((((n & 1) * 4 + c) * 4) * 8 * 64);
        if (lane < 8 && q > 0) {
          const uint32_t ex = slot + (uint32_t)((q * 8 + i) * 64);
          sts_16(ex, make_uint4(He[0], He[1], He[2], He[3]));
          sts_16(ex + 16, make_uint4(He[4], He[5], He[6], He[7]));
          sts_16(ex + 32, make_uint4(Ho[0], Ho[1], Ho[2], Ho[3]));
          sts_16(ex + 48, make_uint4(Ho[4], Ho[5], Ho[6], Ho[7]));
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + c), "r"(128) : "memory");
        // vertical: lanes 0-7 own pooled row 2 q (conv rows 4q .. 4q+2), lanes 16-23 pooled row 2 q + 1 (4q+2, 4q+3, 4q+4)
        uint32_t Re[8], Ro[8];
        const bool second = lane >= 16;
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0;
        if (second && lane < 24 && q < 3) {
          const uint32_t exn = slot + (uint32_t)(((q + 1) * 8 + i) * 64);
          n0 = lds_16(exn); n1 = lds_16(exn + 16); n2 = lds_16(exn + 32); n3 = lds_16(exn + 48);
        }
        const uint32_t Ne[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w}, No[8] = {n2.x, n2.y, n2.z, n2.w, n3.x, n3.y, n3.z, n3.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t e1 = __shfl_down_sync(0xffffffffu, He[k], 8), e2 = __shfl_down_sync(0xffffffffu, He[k], 16);
          const uint32_t o1 = __shfl_down_sync(0xffffffffu, Ho[k], 8), o2 = __shfl_down_sync(0xffffffffu, Ho[k], 16);
          Re[k] = hmax2_u32(hmax2_u32(hmax2_u32(He[k], e1), second ? Ne[k] : e2), 0u);  // ... and ReLU
          Ro[k] = hmax2_u32(hmax2_u32(hmax2_u32(Ho[k], o1), second ? No[k] : o2), 0u);
        }
        const int t = 2 * q + (second ? 1 : 0);
        if ((lane & 8) == 0 && t < S2_PH && py0 + t < Hp) {
          __nv_bfloat16 *orow = out + (((int64_t)b * Hp + py0 + t) * Wp + px0) * 64 + c * 16;
          const int j0 = 2 * i;
          if (px0 + j0 < Wp) {  // j0 <= 14 always
            st_16(orow + (int64_t)j0 * 64, make_uint4(Re[0], Re[1], Re[2], Re[3]));
            st_16(orow + (int64_t)j0 * 64 + 8, make_uint4(Re[4], Re[5], Re[6], Re[7]));
          }
          if (j0 + 1 < S2_PW && px0 + j0 + 1 < Wp) {
            st_16(orow + (int64_t)(j0 + 1) * 64, make_uint4(Ro[0], Ro[1], Ro[2], Ro[3]));
            st_16(orow + (int64_t)(j0 + 1) * 64 + 8, make_uint4(Ro[4], Ro[5], Ro[6], Ro[7]));
          }
        }
      };
      switch (chunk) {
        case 0: body(std::integral_constant<int, 0>{}); break;
        case 1: body(std::integral_constant<int, 1>{}); break;
        case 2: body(std::integral_constant<int, 2>{}); break;
        default: body(std::integral_constant<int, 3>{}); break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  } else if (warp < S2_EPI_WARPS + S2_PROD_WARPS) {
    // ================= producers: input patch -> four phase-shifted bf16 copies =================
    const int pt = tid - S2_EPI_WARPS * 32;  // 0..127
    constexpr int PT = S2_PROD_WARPS * 32;
    auto prefetch_raw = [&](int unit, int rb) {  // u8 path: raw patch rows, zero-filled outside the image
      int b, py0, px0;
      unit_origin(unit, b, py0, px0);
      const int iy0 = 4 * py0 - 5, xa = 4 * px0 - 8;  // patch column p is raw byte 3 + p
      const uint8_t *img = reinterpret_cast<const uint8_t *>(in) + (int64_t)b * H * W;
      for (int k = pt; k < S2_ROWS * S2_RAW_WORDS; k += PT) {
        const int row = k / S2_RAW_WORDS, wd = k - row * S2_RAW_WORDS;
        const int yy = iy0 + row, xx = xa + 4 * wd;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
        cp_async_4_zfill2(raw_u32 + rb * S2_RAW_BYTES + k * 4, ok ? img + (int64_t)yy * W + xx : img, ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int step = gridDim.x;
    if (sizeof(TIn) == 1) {
      if ((int)blockIdx.x < units) prefetch_raw(blockIdx.x, 0); else asm volatile("cp.async.commit_group;" ::: "memory");
      if ((int)blockIdx.x + step < units) prefetch_raw(blockIdx.x + step, 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < units; unit += step, ++n) {
      const uint32_t pb = n & 1;
      const uint32_t dst0 = patch_u32 + pb * S2_PATCH_BYTES;
      if (sizeof(TIn) == 1) {
        // raw buffer n % 3 holds this unit.  After the barrier every producer's share of it has landed AND every producer
        // has finished building unit n - 1, whose raw buffer the prefetch of unit n + 2 reuses.
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        asm volatile("bar.sync 8, %0;" ::"r"(PT) : "memory");
        if (unit + 2 * step < units) prefetch_raw(unit + 2 * step, (n + 2) % 3); else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      mbar_wait_relaxed(&pempty[pb], ((n >> 1) & 1) ^ 1, err, 52);
      if (sizeof(TIn) == 1) {
        const uint32_t rawb = raw_u32 + (n % 3) * S2_RAW_BYTES;
        for (int k = pt; k < S2_ROWS * 32; k += PT) {
          const int u = k & 7, phi = (k >> 3) & 3, y = k >> 5;
          const int o = 8 * u + 2 * phi + 3;  // first raw byte of the unit
          const uint32_t wa = rawb + (uint32_t)(y * S2_RAW_WORDS * 4 + (o & ~3));
          uint32_t w0, w1, w2;
          asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];" : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(wa));
          const uint32_t sh = (uint32_t)(o & 3) * 8;
          const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);  // bytes 0-3, 4-7 of the unit
          uint4 v;
          v.x = pack_bf16((float)(lo & 0xffu), (float)((lo >> 8) & 0xffu));
          v.y = pack_bf16((float)((lo >> 16) & 0xffu), (float)(lo >> 24));
          v.z = pack_bf16((float)(hi & 0xffu), (float)((hi >> 8) & 0xffu));
          v.w = pack_bf16((float)((hi >> 16) & 0xffu), (float)(hi >> 24));
          sts_16(dst0 + (uint32_t)(phi * S2_PHASE_BYTES + y * 128 + u * 16), v);
        }
      } else {
        int b, py0, px0;
        unit_origin(unit, b, py0, px0);
        const int iy0 = 4 * py0 - 5, ix0 = 4 * px0 - 5;
        const TIn *img = in + (int64_t)b * H * W;
        for (int k = pt; k < S2_ROWS * 32; k += PT) {
          const int u = k & 7, phi = (k >> 3) & 3, y = k >> 5;
          const int yy = iy0 + y, x0 = ix0 + 8 * u + 2 * phi;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = (yy >= 0 && yy < H && x0 + e >= 0 && x0 + e < W) ? (float)img[(int64_t)yy * W + x0 + e] : 0.0f;
          sts_16(dst0 + (uint32_t)(phi * S2_PHASE_BYTES + y * 128 + u * 16),
                 make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7])));
        }
      }
      fence_proxy_async();
      mbar_arrive(&pfull[pb]);
    }
  } else {
    // ================= MMA issuer: 4 phase tiles x 4 K steps of UMMA 128 x 64 x 16 per unit =================
    constexpr uint32_t idesc = make_idesc(64);
    const uint64_t bdesc = make_smem_desc(sB);
    uint32_t n = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++n) {
      const uint32_t pb = n & 1, acc = n & 1;
      mbar_wait_relaxed(&tempty[acc], ((n >> 1) & 1) ^ 1, err, 53);
      mbar_wait_relaxed(&pfull[pb], (n >> 1) & 1, err, 54);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int phi = 0; phi < 4; ++phi) {
          const uint32_t a0 = patch_u32 + pb * S2_PATCH_BYTES + phi * S2_PHASE_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + acc * 256 + phi * 64, make_smem_desc_nosw(a0 + (uint32_t)(2 * k * 128), 128, 256), bdesc + (uint64_t)(2 * k), idesc,
                      k != 0 ? 1u : 0u);
        }
        umma_commit(&pempty[pb]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == S2_EPI_WARPS + S2_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_stem_tc2(ocrb_ctx *ctx, const void *in, int is_u8, int B, int H, int W, const float *w, const float *scale_host,
                    const float *shift_host, __nv_bfloat16 *out, int *err) {
  Stem2Consts sc;
  memcpy(sc.scale, scale_host, sizeof(sc.scale));
  memcpy(sc.shift, shift_host, sizeof(sc.shift));
  OCRB_TRY(ensure_dyn_smem(ctx, stem_tc2_kernel<uint8_t>, S2_SMEM));
  OCRB_TRY(ensure_dyn_smem(ctx, stem_tc2_kernel<float>, S2_SMEM));
  const int Hp = H / 4, Wp = W / 4;
  const int64_t units = cdiv(Wp, S2_PW) * cdiv(Hp, S2_PH) * B;
  const int sms = ctx->sm_limit > 0 && ctx->sm_limit < ctx->sm_count ? ctx->sm_limit : ctx->sm_count;
  const int grid = (int)(units < sms ? units : sms);
  if (is_u8)
    stem_tc2_kernel<uint8_t><<<grid, S2_THREADS, S2_SMEM, ctx->stream>>>((const uint8_t *)in, B, H, W, w, sc, out, err);
  else
    stem_tc2_kernel<float><<<grid, S2_THREADS, S2_SMEM, ctx->stream>>>((const float *)in, B, H, W, w, sc, out, err);
  return check_launch(ctx, "tc:stem");
}

}  // namespace ocrb
