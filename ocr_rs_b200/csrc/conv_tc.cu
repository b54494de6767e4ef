// BF16 mode of the detector: implicit-GEMM convolution on the 5th-generation tensor cores.
//
//   * activations NHWC bf16, weights [Cout][R*S*Cin] bf16 (K-major), fp32 accumulation in TMEM
//   * one CTA = one output tile of TH x TW = 5 x 25 = 125 pixels (M = 128 rows of a UMMA tile;
//     200, 100, 50 and 25 are all multiples of 25 and 5) times N_TILE output channels
//   * per K block (one filter tap x 64 input channels) TMA loads the SHIFTED activation box
//     {64 ch, TW, TH, 1} (hardware zero fill = the convolution padding; elementStrides = the
//     convolution stride) and the weight box {64, N_TILE}, both 128B-swizzled, K-major
//   * warp-specialised persistent kernel: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
//     (and TMEM owner), warps 2-5 = epilogue (tcgen05.ld -> BN scale/shift -> +residual ->
//     ReLU -> bf16 -> global), two TMEM accumulator stages so the epilogue of tile i
//     overlaps the MMAs of tile i+1
//   * fused epilogues: FPN lateral "+ up2(x)" dual output, nearest-upsample replicate into a
//     channel slice of the concat buffer, and the whole DB head tail
//     (conv-transpose 2x2 + BN + ReLU + conv-transpose 2x2 + sigmoid + binarize).
//
// reference: model.rs:4-12, 40-55, 75-105, 126-150.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

constexpr int TC_TW = 25, TC_TH = 5, TC_ROWS = TC_TW * TC_TH;  // 125 valid rows of 128
constexpr int TC_A_BYTES = 128 * 128;                           // A stage: 128 rows x 64 bf16
constexpr int TC_A_TX = TC_ROWS * 128;                          // bytes TMA actually writes
// EW epilogue warps (8, or 16 for the store-bound FPN laterals): 4 TMEM lane quarters x EW/4 column parts

// ---------------------------------------------------------------------------------------
// shared-memory carve-up (dynamic part; scale/shift/w2 are static __shared__)
// ---------------------------------------------------------------------------------------
template <int N_TILE, int STAGES, int RING, int EW>
struct TcSmem {
  static constexpr int B_BYTES = N_TILE * 128;
  static constexpr int OFF_B = STAGES * TC_A_BYTES;
  static constexpr int OFF_STG = OFF_B + STAGES * B_BYTES;
  static constexpr int OFF_RING = OFF_STG + EW * 2048;            // per-warp 2 KB staging, then RING x 2 KB per warp (addend prefetch)
  static constexpr int OFF_BAR = OFF_RING + EW * RING * 2048;  // full[STAGES], empty[STAGES], tfull[2], tempty[2]
  static constexpr int OFF_TMEM = OFF_BAR + (2 * STAGES + 8) * 8;  // + hready[2], zfull[2] (EPI_HEAD2)
  static constexpr int TOTAL = OFF_TMEM + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;                 // slack for manual 1024 B alignment
};

template <int N_TILE, int STAGES, int EPI, int RING, int EW>
__global__ void __launch_bounds__((2 + EW) * 32, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p,
               const __grid_constant__ HeadConsts hc) {
  using L = TcSmem<N_TILE, STAGES, RING, EW>;
  constexpr int TC_THREADS = (2 + EW) * 32;
  // DB head tail with 16 epilogue warps: TWO groups of 8, group g owns accumulator stage g and takes every
  // second tile of this CTA — the tail is FP32-issue/latency bound, so two tiles in flight fill the schedulers
  constexpr bool HEAD2 = EPI == EPI_HEAD2 || EPI == EPI_HEAD2_TS;  // tensor-core head tail, second operand in shared memory / in TMEM
  constexpr bool HEAD = EPI == EPI_HEAD || HEAD2;
  // EPI_HEAD2 carve-up: three A stages; the [16][64] weight tile of the second GEMM in the fourth A slot; the conv-transpose-1
  // weights resident in B slot 0; B slots 1..3 + the staging area = eight 16 KB tiles [128 px][64 ch] bf16 (stage x tap) for
  // the A operand of the second GEMM when it is read from shared memory
  constexpr int AS = HEAD2 ? 3 : STAGES;
  static_assert(!HEAD2 || (STAGES == 4 && RING == 0), "EPI_HEAD2 shared-memory carve-up");
  constexpr int GROUPS = (HEAD && EW == 16) ? 2 : 1;
  static_assert(!HEAD2 || (EW == 16 && N_TILE == 256), "tensor-core head tail: two epilogue groups, four taps of 64 channels");
  constexpr int PARTS = EW / 4 / GROUPS;  // column parts
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float s_scale[512], s_shift[512];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;
  uint8_t *sB = smem + L::OFF_B;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::OFF_BAR);
  uint64_t *empty = full + STAGES;
  uint64_t *tfull = empty + STAGES;
  uint64_t *tempty = tfull + 2;
  uint64_t *hready = tempty + 2, *zfull = hready + 2;  // EPI_HEAD2: A operand of the second GEMM written / its result complete
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = (2 * N_TILE <= 32) ? 32 : (2 * N_TILE <= 64) ? 64 : (2 * N_TILE <= 128) ? 128 : (2 * N_TILE <= 256) ? 256 : 512;
  const int num_kb = p.R * p.S * p.cin_chunks;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_m_tiles = tiles_per_img * p.B;
  const int num_tiles = num_m_tiles * p.num_n_tiles;

  // ---- one-time setup ----
  const int n_sc = HEAD ? 0 : (p.Cout < 512 ? p.Cout : 512);
  for (int i = threadIdx.x; i < n_sc; i += TC_THREADS) {
    s_scale[i] = p.scale ? p.scale[i] : 1.0f;
    s_shift[i] = p.shift ? p.shift[i] : 0.0f;
  }
  // rows 125..127 of every A stage are never written by TMA: keep them zero
  for (int i = threadIdx.x; i < STAGES * 3 * 32; i += TC_THREADS) {
    const int s = i / 96, w = i % 96;
    reinterpret_cast<uint32_t *>(sA + s * TC_A_BYTES + TC_ROWS * 128)[w] = 0u;
  }
  if (HEAD2 && threadIdx.x < 128) {
    // B operand of the second GEMM: [16 rows n][64 channels] bf16, K-major, 128B swizzle.  Rows 0..3 = the conv-transpose-2
    // weights of output q = n rounded to bf16, rows 4..7 = what the rounding dropped (the two partial sums are added in the
    // epilogue, so the weights count with 16 significant bits), rows 8..15 = 0 (N = 16 is the narrowest M = 128 shape).
    const int n = threadIdx.x >> 3, j = threadIdx.x & 7;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float f[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float x = hc.w2[(8 * j + 2 * e + h) * 4 + (n & 3)];
        const float hi = __bfloat162float(__float2bfloat16_rn(x));
        f[h] = n < 4 ? hi : (n < 8 ? x - hi : 0.0f);
      }
      w[e] = pack_bf16(f[0], f[1]);
    }
    *reinterpret_cast<uint4 *>(smem + 3 * TC_A_BYTES + n * 128 + ((j ^ (n & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EW / GROUPS); }
    for (int a = 0; a < 2; ++a) { mbar_init(&hready[a], EW / GROUPS); mbar_init(&zfull[a], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_tile = tile / num_m_tiles, m_tile = tile - n_tile * num_m_tiles;
      const int b = m_tile / tiles_per_img, t = m_tile - b * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      const int x_base = tx * TC_TW * p.stride - p.pad, y_base = ty * TC_TH * p.stride - p.pad;
      for (int r = 0; r < p.R; ++r)
        for (int s = 0; s < p.S; ++s)
          for (int ck = 0; ck < p.cin_chunks; ++ck) {
            mbar_wait(&empty[stage], phase ^ 1, p.err, 1);
            // EPI_HEAD2 (one K block per tile): the weight tile is loaded once, into stage 0's slot, and stays there
            const bool load_b = !HEAD2 || tile == (int)blockIdx.x;
            if (elect_one()) {
              mbar_expect_tx(&full[stage], TC_A_TX + (load_b ? L::B_BYTES : 0));
              int a_c = ck * 64;
              if (p.split_nblk) {  // K block -> (term plane, 64-channel chunk) of the split activation tensor (conv_tc.cuh)
                const int chunk = ck / p.split_nblk, j = ck - chunk * p.split_nblk;
                const int plane = p.split_nblk == 3 ? (j == 2 ? 1 : 0) : (j < 3 ? 0 : (j < 5 ? 1 : 2));
                a_c = plane * p.split_cin + chunk * 64;
              }
              tma_load_4d(sA + stage * TC_A_BYTES, &tmA, &full[stage], a_c, x_base + s, y_base + r, b);
              if (load_b) tma_load_2d(sB + stage * L::B_BYTES, &tmB, &full[stage], ((r * p.S + s) * p.cin_chunks + ck) * 64, n_tile * N_TILE);
            }
            __syncwarp();
            if (++stage == AS) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(N_TILE);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (HEAD2) {
      // GEMM 1 of tile i (conv-transpose 1: [125 px][64] x [64][4 taps x 64]) is followed by GEMM 2 of tile i - 1 (per tap
      // [125 px][64 bf16 in TMEM] x [64][16]) as soon as that tile's epilogue group has written its A operand
      constexpr uint32_t idesc2 = make_idesc(16);
      const uint64_t bdesc1 = make_smem_desc(sB);
      const uint64_t bdesc2 = make_smem_desc(smem + 3 * TC_A_BYTES);
      constexpr bool h_in_tmem = EPI == EPI_HEAD2_TS;  // OCRB_HEAD=ts: A operand of the second GEMM in TMEM (128 tensor cycles per MMA: TMEM-read bound)
      auto gemm2 = [&](int j) {
        const int a2 = j & 1;
        mbar_wait(&hready[a2], (uint32_t)((j >> 1) & 1), p.err, 5);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t base = tmem_base + (uint32_t)(a2 * N_TILE);
          if (h_in_tmem) {
            // K step outermost: consecutive MMAs accumulate into different taps' results
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int t = 0; t < 4; ++t)
                umma_bf16_ts(base + (uint32_t)(t * 64 + 32), base + (uint32_t)(t * 64 + 8 * k), bdesc2 + (uint64_t)(2 * k), idesc2, k != 0 ? 1u : 0u);
          } else {
            const uint64_t hdesc = make_smem_desc(sB + L::B_BYTES + a2 * 4 * TC_A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int t = 0; t < 4; ++t)
                umma_bf16(base + (uint32_t)(t * 64 + 32), hdesc + (uint64_t)(t * (TC_A_BYTES >> 4) + 2 * k), bdesc2 + (uint64_t)(2 * k), idesc2, k != 0 ? 1u : 0u);
          }
          umma_commit(&zfull[a2]);
        }
        __syncwarp();
      };
      // Both GEMMs are issued as soon as their own condition holds (polling, no fixed order): GEMM 1 of tile g1 when its
      // accumulator stage is free (the epilogue of tile g1 - 2 has READ its results — the sigmoid and the stores come after
      // the release), GEMM 2 of tile g2 when that tile's operand is written.
      const int n_local = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
      int g1 = 0, g2 = 0;
      long long t0 = clock64();
      while (g2 < n_local) {
        bool progress = false;
        if (g2 < g1 && mbar_test_wait(&hready[g2 & 1], (uint32_t)((g2 >> 1) & 1))) {
          gemm2(g2);
          ++g2;
          progress = true;
        }
        if (g1 < n_local && mbar_test_wait(&tempty[g1 & 1], (uint32_t)(((g1 >> 1) & 1) ^ 1))) {
          mbar_wait(&full[stage], phase, p.err, 3);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sA + stage * TC_A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + (uint32_t)((g1 & 1) * N_TILE), adesc + (uint64_t)(2 * k), bdesc1 + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);
            umma_commit(&tfull[g1 & 1]);
          }
          __syncwarp();
          if (++stage == AS) { stage = 0; phase ^= 1; }
          ++g1;
          progress = true;
        }
        if (progress) {
          t0 = clock64();
        } else if (clock64() - t0 > 4000000000ll) {
          if (p.err) { *reinterpret_cast<volatile int *>(p.err) = 7; __threadfence_system(); }
          __trap();
        }
      }
    }
    for (int tile = blockIdx.x; !HEAD2 && tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1, p.err, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N_TILE);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase, p.err, 3);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(sA + stage * TC_A_BYTES);
        const uint64_t bdesc = make_smem_desc(sB + stage * L::B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (K = 16 bf16 = 32 B) inside the 128 B swizzle atom
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ================= epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = (ew >> 2) % PARTS;  // column part of this warp (0 .. PARTS-1)
    const int group = ew / (EW / GROUPS);
    int acc = GROUPS == 2 ? group : 0;
    uint32_t acc_phase = 0;
    int tile_it = 0;
    EpiParams e;
    e.s_scale = s_scale; e.s_shift = s_shift; e.has_affine = p.scale != nullptr;
    e.addend = p.residual ? p.residual : p.up_src;
    e.add_mode = p.residual ? EPI_ADD_RESIDUAL : (p.sum_out ? EPI_ADD_SUM : EPI_ADD_NONE);
    e.out = p.out; e.sum_out = p.sum_out;
    e.Cout = p.Cout; e.out_ldc = p.out_ldc; e.out_coff = p.out_coff; e.rep = p.rep; e.Wo = p.Wo; e.relu = p.relu;
    const uint32_t stg = smem_u32(smem + L::OFF_STG + ew * 2048);
    if (HEAD2) {
      // DB head tail, contraction on the tensor cores: per tap, BN + ReLU of the 64 conv-transpose-1 channels -> bf16 pairs
      // written back over the accumulator's first 32 columns (lane = pixel, column = channel pair: the TMEM form of an A
      // operand); the issuer warp multiplies them with the [64][16] tile of conv-transpose-2 weights into columns 32..47 of
      // the same tap; this thread then reads its 4 + 4 partial sums, adds the bias, sigmoid, binarize.  Per pixel and tap:
      // 32 packed FMAs + 32 conversions instead of 160 packed FMAs + 64 maxima on a dependent chain.
      // Group g takes the tiles blockIdx.x + (2 i + g) gridDim.x; the tile coordinates advance incrementally (no divisions).
      const int m = quarter * 32 + lane;
      const int yl = m / TC_TW, xl = m - yl * TC_TW;
      const int step = 2 * (int)gridDim.x;
      const int step_b = step / tiles_per_img, step_r = step - step_b * tiles_per_img;
      const int step_y = step_r / p.tiles_x, step_x = step_r - step_y * p.tiles_x;
      int tile = (int)blockIdx.x + group * (int)gridDim.x;
      int b = tile / tiles_per_img, ty, tx;
      { const int t = tile - b * tiles_per_img; ty = t / p.tiles_x; tx = t - ty * p.tiles_x; }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE);
      const uint32_t tap0 = taddr + (uint32_t)(half * 128);  // taps i*2 + j with i = half: two rows of the pixel's 4x4 output block
      const int64_t Wp = (int64_t)p.Wo * 4;
      constexpr bool h_in_tmem = EPI == EPI_HEAD2_TS;
      const uint32_t hrow = smem_u32(sB + L::B_BYTES + (acc * 4 + half * 2) * TC_A_BYTES + m * 128);
      // 16 channels of one tap: raw accumulator registers -> 8 packed bf16 pairs
      auto bn_relu16 = [&](const uint32_t (&r)[16], int co0, uint32_t *h) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const int co = co0 + j;
          const float2 f = ffma2(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), make_float2(hc.scale[co], hc.scale[co + 1]),
                                 make_float2(hc.shift[co], hc.shift[co + 1]));
          h[j >> 1] = pack_bf16_relu(f.x, f.y);
        }
      };
      for (; tile < num_tiles; tile += step) {
        const int y = ty * TC_TH + yl, x = tx * TC_TW + xl;
        const bool valid = m < TC_ROWS && y < p.Ho && x < p.Wo;
        mbar_wait(&tfull[acc], acc_phase, p.err, 4);
        tc_fence_after();
        // the loads of the second tap fly while the first is converted
        uint32_t ra[4][16], rb[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(tap0 + (uint32_t)(c * 16), ra[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_wait16(ra[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld16_nowait(tap0 + (uint32_t)(64 + c * 16), rb[c]);
        if (h_in_tmem) {
          {
            uint32_t h[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) bn_relu16(ra[c], c * 16, h + c * 8);
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_wait16(rb[c]);
            tmem_st32(tap0, h);  // every column of this tap is in registers: the pairs land on its first 32 columns
          }
          {
            uint32_t h[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) bn_relu16(rb[c], c * 16, h + c * 8);
            tmem_st32(tap0 + 64, h);
          }
          tmem_st_wait();
        } else {
          // this pixel's row of the [128][64] bf16 tile of each tap: eight 16-byte chunks, 128B-swizzled like a TMA-written tile
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t h[8];
            bn_relu16(ra[c], c * 16, h);
            sts_16(hrow + (uint32_t)(((2 * c) ^ (m & 7)) << 4), make_uint4(h[0], h[1], h[2], h[3]));
            sts_16(hrow + (uint32_t)(((2 * c + 1) ^ (m & 7)) << 4), make_uint4(h[4], h[5], h[6], h[7]));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_wait16(rb[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t h[8];
            bn_relu16(rb[c], c * 16, h);
            sts_16(hrow + TC_A_BYTES + (uint32_t)(((2 * c) ^ (m & 7)) << 4), make_uint4(h[0], h[1], h[2], h[3]));
            sts_16(hrow + TC_A_BYTES + (uint32_t)(((2 * c + 1) ^ (m & 7)) << 4), make_uint4(h[4], h[5], h[6], h[7]));
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's operand reads
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hready[acc]);
        mbar_wait(&zfull[acc], acc_phase, p.err, 6);
        tc_fence_after();
        float z[2][8];
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) tmem_ld8(tap0 + (uint32_t)(tj * 64 + 32), z[tj]);
        // the accumulator stage is free once its results are in registers: release it before the sigmoid and the stores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        acc_phase ^= 1;
        float o[2][4];
#pragma unroll
        for (int tj = 0; tj < 2; ++tj) {
          // q = i'*2 + j' of conv-transpose 2: output (4y + 2i + i', 4x + 2j + j')
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q >> 1][2 * tj + (q & 1)] = __fdividef(1.0f, 1.0f + __expf(-(z[tj][q] + z[tj][4 + q] + p.b2)));
        }
        if (valid) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int64_t off = ((int64_t)b * p.Ho * 4 + (int64_t)y * 4 + 2 * half + a) * Wp + (int64_t)x * 4;
            *reinterpret_cast<float4 *>(p.prob + off) = make_float4(o[a][0], o[a][1], o[a][2], o[a][3]);
            if (p.bitmap) {
              uint32_t bits = (o[a][0] > p.thresh ? 1u : 0u) | (o[a][1] > p.thresh ? 0x100u : 0u) |
                              (o[a][2] > p.thresh ? 0x10000u : 0u) | (o[a][3] > p.thresh ? 0x1000000u : 0u);
              *reinterpret_cast<uint32_t *>(p.bitmap + off) = bits;
            }
          }
        }
        tx += step_x;
        if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
        ty += step_y;
        if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
        b += step_b;
      }
    }
    for (int tile = blockIdx.x; !HEAD2 && tile < num_tiles; tile += gridDim.x, ++tile_it) {
      if (GROUPS == 2 && (tile_it & 1) != group) continue;
      const int n_tile = tile / num_m_tiles, m_tile = tile - n_tile * num_m_tiles;
      const int b = m_tile / tiles_per_img, t = m_tile - b * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      if (EPI == EPI_STD) {
        constexpr int NH = N_TILE / PARTS, NBLK = NH / 32;
        const int n0 = n_tile * N_TILE + half * NH;
        auto rows_of_tile = [&](int tl, EpiRows &rw) {
          const int nt = tl / num_m_tiles, mt = tl - nt * num_m_tiles;
          const int bb = mt / tiles_per_img, tt = mt - bb * tiles_per_img;
          const int tyy = tt / p.tiles_x, txx = tt - tyy * p.tiles_x;
          rw.valid = 0;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int m = quarter * 32 + it * 8 + (lane >> 2);
            const int yl = m / TC_TW, xl = m - yl * TC_TW;
            const int y = tyy * TC_TH + yl, x = txx * TC_TW + xl;
            rw.opix[it] = (bb * p.Ho + y) * p.Wo + x;
            rw.apix[it] = p.sum_out ? (bb * (p.Ho >> 1) + (y >> 1)) * (p.Wo >> 1) + (x >> 1) : rw.opix[it];
            if (m < TC_ROWS && y < p.Ho && x < p.Wo) rw.valid |= 1u << it;
          }
        };
        EpiRows rw;
        rows_of_tile(tile, rw);
        constexpr bool USE_RING = RING >= NBLK && RING > 0;
        const uint32_t ring = smem_u32(smem + L::OFF_RING + ew * (RING * 2048));
        uint4 pre[4] = {};
        if (e.add_mode != EPI_ADD_NONE) {
          if (USE_RING) {
            if (tile == (int)blockIdx.x) {  // first tile of this CTA: nothing was prefetched yet
#pragma unroll
              for (int blk = 0; blk < NBLK; ++blk) {
                epi_prefetch_addend(ring + blk * 2048, lane, rw, e.addend, e.Cout, n0 + blk * 32);
                cp_async_commit_group();
              }
            }
          } else {
            epi_fetch_addend(pre, lane, rw, e.addend, e.Cout, n0);
          }
        }
        EpiRows rw_next;
        const int next_tile = tile + gridDim.x;
        if (USE_RING && e.add_mode != EPI_ADD_NONE && next_tile < num_tiles) rows_of_tile(next_tile, rw_next);
        mbar_wait(&tfull[acc], acc_phase, p.err, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE + half * NH);
#pragma unroll 1
        for (int blk = 0; blk < NBLK; ++blk) {
          float v[32];
          tmem_ld32(taddr + blk * 32, v);
          if (USE_RING && e.add_mode != EPI_ADD_NONE) {
            // this block's addend was prefetched a whole tile ago; NBLK - 1 younger groups may still fly
            cp_async_wait_group<(NBLK > 0 ? NBLK - 1 : 0)>();
            __syncwarp();
            epi_block32(v, lane, stg, rw, e, n0 + blk * 32, pre, ring + blk * 2048);
            // slot free again: prefetch the same block of the next tile (an empty group keeps the count uniform)
            if (next_tile < num_tiles) {
              const int nn0 = (next_tile / num_m_tiles) * N_TILE + half * NH;
              epi_prefetch_addend(ring + blk * 2048, lane, rw_next, e.addend, e.Cout, nn0 + blk * 32);
            }
            cp_async_commit_group();
          } else {
            uint4 cur[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) cur[it] = pre[it];
            if (e.add_mode != EPI_ADD_NONE && blk + 1 < NBLK) epi_fetch_addend(pre, lane, rw, e.addend, e.Cout, n0 + (blk + 1) * 32);
            epi_block32(v, lane, stg, rw, e, n0 + blk * 32, cur);
          }
        }
      } else if (EPI == EPI_F32) {
        // fp32 in, fp32 out: a thread owns one pixel (TMEM lane) and this warp's share of the channels
        constexpr int NH = N_TILE / PARTS;
        const int n0 = n_tile * N_TILE + half * NH;
        const int m = quarter * 32 + lane;
        const int yl = m / TC_TW, xl = m - yl * TC_TW;
        const int y = ty * TC_TH + yl, x = tx * TC_TW + xl;
        const bool valid = m < TC_ROWS && y < p.Ho && x < p.Wo;
        const int64_t pix = ((int64_t)b * p.Ho + y) * p.Wo + x;
        mbar_wait(&tfull[acc], acc_phase, p.err, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE + half * NH);
#pragma unroll 1
        for (int blk = 0; blk < NH / 32; ++blk) {
          float v[32];
          tmem_ld32(taddr + blk * 32, v);
          if (valid) {
            const int n = n0 + blk * 32;
            float *o = p.out32 + pix * p.Cout + n;
            const float *rr = p.res32 ? p.res32 + pix * p.Cout + n : nullptr;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 r4 = rr ? *reinterpret_cast<const float4 *>(rr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              float4 y4;
              y4.x = v[j + 0] * s_scale[n + j + 0] + s_shift[n + j + 0] + r4.x;
              y4.y = v[j + 1] * s_scale[n + j + 1] + s_shift[n + j + 1] + r4.y;
              y4.z = v[j + 2] * s_scale[n + j + 2] + s_shift[n + j + 2] + r4.z;
              y4.w = v[j + 3] * s_scale[n + j + 3] + s_shift[n + j + 3] + r4.w;
              if (p.relu) { y4.x = fmaxf(y4.x, 0.f); y4.y = fmaxf(y4.y, 0.f); y4.z = fmaxf(y4.z, 0.f); y4.w = fmaxf(y4.w, 0.f); }
              *reinterpret_cast<float4 *>(o + j) = y4;
            }
          }
        }
      } else {
        static_assert(EPI != EPI_HEAD || PARTS == 2, "head epilogue splits the four taps over two warp halves");
        // DB head tail: columns n = tap(i,j)*64 + co of conv-transpose 1; per tap BN+ReLU then
        // the 64 -> 4 dot products of conv-transpose 2, sigmoid; each warp half takes two taps
        // (= two rows of the pixel's 4x4 output block).
        const int m = quarter * 32 + lane;
        const int yl = m / TC_TW, xl = m - yl * TC_TW;
        const int y = ty * TC_TH + yl, x = tx * TC_TW + xl;
        const bool valid = m < TC_ROWS && y < p.Ho && x < p.Wo;
        mbar_wait(&tfull[acc], acc_phase, p.err, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * N_TILE);
        float o[2][4];
        {
          // both taps of this warp half go through the channel loop together: every BN / conv-transpose-2 constant
          // (a uniform-register load per use) serves two taps, and eight independent accumulator chains hide the FMA latency.
          // packed fp32x2 FMAs (FFMA2): two channels at a time through BN, two outputs at a time through the 64 -> 4 contraction
          float2 z01[2], z23[2];
#pragma unroll
          for (int tj = 0; tj < 2; ++tj) { z01[tj] = make_float2(p.b2, p.b2); z23[tj] = make_float2(p.b2, p.b2); }
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[2][32];
            tmem_ld32(taddr + (half * 2 + 0) * 64 + c0, v[0]);  // tap = i*2 + j with i = half
            tmem_ld32(taddr + (half * 2 + 1) * 64 + c0, v[1]);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int co = c0 + j;
              const float2 sc = make_float2(hc.scale[co], hc.scale[co + 1]), sh = make_float2(hc.shift[co], hc.shift[co + 1]);
              const float2 wa01 = make_float2(hc.w2[co * 4 + 0], hc.w2[co * 4 + 1]), wa23 = make_float2(hc.w2[co * 4 + 2], hc.w2[co * 4 + 3]);
              const float2 wb01 = make_float2(hc.w2[co * 4 + 4], hc.w2[co * 4 + 5]), wb23 = make_float2(hc.w2[co * 4 + 6], hc.w2[co * 4 + 7]);
#pragma unroll
              for (int tj = 0; tj < 2; ++tj) {
                const float2 h = ffma2(make_float2(v[tj][j], v[tj][j + 1]), sc, sh);
                const float h0 = fmaxf(h.x, 0.0f), h1 = fmaxf(h.y, 0.0f);
                z01[tj] = ffma2(make_float2(h0, h0), wa01, z01[tj]);
                z23[tj] = ffma2(make_float2(h0, h0), wa23, z23[tj]);
                z01[tj] = ffma2(make_float2(h1, h1), wb01, z01[tj]);
                z23[tj] = ffma2(make_float2(h1, h1), wb23, z23[tj]);
              }
            }
          }
#pragma unroll
          for (int tj = 0; tj < 2; ++tj) {
            const float z[4] = {z01[tj].x, z01[tj].y, z23[tj].x, z23[tj].y};
            // q = i'*2 + j' of conv-transpose 2: output (4y + 2i + i', 4x + 2j + j')
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q >> 1][2 * tj + (q & 1)] = 1.0f / (1.0f + expf(-z[q]));
          }
        }
        if (valid) {
          const int64_t Wp = (int64_t)p.Wo * 4;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int64_t off = ((int64_t)b * p.Ho * 4 + (int64_t)y * 4 + 2 * half + a) * Wp + (int64_t)x * 4;
            *reinterpret_cast<float4 *>(p.prob + off) = make_float4(o[a][0], o[a][1], o[a][2], o[a][3]);
            if (p.bitmap) {
              uint32_t bits = (o[a][0] > p.thresh ? 1u : 0u) | (o[a][1] > p.thresh ? 0x100u : 0u) |
                              (o[a][2] > p.thresh ? 0x10000u : 0u) | (o[a][3] > p.thresh ? 0x1000000u : 0u);
              *reinterpret_cast<uint32_t *>(p.bitmap + off) = bits;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (GROUPS == 2) {
        acc_phase ^= 1;  // this group's stage is used by every second tile
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// activations NHWC bf16 [B][H][W][C] -> 4-D map {C, W, H, B}, box {64, TW*stride, TH*stride, 1}
int make_act_tensor_map(CUtensorMap *map, const void *base, int B, int H, int W, int C, int stride) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return OCRB_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(TC_TW * stride), (cuuint32_t)(TC_TH * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act %dx%dx%dx%d s%d) -> %d", B, H, W, C, stride, (int)r); return OCRB_ERR_CUDA; }
  return OCRB_OK;
}

// activations NHWC bf16 -> 4-D map {C, W, H, B} with an arbitrary (box_w x box_h)-pixel box of 64 channels
int make_act_tensor_map_box(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return OCRB_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act box %dx%d of %dx%dx%dx%d) -> %d", box_w, box_h, B, H, W, C, (int)r); return OCRB_ERR_CUDA; }
  return OCRB_OK;
}

// the same with box_b images per box (glyph batches of the recognition net: rec_tc.cu)
int make_act_tensor_map_box_b(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h, int box_b) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return OCRB_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_b};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act box %dx%dx%d of %dx%dx%dx%d) -> %d", box_w, box_h, box_b, B, H, W, C, (int)r); return OCRB_ERR_CUDA; }
  return OCRB_OK;
}

// a C-channel slice of pixels that are ldc channels apart ([B][H][W][ldc], base already at the slice): TMA store / residual load box
// step > 1: the H x W positions sit on every step-th pixel and row of a (H*step) x (W*step) map; row_px > 0: pixels per
// (full-resolution) row of the buffer when its rows are padded
int make_act_tensor_map_pitched(CUtensorMap *map, const void *base, int B, int H, int W, int C, int ldc, int box_w, int box_h, int step, int row_px) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return OCRB_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t row = (cuuint64_t)(row_px > 0 ? row_px : W * step) * ldc * 2;  // bytes of one full-resolution row
  cuuint64_t strides[3] = {(cuuint64_t)ldc * 2 * step, row * step, row * (cuuint64_t)H * step};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(pitched box %dx%d, ldc %d) -> %d", box_w, box_h, ldc, (int)r); return OCRB_ERR_CUDA; }
  return OCRB_OK;
}

// activations NHWC bf16 sampled every `stride` pixels: box of box_w x box_h SAMPLED positions
int make_act_tensor_map_strided_box(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h, int stride) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return OCRB_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(box_w * stride), (cuuint32_t)(box_h * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  if (box[1] > 256 || box[2] > 256) { set_error("strided TMA box %ux%u exceeds 256", box[1], box[2]); return OCRB_ERR_INVALID; }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(strided box) -> %d", (int)r); return OCRB_ERR_CUDA; }
  return OCRB_OK;
}

// weights [Cout][Ktot] bf16 -> 2-D map {Ktot, Cout}, box {64, n_tile}
int make_weight_tensor_map(CUtensorMap *map, const void *base, int Cout, int Ktot, int n_tile) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return OCRB_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)n_tile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights %dx%d) -> %d", Cout, Ktot, (int)r); return OCRB_ERR_CUDA; }
  return OCRB_OK;
}

template <int N_TILE, int STAGES, int EPI, int RING, int EW>
static int launch_one(ocrb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const ConvTcParams &p, int num_tiles, const char *tag,
                      const HeadConsts &hc) {
  using L = TcSmem<N_TILE, STAGES, RING, EW>;
  auto kern = conv_tc_kernel<N_TILE, STAGES, EPI, RING, EW>;
  OCRB_TRY(ensure_dyn_smem(ctx, kern, L::DYN_BYTES));
  int grid = num_tiles < ctx->sm_budget() ? num_tiles : ctx->sm_budget();
  kern<<<grid, (2 + EW) * 32, L::DYN_BYTES, ctx->stream>>>(tmA, tmB, p, hc);
  return check_launch(ctx, tag);
}

int launch_conv_tc(ocrb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, ConvTcParams p, int n_tile, int epi, const char *tag,
                   const HeadConsts *hcp) {
  static const HeadConsts hc_zero = {};
  const HeadConsts &hc = hcp ? *hcp : hc_zero;
  p.tiles_x = (int)cdiv(p.Wo, TC_TW);
  p.tiles_y = (int)cdiv(p.Ho, TC_TH);
  p.num_n_tiles = p.Cout / n_tile;
  if (p.Cout % n_tile != 0 || p.Cout > 512) { set_error("conv_tc: Cout %d not a multiple of the N tile %d (or > 512)", p.Cout, n_tile); return OCRB_ERR_INVALID; }
  const int num_tiles = p.tiles_x * p.tiles_y * p.B * p.num_n_tiles;
  if (epi == EPI_HEAD) {
    if (n_tile != 256) { set_error("conv_tc head needs N tile 256"); return OCRB_ERR_INVALID; }
    // default: the 64 -> 4 contraction of the tail as a second GEMM (OCRB_HEAD=cuda: on the CUDA cores, round 1's form)
    static const bool head_cuda = getenv("OCRB_HEAD") && !strcmp(getenv("OCRB_HEAD"), "cuda");
    static const bool head_ts = !(getenv("OCRB_HEAD") && !strcmp(getenv("OCRB_HEAD"), "ss"));  // measured: 3.6 ms per 1024 images (ss: 4.3, cuda: 5.0)
    if (!head_cuda && p.R == 1 && p.S == 1 && p.cin_chunks == 1 && p.num_n_tiles == 1 && !p.split_nblk)
      return head_ts ? launch_one<256, 4, EPI_HEAD2_TS, 0, 16>(ctx, tmA, tmB, p, num_tiles, tag, hc)
                     : launch_one<256, 4, EPI_HEAD2, 0, 16>(ctx, tmA, tmB, p, num_tiles, tag, hc);
    static const bool ew8 = getenv("OCRB_HEAD_EW") && atoi(getenv("OCRB_HEAD_EW")) == 8;  // tuning knob
    return ew8 ? launch_one<256, 4, EPI_HEAD, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc)
               : launch_one<256, 4, EPI_HEAD, 0, 16>(ctx, tmA, tmB, p, num_tiles, tag, hc);
  }
  if (epi == EPI_F32) {
    switch (n_tile) {
      case 64: return launch_one<64, 6, EPI_F32, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
      case 128: return launch_one<128, 5, EPI_F32, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
      case 256: return launch_one<256, 4, EPI_F32, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
    }
    set_error("conv_tc: unsupported N tile %d", n_tile);
    return OCRB_ERR_INVALID;
  }
  switch (n_tile) {
    case 64: return launch_one<64, 6, EPI_STD, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
    case 128: return launch_one<128, 5, EPI_STD, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
    case 256:
      // FPN laterals (second output = y + up2(addend)): 1x1 convs with few K blocks, so two
      // operand stages suffice and the shared memory goes to the addend prefetch ring instead
      if (p.sum_out) {
        static const bool ew16 = getenv("OCRB_LATERAL_EW") && atoi(getenv("OCRB_LATERAL_EW")) == 16;  // tuning knob: measured slower (in2 11.4 vs 10.0 ms / 1024 images)
        return ew16 ? launch_one<256, 2, EPI_STD, 2, 16>(ctx, tmA, tmB, p, num_tiles, tag, hc)
                    : launch_one<256, 2, EPI_STD, 4, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
      }
      return launch_one<256, 4, EPI_STD, 0, 8>(ctx, tmA, tmB, p, num_tiles, tag, hc);
  }
  set_error("conv_tc: unsupported N tile %d", n_tile);
  return OCRB_ERR_INVALID;
}

}  // namespace ocrb
