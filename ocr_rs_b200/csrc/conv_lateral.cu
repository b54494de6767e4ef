// FPN lateral 1x1 convolutions with the top-down sum fused (model.rs:126-137):
//     lat = conv1x1(feat)                 [B][Ho][Wo][256]   (raw lateral, optional output)
//     sum = lat + up2(upper)              upper = the raw lateral of the level above, [B][Ho/2][Wo/2][256]
// These layers are pure data movement (K = 64 or 128 against 256 output channels: 64 KB written per
// 16-32 KB read), so the kernel is organised around the memory system, not the tensor pipe:
//   * the whole weight matrix [256][K] stays resident in shared memory (one TMA at kernel start);
//   * one CTA tile = one box {64 ch, 32 px, 4 rows} = 128 GEMM rows, every row a real pixel;
//   * the addend arrives by TMA as the half-resolution box {64 ch, 16, 2} per channel group — a
//     thread reads the row of its 2x2 parent, no per-thread global loads;
//   * outputs leave through TMA: epilogue threads write their own pixel row (bf16, 128B swizzle)
//     into a ring of 16 KB staging tiles, one extra warp issues cp.async.bulk.tensor stores
//     (image borders clipped by the hardware).
// Roles: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer + TMEM owner, warps 2..9 = epilogue
// (4 TMEM lane quarters x 2 halves of the 256 channels), warp 10 = store warp.
// The general engine (conv_tc.cu, per-thread coalesced stores) keeps the laterals this kernel
// does not take (K > 128, odd map sizes).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

constexpr int LT_TW = 32, LT_TH = 4;          // tile = 128 pixels
constexpr int LT_A_BYTES = 128 * 128;         // one 64-channel chunk of a tile
constexpr int LT_W_BYTES = 256 * 128;         // one 64-channel chunk of the weights
constexpr int LT_ADD_BYTES = 4 * 32 * 128;    // addend of a tile: 4 channel groups x (16 x 2) parents
constexpr int LT_OB_BYTES = 128 * 128;        // one staging tile: 128 pixels x 64 channels
constexpr int LT_NOB = 4;                     // staging ring (positions of a tile map to fixed slots)
constexpr int LT_ADD_STAGES = 2;
constexpr int LT_THREADS = 11 * 32;

struct LateralParams {
  int B, Ho, Wo, chunks;  // chunks = Cin / 64 (1 or 2)
  int tiles_x, tiles_y;
  int a_stages;           // tiles of A in flight (each chunks x 16 KB)
  int has_out;            // raw lateral stored too
  int *err;
};

__global__ void __launch_bounds__(LT_THREADS, 1)
conv_lateral_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmAdd,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmSum, const LateralParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_tile_bytes = p.chunks * LT_A_BYTES;
  uint8_t *sW = smem;
  uint8_t *sA = sW + p.chunks * LT_W_BYTES;
  uint8_t *sAdd = sA + p.a_stages * a_tile_bytes;
  uint8_t *sOb = sAdd + LT_ADD_STAGES * LT_ADD_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sOb + LT_NOB * LT_OB_BYTES);
  uint64_t *w_full = bars;
  uint64_t *a_full = w_full + 1, *a_empty = a_full + 4;          // up to 4 A stages
  uint64_t *add_full = a_empty + 4, *add_empty = add_full + LT_ADD_STAGES;
  uint64_t *tfull = add_empty + LT_ADD_STAGES, *tempty = tfull + 2;
  uint64_t *o_ready = tempty + 2, *o_done = o_ready + LT_NOB;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_done + LT_NOB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = tiles_per_img * p.B;
  const int nk = p.has_out ? 2 : 1;  // staging tiles per channel group
  const int NP = 4 * nk;             // staging tiles per CTA tile; position = (j*2 + half)*nk + kind

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmAdd);
    tma_prefetch_desc(&tmSum);
    if (p.has_out) tma_prefetch_desc(&tmOut);
    mbar_init(w_full, 1);
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < LT_ADD_STAGES; ++s) { mbar_init(&add_full[s], 1); mbar_init(&add_empty[s], 8); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    for (int s = 0; s < LT_NOB; ++s) { mbar_init(&o_ready[s], 1); mbar_init(&o_done[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int &b, int &y0, int &x0) {
    b = tile / tiles_per_img;
    const int t = tile - b * tiles_per_img;
    const int ty = t / p.tiles_x;
    y0 = ty * LT_TH;
    x0 = (t - ty * p.tiles_x) * LT_TW;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      mbar_expect_tx(w_full, p.chunks * LT_W_BYTES);
      for (int ck = 0; ck < p.chunks; ++ck) tma_load_2d(sW + ck * LT_W_BYTES, &tmW, w_full, ck * 64, 0);
    }
    __syncwarp();
    int as = 0, ds = 0;
    uint32_t aph = 0, dph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      mbar_wait(&a_empty[as], aph ^ 1, p.err, 31);
      if (elect_one()) {
        mbar_expect_tx(&a_full[as], a_tile_bytes);
        for (int ck = 0; ck < p.chunks; ++ck) tma_load_4d(sA + as * a_tile_bytes + ck * LT_A_BYTES, &tmA, &a_full[as], ck * 64, x0, y0, b);
      }
      __syncwarp();
      if (++as == p.a_stages) { as = 0; aph ^= 1; }
      mbar_wait(&add_empty[ds], dph ^ 1, p.err, 32);
      if (elect_one()) {
        mbar_expect_tx(&add_full[ds], LT_ADD_BYTES);
        for (int gq = 0; gq < 4; ++gq) tma_load_4d(sAdd + ds * LT_ADD_BYTES + gq * 4096, &tmAdd, &add_full[ds], gq * 64, x0 >> 1, y0 >> 1, b);
      }
      __syncwarp();
      if (++ds == LT_ADD_STAGES) { ds = 0; dph ^= 1; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(256);
    mbar_wait(w_full, 0, p.err, 33);
    int as = 0, acc = 0;
    uint32_t aph = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1, p.err, 34);
      mbar_wait(&a_full[as], aph, p.err, 35);
      tc_fence_after();
      if (elect_one()) {
        for (int ck = 0; ck < p.chunks; ++ck) {
          const uint64_t adesc = make_smem_desc(sA + as * a_tile_bytes + ck * LT_A_BYTES);
          const uint64_t bdesc = make_smem_desc(sW + ck * LT_W_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + (uint32_t)(acc * 256), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ck | k) != 0 ? 1u : 0u);
        }
        umma_commit(&a_empty[as]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++as == p.a_stages) { as = 0; aph ^= 1; }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp == 10) {
    // ================= store warp (one thread: bulk groups are per thread) =================
    if (lane == 0) {
      for (int s = 0; s < LT_NOB; ++s) mbar_arrive(&o_ready[s]);  // all staging tiles start free
      int cnt = 0, prev = -1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int b, y0, x0;
        tile_coords(tile, b, y0, x0);
        for (int pos = 0; pos < NP; ++pos, ++cnt) {
          const int slot = cnt & (LT_NOB - 1);
          mbar_wait(&o_done[slot], (uint32_t)(cnt / LT_NOB) & 1u, p.err, 36);
          const int kind = pos % nk, jh = pos / nk;
          const int grp = (jh & 1) * 2 + (jh >> 1);  // jh = j*2 + half, group = half*2 + j
          const bool is_sum = kind == nk - 1;
          tma_store_4d(is_sum ? &tmSum : &tmOut, sOb + slot * LT_OB_BYTES, grp * 64, x0, y0, b);
          bulk_commit_group();
          if (prev >= 0) {
            bulk_wait_group_read<1>();
            mbar_arrive(&o_ready[prev]);
          }
          prev = slot;
        }
      }
      bulk_wait_group<0>();
    }
  } else {
    // ================= epilogue =================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // channel groups 2*half, 2*half + 1
    const int m = quarter * 32 + lane;                     // pixel (yl = quarter, xl = lane) of the tile
    const int arow = (quarter >> 1) * 16 + (lane >> 1);    // its 2x2 parent in the {16 x 2} addend box
    int acc = 0, ds = 0, t_local = 0;
    uint32_t acc_phase = 0, dph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_local) {
      mbar_wait(&add_full[ds], dph, p.err, 37);
      mbar_wait(&tfull[acc], acc_phase, p.err, 38);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        const int grp = half * 2 + j;
        const int cnt0 = t_local * NP + (j * 2 + half) * nk;  // first staging position of this group
        const int slot_sum = (cnt0 + nk - 1) & (LT_NOB - 1), slot_out = cnt0 & (LT_NOB - 1);
        mbar_wait(&o_ready[slot_sum], (uint32_t)((cnt0 + nk - 1) / LT_NOB) & 1u, p.err, 39);
        if (p.has_out) mbar_wait(&o_ready[slot_out], (uint32_t)(cnt0 / LT_NOB) & 1u, p.err, 40);
        const uint32_t t_sum = smem_u32(sOb + slot_sum * LT_OB_BYTES), t_out = smem_u32(sOb + slot_out * LT_OB_BYTES);
        const uint32_t t_add = smem_u32(sAdd + ds * LT_ADD_BYTES + grp * 4096);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + grp * 64 + blk * 32), v);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 a = lds_16(t_add + (uint32_t)(arow * 128 + (((4 * blk + c) ^ (arow & 7)) << 4)));
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
            uint32_t yw[4], sw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              yw[q] = pack_bf16(v[c * 8 + 2 * q], v[c * 8 + 2 * q + 1]);  // the lateral as it is (or would be) materialised
              const float2 y = unpack_bf16(yw[q]), ad = unpack_bf16(aw[q]);
              sw[q] = pack_bf16(y.x + ad.x, y.y + ad.y);
            }
            const uint32_t off = (uint32_t)(m * 128 + (((4 * blk + c) ^ (m & 7)) << 4));
            sts_16(t_sum + off, make_uint4(sw[0], sw[1], sw[2], sw[3]));
            if (p.has_out) sts_16(t_out + off, make_uint4(yw[0], yw[1], yw[2], yw[3]));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&o_done[slot_sum]);
          if (p.has_out) mbar_arrive(&o_done[slot_out]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tempty[acc]);
        mbar_arrive(&add_empty[ds]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      if (++ds == LT_ADD_STAGES) { ds = 0; dph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int make_act_tensor_map_box(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h);

bool lateral_ts_supported(int Cin, int Cout, int Ho, int Wo) {
  static const bool on = !(getenv("OCRB_LATERAL_TS") && atoi(getenv("OCRB_LATERAL_TS")) == 0);
  return on && Cout == 256 && (Cin == 64 || Cin == 128) && Ho % 2 == 0 && Wo % 2 == 0;
}

// in [B][Ho][Wo][Cin], w [256][Cin] (K-major bf16), upper [B][Ho/2][Wo/2][256], out (optional) and sum [B][Ho][Wo][256]
int launch_conv_lateral(ocrb_ctx *ctx, const __nv_bfloat16 *in, const __nv_bfloat16 *w, const __nv_bfloat16 *upper, __nv_bfloat16 *out,
                        __nv_bfloat16 *sum, int B, int Ho, int Wo, int Cin, int *err, const char *tag) {
  if (!lateral_ts_supported(Cin, 256, Ho, Wo) || !in || !w || !upper || !sum) { set_error("conv_lateral: unsupported arguments"); return OCRB_ERR_INVALID; }
  LateralParams p;
  p.B = B; p.Ho = Ho; p.Wo = Wo; p.chunks = Cin / 64;
  p.tiles_x = (int)cdiv(Wo, LT_TW);
  p.tiles_y = (int)cdiv(Ho, LT_TH);
  p.has_out = out != nullptr;
  p.err = err;
  p.a_stages = p.chunks == 1 ? 4 : 2;
  CUtensorMap tmA, tmW, tmAdd, tmOut, tmSum;
  OCRB_TRY(make_act_tensor_map_box(&tmA, in, B, Ho, Wo, Cin, LT_TW, LT_TH));
  OCRB_TRY(make_weight_tensor_map(&tmW, w, 256, Cin, 256));
  OCRB_TRY(make_act_tensor_map_box(&tmAdd, upper, B, Ho / 2, Wo / 2, 256, LT_TW / 2, LT_TH / 2));
  OCRB_TRY(make_act_tensor_map_box(&tmSum, sum, B, Ho, Wo, 256, LT_TW, LT_TH));
  if (out) OCRB_TRY(make_act_tensor_map_box(&tmOut, out, B, Ho, Wo, 256, LT_TW, LT_TH));
  else tmOut = tmSum;
  const int smem = 1024 + p.chunks * LT_W_BYTES + p.a_stages * p.chunks * LT_A_BYTES + LT_ADD_STAGES * LT_ADD_BYTES + LT_NOB * LT_OB_BYTES + 512;
  OCRB_TRY(ensure_dyn_smem(ctx, conv_lateral_kernel, 227 * 1024));
  const int num_tiles = p.tiles_x * p.tiles_y * B;
  const int grid = num_tiles < ctx->sm_budget() ? num_tiles : ctx->sm_budget();
  conv_lateral_kernel<<<grid, LT_THREADS, smem, ctx->stream>>>(tmA, tmW, tmAdd, tmOut, tmSum, p);
  return check_launch(ctx, tag);
}

}  // namespace ocrb
