// 3x3 stride-1 convolution (model.rs:4-12 with k=3, s=1, p=1) as an implicit GEMM on tcgen05
// with the input tile RESIDENT in shared memory across the nine filter taps.
//
// Why: with one TMA box per (tap, 64-channel chunk) the activation tile is fetched nine
// times and a 64-wide output tile moves 24 KB through L2 per 128 MMA cycles — the first
// version of the engine (conv_tc.cu) ran those layers at the L2 bandwidth, not the tensor
// pipe.  Here:
//   * per 64-channel chunk ONE TMA box {64 ch, PW, TH+2 rows} (hardware zero fill = padding)
//     lands the halo'd tile as rows of 128 B, row index = y * PW + x (128B swizzle, K-major);
//   * the GEMM M index runs LINEARLY over that padded tile (pitch PW = TW + 2; the two extra
//     columns per row produce junk outputs that are never stored), so the A operand of tap
//     (r, s) is the same tile shifted by (r * PW + s) rows — a pure start-address offset in
//     the UMMA shared-memory descriptor, no data movement;
//   * one CTA work unit = G sub-tiles of 128 rows x N_TILE channels, G accumulators in TMEM
//     (two sets: the epilogue of unit i overlaps the MMAs of unit i+1), so every streamed
//     weight tile [N_TILE x 64] feeds G * 4 MMAs.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue
// (two warps per TMEM lane quarter, each taking half of the channels).
// Epilogue = folded batch-norm scale/shift (+ residual) (+ ReLU) -> bf16 NHWC, optionally
// replicated x rep into a channel slice of the concat buffer (model.rs:82-97, :140).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "tc_epilogue.cuh"
#include "tc_ptx.cuh"

namespace ocrb {

constexpr int HL_EPI_WARPS = 8;
constexpr int HL_STG_BYTES = HL_EPI_WARPS * 2048;  // per-warp epilogue staging (tc_epilogue.cuh)

struct HaloGeom {
  int PW, TH, TW;      // padded pitch, tile rows, valid tile columns (TW = PW - 2)
  int a_stage_bytes;   // bytes of one A stage (multiple of 1024)
  int a_tx_bytes;      // bytes one halo TMA box writes = (TH + 2) * PW * 128
  int a_stages, b_stages;
  // M rows between consecutive sub-tiles: 128 (sub-tiles tile the padded pitch linearly) or, for the
  // TMA-store epilogue, sub_rows * PW <= 128 so that every sub-tile covers WHOLE tile rows
  int sub_rows, sub_stride;
  int obufs, obuf_bytes;  // TMA-store epilogue: ring of output / residual staging tiles [sub_rows * TW][128 B]
};
constexpr int HL_MAX_OBUFS = 4;
// folded batch-norm of a 64-channel convolution, passed by value: the TMA-store epilogue reads it from the constant
// bank (shared memory is the bottleneck of the N = 64 kernels — operand reads alone exceed its bandwidth)
struct HaloAffine { float scale[64], shift[64]; };

// TS = 1: the epilogue leaves through TMA (N_TILE = 64; N_TILE = 128 is the "pair" mode of the fused neck: two
// 64-channel convolutions of one operand tile, one staging tile and one destination map per half — conv_tc.cuh).  Sub-tiles are row-aligned, so a
// sub-tile's valid outputs are one box {64 ch, TW, sub_rows}; the epilogue threads write their own
// pixel row (bf16, 128B-swizzled) into a staging tile and one extra warp turns full tiles into
// cp.async.bulk.tensor stores (image borders clipped by the hardware) and TMA-loads the residual
// tile of a later sub-tile INTO the freed staging tile (read-modify-write in place).  No
// per-thread global loads/stores, no transposing round trip through shared memory.
template <int N_TILE, int G, int CG, int TS>
__global__ void __launch_bounds__((1 + (G >= 4 ? 2 : 1) + HL_EPI_WARPS + TS) * 32, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const ConvTcParams p, const HaloGeom g,
                 const __grid_constant__ HaloAffine ha) {
  // CG = 2: a CTA PAIR works as one unit (tcgen05 cta_group::2).  Each CTA owns a spatial tile
  // (its A operand, its accumulators) and HALF of every weight tile; the leader CTA issues
  // UMMAs of M = 256 that read both halves.  Per CTA that halves the shared-memory reads and
  // the L2 traffic of the B operand — the single-CTA kernel is bound by exactly those reads
  // ((4 KB A + 32*N B) per N/2 tensor cycles: 192 B/cycle at N = 64, 128 B/cycle at N = 128).
  // N = 64 MMAs last only 32 tensor-pipe cycles: two issuing warps (sub-tiles split between
  // them) keep the pipe fed; N = 128 needs one.
  constexpr int MW = G >= 4 ? 2 : 1;
  constexpr int GW = G / MW;  // sub-tiles per MMA warp
  constexpr int THREADS = (1 + MW + HL_EPI_WARPS + TS) * 32;
  static_assert(!TS || N_TILE == 64 || (N_TILE == 128 && G == 2), "TMA-store epilogue: one 128-byte row per pixel (N = 128: pair mode, two rows)");
  constexpr int NB = N_TILE / CG;  // weight rows held by this CTA
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(16) float s_scale[512], s_shift[512];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int B_BYTES = NB * 128;
  uint8_t *sA = smem;
  uint8_t *sB = smem + g.a_stages * g.a_stage_bytes;
  uint8_t *sStg = sB + g.b_stages * B_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStg + (TS ? g.obufs * g.obuf_bytes : HL_STG_BYTES));
  uint64_t *a_full = bars, *a_empty = a_full + g.a_stages;
  uint64_t *b_full = a_empty + g.a_stages, *b_empty = b_full + g.b_stages;
  uint64_t *tfull = b_empty + g.b_stages, *tempty = tfull + 2;
  uint64_t *o_ready = tempty + 2, *o_done = o_ready + HL_MAX_OBUFS;  // TS only
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_done + HL_MAX_OBUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  constexpr uint32_t TMEM_COLS = 2 * G * N_TILE <= 256 ? 256 : 512;
  static_assert(2 * G * N_TILE <= 512, "accumulators do not fit TMEM");
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_m_tiles = tiles_per_img * p.B;
  const int pairs_per_n = (num_m_tiles + CG - 1) / CG;  // CG spatial tiles per work unit
  const int num_units = pairs_per_n * p.num_n_tiles;
  const int unit0 = blockIdx.x / CG, unit_step = gridDim.x / CG;
  const int chunks = p.cin_chunks;

  for (int i = threadIdx.x; i < p.Cout && i < 512; i += THREADS) {
    s_scale[i] = p.scale ? p.scale[i] : 1.0f;
    s_shift[i] = p.shift ? p.shift[i] : 0.0f;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.ds_chunks) tma_prefetch_desc(&tmD);
    // full barriers live in the leader: one producer arrival (+ its bytes) per CTA of the pair
    for (int s = 0; s < g.a_stages; ++s) { mbar_init(&a_full[s], CG); mbar_init(&a_empty[s], MW); }
    for (int s = 0; s < g.b_stages; ++s) { mbar_init(&b_full[s], CG); mbar_init(&b_empty[s], MW); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], MW); mbar_init(&tempty[a], CG * HL_EPI_WARPS); }
    if (TS) {
      tma_prefetch_desc(&tmO);
      if (p.residual) tma_prefetch_desc(&tmR);
      for (int s = 0; s < g.obufs; ++s) { mbar_init(&o_ready[s], 1); mbar_init(&o_done[s], HL_EPI_WARPS); }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (every CTA loads its own tile and its half of B) =================
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      const int n_tile = unit / pairs_per_n, m_tile = (unit - n_tile * pairs_per_n) * CG + (int)rank;
      // a tile index past the end (odd tile count) reads only out-of-bounds zeros
      const int b = m_tile / tiles_per_img, t = m_tile - b * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      for (int ck = 0; ck < chunks; ++ck) {
        mbar_wait(&a_empty[as], aph ^ 1, p.err, 11);
        if (elect_one()) {
          if (CG == 2) {
            const uint32_t bar = mapa_u32(&a_full[as], 0);
            mbar_expect_tx_cluster(bar, g.a_tx_bytes);
            tma_load_4d_2sm(sA + as * g.a_stage_bytes, &tmA, bar, ck * 64, tx * g.TW - 1 + p.in_x_off, ty * g.TH - 1, b);
          } else {
            mbar_expect_tx(&a_full[as], g.a_tx_bytes);
            tma_load_4d(sA + as * g.a_stage_bytes, &tmA, &a_full[as], ck * 64, tx * g.TW - 1 + p.in_x_off, ty * g.TH - 1, b);
          }
        }
        __syncwarp();
        if (++as == g.a_stages) { as = 0; aph ^= 1; }
        for (int tap = 0; tap < 9; ++tap) {
          if (!((p.tap_mask >> tap) & 1)) continue;
          mbar_wait(&b_empty[bs], bph ^ 1, p.err, 12);
          if (elect_one()) {
            if (CG == 2) {
              const uint32_t bar = mapa_u32(&b_full[bs], 0);
              mbar_expect_tx_cluster(bar, B_BYTES);
              tma_load_2d_2sm(sB + bs * B_BYTES, &tmB, bar, (tap * chunks + ck) * 64, n_tile * N_TILE + (int)rank * NB);
            } else {
              mbar_expect_tx(&b_full[bs], B_BYTES);
              tma_load_2d(sB + bs * B_BYTES, &tmB, &b_full[bs], (tap * chunks + ck) * 64, n_tile * N_TILE);
            }
          }
          __syncwarp();
          if (++bs == g.b_stages) { bs = 0; bph ^= 1; }
        }
      }
      // fused downsample input: TH x PW positions of the stride-2 sampled block input, 1 "tap"
      for (int dk = 0; dk < p.ds_chunks; ++dk) {
        mbar_wait(&a_empty[as], aph ^ 1, p.err, 17);
        if (elect_one()) {
          const int ds_bytes = g.TH * g.PW * 128;
          if (CG == 2) {
            const uint32_t bar = mapa_u32(&a_full[as], 0);
            mbar_expect_tx_cluster(bar, ds_bytes);
            tma_load_4d_2sm(sA + as * g.a_stage_bytes, &tmD, bar, dk * 64, 2 * tx * g.TW, 2 * ty * g.TH, b);
          } else {
            mbar_expect_tx(&a_full[as], ds_bytes);
            tma_load_4d(sA + as * g.a_stage_bytes, &tmD, &a_full[as], dk * 64, 2 * tx * g.TW, 2 * ty * g.TH, b);
          }
        }
        __syncwarp();
        if (++as == g.a_stages) { as = 0; aph ^= 1; }
        mbar_wait(&b_empty[bs], bph ^ 1, p.err, 18);
        if (elect_one()) {
          if (CG == 2) {
            const uint32_t bar = mapa_u32(&b_full[bs], 0);
            mbar_expect_tx_cluster(bar, B_BYTES);
            tma_load_2d_2sm(sB + bs * B_BYTES, &tmB, bar, (9 * chunks + dk) * 64, n_tile * N_TILE + (int)rank * NB);
          } else {
            mbar_expect_tx(&b_full[bs], B_BYTES);
            tma_load_2d(sB + bs * B_BYTES, &tmB, &b_full[bs], (9 * chunks + dk) * 64, n_tile * N_TILE);
          }
        }
        __syncwarp();
        if (++bs == g.b_stages) { bs = 0; bph ^= 1; }
      }
    }
  } else if (warp <= MW) {
    // ================= MMA issuers (leader CTA only) =================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(N_TILE, 128 * CG);
      const int g0 = (warp - 1) * GW;  // first sub-tile of this warp
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      const int first_tap = __ffs(p.tap_mask) - 1, last_tap = 31 - __clz(p.tap_mask);
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        mbar_wait(&tempty[acc], acc_phase ^ 1, p.err, 13);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)((acc * G + g0) * N_TILE);
        for (int ck = 0; ck < chunks; ++ck) {
          mbar_wait(&a_full[as], aph, p.err, 14);
          // start address advances by whole 128 B rows: (gi*128 + row_off) * 128 B >> 4
          const uint64_t a0 = make_smem_desc(sA + as * g.a_stage_bytes) + (uint64_t)(g0 * g.sub_stride * 8);
          for (int tap = 0; tap < 9; ++tap) {
            if (!((p.tap_mask >> tap) & 1)) continue;
            mbar_wait(&b_full[bs], bph, p.err, 15);
            tc_fence_after();
            const int r = tap / 3, s = tap - 3 * r;
            const uint64_t bdesc = make_smem_desc(sB + bs * B_BYTES);
            const uint64_t at = a0 + (uint64_t)((r * g.PW + s) * 8);
            const uint32_t first = (ck != 0 || tap != first_tap) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int gi = 0; gi < GW; ++gi) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (CG == 2)
                    umma_bf16_2sm(d_base + (uint32_t)(gi * N_TILE), at + (uint64_t)(gi * g.sub_stride * 8 + 2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                  k != 0 ? 1u : first);
                  else
                    umma_bf16(d_base + (uint32_t)(gi * N_TILE), at + (uint64_t)(gi * g.sub_stride * 8 + 2 * k), bdesc + (uint64_t)(2 * k), idesc,
                              k != 0 ? 1u : first);
                }
              }
              if (CG == 2) {
                umma_commit_2sm(&b_empty[bs]);
                if (tap == last_tap) {
                  umma_commit_2sm(&a_empty[as]);
                  if (ck == chunks - 1 && p.ds_chunks == 0) umma_commit_2sm(&tfull[acc]);
                }
              } else {
                umma_commit(&b_empty[bs]);
                if (tap == last_tap) {
                  umma_commit(&a_empty[as]);
                  if (ck == chunks - 1 && p.ds_chunks == 0) umma_commit(&tfull[acc]);
                }
              }
            }
            __syncwarp();
            if (++bs == g.b_stages) { bs = 0; bph ^= 1; }
          }
          if (++as == g.a_stages) { as = 0; aph ^= 1; }
        }
        for (int dk = 0; dk < p.ds_chunks; ++dk) {  // fused 1x1 stride-2 downsample: no tap offset
          mbar_wait(&a_full[as], aph, p.err, 19);
          mbar_wait(&b_full[bs], bph, p.err, 20);
          tc_fence_after();
          const uint64_t a0 = make_smem_desc(sA + as * g.a_stage_bytes) + (uint64_t)(g0 * g.sub_stride * 8);
          const uint64_t bdesc = make_smem_desc(sB + bs * B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int gi = 0; gi < GW; ++gi) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (CG == 2) umma_bf16_2sm(d_base + (uint32_t)(gi * N_TILE), a0 + (uint64_t)(gi * g.sub_stride * 8 + 2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                else umma_bf16(d_base + (uint32_t)(gi * N_TILE), a0 + (uint64_t)(gi * g.sub_stride * 8 + 2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
              }
            }
            if (CG == 2) {
              umma_commit_2sm(&b_empty[bs]);
              umma_commit_2sm(&a_empty[as]);
              if (dk == p.ds_chunks - 1) umma_commit_2sm(&tfull[acc]);
            } else {
              umma_commit(&b_empty[bs]);
              umma_commit(&a_empty[as]);
              if (dk == p.ds_chunks - 1) umma_commit(&tfull[acc]);
            }
          }
          __syncwarp();
          if (++bs == g.b_stages) { bs = 0; bph ^= 1; }
          if (++as == g.a_stages) { as = 0; aph ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (TS && warp == 1 + MW + HL_EPI_WARPS) {
    // ================= TMA store / residual-load warp (one thread: bulk groups are per thread) =================
    if (lane == 0) {
      const int obuf_tx = g.sub_rows * g.TW * 128;
      auto coords = [&](int unit, int gi, int &c0, int &c1, int &c2, int &c3) -> bool {
        const int n_tile = unit / pairs_per_n, m_tile = (unit - n_tile * pairs_per_n) * CG + (int)rank;
        const int b = m_tile / tiles_per_img, t = m_tile - b * tiles_per_img;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        c0 = n_tile * N_TILE; c1 = tx * g.TW; c2 = ty * g.TH + gi * g.sub_rows; c3 = b;
        return m_tile < num_m_tiles;
      };
      // hand staging tile `buf` to the epilogue for sub-tile (unit, gi): with its residual tile in it, or just free
      auto prepare = [&](int buf, int unit, int gi) {
        if (unit >= num_units) return;
        if (p.residual) {
          int c0, c1, c2, c3;
          coords(unit, gi, c0, c1, c2, c3);  // past-the-end tiles read zeros (batch index out of bounds)
          mbar_expect_tx(&o_ready[buf], obuf_tx);
          tma_load_4d(sStg + buf * g.obuf_bytes, &tmR, &o_ready[buf], c0, c1, c2, c3);
        } else {
          mbar_arrive(&o_ready[buf]);
        }
      };
      int pu = unit0, pg = 0;  // next sub-tile to prepare
      for (int j = 0; j < g.obufs; ++j) {
        prepare(j, pu, pg);
        if (++pg == G) { pg = 0; pu += unit_step; }
      }
      int buf = 0, prev = -1;
      uint32_t dph = 0;
      for (int unit = unit0; unit < num_units; unit += unit_step) {
        for (int gi = 0; gi < G; ++gi) {
          mbar_wait(&o_done[buf], dph, p.err, 21);
          int c0, c1, c2, c3;
          if (coords(unit, gi, c0, c1, c2, c3)) {
            if (N_TILE == 128) {  // pair mode: half 0 -> `out` (tmO), half 1 -> `out2` (tmR's slot); the one-column shift between them is in the maps' base pointers
              tma_store_4d(&tmO, sStg + buf * g.obuf_bytes, 0, c1, c2, c3);
              tma_store_4d(&tmR, sStg + buf * g.obuf_bytes + g.obuf_bytes / 2, 0, c1, c2, c3);
            } else {
              tma_store_4d(&tmO, sStg + buf * g.obuf_bytes, c0, c1, c2, c3);
            }
          }
          bulk_commit_group();
          if (prev >= 0) {
            bulk_wait_group_read<1>();  // the previous sub-tile's store has left its staging tile
            prepare(prev, pu, pg);
            if (++pg == G) { pg = 0; pu += unit_step; }
          }
          prev = buf;
          if (++buf == g.obufs) { buf = 0; dph ^= 1; }
        }
      }
      bulk_wait_group<0>();
    }
  } else if (TS && N_TILE == 128) {
    // ================= pair-mode epilogue: plain bf16 conversion, one staging tile per 64-channel half =================
    const int ew = warp - 1 - MW;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which of the two convolutions
    const int m_local = quarter * 32 + lane;
    const int ysub = m_local / g.PW, xl = m_local - ysub * g.PW;
    const bool row_ok = m_local < g.sub_stride && xl < g.TW;
    const int r = ysub * g.TW + xl;
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0, rph = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      mbar_wait(&tfull[acc], acc_phase, p.err, 16);
      tc_fence_after();
#pragma unroll 1
      for (int gi = 0; gi < G; ++gi) {
        mbar_wait(&o_ready[buf], rph, p.err, 22);
        const uint32_t tile = smem_u32(sStg + buf * g.obuf_bytes + half * (g.obuf_bytes / 2));
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * G + gi) * N_TILE + half * 64 + blk * 32), v);
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              sts_16(tile + (uint32_t)(r * 128 + (((4 * blk + c) ^ (r & 7)) << 4)),
                     make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]), pack_bf16(v[c * 8 + 4], v[c * 8 + 5]),
                                pack_bf16(v[c * 8 + 6], v[c * 8 + 7])));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_done[buf]);
        if (++buf == g.obufs) { buf = 0; rph ^= 1; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
        else mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (TS) {
    // ================= epilogue through TMA =================
    const int ew = warp - 1 - MW;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which 32 of the 64 channels
    const int m_local = quarter * 32 + lane;
    const int ysub = m_local / g.PW, xl = m_local - ysub * g.PW;
    const bool row_ok = m_local < g.sub_stride && xl < g.TW;  // else a junk row: pitch padding / overlap with the next sub-tile
    const int r = ysub * g.TW + xl;                           // row of the staging tile = box-linear pixel index
    uint32_t own[4];                                          // this thread's four 16-byte chunks (128B swizzle, like TMA's)
#pragma unroll
    for (int c = 0; c < 4; ++c) own[c] = (uint32_t)(r * 128 + (((4 * half + c) ^ (r & 7)) << 4));
    const bool affine = p.scale != nullptr, has_res = p.residual != nullptr;
    const bool affine_c = affine && p.scale_host != nullptr && p.num_n_tiles == 1;  // constants valid (Cout = 64)
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0, rph = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      const int n0 = (unit / pairs_per_n) * N_TILE + half * 32;
      mbar_wait(&tfull[acc], acc_phase, p.err, 16);
      tc_fence_after();
#pragma unroll 1
      for (int gi = 0; gi < G; ++gi) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * G + gi) * N_TILE + half * 32), v);
        if (affine_c) {
          if (half == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], ha.scale[j], ha.shift[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], ha.scale[32 + j], ha.shift[32 + j]);
          }
        } else if (affine) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sc = *reinterpret_cast<const float4 *>(s_scale + n0 + 4 * j);
            const float4 sh = *reinterpret_cast<const float4 *>(s_shift + n0 + 4 * j);
            v[4 * j + 0] = fmaf(v[4 * j + 0], sc.x, sh.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], sc.y, sh.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], sc.z, sh.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], sc.w, sh.w);
          }
        }
        mbar_wait(&o_ready[buf], rph, p.err, 22);
        const uint32_t tile = smem_u32(sStg + buf * g.obuf_bytes);
        if (row_ok) {
          if (has_res) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 u = lds_16(tile + own[c]);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(w[j]); v[c * 8 + 2 * j] += f.x; v[c * 8 + 2 * j + 1] += f.y; }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts_16(tile + own[c], make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]),
                                             pack_bf16(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16(v[c * 8 + 6], v[c * 8 + 7])));
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_done[buf]);
        if (++buf == g.obufs) { buf = 0; rph ^= 1; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
        else mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ================= epilogue =================
    const int ew = warp - 1 - MW;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int half = ew >> 2;      // which half of the N_TILE columns
    constexpr int NH = N_TILE / 2;
    constexpr int NBLK = NH / 32;  // 32-channel blocks per sub-tile for this warp
    const uint32_t stg = smem_u32(sStg + ew * 2048);
    EpiParams e;
    e.s_scale = s_scale; e.s_shift = s_shift; e.has_affine = p.scale != nullptr;
    e.addend = p.residual; e.add_mode = p.residual ? EPI_ADD_RESIDUAL : EPI_ADD_NONE;
    e.out = p.out; e.sum_out = nullptr;
    e.Cout = p.Cout; e.out_ldc = p.out_ldc; e.out_coff = p.out_coff; e.rep = p.rep; e.Wo = p.Wo; e.relu = p.relu;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = unit0; unit < num_units; unit += unit_step) {
      const int n_tile = unit / pairs_per_n, m_tile = (unit - n_tile * pairs_per_n) * CG + (int)rank;
      const bool real_tile = m_tile < num_m_tiles;
      const int b = m_tile / tiles_per_img, t = m_tile - b * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      const int n0 = n_tile * N_TILE + half * NH;
      // coalesced-view rows of sub-tile gi: row it*8 + (lane >> 2) of this warp's 32
      auto rows_of = [&](int gi, EpiRows &rw) {
        rw.valid = 0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int m = gi * 128 + quarter * 32 + it * 8 + (lane >> 2);
          const int yl = m / g.PW, xl = m - yl * g.PW;
          const int y = ty * g.TH + yl, x = tx * g.TW + xl;
          const bool ok = real_tile && yl < g.TH && xl < g.TW && y < p.Ho && x < p.Wo;
          rw.opix[it] = (b * p.Ho + y) * p.Wo + x;
          rw.apix[it] = rw.opix[it];
          if (ok) rw.valid |= 1u << it;
        }
      };
      EpiRows rw, rw_next;
      uint4 pre[4] = {};
      rows_of(0, rw);
      if (e.add_mode != EPI_ADD_NONE) epi_fetch_addend(pre, lane, rw, e.addend, e.Cout, n0);
      mbar_wait(&tfull[acc], acc_phase, p.err, 16);
      tc_fence_after();
#pragma unroll 1
      for (int gi = 0; gi < G; ++gi) {
        if (gi + 1 < G) rows_of(gi + 1, rw_next);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * G + gi) * N_TILE + half * NH);
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
          float v[32];
          tmem_ld32(taddr + blk * 32, v);
          uint4 cur[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) cur[it] = pre[it];
          if (e.add_mode != EPI_ADD_NONE) {  // fetch the next block's addend now
            if (blk + 1 < NBLK) epi_fetch_addend(pre, lane, rw, e.addend, e.Cout, n0 + (blk + 1) * 32);
            else if (gi + 1 < G) epi_fetch_addend(pre, lane, rw_next, e.addend, e.Cout, n0);
          }
          epi_block32(v, lane, stg, rw, e, n0 + blk * 32, cur);
        }
        rw = rw_next;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));  // the issuer waits in the leader
        else mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// Picks the padded pitch / tile height for a Wo-wide map: TH * PW <= G * 128, valid fraction
// (TW / PW) * (TH * PW / (G*128)) * (coverage of Wo and Ho by whole tiles) maximised.
static void pick_geom(int Ho, int Wo, int G, int max_tw, int *PW, int *TH) {
  double best = -1.0;
  for (int tw = 8; tw <= max_tw && tw <= Wo + 7; ++tw) {
    const int pw = tw + 2;
    int th = G * 128 / pw;
    if (th > 254) th = 254;
    if (th < 1) continue;
    if (th > Ho) th = Ho;
    const int tx = (Wo + tw - 1) / tw, ty = (Ho + th - 1) / th;
    const double eff = (double)Ho * Wo / ((double)tx * ty * G * 128);
    // A traffic per output grows with the halo: prefer taller tiles at equal efficiency
    static const double halo_w = getenv("OCRB_HALO_W") ? atof(getenv("OCRB_HALO_W")) : 0.3;  // tuning knob
    const double score = eff - halo_w * ((double)(th + 2) * pw / ((double)th * tw) - 1.0);
    if (score > best) { best = score; *PW = pw; *TH = th; }
  }
}

// Row-aligned variant (TMA-store epilogue): each sub-tile = sub_rows whole rows of pitch PW.
static void pick_geom_rows(int Ho, int Wo, int G, int max_tw, int *PW, int *sub_rows) {
  double best = -1.0;
  for (int tw = 8; tw <= max_tw && tw <= 126 && tw <= Wo + 7; ++tw) {
    const int pw = tw + 2;
    int rps = 128 / pw;
    const int need = (Ho + G - 1) / G;
    if (rps > need) rps = need;
    const int th = G * rps;
    const int tx = (Wo + tw - 1) / tw, ty = (Ho + th - 1) / th;
    const double eff = (double)Ho * Wo / ((double)tx * ty * G * 128);
    static const double halo_w = getenv("OCRB_HALO_W") ? atof(getenv("OCRB_HALO_W")) : 0.3;
    const double score = eff - halo_w * ((double)(th + 2) * pw / ((double)th * tw) - 1.0);
    if (score > best) { best = score; *PW = pw; *sub_rows = rps; }
  }
}

// TMA-store epilogue for 64-channel convolutions without replication (OCRB_HALO_TS=0 turns it off)
bool halo_use_ts(int n_tile, int rep) {
  static const bool on = !(getenv("OCRB_HALO_TS") && atoi(getenv("OCRB_HALO_TS")) == 0);
  return on && n_tile == 64 && rep == 1;
}

static int halo_geometry_try(int Ho, int Wo, int n_tile, int G, int ts, int max_tw, HaloGeom *out);

// A very wide, very flat map (one row of 240 pixels, say) would pick a tile whose halo'd operand stage leaves no room
// for the weight ring: retry with narrower tiles until everything fits.
int halo_geometry(int Ho, int Wo, int n_tile, int G, int ts, HaloGeom *out) {  // n_tile = weight rows per CTA
  for (int max_tw = 254; max_tw >= 8; max_tw /= 2)
    if (halo_geometry_try(Ho, Wo, n_tile, G, ts, max_tw, out) == OCRB_OK) return OCRB_OK;
  set_error("conv_halo: no tiling of a %dx%d map fits shared memory", Wo, Ho);
  return OCRB_ERR_INTERNAL;
}

static int halo_geometry_try(int Ho, int Wo, int n_tile, int G, int ts, int max_tw, HaloGeom *out) {
  HaloGeom g;
  if (ts) {
    pick_geom_rows(Ho, Wo, G, max_tw, &g.PW, &g.sub_rows);
    g.TH = G * g.sub_rows;
    g.sub_stride = g.sub_rows * g.PW;
  } else {
    // PW <= 128: the fused-downsample input is a stride-2 TMA box of 2 * PW pixels (box dimensions are limited to 256)
    pick_geom(Ho, Wo, G, max_tw < 126 ? max_tw : 126, &g.PW, &g.TH);
    g.sub_rows = 0;
    g.sub_stride = 128;
  }
  g.TW = g.PW - 2;
  g.obuf_bytes = ts ? ((g.sub_rows * g.TW * 128 + 1023) / 1024) * 1024 * (ts == 2 ? 2 : 1) : 0;  // ts = 2: pair mode, one tile per half
  g.obufs = ts ? 3 : 0;
  const int halo_rows = (g.TH + 2) * g.PW;
  const int read_rows = (G - 1) * g.sub_stride + 128 + 2 * g.PW + 2;  // last sub-tile, tap (2,2)
  int rows = halo_rows > read_rows ? halo_rows : read_rows;
  g.a_stage_bytes = ((rows * 128 + 1023) / 1024) * 1024;
  g.a_tx_bytes = halo_rows * 128;
  const int b_bytes = n_tile * 128;
  int budget = 227 * 1024 - 1024 /*align*/ - 4096 /*static scale/shift*/ - 512 /*barriers*/ - (ts ? g.obufs * g.obuf_bytes : HL_STG_BYTES);
  g.a_stages = 2;
  g.b_stages = (budget - g.a_stages * g.a_stage_bytes) / b_bytes;
  if (ts && g.b_stages < 4) {  // wide tiles with whole weight tiles per CTA (single-CTA mode): two staging tiles instead of three
    budget += g.obuf_bytes;
    g.obufs = 2;
    g.b_stages = (budget - g.a_stages * g.a_stage_bytes) / b_bytes;
  }
  if (g.b_stages > 8) {
    // room to spare: a third A stage helps layers with many input chunks
    if (budget - 3 * g.a_stage_bytes >= 6 * b_bytes) { g.a_stages = 3; g.b_stages = (budget - 3 * g.a_stage_bytes) / b_bytes; }
    if (g.b_stages > 8) g.b_stages = 8;
  }
  if (g.b_stages < 2) return OCRB_ERR_INTERNAL;  // the caller retries with narrower tiles
  *out = g;
  return OCRB_OK;
}

int make_act_tensor_map_box(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h);
int make_act_tensor_map_pitched(CUtensorMap *map, const void *base, int B, int H, int W, int C, int ldc, int box_w, int box_h, int step = 1, int row_px = 0);


template <int N_TILE, int G, int CG, int TS>
static int launch_halo_one(ocrb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmD, const CUtensorMap &tmO,
                           const CUtensorMap &tmR, const ConvTcParams &p, const HaloGeom &g, int num_units, const char *tag) {
  HaloAffine ha = {};
  ConvTcParams pk = p;
  if (TS && p.scale_host && p.shift_host && p.Cout == 64) {
    memcpy(ha.scale, p.scale_host, sizeof(ha.scale));
    memcpy(ha.shift, p.shift_host, sizeof(ha.shift));
  } else {
    pk.scale_host = pk.shift_host = nullptr;
  }
  auto kern = conv_halo_kernel<N_TILE, G, CG, TS>;
  OCRB_TRY(ensure_dyn_smem(ctx, kern, 227 * 1024 - 4096));
  const int smem = 1024 + g.a_stages * g.a_stage_bytes + g.b_stages * (N_TILE / CG) * 128 + (TS ? g.obufs * g.obuf_bytes : HL_STG_BYTES) + 512;
  int grid = num_units * CG < ctx->sm_budget() ? num_units * CG : (ctx->sm_budget() / CG) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((1 + (G >= 4 ? 2 : 1) + HL_EPI_WARPS + TS) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OCRB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmD, tmO, tmR, pk, g, ha));
  return check_launch(ctx, tag);
}

static int halo_cg() {  // CTA-pair mode unless OCRB_HALO_CG=1
  static const int cg = (getenv("OCRB_HALO_CG") && atoi(getenv("OCRB_HALO_CG")) == 1) ? 1 : 2;
  return cg;
}

int make_halo_act_map(CUtensorMap *map, const void *base, int B, int H, int W, int C, int n_tile, int G, int rep, int pair) {
  HaloGeom g;
  // pair mode: one output column more than input columns, geometry of the launch (launch_conv_halo) must match
  OCRB_TRY(halo_geometry(H, W + (pair ? 1 : 0), n_tile / halo_cg(), G, pair ? 2 : (int)halo_use_ts(n_tile, rep), &g));
  return make_act_tensor_map_box(map, base, B, H, W, C, g.PW, g.TH + 2);
}
int halo_weight_box_rows(int n_tile) { return n_tile / halo_cg(); }

// 3x3 / stride 1 / pad 1 only; n_tile in {64 (G = 4), 128 (G = 2)}
// the fused downsample input [B][H][W][C] sampled at stride 2: box of PW x TH positions (pitch PW like the main tile)
int make_act_tensor_map_strided_box(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_w, int box_h, int stride);
int make_halo_ds_map(CUtensorMap *map, const void *base, int B, int H, int W, int C, int Ho, int Wo, int n_tile, int G) {
  HaloGeom g;
  OCRB_TRY(halo_geometry(Ho, Wo, n_tile / halo_cg(), G, 0, &g));  // fused downsample: 128-channel tiles only
  return make_act_tensor_map_strided_box(map, base, B, H, W, C, g.PW, g.TH, 2);
}

int launch_conv_halo(ocrb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, ConvTcParams p, int n_tile, const char *tag,
                     const CUtensorMap *tmDp) {
  const CUtensorMap &tmD = tmDp ? *tmDp : tmA;
  if ((p.ds_chunks != 0) != (tmDp != nullptr)) { set_error("conv_halo: downsample chunks without a tensor map (or vice versa)"); return OCRB_ERR_INVALID; }
  if (p.R != 3 || p.S != 3 || p.stride != 1 || p.pad != 1 || p.sum_out || !p.out) { set_error("conv_halo: unsupported convolution"); return OCRB_ERR_INVALID; }
  if (p.Cout % n_tile != 0 || p.Cout > 512) { set_error("conv_halo: Cout %d vs N tile %d", p.Cout, n_tile); return OCRB_ERR_INVALID; }
  const int G = n_tile == 64 ? 4 : 2, CG = halo_cg();
  const int ts = p.pair_mode ? 2 : (int)halo_use_ts(n_tile, p.rep);
  if (p.pair_mode && (n_tile != 128 || p.Cout != 128 || p.rep != 1 || p.residual || p.scale || p.relu || !p.out2 || p.ds_chunks || !halo_use_ts(64, 1))) {
    set_error("conv_halo: bad pair-mode arguments");
    return OCRB_ERR_INVALID;
  }
  if (!ts && (p.out_step != 1 || (p.tap_mask & 0x1ff) != 0x1ff)) { set_error("conv_halo: tap mask / output step need the TMA-store path"); return OCRB_ERR_INVALID; }
  if ((p.tap_mask & 0x1ff) == 0) { set_error("conv_halo: empty tap mask"); return OCRB_ERR_INVALID; }
  p.tap_mask &= 0x1ff;
  if (ts && p.ds_chunks) { set_error("conv_halo: fused downsample with 64-channel tiles"); return OCRB_ERR_INVALID; }
  HaloGeom g;
  OCRB_TRY(halo_geometry(p.Ho, p.Wo, n_tile / CG, G, ts, &g));
  p.tiles_x = (int)cdiv(p.Wo, g.TW);
  p.tiles_y = (int)cdiv(p.Ho, g.TH);
  p.num_n_tiles = p.Cout / n_tile;
  const int num_units = (int)cdiv((int64_t)p.tiles_x * p.tiles_y * p.B, CG) * p.num_n_tiles;
  if (ts == 2) {
    // two 64-channel outputs, Wo columns each (the caller's buffers have the padding for the surplus end columns)
    if (p.out_row_px <= 0) { set_error("conv_halo: pair mode needs padded output rows"); return OCRB_ERR_INVALID; }
    CUtensorMap tmO, tmO2;
    OCRB_TRY(make_act_tensor_map_pitched(&tmO, p.out + p.out_coff, p.B, p.Ho, p.Wo, 64, p.out_ldc, g.TW, g.sub_rows, p.out_step, p.out_row_px));
    OCRB_TRY(make_act_tensor_map_pitched(&tmO2, p.out2 + p.out_coff, p.B, p.Ho, p.Wo, 64, p.out_ldc, g.TW, g.sub_rows, p.out_step, p.out_row_px));
    return CG == 2 ? launch_halo_one<128, 2, 2, 1>(ctx, tmA, tmB, tmD, tmO, tmO2, p, g, num_units, tag)
                   : launch_halo_one<128, 2, 1, 1>(ctx, tmA, tmB, tmD, tmO, tmO2, p, g, num_units, tag);
  }
  if (ts) {
    // output slice [B][Ho][Wo][Cout] at channel offset out_coff of an out_ldc-wide buffer; residual [B][Ho][Wo][Cout]
    CUtensorMap tmO, tmR;
    OCRB_TRY(make_act_tensor_map_pitched(&tmO, p.out + p.out_coff, p.B, p.Ho, p.Wo, p.Cout, p.out_ldc, g.TW, g.sub_rows, p.out_step, p.out_row_px));
    if (p.residual) OCRB_TRY(make_act_tensor_map_pitched(&tmR, p.residual, p.B, p.Ho, p.Wo, p.Cout, p.Cout, g.TW, g.sub_rows, 1, p.res_row_px));
    else tmR = tmO;
    return CG == 2 ? launch_halo_one<64, 4, 2, 1>(ctx, tmA, tmB, tmD, tmO, tmR, p, g, num_units, tag)
                   : launch_halo_one<64, 4, 1, 1>(ctx, tmA, tmB, tmD, tmO, tmR, p, g, num_units, tag);
  }
  if (n_tile == 64) return CG == 2 ? launch_halo_one<64, 4, 2, 0>(ctx, tmA, tmB, tmD, tmA, tmA, p, g, num_units, tag) : launch_halo_one<64, 4, 1, 0>(ctx, tmA, tmB, tmD, tmA, tmA, p, g, num_units, tag);
  if (n_tile == 128) return CG == 2 ? launch_halo_one<128, 2, 2, 0>(ctx, tmA, tmB, tmD, tmA, tmA, p, g, num_units, tag) : launch_halo_one<128, 2, 1, 0>(ctx, tmA, tmB, tmD, tmA, tmA, p, g, num_units, tag);
  set_error("conv_halo: unsupported N tile %d", n_tile);
  return OCRB_ERR_INVALID;
}

// host-side view of the tiling for tests (include/ocrb.h)
int debug_conv_geometry(int Ho, int Wo, int mode, int *out) {
  const int cg = halo_cg();
  const int n_tile = mode == 1 ? 64 : 128, G = mode == 1 ? 4 : 2;
  HaloGeom g;
  OCRB_TRY(halo_geometry(Ho, Wo + (mode == 2 ? 1 : 0), n_tile / cg, G, mode, &g));
  const int stg = mode ? g.obufs * g.obuf_bytes : HL_STG_BYTES;
  const int smem = 1024 + g.a_stages * g.a_stage_bytes + g.b_stages * (n_tile / cg) * 128 + stg + 512;
  const int v[10] = {g.PW, g.TH, g.TW, g.sub_rows, g.sub_stride, g.a_stage_bytes, g.a_stages, g.b_stages, stg, smem};
  for (int i = 0; i < 10; ++i) out[i] = v[i];
  return OCRB_OK;
}

}  // namespace ocrb
