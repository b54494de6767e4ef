// Interface of the tcgen05 implicit-GEMM convolution engine (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace ocrb {

// EPI_HEAD2: the head tail with its 64 -> 4 contraction as a second tcgen05 GEMM (A operand written back to TMEM); EPI_HEAD: the
// same tail on the CUDA cores (OCRB_HEAD=cuda)
enum { EPI_STD = 0, EPI_HEAD = 1, EPI_F32 = 2, EPI_HEAD2 = 3, EPI_HEAD2_TS = 4 };

struct ConvTcParams {
  // problem geometry (output side); tiles_x/tiles_y/num_n_tiles are filled by the launcher
  int B = 0, Ho = 0, Wo = 0, Cout = 0;
  int R = 1, S = 1, cin_chunks = 1, stride = 1, pad = 0;
  // conv_halo only: a fused 1x1 stride-2 side input (ResNet downsample, model.rs:30-38): ds_chunks extra
  // K blocks of 64 channels read from a second tensor map, weights appended after the 9 taps
  int ds_chunks = 0;
  const __nv_bfloat16 *ds_src = nullptr;    // host-side only: base of that side input
  int tiles_x = 0, tiles_y = 0, num_n_tiles = 0;
  // standard epilogue: y = acc * scale[c] + shift[c] (+ residual) (ReLU)
  const float *scale = nullptr, *shift = nullptr;
  const float *scale_host = nullptr, *shift_host = nullptr;  // host copies of the same (optional; launcher-side only)
  const __nv_bfloat16 *residual = nullptr;  // [B][Ho][Wo][Cout]
  int relu = 0;
  __nv_bfloat16 *out = nullptr;             // [B][Ho*rep][Wo*rep][out_ldc], channel offset out_coff
  int out_ldc = 0, out_coff = 0, rep = 1;
  // conv_halo TMA-store path only: 3x3 taps actually present (bit r*3+s; absent taps are skipped, their weights
  // never read) and the pixel step of the output (2: results land on every second pixel / row of a map twice
  // as large — `out` then points at the first of them)
  int tap_mask = 0x1ff, out_step = 1;
  // conv_halo "pair" mode (N tile 128 through the TMA-store epilogue): the two 64-channel halves are two different
  // convolutions of the same input patch whose results belong to neighbouring output columns — half 0 goes to `out` at
  // column X - 1, half 1 to `out2` at column X (X = output column of the tile, Wo = input width + 1, the input tile
  // starts one column further left: in_x_off = -1).  Used for the parity-class convolutions of the fused neck.
  int pair_mode = 0, in_x_off = 0;
  __nv_bfloat16 *out2 = nullptr;
  // TMA-store path: pixels per full-resolution row of the out / residual buffers when their rows are padded (0 = dense).
  // TMA stores take no negative coordinates, so pair mode writes its column -1 (and Wo - 1 of `out2`) into padding:
  // both outputs are Wo columns wide.
  int out_row_px = 0, res_row_px = 0;
  const __nv_bfloat16 *up_src = nullptr;    // [B][Ho/2][Wo/2][Cout]
  __nv_bfloat16 *sum_out = nullptr;         // [B][Ho][Wo][Cout] = y + up2(up_src)
  // DB head tail
  const float *w2 = nullptr;                // [4][64]
  float b2 = 0.f, thresh = 0.6f;
  float *prob = nullptr;                    // [B][4*Ho][4*Wo]
  uint8_t *bitmap = nullptr;                // optional
  int *err = nullptr;                       // device flag set before a pipeline-timeout trap
  // FP32-accuracy mode on the tensor cores (EPI_F32): every fp32 operand is split into bf16 terms (x = hi + mid (+ lo)), the
  // activation tensor holds the terms as channel planes [hi | mid (| lo)] of split_cin channels each, the weight rows hold the
  // matching blocks, and a product of split numbers becomes split_nblk K blocks per (tap, 64-channel chunk):
  //   2 terms, 3 blocks: (a_hi, w_hi) (a_hi, w_mid) (a_mid, w_hi)                          — 16 significant bits
  //   3 terms, 6 blocks: (a_hi, w_hi) (a_hi, w_mid) (a_hi, w_lo) (a_mid, w_hi) (a_mid, w_mid) (a_lo, w_hi) — 24 bits
  // cin_chunks counts ALL K blocks of a tap (split_nblk * split_cin / 64).  Accumulation, affine, residual and output are fp32.
  int split_nblk = 0, split_cin = 0;
  const float *res32 = nullptr;             // [B][Ho][Wo][Cout]
  float *out32 = nullptr;                   // [B][Ho][Wo][Cout]
};

// DB head tail constants, passed by value (kernel parameter = constant bank: the fully unrolled
// epilogue reads them as immediate constant operands, no shared-memory traffic)
struct HeadConsts {
  float scale[64], shift[64];  // bin_bn2 folded with the conv-transpose-1 bias
  float w2[256];               // conv-transpose-2 weights [co][q] (q = 2*i' + j' fastest)
};

int make_act_tensor_map(CUtensorMap *map, const void *base, int B, int H, int W, int C, int stride);
int make_weight_tensor_map(CUtensorMap *map, const void *base, int Cout, int Ktot, int n_tile);
int launch_conv_tc(ocrb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, ConvTcParams p, int n_tile, int epi, const char *tag = "tc:conv",
                   const HeadConsts *hc = nullptr);

// conv_halo.cu: 3x3 stride-1 convolutions with the halo'd input tile resident in shared memory
int make_halo_act_map(CUtensorMap *map, const void *base, int B, int H, int W, int C, int n_tile, int G, int rep, int pair = 0);
bool halo_use_ts(int n_tile, int rep);  // TMA-store epilogue in use for this tile shape
int halo_weight_box_rows(int n_tile);  // rows of the weight TMA box (half the N tile in CTA-pair mode)
int make_halo_ds_map(CUtensorMap *map, const void *base, int B, int H, int W, int C, int Ho, int Wo, int n_tile, int G);
int launch_conv_halo(ocrb_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, ConvTcParams p, int n_tile, const char *tag,
                     const CUtensorMap *tmD = nullptr);

// conv_lateral.cu: FPN lateral 1x1 conv + "up2 + lateral" sum, TMA in / TMA out
bool lateral_ts_supported(int Cin, int Cout, int Ho, int Wo);
int launch_conv_lateral(ocrb_ctx *ctx, const __nv_bfloat16 *in, const __nv_bfloat16 *w, const __nv_bfloat16 *upper, __nv_bfloat16 *out,
                        __nv_bfloat16 *sum, int B, int Ho, int Wo, int Cin, int *err, const char *tag);

}  // namespace ocrb
