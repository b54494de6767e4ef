// Shared epilogue of the tcgen05 convolution kernels: one warp turns a 32-row x 32-channel
// block of fp32 accumulators (row = TMEM lane = thread) into bf16 NHWC global stores.
//
// The accumulator layout gives every thread one pixel's channels, so direct stores would hit
// 32 different 128-byte lines per instruction (the first version did: L1TEX-bound).  Here
// every global access is re-mapped through a 2 KB per-warp shared-memory block so that 4
// consecutive lanes cover 64 contiguous bytes of one pixel and a warp instruction touches 8
// lines: residual / FPN addend loads and all output stores are 16-byte, sector-exact.
#pragma once
#include <cuda_bf16.h>

#include <cstdint>

#include "tc_ptx.cuh"

namespace ocrb {

struct EpiRows {
  // "coalesced view": in iteration it (0..3) this lane serves row it*8 + (lane >> 2),
  // 16-byte chunk (lane & 3) of that row's 64-byte block
  int32_t opix[4];   // output pixel index of that row (valid rows only; B*Ho*Wo < 2^31)
  int32_t apix[4];   // addend pixel index (residual: = opix; FPN: the half-resolution pixel)
  uint32_t valid;    // bit it = row is a real output pixel
};

// staging address of (row, 16-byte chunk): XOR swizzle keeps both views conflict-free
__device__ __forceinline__ uint32_t epi_stg_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ uint4 ld_nc_16(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_16(void *p, const uint4 &v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_16(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts_16(uint32_t saddr, const uint4 &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// coalesced fetch of the addend block (32 rows x 64 B at channel n) into registers; issued one
// block ahead of its use so the global latency hides behind the previous block's work
__device__ __forceinline__ void epi_fetch_addend(uint4 (&pre)[4], int lane, const EpiRows &rw, const __nv_bfloat16 *add, int add_ld, int n) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    pre[it] = make_uint4(0, 0, 0, 0);
    if (rw.valid & (1u << it)) pre[it] = ld_nc_16(add + (int64_t)rw.apix[it] * add_ld + n + (lane & 3) * 8);
  }
}
__device__ __forceinline__ void epi_stage_addend(uint32_t stg, int lane, const uint4 (&pre)[4]) {
#pragma unroll
  for (int it = 0; it < 4; ++it) sts_16(stg + epi_stg_off(it * 8 + (lane >> 2), lane & 3), pre[it]);
}
// this thread's own row (row = lane): 32 bf16 addend values added into v
__device__ __forceinline__ void epi_add_own_row(uint32_t stg, int lane, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 u = lds_16(stg + epi_stg_off(lane, c));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(w[j]); v[c * 8 + 2 * j] += f.x; v[c * 8 + 2 * j + 1] += f.y; }
  }
}
__device__ __forceinline__ void epi_put_own_row(uint32_t stg, int lane, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    sts_16(stg + epi_stg_off(lane, c), make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]),
                                                  pack_bf16(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16(v[c * 8 + 6], v[c * 8 + 7])));
}
// coalesced store of the staged block to out[(opix) * ldc + coff + n ...]; rep > 1 replicates
// every pixel into a rep x rep block of the (Wo*rep)-wide map (nearest upsample, model.rs:82-97)
__device__ __forceinline__ void epi_store(uint32_t stg, int lane, const EpiRows &rw, __nv_bfloat16 *out, int ldc, int coff_n, int rep, int Wo) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (!(rw.valid & (1u << it))) continue;
    const uint4 u = lds_16(stg + epi_stg_off(it * 8 + (lane >> 2), lane & 3));
    if (rep == 1) {
      st_16(out + (int64_t)rw.opix[it] * ldc + coff_n + (lane & 3) * 8, u);
    } else {
      // opix = (b*Ho + y)*Wo + x  ->  ((b*Ho + y)*rep + ry) * Wo*rep + x*rep + rx
      const int64_t row = rw.opix[it] / Wo;
      const int x = rw.opix[it] - (int)row * Wo;
      const int64_t Wr = (int64_t)Wo * rep;
      for (int ry = 0; ry < rep; ++ry)
        for (int rx = 0; rx < rep; ++rx)
          st_16(out + ((row * rep + ry) * Wr + (int64_t)x * rep + rx) * ldc + coff_n + (lane & 3) * 8, u);
    }
  }
}

// ---- addend prefetch ring (cp.async straight into shared memory, a whole tile ahead) -------
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void *src, bool ok) {
  const int n = ok ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one 32-row x 64-byte addend block -> ring slot (same swizzled layout as the staging block)
__device__ __forceinline__ void epi_prefetch_addend(uint32_t slot, int lane, const EpiRows &rw, const __nv_bfloat16 *add, int add_ld, int n) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const bool ok = (rw.valid >> it) & 1u;
    cp_async_16_zfill(slot + epi_stg_off(it * 8 + (lane >> 2), lane & 3), ok ? add + (int64_t)rw.apix[it] * add_ld + n + (lane & 3) * 8 : add, ok);
  }
}

enum { EPI_ADD_NONE = 0, EPI_ADD_RESIDUAL = 1, EPI_ADD_SUM = 2 };

struct EpiParams {
  const float *s_scale, *s_shift;  // shared-memory copies, indexed by absolute channel
  int has_affine;                  // 0: scale = 1, shift = 0 (convolutions without batch-norm)
  // addend [..][Cout] bf16: RESIDUAL: y += addend[opix] before ReLU (model.rs:47-53);
  // SUM: second output sum_out = y + addend[apix] (FPN "up2 + lateral", model.rs:126-137)
  const __nv_bfloat16 *addend;
  int add_mode;
  __nv_bfloat16 *out, *sum_out;
  int Cout, out_ldc, out_coff, rep, Wo, relu;
};

// v: this thread's 32 accumulators for channels [n, n + 32); stg: this warp's 2 KB block;
// pre: the addend block fetched by epi_fetch_addend (ignored when add_mode == NONE), or, when
// add_slot != 0, the addend block sits in that shared-memory ring slot (epi_prefetch_addend).
__device__ __forceinline__ void epi_block32(float (&v)[32], int lane, uint32_t stg, const EpiRows &rw, const EpiParams &e, int n,
                                            const uint4 (&pre)[4], uint32_t add_slot = 0) {
  if (e.has_affine) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 128-bit broadcast loads: 16 shared-memory wavefronts instead of 64
      const float4 sc = *reinterpret_cast<const float4 *>(e.s_scale + n + 4 * j);
      const float4 sh = *reinterpret_cast<const float4 *>(e.s_shift + n + 4 * j);
      v[4 * j + 0] = fmaf(v[4 * j + 0], sc.x, sh.x);
      v[4 * j + 1] = fmaf(v[4 * j + 1], sc.y, sh.y);
      v[4 * j + 2] = fmaf(v[4 * j + 2], sc.z, sh.z);
      v[4 * j + 3] = fmaf(v[4 * j + 3], sc.w, sh.w);
    }
  }
  if (e.add_mode == EPI_ADD_RESIDUAL) {
    epi_stage_addend(stg, lane, pre);
    __syncwarp();
    epi_add_own_row(stg, lane, v);
    __syncwarp();
  }
  if (e.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (e.add_mode == EPI_ADD_SUM) {
    // lateral + FPN sum: y is staged once; the coalesced view stores y itself (if wanted) and
    // y + addend, adding the prefetched addend registers in place (no second smem round trip;
    // y is rounded to bf16 before the add, as the materialised lateral would be)
    epi_put_own_row(stg, lane, v);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (!(rw.valid & (1u << it))) continue;
      const uint4 u = lds_16(stg + epi_stg_off(it * 8 + (lane >> 2), lane & 3));
      if (e.out) st_16(e.out + (int64_t)rw.opix[it] * e.out_ldc + e.out_coff + n + (lane & 3) * 8, u);
      uint4 av = pre[it];
      if (add_slot) av = lds_16(add_slot + epi_stg_off(it * 8 + (lane >> 2), lane & 3));
      const uint32_t yw[4] = {u.x, u.y, u.z, u.w}, aw[4] = {av.x, av.y, av.z, av.w};
      uint32_t sw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(yw[j]), b = unpack_bf16(aw[j]);
        sw[j] = pack_bf16(a.x + b.x, a.y + b.y);
      }
      st_16(e.sum_out + (int64_t)rw.opix[it] * e.Cout + n + (lane & 3) * 8, make_uint4(sw[0], sw[1], sw[2], sw[3]));
    }
    __syncwarp();
    return;
  }
  if (e.out) {
    epi_put_own_row(stg, lane, v);
    __syncwarp();
    epi_store(stg, lane, rw, e.out, e.out_ldc, e.out_coff + n, e.rep, e.Wo);
    __syncwarp();
  }
}

}  // namespace ocrb
