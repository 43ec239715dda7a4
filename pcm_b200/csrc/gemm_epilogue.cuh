// Epilogue of one 128 x block_n accumulator tile, shared by the 1-CTA and 2-CTA GEMM kernels.
// Two phases per 32-column chunk so that global traffic is coalesced:
//   1. each thread owns one accumulator row (TMEM lane): TMEM -> registers -> fp32 staging tile in
//      shared memory (16-byte chunks XOR-swizzled by row to stay conflict free);
//   2. the 128 threads of an epilogue group re-map to (row, 8-column group): 4 neighbouring lanes
//      cover 64 contiguous output bytes of one row, add bias / row vector / residual, apply the
//      activation, store 16 bytes.
// Two groups of 4 warps take the even / odd chunks of the tile (grp = 0 / 1).
//
// Small-K GEMMs are bound by this code, and it is instruction-fetch sensitive (ncu: 35 % of the
// samples were `stall_no_inst` with a 14 k-instruction kernel), so the common case - bf16 output,
// full 32-column chunk, no activation, no split-K - is a compact straight-line path and everything else (ragged N,
// fp32 output, split-K partial sums) lives in one out-of-line routine with rolled loops.
#pragma once
#include "gemm_params.h"

namespace pcm {

// Generic phase 2 of one chunk for the calling thread's 4 rows: ragged N, fp32 output, split-K.
static __device__ __noinline__ void gemm_epilogue_rows_generic(const GemmParams& p, const float* sb, int tm,
                                                        int n, int r0, int cg, float* ws) {
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int rr = r0 + 32 * i;
    const int m = tm * 128 + rr;
    if (m >= p.M) continue;
    const float* srow = sb + rr * 32;
    if (ws) {  // split-K partial sums: plain stores into this split's slice of the workspace (the
               // finalize kernel adds the slices in split order and applies the epilogue)
      const float4 a0 = *reinterpret_cast<const float4*>(srow + ((2 * cg) ^ (rr & 7)) * 4);
      const float4 a1 = *reinterpret_cast<const float4*>(srow + ((2 * cg + 1) ^ (rr & 7)) * 4);
      const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float* wp = ws + static_cast<long long>(m) * p.N + n;
      if (n + 8 <= p.N && (p.N & 3) == 0) {   // two 16-byte stores (rows are 16-byte aligned)
        *reinterpret_cast<float4*>(wp) = make_float4(f[0] * p.alpha, f[1] * p.alpha, f[2] * p.alpha, f[3] * p.alpha);
        *reinterpret_cast<float4*>(wp + 4) = make_float4(f[4] * p.alpha, f[5] * p.alpha, f[6] * p.alpha, f[7] * p.alpha);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (n + e < p.N) wp[e] = f[e] * p.alpha;
      }
      continue;
    }
    const int b = m / p.epiHW;
    const int r = m - b * p.epiHW;
    const int h = r / p.epiW;
    const int w = r - h * p.epiW;
    const long long o = b * p.osB + h * p.osH + w * p.osW;
#pragma unroll 1
    for (int e = 0; e < 8 && n + e < p.N; ++e) {
      float x = srow[((2 * cg + (e >> 2)) ^ (rr & 7)) * 4 + (e & 3)] * p.alpha;
      if (p.bias) x += p.bias[n + e];
      if (p.rowvec) x += __bfloat162float(p.rowvec[b * p.rowvec_ld + n + e]);
      if (p.residual) x += __bfloat162float(p.residual[o + n + e]);
      if (p.act == 1) x = silu_f(x);
      if (p.out_fp32) {
        if (p.round_bf16) x = __bfloat162float(__float2bfloat16_rn(x));
        reinterpret_cast<float*>(p.out)[o + n + e] = x;
      } else {
        reinterpret_cast<bf16*>(p.out)[o + n + e] = __float2bfloat16_rn(x);
      }
    }
  }
}

__device__ __forceinline__ void add_bf16x8(float (&f)[8], const uint4& u) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] += t.x; f[1] += t.y;
  t = unpack_bf16x2(u.y); f[2] += t.x; f[3] += t.y;
  t = unpack_bf16x2(u.z); f[4] += t.x; f[5] += t.y;
  t = unpack_bf16x2(u.w); f[6] += t.x; f[7] += t.y;
}

template <class Release>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, int tm, int n0, uint32_t taddr,
                                                   float* sb, int lane, int row, int grp, int cg,
                                                   int r0, uint64_t* tfull, uint32_t tfull_phase,
                                                   Release release, float* ws = nullptr) {
  const bool has_bias = p.bias != nullptr, has_rv = p.rowvec != nullptr;
  // split-K, fp32 output and the SiLU epilogue (time-embedding MLP, M = batch) take the generic path
  const bool fast_ok = p.ws == nullptr && !p.out_fp32 && p.act == 0;
  const bool has_res = p.residual != nullptr && fast_ok;
  long long off[4];
  const bf16* rvp[4];
  bool valid[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = tm * 128 + r0 + 32 * i;
    valid[i] = m < p.M;
    off[i] = 0;
    rvp[i] = nullptr;
    if (valid[i]) {
      const int b = m / p.epiHW;
      const int r = m - b * p.epiHW;
      const int h = r / p.epiW;
      const int w = r - h * p.epiW;
      off[i] = b * p.osB + h * p.osH + w * p.osW;
      if (has_rv) rvp[i] = p.rowvec + b * p.rowvec_ld;
    }
  }
  const int nchunks = p.block_n >> 5;
  const uint32_t sb_u32 = smem_u32(sb);
  const uint32_t st_row = sb_u32 + row * 128, st_x = (row & 7) << 4;            // phase 1 (own row)
  const uint32_t ld_a0 = sb_u32 + r0 * 128 + (((2 * cg) ^ (r0 & 7)) << 4);        // phase 2
  const uint32_t ld_a1 = sb_u32 + r0 * 128 + (((2 * cg + 1) ^ (r0 & 7)) << 4);
  // Residual reads come from HBM (~1 us): a group keeps the reads of TWO of its chunks in flight.
  // Chunks 0 and 1 are issued here, before the accumulator is even complete; chunk jj + 2 is issued
  // as soon as chunk jj's registers are consumed.  (out may alias residual for in-place accumulation:
  // every element is read and later written by the same thread, tiles are disjoint.)
  uint4 rcur[4], rnext[4];
  auto load_res = [&](uint4 (&dst)[4], int jj) {
    const int n = n0 + (grp + 2 * jj) * 32 + cg * 8;
    if (grp + 2 * jj < nchunks && n + 8 <= p.N) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (valid[i]) dst[i] = *reinterpret_cast<const uint4*>(p.residual + off[i] + n);
    }
  };
  if (has_res) {
    load_res(rcur, 0);
    load_res(rnext, 1);
  }
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  const int last_j = ((nchunks - 1 - grp) & ~1) + grp;  // last chunk this group handles
  if (grp >= nchunks) {  // block_n == 32: group 1 has no chunk, still releases the accumulator
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release();
  }
#pragma unroll 1
  for (int jj = 0; grp + 2 * jj < nchunks; ++jj) {
    const int j = grp + 2 * jj;
    const int n = n0 + j * 32 + cg * 8;
    const bool fast = fast_ok && (n0 + j * 32 + 32 <= p.N);  // uniform over the group
    // bias / row vectors are L2 resident: issued per chunk, ahead of the TMEM load
    float4 bia0 = make_float4(0.f, 0.f, 0.f, 0.f), bia1 = bia0;
    uint4 rrv[4];
    if (fast) {
      if (has_bias) {
        bia0 = *reinterpret_cast<const float4*>(p.bias + n);
        bia1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
      }
      if (has_rv) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (valid[i]) rrv[i] = *reinterpret_cast<const uint4*>(rvp[i] + n);
      }
    }
    {
      uint32_t v[32];
      tmem_ld_32x32(taddr + j * 32, v);
      tmem_ld_wait();
      if (j == last_j) {
        // all TMEM reads of this group for this tile are done: release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(st_row + ((c << 4) ^ st_x), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    // group-local barrier (ids 1 / 2): staging tile written
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    if (fast) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!valid[i]) continue;
        // rows r0 + 32 i share (row & 7): same swizzle, 4 KB apart
        const float4 a0 = lds128(ld_a0 + i * 4096);
        const float4 a1 = lds128(ld_a1 + i * 4096);
        float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (p.alpha != 1.0f) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] *= p.alpha;
        }
        f[0] += bia0.x; f[1] += bia0.y; f[2] += bia0.z; f[3] += bia0.w;
        f[4] += bia1.x; f[5] += bia1.y; f[6] += bia1.z; f[7] += bia1.w;
        if (has_rv) add_bf16x8(f, rrv[i]);
        if (has_res) add_bf16x8(f, rcur[i]);
        uint4 u;
        u.x = pack_bf16x2(f[0], f[1]);
        u.y = pack_bf16x2(f[2], f[3]);
        u.z = pack_bf16x2(f[4], f[5]);
        u.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off[i] + n) = u;
      }
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) rcur[i] = rnext[i];
        load_res(rnext, jj + 2);
      }
    } else if (n < p.N) {
      gemm_epilogue_rows_generic(p, sb, tm, n, r0, cg, ws);
    }
    // staging tile consumed: the group may overwrite it in its next chunk
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
  }
}

}  // namespace pcm
