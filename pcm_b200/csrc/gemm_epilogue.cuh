// Epilogue of one 128 x block_n accumulator tile, shared by the 1-CTA and 2-CTA GEMM kernels.
// Two phases per 32-column chunk so that global traffic is coalesced:
//   1. each thread owns one accumulator row (TMEM lane): TMEM -> registers -> fp32 staging tile in
//      shared memory (16-byte chunks XOR-swizzled by row to stay conflict free);
//   2. the 128 threads of an epilogue group re-map to (row, 8-column group): 4 neighbouring lanes
//      cover 64 contiguous output bytes of one row, add bias / row vector / residual (all issued
//      before the TMEM load so their latency is hidden), apply the activation, store 16 bytes.
// Two groups of 4 warps take the even / odd chunks of the tile (grp = 0 / 1).
#pragma once
#include "gemm_params.h"

namespace pcm {

template <class Release>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, int tm, int n0, uint32_t taddr,
                                                   float* sb, int lane, int row, int grp, int cg,
                                                   int r0, uint64_t* tfull, uint32_t tfull_phase,
                                                   Release release) {
  const bool has_bias = p.bias != nullptr, has_res = p.residual != nullptr;
  const bool has_rv = p.rowvec != nullptr, has_alpha = p.alpha != 1.0f;
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
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  const int nchunks = p.block_n >> 5;
  const int last_j = ((nchunks - 1 - grp) & ~1) + grp;  // last chunk this group handles
  if (grp >= nchunks) {  // block_n == 32: group 1 has no chunk, still releases the accumulator
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release();
  }
  for (int j = grp; j < nchunks; j += 2) {
    const int n = n0 + j * 32 + cg * 8;
    const bool full8 = n + 8 <= p.N;
    // Issue every global read of this chunk's phase 2 up front (bias, row vectors, residual):
    // their latency overlaps the TMEM load / staging / barrier, and nothing is re-read after a
    // store (out may alias residual for in-place accumulation, each element by the same thread).
    float4 bia0 = make_float4(0.f, 0.f, 0.f, 0.f), bia1 = bia0;
    uint4 rres[4], rrv[4];
    if (full8) {
      if (has_bias) {
        bia0 = *reinterpret_cast<const float4*>(p.bias + n);
        bia1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
      }
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (valid[i]) rres[i] = *reinterpret_cast<const uint4*>(p.residual + off[i] + n);
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
      float* srow = sb + row * 32;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cs = c ^ (row & 7);
        *reinterpret_cast<float4*>(srow + cs * 4) =
            make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                        __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
      }
    }
    // group-local barrier (ids 1 / 2): staging tile written
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    if (n < p.N) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!valid[i]) continue;
        const int rr = r0 + 32 * i;
        const float* srow = sb + rr * 32;
        const float4 a0 = *reinterpret_cast<const float4*>(srow + ((2 * cg) ^ (rr & 7)) * 4);
        const float4 a1 = *reinterpret_cast<const float4*>(srow + ((2 * cg + 1) ^ (rr & 7)) * 4);
        float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (has_alpha) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] *= p.alpha;
        }
        const long long o = off[i];
        if (p.ws) {  // split-K partial sums (fp32 atomics; finalize kernel applies the epilogue)
          float* wp = p.ws + static_cast<long long>(tm * 128 + rr) * p.N + n;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (n + e < p.N) atomicAdd(wp + e, f[e]);
          continue;
        }
        if (full8) {
          if (has_bias) {
            f[0] += bia0.x; f[1] += bia0.y; f[2] += bia0.z; f[3] += bia0.w;
            f[4] += bia1.x; f[5] += bia1.y; f[6] += bia1.z; f[7] += bia1.w;
          }
          if (has_rv) {
            float2 t;
            t = unpack_bf16x2(rrv[i].x); f[0] += t.x; f[1] += t.y;
            t = unpack_bf16x2(rrv[i].y); f[2] += t.x; f[3] += t.y;
            t = unpack_bf16x2(rrv[i].z); f[4] += t.x; f[5] += t.y;
            t = unpack_bf16x2(rrv[i].w); f[6] += t.x; f[7] += t.y;
          }
          if (has_res) {
            float2 t;
            t = unpack_bf16x2(rres[i].x); f[0] += t.x; f[1] += t.y;
            t = unpack_bf16x2(rres[i].y); f[2] += t.x; f[3] += t.y;
            t = unpack_bf16x2(rres[i].z); f[4] += t.x; f[5] += t.y;
            t = unpack_bf16x2(rres[i].w); f[6] += t.x; f[7] += t.y;
          }
          if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = silu_f(f[e]);
          }
          if (p.out_fp32) {
            if (p.round_bf16) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16_rn(f[e]));
            }
            float* op = reinterpret_cast<float*>(p.out) + o + n;
            *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            uint4 u;
            u.x = pack_bf16x2(f[0], f[1]);
            u.y = pack_bf16x2(f[2], f[3]);
            u.z = pack_bf16x2(f[4], f[5]);
            u.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + o + n) = u;
          }
        } else {
          // ragged N tail (e.g. conv_out, N = 4): scalar path
          const bf16* rv = rvp[i];
          for (int e = 0; e < 8 && n + e < p.N; ++e) {
            float x = f[e];
            if (has_bias) x += p.bias[n + e];
            if (rv) x += __bfloat162float(rv[n + e]);
            if (has_res) x += __bfloat162float(p.residual[o + n + e]);
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
    }
    // staging tile consumed: the group may overwrite it in its next chunk
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
  }
}

}  // namespace pcm
