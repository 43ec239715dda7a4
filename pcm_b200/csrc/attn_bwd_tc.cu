// tcgen05 flash-attention backward for head sizes d <= 64 (SD1.5: d = 40 at the 64x64 level, which
// carries 85 % of the attention FLOPs; larger heads keep the mma.sync kernels in attn.cu).
//
// Two kernels, both recompute the scores from Q, K and the forward's log-sum-exp (no S x S tensor):
//   attn_bwd_dkdv_tc : CTA = 128 keys of one (batch, head), loops over 128-query tiles
//        S^T = K Q^T, dP^T = V dO^T        (TMEM, M = keys)
//        P^T = exp2(c S^T - L[q]), dS^T = P^T o (dP^T - delta[q])   -> bf16 smem (K-major A operands)
//        dV += P^T dO,  dK += dS^T Q       (TMEM accumulators across the loop; dO / Q are read as
//                                           MN-major B operands from the same TMA tiles)
//   attn_bwd_dq_tc   : CTA = 128 queries, loops over 128-key tiles
//        S = Q K^T, dP = dO V^T ; dS = P o (dP - delta[row]) -> smem ; dQ += dS K (TMEM)
// The backward needs no row reduction, so the element-wise stage is split by COLUMNS across two
// warpgroups (8 warps, two per SM sub-partition): warpgroup g handles score columns [64g, 64g+64).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 4-11 element-wise + epilogue.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "host_common.h"
#include "../../include/pcm_b200.h"

namespace pcm {

struct AttnBwdTcParams {
  CUtensorMap q_map, k_map, v_map, do_map;  // (C, S, B) views, box (64, 128, 1)
  bf16 *dq, *dk, *dv;
  const float *lse, *delta;
  int B, H, Sq, Skv, D;
  long long ldq, ldk, ldv;
  float scale;
};

constexpr int kBwdThreads = 384;
constexpr int kBoxB = 128 * 128;

__device__ __forceinline__ float fexp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write 8 bf16 (one 16-byte chunk `ch` of row `row`) into a K-major SWIZZLE_128B operand tile pair
// (base = shared-space address of the tile pair)
__device__ __forceinline__ void st_operand_chunk(uint32_t base, int row, int ch, const float (&v)[8]) {
  sts128(base + (ch >> 3) * kBoxB + row * 128 + (((ch & 7) ^ (row & 7)) << 4), pack_bf16x2(v[0], v[1]),
         pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ------------------------------------------------------------------------------------------
// dK, dV
// smem: K, V (1 box each) | stages x {Q, dO} | P^T (2 boxes) | dS^T (2 boxes) | L, delta per stage
// TMEM: S^T [0,128) dP^T [128,256) dV [256,384) dK [384,512)
// ------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_dkdv_tc_kernel(const __grid_constant__ AttnBwdTcParams p) {
  constexpr int KSTEPS = DP / 16;
  constexpr int ST = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kBoxB;
  uint8_t* sQ = sV + kBoxB;            // [ST]
  uint8_t* sdO = sQ + ST * kBoxB;      // [ST]
  uint8_t* sP = sdO + ST * kBoxB;      // 2 boxes
  uint8_t* sdS = sP + 2 * kBoxB;       // 2 boxes
  float* sL = reinterpret_cast<float*>(sdS + 2 * kBoxB);  // [ST][128]
  float* sD = sL + ST * 128;                              // [ST][128]
  __shared__ __align__(8) uint64_t kv_full, kv_ready, q_full[ST], q_free[ST], ld_full[ST], s_full,
      s_free, p_full, p_free, acc_full;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, n0 = blockIdx.x * 128;
  const int ntiles = (p.Sq + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    tma_prefetch_desc(&p.v_map);
    tma_prefetch_desc(&p.do_map);
    mbar_init(&kv_full, 1);
    mbar_init(&kv_ready, 128);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_free[i], 1);
      mbar_init(&ld_full[i], 128);
    }
    mbar_init(&s_full, 1);
    mbar_init(&s_free, 256);
    mbar_init(&p_full, 256);
    mbar_init(&p_free, 1);
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  griddep_sync();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&kv_full, 2 * kBoxB);
      tma_load_4d(sK, &p.k_map, &kv_full, h * p.D, n0, b, 0);
      tma_load_4d(sV, &p.v_map, &kv_full, h * p.D, n0, b, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % ST;
        mbar_wait(&q_free[st], ((j / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[st], 2 * kBoxB);
        tma_load_4d(sQ + st * kBoxB, &p.q_map, &q_full[st], h * p.D, j * 128, b, 0);
        tma_load_4d(sdO + st * kBoxB, &p.do_map, &q_full[st], h * p.D, j * 128, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_g = umma_idesc_bf16(128, DP, 0, 1);  // B (dO / Q) MN-major
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      mbar_wait(&kv_ready, 0);  // K, V landed and their stray head columns are zeroed
      // S^T / dP^T of tile j+1 are issued as soon as the element-wise warps have copied tile j to
      // registers, so the score GEMMs overlap the exp / dS math of the previous tile
      auto issue_scores = [&](int j) {
        const int st = j % ST;
        mbar_wait(&q_full[st], (j / ST) & 1);
        mbar_wait(&s_free, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(sQ + st * kBoxB), do_addr = smem_u32(sdO + st * kBoxB);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(k_addr + k * 32, 16, 1024),
                   umma_desc_sw128(q_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_f16(tmem_base + 128, umma_desc_sw128(v_addr + k * 32, 16, 1024),
                   umma_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full);
      };
      issue_scores(0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % ST;
        if (j + 1 < ntiles) issue_scores(j + 1);
        const uint32_t q_addr = smem_u32(sQ + st * kBoxB), do_addr = smem_u32(sdO + st * kBoxB);
        mbar_wait(&p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 queries per step
          const uint32_t ao = (k >> 2) * kBoxB + (k & 3) * 32;
          umma_f16(tmem_base + 256, umma_desc_sw128(p_addr + ao, 16, 1024),
                   umma_desc_sw128(do_addr + k * 2048, kBoxB, 1024), idesc_g, (j | k) != 0 ? 1u : 0u);
          umma_f16(tmem_base + 384, umma_desc_sw128(ds_addr + ao, 16, 1024),
                   umma_desc_sw128(q_addr + k * 2048, kBoxB, 1024), idesc_g, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&p_free);
        umma_commit(&q_free[st]);
      }
      umma_commit(&acc_full);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== element-wise stage: thread = key row, warpgroup = column half ============
    const int g = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    const bool kvalid = n0 + row < p.Skv;
    const float* Lg = p.lse + (static_cast<long long>(b) * p.H + h) * p.Sq;
    const float* Dg = p.delta + (static_cast<long long>(b) * p.H + h) * p.Sq;
    const int et = threadIdx.x - 128;  // 0..255
    // Head slices are 64-column boxes at column h*d: zero the stray columns [D, DP) of K and V (once)
    // so that the k-extent DP of S^T = K Q^T and dP^T = V dO^T only sees this head.
    if (g == 0) {
      mbar_wait(&kv_full, 0);
      if (p.D < DP) {
#pragma unroll
        for (int ch = 0; ch < DP / 8; ++ch) {
          if (ch * 8 < p.D) continue;
          const int o = row * 128 + (((ch & 7) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(sK + o) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sV + o) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
      }
      mbar_arrive(&kv_ready);
    }
    const uint32_t sP_u = smem_u32(sP), sdS_u = smem_u32(sdS);
    const uint32_t sL_u = smem_u32(sL), sD_u = smem_u32(sD);
    // per-query statistics (first 128 threads stage them): fetched one tile ahead so that the HBM
    // latency never sits between two tiles; L = +inf past the end
    float l_pre = INFINITY, d_pre = 0.f;
    if (et < 128 && et < p.Sq) {
      l_pre = Lg[et];
      d_pre = Dg[et];
    }
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % ST;
      if (et < 128) {
        sL[st * 128 + et] = l_pre;
        sD[st * 128 + et] = d_pre;
        mbar_arrive(&ld_full[st]);
        const int qn = (j + 1) * 128 + et;
        const bool vn = j + 1 < ntiles && qn < p.Sq;
        l_pre = vn ? Lg[qn] : INFINITY;
        d_pre = vn ? Dg[qn] : 0.f;
      }
      mbar_wait(&ld_full[st], (j / ST) & 1);
      mbar_wait(&s_full, j & 1);
      tc_fence_after();
      // copy this warpgroup's 64 score / dP columns to registers, then hand the TMEM tiles back so
      // the next tile's score GEMMs run under the math below
      uint32_t sv[2][32], dv[2][32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        tmem_ld_32x32(tmem_base + lane_off + g * 64 + cc * 32, sv[cc]);
        tmem_ld_32x32(tmem_base + 128 + lane_off + g * 64 + cc * 32, dv[cc]);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free);
      if (j > 0) mbar_wait(&p_free, (j - 1) & 1);  // previous P^T / dS^T consumed by the MMAs
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col0 = g * 64 + cc * 32;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float pv[8], gv[8];
          // the 8 queries' statistics: two broadcast 16-byte shared loads each
          const uint32_t so = (st * 128 + col0 + ch * 8) * 4;
          const float4 l0 = lds128(sL_u + so), l1 = lds128(sL_u + so + 16);
          const float4 d0 = lds128(sD_u + so), d1 = lds128(sD_u + so + 16);
          const float lq[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          const float dq8[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float pe = kvalid ? fexp2(fmaf(__uint_as_float(sv[cc][ch * 8 + e]), c, -lq[e])) : 0.f;
            pv[e] = pe;
            gv[e] = pe * (__uint_as_float(dv[cc][ch * 8 + e]) - dq8[e]);
          }
          st_operand_chunk(sP_u, row, (col0 >> 3) + ch, pv);
          st_operand_chunk(sdS_u, row, (col0 >> 3) + ch, gv);
        }
      }
      fence_proxy_async();
      mbar_arrive(&p_full);
    }
    // epilogue: warpgroup 0 writes dV, warpgroup 1 writes dK
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const int key = n0 + row;
    const uint32_t tacc = tmem_base + 256 + g * 128 + lane_off;
    bf16* dst = (g == 0 ? p.dv : p.dk) + (static_cast<long long>(b) * p.Skv + key) * (g == 0 ? p.ldv : p.ldk) + h * p.D;
    const float sc = g == 0 ? 1.f : p.scale;
#pragma unroll
    for (int cc = 0; cc < DP / 16; ++cc) {
      uint32_t v[16];
      tmem_ld_32x16(tacc + cc * 16, v);
      tmem_ld_wait();
      if (key < p.Skv) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int col = cc * 16 + hh * 8;
          if (col < p.D) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[hh * 8 + 0]) * sc, __uint_as_float(v[hh * 8 + 1]) * sc);
            u.y = pack_bf16x2(__uint_as_float(v[hh * 8 + 2]) * sc, __uint_as_float(v[hh * 8 + 3]) * sc);
            u.z = pack_bf16x2(__uint_as_float(v[hh * 8 + 4]) * sc, __uint_as_float(v[hh * 8 + 5]) * sc);
            u.w = pack_bf16x2(__uint_as_float(v[hh * 8 + 6]) * sc, __uint_as_float(v[hh * 8 + 7]) * sc);
            *reinterpret_cast<uint4*>(dst + col) = u;
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// dQ
// smem: Q, dO (1 box each) | stages x {K, V} | dS (2 boxes)
// TMEM: S [0,128) dP [128,256) dQ [256,384)
// ------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(kBwdThreads, 1) attn_bwd_dq_tc_kernel(const __grid_constant__ AttnBwdTcParams p) {
  constexpr int KSTEPS = DP / 16;
  constexpr int ST = 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + kBoxB;
  uint8_t* sK = sdO + kBoxB;          // [ST]
  uint8_t* sV = sK + ST * kBoxB;      // [ST]
  uint8_t* sdS = sV + ST * kBoxB;     // [2 buffers][2 boxes]: tile j + 1's dS is written while
                                      // the dQ MMAs of tile j still read theirs
  __shared__ __align__(8) uint64_t q_full, q_ready, kv_full[ST], kv_free[ST], s_full, s_free, p_full[2],
      p_free[2], acc_full;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int ntiles = (p.Skv + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    tma_prefetch_desc(&p.v_map);
    tma_prefetch_desc(&p.do_map);
    mbar_init(&q_full, 1);
    mbar_init(&q_ready, 128);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], 1);
    }
    mbar_init(&s_full, 1);
    mbar_init(&s_free, 256);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 256);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  griddep_sync();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, 2 * kBoxB);
      tma_load_4d(sQ, &p.q_map, &q_full, h * p.D, q0, b, 0);
      tma_load_4d(sdO, &p.do_map, &q_full, h * p.D, q0, b, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % ST;
        mbar_wait(&kv_free[st], ((j / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kBoxB);
        tma_load_4d(sK + st * kBoxB, &p.k_map, &kv_full[st], h * p.D, j * 128, b, 0);
        tma_load_4d(sV + st * kBoxB, &p.v_map, &kv_full[st], h * p.D, j * 128, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_g = umma_idesc_bf16(128, DP, 0, 1);  // B = K MN-major
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), ds_addr = smem_u32(sdS);
      mbar_wait(&q_ready, 0);  // Q, dO landed and their stray head columns are zeroed
      auto issue_scores = [&](int j) {
        const int st = j % ST;
        mbar_wait(&kv_full[st], (j / ST) & 1);
        mbar_wait(&s_free, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * kBoxB), v_addr = smem_u32(sV + st * kBoxB);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(q_addr + k * 32, 16, 1024),
                   umma_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_f16(tmem_base + 128, umma_desc_sw128(do_addr + k * 32, 16, 1024),
                   umma_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full);
      };
      issue_scores(0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % ST;
        if (j + 1 < ntiles) issue_scores(j + 1);
        const uint32_t k_addr = smem_u32(sK + st * kBoxB);
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t ds_j = ds_addr + (j & 1) * 2 * kBoxB;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 keys per step
          umma_f16(tmem_base + 256, umma_desc_sw128(ds_j + (k >> 2) * kBoxB + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(k_addr + k * 2048, kBoxB, 1024), idesc_g, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&p_free[j & 1]);
        umma_commit(&kv_free[st]);
      }
      umma_commit(&acc_full);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = (warp - 4) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    const int qrow = q0 + row;
    const long long li = (static_cast<long long>(b) * p.H + h) * p.Sq + qrow;
    const float lrow = qrow < p.Sq ? p.lse[li] : INFINITY;
    const float drow = qrow < p.Sq ? p.delta[li] : 0.f;
    if (g == 0) {
      mbar_wait(&q_full, 0);
      if (p.D < DP) {
#pragma unroll
        for (int ch = 0; ch < DP / 8; ++ch) {
          if (ch * 8 < p.D) continue;
          const int o = row * 128 + (((ch & 7) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(sQ + o) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sdO + o) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
      }
      mbar_arrive(&q_ready);
    }
    const uint32_t sdS_u = smem_u32(sdS);
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&s_full, j & 1);
      tc_fence_after();
      uint32_t sv[2][32], dv[2][32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        tmem_ld_32x32(tmem_base + lane_off + g * 64 + cc * 32, sv[cc]);
        tmem_ld_32x32(tmem_base + 128 + lane_off + g * 64 + cc * 32, dv[cc]);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free);
      if (j >= 2) mbar_wait(&p_free[j & 1], ((j >> 1) - 1) & 1);  // dQ MMAs of tile j - 2 done
      const uint32_t ds_u = sdS_u + (j & 1) * 2 * kBoxB;
      auto tile_math = [&](auto masked) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int col0 = g * 64 + cc * 32;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float gv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float pe = fexp2(fmaf(__uint_as_float(sv[cc][ch * 8 + e]), c, -lrow));
              if constexpr (decltype(masked)::value) {
                if (j * 128 + col0 + ch * 8 + e >= p.Skv) pe = 0.f;
              }
              gv[e] = pe * (__uint_as_float(dv[cc][ch * 8 + e]) - drow);
            }
            st_operand_chunk(ds_u, row, (col0 >> 3) + ch, gv);
          }
        }
      };
      if ((j + 1) * 128 <= p.Skv) tile_math(std::false_type{});
      else tile_math(std::true_type{});
      fence_proxy_async();
      mbar_arrive(&p_full[j & 1]);
    }
    // epilogue: warpgroup g writes dQ columns [g * DP/2 ...) -- split by 16-column TMEM chunks
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    bf16* dst = p.dq + (static_cast<long long>(b) * p.Sq + qrow) * p.ldq + h * p.D;
#pragma unroll
    for (int cc = 0; cc < DP / 16; ++cc) {
      if ((cc & 1) != g) continue;
      uint32_t v[16];
      tmem_ld_32x16(tmem_base + 256 + lane_off + cc * 16, v);
      tmem_ld_wait();
      if (qrow < p.Sq) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int col = cc * 16 + hh * 8;
          if (col < p.D) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[hh * 8 + 0]) * p.scale, __uint_as_float(v[hh * 8 + 1]) * p.scale);
            u.y = pack_bf16x2(__uint_as_float(v[hh * 8 + 2]) * p.scale, __uint_as_float(v[hh * 8 + 3]) * p.scale);
            u.z = pack_bf16x2(__uint_as_float(v[hh * 8 + 4]) * p.scale, __uint_as_float(v[hh * 8 + 5]) * p.scale);
            u.w = pack_bf16x2(__uint_as_float(v[hh * 8 + 6]) * p.scale, __uint_as_float(v[hh * 8 + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dst + col) = u;
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int enc_map(CUtensorMap* map, const void* ptr, int C, int S, int B, long long ld) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(S),
                        static_cast<cuuint64_t>(B), 1};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * S,
                           static_cast<cuuint64_t>(ld) * 2 * S * B};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_tmap(map, ptr, 4, dims, strides, box, estr);
}

template <int DP>
static int launch_bwd_tc(const AttnBwdTcParams& p, cudaStream_t stream) {
  const size_t smem1 = static_cast<size_t>(2 + 2 * 2 + 4) * kBoxB + 4 * 128 * sizeof(float) + 1024;
  const size_t smem2 = static_cast<size_t>(2 + 2 * 3 + 4) * kBoxB + 1024;  // dS double-buffered
  static bool set = false;
  if (!set) {
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem1)));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem2)));
    set = true;
  }
  CUDA_TRY(launch_pdl(attn_bwd_dkdv_tc_kernel<DP>, dim3((p.Skv + 127) / 128, p.H, p.B), dim3(kBwdThreads),
                      smem1, stream, p));
  CUDA_TRY(launch_pdl(attn_bwd_dq_tc_kernel<DP>, dim3((p.Sq + 127) / 128, p.H, p.B), dim3(kBwdThreads),
                      smem2, stream, p));
  return 0;
}

// returns 1 if the shape is not covered (caller falls back to the mma.sync kernels); delta must
// already hold rowsum(dO * O)
int attn_bwd_tc(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                const float* delta, void* dq, void* dk, void* dv, int B, int H, int Sq, int Skv, int D,
                long long ldq, long long ldk, long long ldv, long long ldo, float scale,
                cudaStream_t stream) {
  if (D % 8 != 0 || D > 64) return 1;
  static AttnBwdTcParams p;
  memset(&p, 0, sizeof(p));
  if (int rc = enc_map(&p.q_map, q, H * D, Sq, B, ldq)) return rc;
  if (int rc = enc_map(&p.k_map, k, H * D, Skv, B, ldk)) return rc;
  if (int rc = enc_map(&p.v_map, v, H * D, Skv, B, ldv)) return rc;
  if (int rc = enc_map(&p.do_map, dout, H * D, Sq, B, ldo)) return rc;
  p.dq = reinterpret_cast<bf16*>(dq);
  p.dk = reinterpret_cast<bf16*>(dk);
  p.dv = reinterpret_cast<bf16*>(dv);
  p.lse = lse;
  p.delta = delta;
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.D = D;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.scale = scale;
  const int dp = (D + 15) / 16 * 16;
  switch (dp) {
    case 16: return launch_bwd_tc<16>(p, stream);
    case 32: return launch_bwd_tc<32>(p, stream);
    case 48: return launch_bwd_tc<48>(p, stream);
    case 64: return launch_bwd_tc<64>(p, stream);
    default: return 1;
  }
}

}  // namespace pcm
