// tcgen05 flash-attention forward for the SD1.5 head sizes d <= 80 (d = 40 at 64x64 / d = 80 at
// 32x32 carry 96 % of the attention FLOPs; d = 160 keeps the mma.sync kernel in attn.cu).
//
// One CTA = 128 query rows of one (batch, head); KV tiles of 128 keys.
//   warp 0      TMA producer: Q once, (K_j, V_j) through a 2-stage ring; 3-D tensor maps
//               (C, S, B) so rows past the sequence end are zero filled per batch
//   warp 1      tcgen05.mma issuer: S_j = Q K_j^T into TMEM (two S buffers), O_j = P_j V_j into TMEM
//   warps 2-5   softmax: thread r owns query row r == TMEM lane r: tcgen05.ld its 128 scores, online
//               max / exp2 / sum in registers, P_j -> shared memory as the K-major SWIZZLE_128B A
//               operand of the PV product, then folds O_j (TMEM) into its fp32 register accumulator
//               with the running rescale.  No shuffles: a row never leaves its thread.
// Head slices are 64-column TMA boxes at column h*d; for d = 40 the QK^T product runs k = 0..47
// and the 8 stray columns of Q (next head) are zeroed in shared memory once, so K needs no fix-up;
// V's stray columns only produce output columns that are never stored.
#include <stdlib.h>
#include "common.cuh"
#include "host_common.h"
#include "../../include/pcm_b200.h"

namespace pcm {

struct AttnTcParams {
  CUtensorMap q_map, k_map, v_map;  // (C, S, B), box (64, 128, 1)
  bf16* out;
  float* lse;
  int B, H, Sq, Skv, D;
  long long ldo;
  float scale;
};

constexpr int kTcThreads = 192;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr int kBoxBytes = 128 * 128;  // 128 rows x 64 bf16

// dynamic smem: [Q: NBOX boxes][K stage 0..1: NBOX boxes each][V stage 0..1][P: 2 boxes]
template <int DP>  // padded head dim (multiple of 16): MMA k extent of QK^T and N of PV
__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  constexpr int NBOX = (DP + 63) / 64;
  constexpr int KSTEPS = DP / 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NBOX * kBoxBytes;      // [2][NBOX]
  uint8_t* sV = sK + 2 * NBOX * kBoxBytes;  // [2][NBOX]
  uint8_t* sP = sV + 2 * NBOX * kBoxBytes;  // 2 boxes (keys 0-63, 64-127)
  __shared__ __align__(8) uint64_t q_full, q_ready, kv_full[2], kv_free[2], s_full[2], s_free[2],
      p_full, p_free, o_full, o_free;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int ntiles = (p.Skv + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    tma_prefetch_desc(&p.v_map);
    mbar_init(&q_full, 1);
    mbar_init(&q_ready, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
    }
    mbar_init(&p_full, 128);
    mbar_init(&p_free, 1);
    mbar_init(&o_full, 1);
    mbar_init(&o_free, 128);
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
  griddep_sync();  // PDL: everything above overlapped the previous kernel's tail
  const uint32_t tS = tmem_base;        // S buffers at columns 0 and 128
  const uint32_t tO = tmem_base + 256;  // O tile

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, NBOX * kBoxBytes);
      for (int x = 0; x < NBOX; ++x)
        tma_load_4d(sQ + x * kBoxBytes, &p.q_map, &q_full, h * p.D + x * 64, q0, b, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_free[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * NBOX * kBoxBytes);
        for (int x = 0; x < NBOX; ++x) {
          tma_load_4d(sK + (st * NBOX + x) * kBoxBytes, &p.k_map, &kv_full[st], h * p.D + x * 64,
                      j * 128, b, 0);
          tma_load_4d(sV + (st * NBOX + x) * kBoxBytes, &p.v_map, &kv_full[st], h * p.D + x * 64,
                      j * 128, b, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, DP, 0, 1);  // B = V is MN-major
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        mbar_wait(&s_free[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * NBOX * kBoxBytes);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          const uint32_t o = (k >> 2) * kBoxBytes + (k & 3) * 32;
          umma_f16(tS + st * 128, umma_desc_sw128(q_addr + o, 16, 1024),
                   umma_desc_sw128(k_addr + o, 16, 1024), idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
      };
      mbar_wait(&q_ready, 0);  // Q landed and its stray columns are zeroed (softmax warps)
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_qk(j + 1);
        const int st = j & 1;
        mbar_wait(&p_full, j & 1);
        mbar_wait(&o_free, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + st * NBOX * kBoxBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 keys per step
          const uint64_t ad = umma_desc_sw128(p_addr + (k >> 2) * kBoxBytes + (k & 3) * 32, 16, 1024);
          const uint64_t bd = umma_desc_sw128(v_addr + k * 2048, kBoxBytes, 1024);
          umma_f16(tO, ad, bd, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(&o_full);
        umma_commit(&p_free);
        umma_commit(&kv_free[st]);
      }
    }
  } else {
    // ===================== softmax / correction: thread = query row =====================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    float oacc[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) oacc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    // zero Q's stray columns [D, DP) (d = 40: columns 40..47) once Q has landed
    mbar_wait(&q_full, 0);
    if (p.D < DP) {
      const int c0 = p.D >> 3;  // first stray 16-byte chunk (D % 8 == 0)
#pragma unroll
      for (int ch = 0; ch < DP / 8; ++ch) {
        if (ch < c0) continue;
        uint8_t* dst = sQ + (ch >> 3) * kBoxBytes + row * 128 + (((ch & 7) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async();
    }
    mbar_arrive(&q_ready);  // the MMA warp's first QK^T waits for all 128 rows

    for (int j = 0; j < ntiles; ++j) {
      const int st = j & 1;
      // s_full[st] completes once per use of buffer st
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      float s[128];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_off + st * 128 + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[cc * 32 + i] = __uint_as_float(v[i]);
      }
      tc_fence_before();
      mbar_arrive(&s_free[st]);  // buffer st may be overwritten by QK^T of tile j + 2
      const int kbase = j * 128;
      // 4 independent max / sum chains: a single 128-long dependent chain costs ~4 cycles per link
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        s[i] = (kbase + i < p.Skv) ? s[i] * c : -INFINITY;
        mx4[i & 3] = fmaxf(mx4[i & 3], s[i]);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2(m_run - m_new);
      m_run = m_new;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      if (j > 0) mbar_wait(&p_free, (j - 1) & 1);  // PV of tile j-1 finished reading P
      const uint32_t sp_row = smem_u32(sP) + row * 128;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = fast_exp2(s[ch * 8 + e] - m_new);
          rs4[e & 3] += pv[e];
        }
        uint4 u;
        u.x = pack_bf16x2(pv[0], pv[1]);
        u.y = pack_bf16x2(pv[2], pv[3]);
        u.z = pack_bf16x2(pv[4], pv[5]);
        u.w = pack_bf16x2(pv[6], pv[7]);
        sts128(sp_row + (ch >> 3) * kBoxBytes + (((ch & 7) ^ (row & 7)) << 4), u.x, u.y, u.z, u.w);
      }
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      fence_proxy_async();
      mbar_arrive(&p_full);
      // fold the previous tile's O (computed against the previous max) into the accumulator
      if (j > 0) {
        mbar_wait(&o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < DP / 16; ++cc) {
          uint32_t v[16];
          tmem_ld_32x16(tO + lane_off + cc * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            oacc[cc * 16 + i] = oacc[cc * 16 + i] * alpha_prev + __uint_as_float(v[i]);
        }
        tc_fence_before();
        mbar_arrive(&o_free);
      }
      alpha_prev = alpha;
    }
    // last tile
    mbar_wait(&o_full, (ntiles - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < DP / 16; ++cc) {
      uint32_t v[16];
      tmem_ld_32x16(tO + lane_off + cc * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        oacc[cc * 16 + i] = oacc[cc * 16 + i] * alpha_prev + __uint_as_float(v[i]);
    }
    tc_fence_before();
    const int qrow = q0 + row;
    if (qrow < p.Sq) {
      const float inv = 1.f / l_run;
      bf16* orow = p.out + (static_cast<long long>(b) * p.Sq + qrow) * p.ldo + h * p.D;
#pragma unroll
      for (int ch = 0; ch < DP / 8; ++ch) {
        if (ch * 8 >= p.D) break;
        uint4 u;
        u.x = pack_bf16x2(oacc[ch * 8] * inv, oacc[ch * 8 + 1] * inv);
        u.y = pack_bf16x2(oacc[ch * 8 + 2] * inv, oacc[ch * 8 + 3] * inv);
        u.z = pack_bf16x2(oacc[ch * 8 + 4] * inv, oacc[ch * 8 + 5] * inv);
        u.w = pack_bf16x2(oacc[ch * 8 + 6] * inv, oacc[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = u;
      }
      if (p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qrow] = m_run + log2f(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------
// v2 (d <= 64): 256 query rows per CTA as two independent 128-row tiles that share every K/V
// tile.  Two softmax warpgroups (one per query tile) run concurrently so each SM sub-partition
// always has two warps to interleave (MUFU.EX2 latency is hidden), and K/V smem traffic per
// query halves.  setmaxnreg moves registers from the TMA/MMA warpgroup to the softmax warps
// (each thread keeps a full 128-score row plus its output row in registers).
//   warpgroup 0: warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 idle
//   warpgroup 1: softmax for query tile 0      warpgroup 2: softmax for query tile 1
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512).
// ------------------------------------------------------------------------------------------
constexpr int kV2Threads = 384;
constexpr int kV2Stages = 3;

template <int DP>
__global__ void __launch_bounds__(kV2Threads, 1) attn_fwd_tc2_kernel(const __grid_constant__ AttnTcParams p) {
  static_assert(DP <= 64, "v2 handles one 64-column box per operand");
  constexpr int KSTEPS = DP / 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                              // [2] boxes (query tile 0 / 1)
  uint8_t* sK = sQ + 2 * kBoxBytes;                // [kV2Stages]
  uint8_t* sV = sK + kV2Stages * kBoxBytes;        // [kV2Stages]
  uint8_t* sP = sV + kV2Stages * kBoxBytes;        // [2 query tiles][2 boxes]
  __shared__ __align__(8) uint64_t q_full, q_ready, kv_full[kV2Stages], kv_free[kV2Stages];
  __shared__ __align__(8) uint64_t s_full[2], s_free[2], p_full[2], p_free[2], o_full[2], o_free[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 256;
  const int ntiles = (p.Skv + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.k_map);
    tma_prefetch_desc(&p.v_map);
    mbar_init(&q_full, 1);
    mbar_init(&q_ready, 256);
    for (int i = 0; i < kV2Stages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_free[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 128);
    }
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
  griddep_sync();  // PDL: everything above overlapped the previous kernel's tail

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(&q_full, 2 * kBoxBytes);
      tma_load_4d(sQ, &p.q_map, &q_full, h * p.D, q0, b, 0);
      tma_load_4d(sQ + kBoxBytes, &p.q_map, &q_full, h * p.D, q0 + 128, b, 0);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&kv_free[st], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kBoxBytes);
        tma_load_4d(sK + st * kBoxBytes, &p.k_map, &kv_full[st], h * p.D, j * 128, b, 0);
        tma_load_4d(sV + st * kBoxBytes, &p.v_map, &kv_full[st], h * p.D, j * 128, b, 0);
        if (++st == kV2Stages) { st = 0; ph ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, DP, 0, 1);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      mbar_wait(&q_ready, 0);
      int st = 0, stv = 0;     // K stage of tile j, V stage of tile j-1
      uint32_t ph = 0;
      for (int j = 0; j <= ntiles; ++j) {
        if (j < ntiles) {
          mbar_wait(&kv_full[st], ph);
          const uint32_t k_addr = smem_u32(sK + st * kBoxBytes);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&s_free[t], (j & 1) ^ 1);   // softmax t copied S(j-1) to registers
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_f16(tmem_base + t * 128, umma_desc_sw128(q_addr + t * kBoxBytes + k * 32, 16, 1024),
                       umma_desc_sw128(k_addr + k * 32, 16, 1024), idesc_qk, k != 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
          }
        }
        if (j > 0) {
          const int jj = j - 1;
          const uint32_t v_addr = smem_u32(sV + stv * kBoxBytes);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&p_full[t], jj & 1);
            mbar_wait(&o_free[t], (jj & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t ad = umma_desc_sw128(
                  p_addr + t * 2 * kBoxBytes + (k >> 2) * kBoxBytes + (k & 3) * 32, 16, 1024);
              const uint64_t bd = umma_desc_sw128(v_addr + k * 2048, kBoxBytes, 1024);
              umma_f16(tmem_base + 256 + t * 128, ad, bd, idesc_pv, k != 0 ? 1u : 0u);
            }
            umma_commit(&o_full[t]);
            umma_commit(&p_free[t]);
          }
          umma_commit(&kv_free[stv]);
          if (++stv == kV2Stages) stv = 0;
        }
        if (j < ntiles) {
          if (++st == kV2Stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===================== softmax warpgroups =====================
    const int t = (warp - 4) >> 2;  // query tile
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_off;
    uint8_t* sPt = sP + t * 2 * kBoxBytes;
    const float c = p.scale * 1.4426950408889634f;
    float oacc[DP];
#pragma unroll
    for (int i = 0; i < DP; ++i) oacc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    mbar_wait(&q_full, 0);
    if (p.D < DP) {
      const int c0 = p.D >> 3;
#pragma unroll
      for (int ch = 0; ch < DP / 8; ++ch) {
        if (ch < c0) continue;
        uint8_t* dst = sQ + t * kBoxBytes + row * 128 + (((ch & 7) ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async();
    }
    mbar_arrive(&q_ready);

    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float s[128];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32(tS + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[cc * 32 + i] = __uint_as_float(v[i]);
      }
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      const int kbase = j * 128;
      if (kbase + 128 > p.Skv) {  // partial last tile: mask keys past the sequence end
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kbase + i >= p.Skv) s[i] = -INFINITY;
      }
      // 4 independent max / sum chains: a single 128-long dependent chain costs ~4 cycles per link
      float mx4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int i = 4; i < 128; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], s[i]);
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m_run, mx);        // raw-score domain (c > 0)
      const float alpha = fast_exp2((m_run - m_new) * c);
      m_run = m_new;
      const float mc = m_new * c;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      if (j > 0) mbar_wait(&p_free[t], (j - 1) & 1);
      const uint32_t sp_row = smem_u32(sPt) + row * 128;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = fast_exp2(fmaf(s[ch * 8 + e], c, -mc));
          rs4[e & 3] += pv[e];
        }
        uint4 u;
        u.x = pack_bf16x2(pv[0], pv[1]);
        u.y = pack_bf16x2(pv[2], pv[3]);
        u.z = pack_bf16x2(pv[4], pv[5]);
        u.w = pack_bf16x2(pv[6], pv[7]);
        sts128(sp_row + (ch >> 3) * kBoxBytes + (((ch & 7) ^ (row & 7)) << 4), u.x, u.y, u.z, u.w);
      }
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      fence_proxy_async();
      mbar_arrive(&p_full[t]);
      if (j > 0) {
        mbar_wait(&o_full[t], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < DP / 16; ++cc) {
          uint32_t v[16];
          tmem_ld_32x16(tO + cc * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            oacc[cc * 16 + i] = oacc[cc * 16 + i] * alpha_prev + __uint_as_float(v[i]);
        }
        tc_fence_before();
        mbar_arrive(&o_free[t]);
      }
      alpha_prev = alpha;
    }
    mbar_wait(&o_full[t], (ntiles - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < DP / 16; ++cc) {
      uint32_t v[16];
      tmem_ld_32x16(tO + cc * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        oacc[cc * 16 + i] = oacc[cc * 16 + i] * alpha_prev + __uint_as_float(v[i]);
    }
    tc_fence_before();
    const int qrow = q0 + t * 128 + row;
    if (qrow < p.Sq) {
      const float inv = 1.f / l_run;
      bf16* orow = p.out + (static_cast<long long>(b) * p.Sq + qrow) * p.ldo + h * p.D;
#pragma unroll
      for (int ch = 0; ch < DP / 8; ++ch) {
        if (ch * 8 >= p.D) break;
        uint4 u;
        u.x = pack_bf16x2(oacc[ch * 8] * inv, oacc[ch * 8 + 1] * inv);
        u.y = pack_bf16x2(oacc[ch * 8 + 2] * inv, oacc[ch * 8 + 3] * inv);
        u.z = pack_bf16x2(oacc[ch * 8 + 4] * inv, oacc[ch * 8 + 5] * inv);
        u.w = pack_bf16x2(oacc[ch * 8 + 6] * inv, oacc[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = u;
      }
      if (p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qrow] = m_run * c + log2f(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int DP>
static int launch_tc2(const AttnTcParams& p, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(2 + 2 * kV2Stages + 4) * kBoxBytes + 1024;
  static bool set = false;
  if (!set) {
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc2_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    set = true;
  }
  dim3 grid((p.Sq + 255) / 256, p.H, p.B);
  CUDA_TRY(launch_pdl(attn_fwd_tc2_kernel<DP>, dim3(grid), dim3(kV2Threads), smem, stream, p));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

template <int DP>
static int launch_tc(const AttnTcParams& p, cudaStream_t stream) {
  constexpr int NBOX = (DP + 63) / 64;
  const size_t smem = static_cast<size_t>(5 * NBOX + 2) * kBoxBytes + 1024;
  static bool set = false;
  if (!set) {
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    set = true;
  }
  dim3 grid((p.Sq + 127) / 128, p.H, p.B);
  CUDA_TRY(launch_pdl(attn_fwd_tc_kernel<DP>, dim3(grid), dim3(kTcThreads), smem, stream, p));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int encode_seq_map(CUtensorMap* map, const void* ptr, int C, int S, int B, long long ld) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(S),
                        static_cast<cuuint64_t>(B), 1};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * S,
                           static_cast<cuuint64_t>(ld) * 2 * S * B};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_tmap(map, ptr, 4, dims, strides, box, estr);
}

// returns 1 if this shape is not handled by the tcgen05 kernel (caller falls back to attn.cu)
int attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H,
                int Sq, int Skv, int D, long long ldq, long long ldk, long long ldv, long long ldo,
                float scale, cudaStream_t stream) {
  if (D % 8 != 0 || D > 80) return 1;
  static AttnTcParams p;
  memset(&p, 0, sizeof(p));
  // logical row width seen by the maps: all heads (columns past it are zero filled)
  if (int rc = encode_seq_map(&p.q_map, q, H * D, Sq, B, ldq)) return rc;
  if (int rc = encode_seq_map(&p.k_map, k, H * D, Skv, B, ldk)) return rc;
  if (int rc = encode_seq_map(&p.v_map, v, H * D, Skv, B, ldv)) return rc;
  p.out = reinterpret_cast<bf16*>(out);
  p.lse = lse;
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.D = D;
  p.ldo = ldo;
  p.scale = scale;
  const int dp = (D + 15) / 16 * 16;
  static const bool v1only = getenv("PCM_ATTN_V1") != nullptr;
  if (!v1only && Sq >= 256) {
    switch (dp) {
      case 16: return launch_tc2<16>(p, stream);
      case 32: return launch_tc2<32>(p, stream);
      case 48: return launch_tc2<48>(p, stream);
      case 64: return launch_tc2<64>(p, stream);
      default: break;
    }
  }
  switch (dp) {
    case 16: return launch_tc<16>(p, stream);
    case 32: return launch_tc<32>(p, stream);
    case 48: return launch_tc<48>(p, stream);
    case 64: return launch_tc<64>(p, stream);
    case 80: return launch_tc<80>(p, stream);
    default: return 1;
  }
}

}  // namespace pcm
