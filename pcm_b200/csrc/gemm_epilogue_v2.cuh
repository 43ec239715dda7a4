// Epilogue v2 of the tcgen05 implicit GEMM: thread = accumulator row end to end, no transposition.
//
// Each of the 8 epilogue warps owns 32 accumulator rows (its TMEM lane quarter); the two groups of 4
// warps take the even / odd 32-column chunks.  Per chunk a warp
//   * reads its [32 x 32] fp32 block from TMEM (tcgen05.ld 32x32b.x32, one row per lane),
//   * adds alpha / bias / per-image row vector / residual, packs to bf16,
//   * writes its 64-byte row into a SWIZZLE_64B staging box and lets ONE lane store the box with TMA
//     (cp.async.bulk.tensor, bulk groups double buffered) - rows past M / columns past N are clipped
//     by the tensor map, so there is no tail code.
// The residual block arrives the same way: a TMA load into a per-warp [32 x 32] box, requested TWO
// chunks ahead across tiles, completion on a per-warp mbarrier - its HBM latency never reaches a
// register scoreboard.  No CTA-level barrier is used; ~130 instructions per chunk instead of ~350 in
// the two-phase v1 (which bounds the K <= 1472 Linear layers).  Used for bf16 outputs without
// activation / split-K whose row mapping fits a 32-row TMA box (see launch_gemm).
#pragma once
#include "gemm_params.h"

namespace pcm {

constexpr int kEpi2BytesPerWarp = 4 * 2048;  // 2 store boxes + 2 residual boxes of 32 x 64 B

// byte offset of 16-byte chunk c of row r inside a [32 rows x 64 B] SWIZZLE_64B box
__device__ __forceinline__ uint32_t sw64(int r, int c) {
  return static_cast<uint32_t>(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

struct Epi2Iter {  // enumerates the chunks this warp's group handles, tile after tile
  int item, jj;
  int my;              // chunks of the group in tile `item`
  int n_first;         // first column of the group's first chunk
  int cw, ch, cb;      // TMA coordinates of the warp's 32-row block in that tile
};

template <class Release>
__device__ __forceinline__ void gemm_epilogue_v2(const GemmParams& p, int q, int grp, int lane,
                                                 uint32_t tmem_base, uint8_t* ebuf, uint64_t* rbar,
                                                 uint64_t* tfull_bar, Release release) {
  // q = TMEM lane quarter this warp may access (warp id & 3), grp = chunk parity (0 / 1)
  const uint32_t st_u = smem_u32(ebuf), rs_u = st_u + 2 * 2048;
  uint8_t* rs_ptr = ebuf + 2 * 2048;
  const bool has_bias = p.bias != nullptr, has_rv = p.rowvec != nullptr, has_res = p.residual != nullptr;
  const int num_items = p.tiles_m * p.tiles_n;  // ksplit == 1

  // per-tile quantities (integer divisions) are computed once per tile, not per chunk
  auto load_tile = [&](Epi2Iter& it) {
    if (it.item >= num_items) {
      it.my = 0;
      return;
    }
    int tm = it.item / p.tiles_n;
    const int tn = it.item - tm * p.tiles_n;
    if (p.dep_a_map >= 0) tm = p.tiles_m - 1 - tm;  // same visiting order as the producer / MMA warps
    const int ncols = min(p.block_n, p.N - tn * p.block_n);
    const int nch = (ncols + 31) >> 5;
    it.my = nch > grp ? (nch - grp + 1) >> 1 : 0;
    it.n_first = tn * p.block_n + grp * 32;
    const int m = tm * 128 + q * 32;
    it.cb = m / p.epiHW;
    const int r = m - it.cb * p.epiHW;
    it.ch = r / p.epiW;
    it.cw = r - it.ch * p.epiW;
  };
  auto next_chunk = [&](Epi2Iter& it) {  // advance to the group's next chunk (skipping empty tiles)
    ++it.jj;
    while (it.item < num_items && it.jj >= it.my) {
      it.item += gridDim.x;
      it.jj = 0;
      load_tile(it);
    }
  };
  auto issue_res = [&](const Epi2Iter& it, int seq) {
    if (it.item >= num_items) return;
    if (lane == 0) {
      mbar_arrive_expect_tx(&rbar[seq & 1], 2048);
      tma_load_4d(rs_ptr + (seq & 1) * 2048, &p.res_map, &rbar[seq & 1], it.n_first + it.jj * 64, it.cw, it.ch,
                  it.cb);
    }
  };

  Epi2Iter pf;
  pf.item = blockIdx.x;
  pf.jj = -1;
  load_tile(pf);
  next_chunk(pf);  // -> first chunk of the group
  if (has_res) {   // two residual boxes in flight from the start
    issue_res(pf, 0);
    next_chunk(pf);
    issue_res(pf, 1);
    next_chunk(pf);
  }

  int seq = 0, acc = 0;
  uint32_t acc_phase = 0;
  Epi2Iter cur;
  for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
    cur.item = item;
    cur.jj = 0;
    load_tile(cur);
    int tm = item / p.tiles_n;
    if (p.dep_a_map >= 0) tm = p.tiles_m - 1 - tm;
    const int my = cur.my;
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
    if (my == 0) {  // nothing to read for this group: hand the accumulator back (paced by tfull)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release(acc);
    }
    const int m_row = tm * 128 + q * 32 + lane;
    const bool valid = m_row < p.M;
    const bf16* rvp = nullptr;
    if (has_rv && valid) rvp = p.rowvec + static_cast<long long>(m_row / p.epiHW) * p.rowvec_ld;
#pragma unroll 1
    for (int jj = 0; jj < my; ++jj) {
      const int n = cur.n_first + jj * 64;
      const int cw = cur.cw, ch = cur.ch, cb = cur.cb;
      uint32_t v[32];
      tmem_ld_32x32(taddr + (grp + 2 * jj) * 32, v);
      // operands that do not depend on the accumulator: issued under the TMEM load
      float4 bia[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bia[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_bias && n + 4 * i + 4 <= p.N) bia[i] = *reinterpret_cast<const float4*>(p.bias + n + 4 * i);
      }
      uint4 rv4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rv4[i] = make_uint4(0, 0, 0, 0);
        if (rvp != nullptr && n + 8 * i + 8 <= p.N) rv4[i] = *reinterpret_cast<const uint4*>(rvp + n + 8 * i);
      }
      tmem_ld_wait();
      if (jj == my - 1) {  // last TMEM read of this warp for the tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release(acc);
      }
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[4 * i] += bia[i].x; f[4 * i + 1] += bia[i].y; f[4 * i + 2] += bia[i].z; f[4 * i + 3] += bia[i].w;
      }
      if (has_rv) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 t;
          t = unpack_bf16x2(rv4[i].x); f[8 * i] += t.x; f[8 * i + 1] += t.y;
          t = unpack_bf16x2(rv4[i].y); f[8 * i + 2] += t.x; f[8 * i + 3] += t.y;
          t = unpack_bf16x2(rv4[i].z); f[8 * i + 4] += t.x; f[8 * i + 5] += t.y;
          t = unpack_bf16x2(rv4[i].w); f[8 * i + 6] += t.x; f[8 * i + 7] += t.y;
        }
      }
      if (has_res) {
        mbar_wait(&rbar[seq & 1], (seq >> 1) & 1);  // residual box of this chunk landed
        const uint32_t rb = rs_u + (seq & 1) * 2048;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 raw = lds128(rb + sw64(lane, i));
          float2 t;
          t = unpack_bf16x2(__float_as_uint(raw.x)); f[8 * i] += t.x; f[8 * i + 1] += t.y;
          t = unpack_bf16x2(__float_as_uint(raw.y)); f[8 * i + 2] += t.x; f[8 * i + 3] += t.y;
          t = unpack_bf16x2(__float_as_uint(raw.z)); f[8 * i + 4] += t.x; f[8 * i + 5] += t.y;
          t = unpack_bf16x2(__float_as_uint(raw.w)); f[8 * i + 6] += t.x; f[8 * i + 7] += t.y;
        }
      }
      // the store box of chunk seq-2 must have been read out before it is overwritten
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      const uint32_t sb = st_u + (seq & 1) * 2048;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        sts128(sb + sw64(lane, i), pack_bf16x2(f[8 * i], f[8 * i + 1]), pack_bf16x2(f[8 * i + 2], f[8 * i + 3]),
               pack_bf16x2(f[8 * i + 4], f[8 * i + 5]), pack_bf16x2(f[8 * i + 6], f[8 * i + 7]));
      fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&p.out_map, sb, n, cw, ch, cb);
        tma_store_commit();
      }
      if (has_res) {  // every lane has consumed residual box seq: refill it with chunk seq + 2
        issue_res(pf, seq + 2);
        next_chunk(pf);
      }
      ++seq;
    }
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (lane == 0) tma_store_wait_all();  // stores complete before the CTA (and its smem) goes away
}

}  // namespace pcm
