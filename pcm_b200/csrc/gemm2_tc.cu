// 2-CTA (cta_group::2) variant of the tcgen05 implicit GEMM for the large layers.
//
// A CTA pair (cluster 2x1x1, two SMs of one TPC) owns a 256 x block_n output tile: CTA r loads its
// own 128 activation rows and only HALF of the weight tile (rows n0 + r*block_n/2 ...); the leader
// CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads A from each CTA's own shared memory
// and the two halves of B from both.  Per SM this cuts the L2 -> SM operand traffic per MMA from
// (128 + bn) to (128 + bn/2) rows per k-block -- the 1-CTA kernel is bound by that traffic
// (~40 B/clk/SM) on the big convolutions -- and every CTA keeps its 128 accumulator rows in its own
// TMEM, so the epilogue (gemm_epilogue.cuh) is unchanged.
//   warp 0      TMA producer (each CTA; both signal the LEADER's full barrier)
//   warp 1      MMA issuer (leader CTA only); commits are multicast to both CTAs' barriers
//   warps 2-9   epilogue (each CTA on its own accumulator half)
#include <stdlib.h>
#include "common.cuh"
#include "host_common.h"
#include "gemm_params.h"
#include "gemm_epilogue.cuh"

namespace pcm {

constexpr int kGemm2Threads = 320;
constexpr int kMaxStages2 = 8;
constexpr int kStaging2 = 2 * 128 * 32 * 4;
constexpr int kSmemLimit2 = 227 * 1024 - 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a local shared address) inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of BOTH CTAs once all prior MMAs of this thread are complete
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's barrier (peer bit cleared)
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0,
                                             int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)),
        "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const void* desc, uint64_t* bar, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)),
        "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// GemmParams conventions for this kernel: b_maps have box rows block_n / 2; tiles_m counts PAIR tiles
// (256 rows); num_stages / block_n as usual; ksplit must be 1.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
pcm_gemm2_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages2];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages2];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int S = p.num_stages;
  const int half_n = p.block_n >> 1;
  const uint32_t stage_bytes = kATileBytes + half_n * 128;
  const int num_tiles = p.tiles_m * p.tiles_n;  // pair tiles
  const int ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < PCM_MAX_ASRC; ++i) tma_prefetch_desc(&p.a_maps[i]);
    for (int i = 0; i < PCM_MAX_BSRC; ++i) tma_prefetch_desc(&p.b_maps[i]);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);   // leader: one arrive.expect_tx per stage (covers both CTAs' bytes)
      mbar_init(&empty_bar[i], 1);  // one multicast commit per stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 16);  // leader: 8 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(&tmem_base_smem, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  griddep_sync();

  if (warp == 0) {
    // ===================== TMA producer (each CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cl; tile < num_tiles; tile += ncl) {
        const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
        const int m0 = (tm * 2 + static_cast<int>(rank)) * 128;
        const int n0 = tn * p.block_n + static_cast<int>(rank) * half_n;
        int b0 = 0, h0 = 0;
        if (!p.lin) {
          b0 = m0 / p.geoHW;
          h0 = (m0 - b0 * p.geoHW) / p.geoW;
        }
        for (int e = 0; e < p.num_prog; ++e) {
          const KEntry en = p.prog[e];
          for (int c = 0; c < en.nchunks; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
            uint8_t* sa = smem + stage * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            if (p.lin)
              tma2_load_4d(sa, &p.a_maps[en.a_map], &full_bar[stage], en.a_c0 + c * 64, m0, 0, 0);
            else
              tma2_load_4d(sa, &p.a_maps[en.a_map], &full_bar[stage], en.a_c0 + c * 64, en.dw,
                           h0 + en.dh, b0);
            tma2_load_2d(sb, &p.b_maps[en.b_map], &full_bar[stage], en.b_k0 + c * 64, n0);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cl; tile < num_tiles; tile += ncl) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bd = umma_desc_sw128(b_addr + k * 32, 16, 1024);
            umma2_f16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma2_commit_mc(&empty_bar[stage]);  // frees the stage in BOTH CTAs
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc(&tfull_bar[acc]);  // accumulator ready in BOTH CTAs
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (each CTA on its own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int grp = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;
    const int cg = et & 3;
    const int r0 = et >> 2;
    float* sb = reinterpret_cast<float*>(smem + S * stage_bytes) + grp * (128 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cl; tile < num_tiles; tile += ncl) {
      const int tmp = tile / p.tiles_n, tn = tile - tmp * p.tiles_n;
      const int tm = tmp * 2 + static_cast<int>(rank);
      const int n0 = tn * p.block_n;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      // release: arrive on the LEADER's tmem-empty barrier (16 arrivals per phase)
      const uint32_t remote = mapa(smem_u32(&tempty_bar[acc]), 0);
      gemm_epilogue_tile(p, tm, n0, taddr, sb, lane, row, grp, cg, r0, &tfull_bar[acc], acc_phase,
                         [remote]() { mbar_arrive_cluster(remote); });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// host: returns 1 when the problem should use the 1-CTA kernel
int launch_gemm2(GemmParams& p, const pcm_gemm_desc* d, cudaStream_t stream) {
  // Opt-in (PCM_2CTA=1): parity-green, but on B200 it is not yet faster than the 1-CTA kernel --
  // both are bound by the TMA pipeline depth that fits in shared memory.
  static const bool enabled = getenv("PCM_2CTA") != nullptr;
  if (!enabled || p.ksplit > 1) return 1;
  for (int i = 0; i < d->num_b; ++i)
    if (d->b[i].kblocked) return 1;   // the pair kernel keeps 2-D row-major weight maps
  const int bn = p.block_n;
  if (bn < 64 || (bn % 32) != 0) return 1;
  const int tiles_m1 = (p.M + 127) / 128;
  if (tiles_m1 < 2) return 1;
  // worth it only for operand-traffic-bound launches: enough K and enough tiles
  if (p.num_kblocks < 8) return 1;
  const int pair_m = (tiles_m1 + 1) / 2;
  if (pair_m * p.tiles_n < 32) return 1;
  // re-encode the B maps with half-height boxes
  GemmParams q = p;
  for (int i = 0; i < PCM_MAX_BSRC; ++i) {
    const pcm_bsrc& b = d->b[i < d->num_b ? i : 0];
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(b.K), static_cast<cuuint64_t>(b.N)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(b.ld) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(bn / 2)};
    cuuint32_t estr[2] = {1, 1};
    if (int rc = encode_tmap(&q.b_maps[i], b.ptr, 2, dims, strides, box, estr)) return rc;
  }
  q.tiles_m = pair_m;
  const int stage_bytes = kATileBytes + (bn / 2) * 128;
  int S = (kSmemLimit2 - 1024 - kStaging2) / stage_bytes;
  if (S > kMaxStages2) S = kMaxStages2;
  q.num_stages = S;
  const size_t smem = static_cast<size_t>(S) * stage_bytes + kStaging2 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(pcm_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSmemLimit2));
    attr_set = true;
  }
  const int tiles = q.tiles_m * q.tiles_n;
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // cluster shape comes from __cluster_dims__(2, 1, 1) on the kernel
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, pcm_gemm2_kernel, q));
  return 0;
}

}  // namespace pcm
