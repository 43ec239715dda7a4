// GroupNorm(+SiLU) and LayerNorm, forward and input-gradient, on NHWC bf16 activations.
// HBM-bound kernels: 16-byte vector loads along the contiguous channel dimension, fp32 statistics,
// warp/shared-memory reductions, one bf16 rounding at the output.
//
// Replaces the ATen group_norm / layer_norm / silu launches inside diffusers' ResnetBlock2D,
// Transformer2DModel and BasicTransformerBlock (called from train_pcm_lora_sd15.py:1192-1198,
// 1219-1244, 1263-1268) and their autograd backward (:1296).  GroupNorm reads an optional second
// source so the up-block skip concat torch.cat([h, res], dim=1) is never materialised twice.
#include "common.cuh"
#include "host_common.h"
#include "../../include/pcm_b200.h"

namespace pcm {

// ------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[b, g] = (sum, sumsq) over HW x (C/G) elements
// ------------------------------------------------------------------------------------------
constexpr int kGnMaxC = 2560;
constexpr int kGnStage = 4096;   // ny * C <= 4096 floats of per-thread partials (see gn_launch_cfg)

__device__ __forceinline__ const bf16* gn_src(const bf16* x1, const bf16* x2, int C1, int C2,
                                              long long pix, int c) {
  return c < C1 ? x1 + pix * C1 + c : x2 + pix * C2 + (c - C1);
}

__global__ void gn_stats_kernel(const bf16* __restrict__ x1, const bf16* __restrict__ x2, int C1,
                                int C2, int HW, int G, int pix_per_block,
                                float* __restrict__ stats) {
  griddep_sync();
  // per-thread partials are staged as [ty][channel] (plain stores) and tree-summed: shared-memory
  // atomics cost ~2 cycles per lane and dominated this kernel
  __shared__ float s_sum[kGnStage];
  __shared__ float s_sq[kGnStage];
  const int C = C1 + C2;
  const int b = blockIdx.y;
  const int nvec = C >> 3;
  const int tx = threadIdx.x % nvec, ty = threadIdx.x / nvec;
  const int ny = blockDim.x / nvec;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  float a[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = q[i] = 0.f;
  if (ty < ny) {
    const int c = tx * 8;
    // four independent 16-byte loads in flight per thread
    for (int p = p0 + ty; p < p1; p += 4 * ny) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pp = p + k * ny;
        u[k] = make_uint4(0, 0, 0, 0);
        if (pp < p1)
          u[k] = *reinterpret_cast<const uint4*>(
              gn_src(x1, x2, C1, C2, static_cast<long long>(b) * HW + pp, c));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_bf16x2(w[i]);
          a[2 * i] += f.x; q[2 * i] += f.x * f.x;
          a[2 * i + 1] += f.y; q[2 * i + 1] += f.y * f.y;
        }
      }
    }
    *reinterpret_cast<float4*>(&s_sum[ty * C + c]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(&s_sum[ty * C + c + 4]) = make_float4(a[4], a[5], a[6], a[7]);
    *reinterpret_cast<float4*>(&s_sq[ty * C + c]) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(&s_sq[ty * C + c + 4]) = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int y = 0; y < ny; ++y) {
      s += s_sum[y * C + ch];
      ss += s_sq[y * C + ch];
    }
    s_sum[ch] = s;   // row 0 now holds the block totals (each thread only touches its own column)
    s_sq[ch] = ss;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int i = 0; i < cpg; ++i) {
      s += s_sum[g * cpg + i];
      ss += s_sq[g * cpg + i];
    }
    atomicAdd(&stats[(b * G + g) * 2], s);
    atomicAdd(&stats[(b * G + g) * 2 + 1], ss);
  }
}

// out = [silu]( (x - mean) * rstd * gamma + beta ), bf16.  Same thread -> channel-vector mapping as
// gn_stats_kernel: each thread folds the statistics of its 8 channels into (scale, shift) once and
// then streams pixels (16-byte load, 8 FMAs, 16-byte store), two pixels in flight.
__global__ void gn_apply_kernel(const bf16* __restrict__ x1, const bf16* __restrict__ x2, int C1,
                                int C2, int HW, int G, int pix_per_block,
                                const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int silu,
                                bf16* __restrict__ out) {
  griddep_sync();
  const int C = C1 + C2;
  const int b = blockIdx.y;
  const int nvec = C >> 3;
  const int tx = threadIdx.x % nvec, ty = threadIdx.x / nvec;
  const int ny = blockDim.x / nvec;
  const int cpg = C / G;
  const float inv_n = 1.f / (static_cast<float>(HW) * cpg);
  const int c = tx * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c + i) / cpg;
    const float mean = stats[(b * G + g) * 2] * inv_n;
    const float var = fmaxf(stats[(b * G + g) * 2 + 1] * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    sc[i] = rstd * gamma[c + i];
    sh[i] = beta[c + i] - mean * sc[i];
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  constexpr int U = 4;  // independent 16-byte loads in flight per thread
  for (int p = p0 + ty; p < p1; p += U * ny) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int pp = p + k * ny;
      if (pp < p1)
        u[k] = *reinterpret_cast<const uint4*>(gn_src(x1, x2, C1, C2, static_cast<long long>(b) * HW + pp, c));
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int pp = p + k * ny;
      if (pp >= p1) break;
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = unpack_bf16x2(w[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = f[i] * sc[i] + sh[i];
        if (silu) y = silu_f(y);
        f[i] = y;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * HW + pp) * C + c) = o;
    }
  }
}

// backward reductions: red[b, g] = (sum gamma*dyh, sum gamma*dyh*xhat), dyh = dy * silu'(pre)
__global__ void gn_bwd_stats_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x1,
                                    const bf16* __restrict__ x2, int C1, int C2, int HW, int G,
                                    int pix_per_block, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float eps, int silu, float* __restrict__ red) {
  griddep_sync();
  __shared__ float s_a[kGnStage];
  __shared__ float s_b[kGnStage];
  const int C = C1 + C2;
  const int b = blockIdx.y;
  const int nvec = C >> 3;
  const int tx = threadIdx.x % nvec, ty = threadIdx.x / nvec;
  const int ny = blockDim.x / nvec;
  const int cpg = C / G;
  const float inv_n = 1.f / (static_cast<float>(HW) * cpg);
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  if (ty < ny) {
    const int c = tx * 8;
    float mean[8], rstd[8], gm[8], bt[8], a[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cpg;
      mean[i] = stats[(b * G + g) * 2] * inv_n;
      const float var = fmaxf(stats[(b * G + g) * 2 + 1] * inv_n - mean[i] * mean[i], 0.f);
      rstd[i] = rsqrtf(var + eps);
      gm[i] = gamma[c + i];
      bt[i] = beta[c + i];
      a[i] = q[i] = 0.f;
    }
    for (int p = p0 + ty; p < p1; p += ny) {
      const long long pix = static_cast<long long>(b) * HW + p;
      const uint4 u = *reinterpret_cast<const uint4*>(gn_src(x1, x2, C1, C2, pix, c));
      const uint4 d = *reinterpret_cast<const uint4*>(dy + pix * C + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 xf = unpack_bf16x2(w[i]);
        const float2 df = unpack_bf16x2(dw[i]);
        const float xs[2] = {xf.x, xf.y}, ds[2] = {df.x, df.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = 2 * i + j;
          const float xh = (xs[j] - mean[k]) * rstd[k];
          float g = ds[j];
          if (silu) g *= dsilu_f(xh * gm[k] + bt[k]);
          g *= gm[k];
          a[k] += g;
          q[k] += g * xh;
        }
      }
    }
    *reinterpret_cast<float4*>(&s_a[ty * C + c]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(&s_a[ty * C + c + 4]) = make_float4(a[4], a[5], a[6], a[7]);
    *reinterpret_cast<float4*>(&s_b[ty * C + c]) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(&s_b[ty * C + c + 4]) = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int y = 0; y < ny; ++y) {
      s += s_a[y * C + ch];
      ss += s_b[y * C + ch];
    }
    s_a[ch] = s;
    s_b[ch] = ss;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f, ss = 0.f;
    for (int i = 0; i < cpg; ++i) {
      s += s_a[g * cpg + i];
      ss += s_b[g * cpg + i];
    }
    atomicAdd(&red[(b * G + g) * 2], s);
    atomicAdd(&red[(b * G + g) * 2 + 1], ss);
  }
}

// dx = rstd * (gamma*dyh - (s1 + xhat*s2)/n) (+ add); written to dx1 (first C1 channels) and
// dx2 (remaining C2 channels).  Per-thread channel constants hoisted like gn_apply_kernel.
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x1,
                                    const bf16* __restrict__ x2, int C1, int C2, int HW, int G,
                                    int pix_per_block, const float* __restrict__ stats,
                                    const float* __restrict__ red, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float eps, int silu,
                                    const bf16* __restrict__ add, bf16* __restrict__ dx1,
                                    bf16* __restrict__ dx2, float* __restrict__ colsum) {
  griddep_sync();
  __shared__ float s_cs[kGnMaxC];
  const int C = C1 + C2;
  const int b = blockIdx.y;
  const int nvec = C >> 3;
  const int tx = threadIdx.x % nvec, ty = threadIdx.x / nvec;
  const int ny = blockDim.x / nvec;
  const int cpg = C / G;
  const float inv_n = 1.f / (static_cast<float>(HW) * cpg);
  const int c = tx * 8;
  if (colsum) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_cs[i] = 0.f;
    __syncthreads();
  }
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float mean[8], rstd[8], gm[8], bt[8], s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c + i) / cpg;
    mean[i] = stats[(b * G + g) * 2] * inv_n;
    const float var = fmaxf(stats[(b * G + g) * 2 + 1] * inv_n - mean[i] * mean[i], 0.f);
    rstd[i] = rsqrtf(var + eps);
    gm[i] = gamma[c + i];
    bt[i] = beta[c + i];
    s1[i] = red[(b * G + g) * 2] * inv_n;
    s2[i] = red[(b * G + g) * 2 + 1] * inv_n;
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  for (int p = p0 + ty; p < p1; p += ny) {
    const long long pix = static_cast<long long>(b) * HW + p;
    const uint4 u = *reinterpret_cast<const uint4*>(gn_src(x1, x2, C1, C2, pix, c));
    const uint4 d = *reinterpret_cast<const uint4*>(dy + pix * C + c);
    uint4 ad = make_uint4(0, 0, 0, 0);
    if (add) ad = *reinterpret_cast<const uint4*>(add + pix * C + c);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
    const uint32_t aw[4] = {ad.x, ad.y, ad.z, ad.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 xf = unpack_bf16x2(w[i]);
      const float2 df = unpack_bf16x2(dw[i]);
      const float2 af = unpack_bf16x2(aw[i]);
      const float xs[2] = {xf.x, xf.y}, ds[2] = {df.x, df.y}, as[2] = {af.x, af.y};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = 2 * i + j;
        const float xh = (xs[j] - mean[k]) * rstd[k];
        float gg = ds[j];
        if (silu) gg *= dsilu_f(xh * gm[k] + bt[k]);
        gg *= gm[k];
        o[k] = rstd[k] * (gg - s1[k] - xh * s2[k]) + as[j];
      }
    }
    uint4 ov;
    ov.x = pack_bf16x2(o[0], o[1]);
    ov.y = pack_bf16x2(o[2], o[3]);
    ov.z = pack_bf16x2(o[4], o[5]);
    ov.w = pack_bf16x2(o[6], o[7]);
    if (c < C1)
      *reinterpret_cast<uint4*>(dx1 + pix * C1 + c) = ov;
    else
      *reinterpret_cast<uint4*>(dx2 + pix * C2 + (c - C1)) = ov;
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] += o[k];
  }
  if (colsum) {  // per-image column sums of dx (time-embedding gradient), fp32
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_cs[c + k], cs[k]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[b * C + i], s_cs[i]);
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the channel dimension.  LPR lanes cooperate on one token row (C/8 <= 5 * LPR
// 16-byte vectors, 5 independent loads per lane), 32 / LPR rows per warp; group reductions by
// xor-shuffles below LPR.  SD1.5: C = 320 / 640 / 1280 -> LPR = 8 / 16 / 32, perfectly balanced.
// ------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 5;

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR>
__global__ void ln_fwd_kernel(const bf16* __restrict__ x, int M, int C,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              float eps, bf16* __restrict__ out, float* __restrict__ stats) {
  griddep_sync();
  constexpr int RPW = 32 / LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int r = warp * RPW + lane / LPR;
  const bool rvalid = r < M;
  const int nvec = C >> 3;
  const bf16* row = x + static_cast<long long>(rvalid ? r : 0) * C;
  float f[kLnMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + v * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16x2(w[j]);
        f[i][2 * j] = t.x;
        f[i][2 * j + 1] = t.y;
        s += t.x + t.y;
      }
    }
  }
  const float mean = group_sum<LPR>(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(group_sum<LPR>(ss) / C + eps);
  if (!rvalid) return;
  if (sub == 0 && stats) {
    stats[r * 2] = mean;
    stats[r * 2 + 1] = rstd;
  }
  bf16* orow = out + static_cast<long long>(r) * C;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * gm[j] + bt[j];
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]);
      u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(orow + v * 8) = u;
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ add)
template <int LPR>
__global__ void ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, int M, int C,
                              const float* __restrict__ gamma, const float* __restrict__ stats,
                              const bf16* __restrict__ add, bf16* __restrict__ dx) {
  griddep_sync();
  constexpr int RPW = 32 / LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int r = warp * RPW + lane / LPR;
  const bool rvalid = r < M;
  const int nvec = C >> 3;
  const long long base = static_cast<long long>(rvalid ? r : 0) * C;
  const float mean = stats[(rvalid ? r : 0) * 2], rstd = stats[(rvalid ? r : 0) * 2 + 1];
  float g[kLnMaxVec][8], xh[kLnMaxVec][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + base + v * 8);
      const uint4 d = *reinterpret_cast<const uint4*>(dy + base + v * 8);
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = unpack_bf16x2(w[j]);
        const float2 df = unpack_bf16x2(dw[j]);
        xh[i][2 * j] = (xf.x - mean) * rstd;
        xh[i][2 * j + 1] = (xf.y - mean) * rstd;
        g[i][2 * j] = df.x * gm[2 * j];
        g[i][2 * j + 1] = df.y * gm[2 * j + 1];
        s1 += g[i][2 * j] + g[i][2 * j + 1];
        s2 += g[i][2 * j] * xh[i][2 * j] + g[i][2 * j + 1] * xh[i][2 * j + 1];
      }
    }
  }
  s1 = group_sum<LPR>(s1) / C;
  s2 = group_sum<LPR>(s2) / C;
  if (!rvalid) return;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
      float o[8];
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (add) {
        const uint4 u = *reinterpret_cast<const uint4*>(add + base + v * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_bf16x2(w[j]);
          a[2 * j] = t.x;
          a[2 * j + 1] = t.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - s1 - xh[i][j] * s2) + a[j];
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]);
      u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + base + v * 8) = u;
    }
  }
}

static inline int ln_lpr(int C) {
  const int nvec = C / 8;
  if (nvec <= 8 * kLnMaxVec) return 8;
  if (nvec <= 16 * kLnMaxVec) return 16;
  return 32;
}

static int gn_launch_cfg(int C, int HW, int B, int* threads, int* ppb, int* nblk) {
  const int nvec = C / 8;
  if (C % 8 != 0 || C > kGnMaxC || nvec > 1024) return set_error("groupnorm: unsupported C");
  int ny = 512 / nvec;      // ny * nvec <= 512 threads  =>  ny * C <= 4096 staged floats
  if (ny < 1) ny = 1;
  *threads = nvec * ny;
  // ~4 waves of blocks over the chip
  int target_blocks = (4 * num_sms() + B - 1) / B;
  int p = (HW + target_blocks - 1) / target_blocks;
  if (p < ny * 4) p = ny * 4;
  *ppb = p;
  *nblk = (HW + p - 1) / p;
  return 0;
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_groupnorm_fwd(const void* x1, const void* x2, int C1, int C2, int B, int HW,
                                 int G, const float* gamma, const float* beta, float eps, int silu,
                                 void* out, float* stats, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int C = C1 + C2;
  if (C % G != 0 || C1 % 8 != 0 || C2 % 8 != 0) return set_error("groupnorm: bad channel split");
  int threads, ppb, nblk;
  if (int rc = gn_launch_cfg(C, HW, B, &threads, &ppb, &nblk)) return rc;
  CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * B * G, stream));
  CUDA_TRY(launch_pdl(gn_stats_kernel, dim3(dim3(nblk, B)), dim3(threads), 0, stream, reinterpret_cast<const bf16*>(x1), reinterpret_cast<const bf16*>(x2), C1, C2, HW, G, ppb,
      stats));
  CUDA_TRY(launch_pdl(gn_apply_kernel, dim3(dim3(nblk, B)), dim3(threads), 0, stream, reinterpret_cast<const bf16*>(x1), reinterpret_cast<const bf16*>(x2), C1, C2, HW, G, ppb, stats,
      gamma, beta, eps, silu, reinterpret_cast<bf16*>(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int pcm_groupnorm_bwd(const void* dy, const void* x1, const void* x2, int C1, int C2,
                                 int B, int HW, int G, const float* gamma, const float* beta,
                                 float eps, int silu, const float* stats, float* red,
                                 const void* add, void* dx1, void* dx2, float* colsum,
                                 void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int C = C1 + C2;
  int threads, ppb, nblk;
  if (int rc = gn_launch_cfg(C, HW, B, &threads, &ppb, &nblk)) return rc;
  CUDA_TRY(cudaMemsetAsync(red, 0, sizeof(float) * 2 * B * G, stream));
  if (colsum) CUDA_TRY(cudaMemsetAsync(colsum, 0, sizeof(float) * B * C, stream));
  CUDA_TRY(launch_pdl(gn_bwd_stats_kernel, dim3(dim3(nblk, B)), dim3(threads), 0, stream, reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x1),
      reinterpret_cast<const bf16*>(x2), C1, C2, HW, G, ppb, stats, gamma, beta, eps, silu, red));
  CUDA_TRY(launch_pdl(gn_bwd_apply_kernel, dim3(dim3(nblk, B)), dim3(threads), 0, stream, reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x1),
      reinterpret_cast<const bf16*>(x2), C1, C2, HW, G, ppb, stats, red, gamma, beta, eps, silu,
      reinterpret_cast<const bf16*>(add), reinterpret_cast<bf16*>(dx1), reinterpret_cast<bf16*>(dx2),
      colsum));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int pcm_layernorm_fwd(const void* x, int M, int C, const float* gamma,
                                 const float* beta, float eps, void* out, float* stats,
                                 void* stream_) {
  if (C % 8 != 0 || C > kLnMaxVec * 256) return set_error("layernorm: unsupported C");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int wpb = 8, lpr = ln_lpr(C);
  const int rows_per_block = wpb * (32 / lpr);
  const int grid = (M + rows_per_block - 1) / rows_per_block;
  const bf16* xp = reinterpret_cast<const bf16*>(x);
  bf16* op = reinterpret_cast<bf16*>(out);
  if (lpr == 8) CUDA_TRY(launch_pdl(ln_fwd_kernel<8>, dim3(grid), dim3(wpb * 32), 0, stream, xp, M, C, gamma, beta, eps, op, stats));
  else if (lpr == 16) CUDA_TRY(launch_pdl(ln_fwd_kernel<16>, dim3(grid), dim3(wpb * 32), 0, stream, xp, M, C, gamma, beta, eps, op, stats));
  else CUDA_TRY(launch_pdl(ln_fwd_kernel<32>, dim3(grid), dim3(wpb * 32), 0, stream, xp, M, C, gamma, beta, eps, op, stats));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int pcm_layernorm_bwd(const void* dy, const void* x, int M, int C, const float* gamma,
                                 const float* stats, const void* add, void* dx, void* stream_) {
  if (C % 8 != 0 || C > kLnMaxVec * 256) return set_error("layernorm: unsupported C");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int wpb = 8, lpr = ln_lpr(C);
  const int rows_per_block = wpb * (32 / lpr);
  const int grid = (M + rows_per_block - 1) / rows_per_block;
  const bf16 *dyp = reinterpret_cast<const bf16*>(dy), *xp = reinterpret_cast<const bf16*>(x);
  const bf16* ap = reinterpret_cast<const bf16*>(add);
  bf16* dxp = reinterpret_cast<bf16*>(dx);
  if (lpr == 8) CUDA_TRY(launch_pdl(ln_bwd_kernel<8>, dim3(grid), dim3(wpb * 32), 0, stream, dyp, xp, M, C, gamma, stats, ap, dxp));
  else if (lpr == 16) CUDA_TRY(launch_pdl(ln_bwd_kernel<16>, dim3(grid), dim3(wpb * 32), 0, stream, dyp, xp, M, C, gamma, stats, ap, dxp));
  else CUDA_TRY(launch_pdl(ln_bwd_kernel<32>, dim3(grid), dim3(wpb * 32), 0, stream, dyp, xp, M, C, gamma, stats, ap, dxp));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
