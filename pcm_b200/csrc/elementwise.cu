// HBM-bound glue kernels of the UNet: GEGLU, nearest up-sampling, the 4-channel edge
// convolutions (conv_in forward / conv_out input-gradient), timestep sinusoid, per-image column
// sums (time-embedding gradient), bf16 adds.  All use 16-byte accesses on the contiguous channel
// dimension and grid-stride loops sized to the SM count.
//
// Replaces ATen elementwise launches inside diffusers' UNet2DConditionModel (GEGLU.forward,
// Upsample2D F.interpolate, get_timestep_embedding, conv_in) reached from
// train_pcm_lora_sd15.py:1192-1198, 1219-1244, 1263-1268 and their autograd twins (:1296).
#include "common.cuh"
#include "host_common.h"
#include "../../include/pcm_b200.h"

namespace pcm {

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16x2(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) +
         x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// u [M, 2F] -> out [M, F] = u[:, :F] * gelu(u[:, F:])
__global__ void geglu_fwd_kernel(const bf16* __restrict__ u, long long M, int F,
                                 bf16* __restrict__ out) {
  griddep_sync();
  const int nvec = F >> 3;
  const long long total = M * nvec;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = v / nvec;
    const int c = static_cast<int>(v - m * nvec) * 8;
    float a[8], g[8];
    load8(u + m * 2 * F + c, a);
    load8(u + m * 2 * F + F + c, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] *= gelu_erf(g[i]);
    store8(out + m * F + c, a);
  }
}

// du[:, :F] = dgg * gelu(g) ; du[:, F:] = dgg * a * gelu'(g)
__global__ void geglu_bwd_kernel(const bf16* __restrict__ dgg, const bf16* __restrict__ u,
                                 long long M, int F, bf16* __restrict__ du) {
  griddep_sync();
  const int nvec = F >> 3;
  const long long total = M * nvec;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = v / nvec;
    const int c = static_cast<int>(v - m * nvec) * 8;
    float a[8], g[8], d[8], da[8], dg[8];
    load8(u + m * 2 * F + c, a);
    load8(u + m * 2 * F + F + c, g);
    load8(dgg + m * F + c, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      da[i] = d[i] * gelu_erf(g[i]);
      dg[i] = d[i] * a[i] * dgelu_erf(g[i]);
    }
    store8(du + m * 2 * F + c, da);
    store8(du + m * 2 * F + F + c, dg);
  }
}

// nearest 2x: in [B, H, W, C] -> out [B, 2H, 2W, C]
__global__ void upsample2x_kernel(const bf16* __restrict__ in, int B, int H, int W, int C,
                                  bf16* __restrict__ out) {
  griddep_sync();
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(B) * 4 * H * W * nvec;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(v % nvec);
    long long p = v / nvec;
    const int x = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int y = static_cast<int>(p % (2 * H));
    const int b = static_cast<int>(p / (2 * H));
    const uint4 u = *reinterpret_cast<const uint4*>(
        in + ((static_cast<long long>(b) * H + (y >> 1)) * W + (x >> 1)) * C + cv * 8);
    *reinterpret_cast<uint4*>(out + v * 8) = u;
  }
}

// gradient of nearest 2x: dout [B, 2H, 2W, C] -> din [B, H, W, C] (sum of the 2x2 block)
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dout, int B, int H, int W, int C,
                                      bf16* __restrict__ din) {
  griddep_sync();
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(B) * H * W * nvec;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(v % nvec);
    long long p = v / nvec;
    const int x = static_cast<int>(p % W);
    p /= W;
    const int y = static_cast<int>(p % H);
    const int b = static_cast<int>(p / H);
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        load8(dout + ((static_cast<long long>(b) * 2 * H + 2 * y + dy) * 2 * W + 2 * x + dx) * C +
                  cv * 8,
              f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += f[i];
      }
    store8(din + v * 8, s);
  }
}

// 3x3 pad-1 convolution from a 4-channel fp32 NHWC image to C bf16 channels:
//   out[b,y,x,c] = bias[c] + sum_{kh,kw,i} in[b, y+sgn*(kh-1), x+sgn*(kw-1), i] * w[c][kh][kw][i]
// sgn = +1: conv_in forward (w = conv_in.weight as [C][3][3][4]);
// sgn = -1: conv_out input gradient (w[c][kh][kw][o] = conv_out.weight[o][c][kh][kw]).
__global__ void conv3x3_c4_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                  const bf16* __restrict__ w, const float* __restrict__ bias,
                                  int sgn, int round_in, bf16* __restrict__ out) {
  griddep_sync();
  extern __shared__ float s_w[];  // [36][C]: consecutive threads read consecutive channels
  for (int i = threadIdx.x; i < C * 36; i += blockDim.x) {
    const int ch = i / 36, k = i - ch * 36;
    s_w[k * C + ch] = __bfloat162float(w[i]);
  }
  __syncthreads();
  const int nvec = C >> 3;
  const long long total = static_cast<long long>(B) * H * W * nvec;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(v % nvec);
    long long p = v / nvec;
    const int x = static_cast<int>(p % W);
    const int y = static_cast<int>((p / W) % H);
    const int b = static_cast<int>(p / (static_cast<long long>(W) * H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias ? bias[cv * 8 + i] : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + sgn * (kh - 1);
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int xx = x + sgn * (kw - 1);
        if (xx < 0 || xx >= W) continue;
        float4 t = *reinterpret_cast<const float4*>(
            in + ((static_cast<long long>(b) * H + yy) * W + xx) * 4);
        if (round_in) {
          t.x = __bfloat162float(__float2bfloat16_rn(t.x));
          t.y = __bfloat162float(__float2bfloat16_rn(t.y));
          t.z = __bfloat162float(__float2bfloat16_rn(t.z));
          t.w = __bfloat162float(__float2bfloat16_rn(t.w));
        }
        const float tin[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float* ww = s_w + ((kh * 3 + kw) * 4 + k) * C + cv * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(ww);
          const float4 w1 = *reinterpret_cast<const float4*>(ww + 4);
          acc[0] += tin[k] * w0.x; acc[1] += tin[k] * w0.y; acc[2] += tin[k] * w0.z; acc[3] += tin[k] * w0.w;
          acc[4] += tin[k] * w1.x; acc[5] += tin[k] * w1.y; acc[6] += tin[k] * w1.z; acc[7] += tin[k] * w1.w;
        }
      }
    }
    store8(out + v * 8, acc);
  }
}

// get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]
__global__ void timestep_embed_kernel(const long long* __restrict__ t, int B, int C,
                                      bf16* __restrict__ out) {
  griddep_sync();
  const int half = C >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float f = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
  const float e = static_cast<float>(t[b]) * f;
  float s, c;
  sincosf(e, &s, &c);
  out[b * C + k] = __float2bfloat16_rn(c);
  out[b * C + half + k] = __float2bfloat16_rn(s);
}

// out[b, c] = sum over HW rows of x[b, :, c]   (bf16 in, bf16 out, fp32 accumulate)
__global__ void colsum_kernel(const bf16* __restrict__ x, int HW, int C, bf16* __restrict__ out) {
  griddep_sync();
  // grid (C/8 vectors / blockDim.x chunks, B); blockDim (32 vectors, 8 row-lanes)
  __shared__ float s[8][32][9];
  const int b = blockIdx.y;
  const int cv = blockIdx.x * 32 + threadIdx.x;
  const int nvec = C >> 3;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cv < nvec) {
    for (int r = threadIdx.y; r < HW; r += 8) {
      float f[8];
      load8(x + (static_cast<long long>(b) * HW + r) * C + cv * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && cv < nvec) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
      for (int j = 0; j < 8; ++j) t += s[j][threadIdx.x][i];
      acc[i] = t;
    }
    store8(out + static_cast<long long>(b) * C + cv * 8, acc);
  }
}

__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                long long nvec, bf16* __restrict__ out) {
  griddep_sync();
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    load8(a + v * 8, x);
    load8(b + v * 8, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += y[i];
    store8(out + v * 8, x);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, long long n, bf16* __restrict__ out) {
  griddep_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

static inline int ew_grid(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace pcm

using namespace pcm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<bf16*>(p)
#define CBF(p) reinterpret_cast<const bf16*>(p)

extern "C" int pcm_geglu_fwd(const void* u, int64_t M, int F, void* out, void* stream) {
  if (F % 8) return set_error("geglu: F % 8 != 0");
  CUDA_TRY(launch_pdl(geglu_fwd_kernel, dim3(ew_grid(M * (F / 8), 256)), dim3(256), 0, ST(stream), CBF(u), M, F, BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_geglu_bwd(const void* dgg, const void* u, int64_t M, int F, void* du,
                             void* stream) {
  if (F % 8) return set_error("geglu: F % 8 != 0");
  CUDA_TRY(launch_pdl(geglu_bwd_kernel, dim3(ew_grid(M * (F / 8), 256)), dim3(256), 0, ST(stream), CBF(dgg), CBF(u), M, F,
                                                                       BF(du)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_upsample2x_fwd(const void* in, int B, int H, int W, int C, void* out,
                                  void* stream) {
  if (C % 8) return set_error("upsample: C % 8 != 0");
  const long long total = static_cast<long long>(B) * 4 * H * W * (C / 8);
  CUDA_TRY(launch_pdl(upsample2x_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), CBF(in), B, H, W, C, BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_upsample2x_bwd(const void* dout, int B, int H, int W, int C, void* din,
                                  void* stream) {
  if (C % 8) return set_error("upsample: C % 8 != 0");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  CUDA_TRY(launch_pdl(upsample2x_bwd_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), CBF(dout), B, H, W, C,
                                                                      BF(din)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_conv3x3_c4(const float* in, int B, int H, int W, int C, const void* w,
                              const float* bias, int sgn, int round_in, void* out, void* stream) {
  if (C % 8 || C > 320) return set_error("conv3x3_c4: C must be a multiple of 8, <= 320");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  const size_t smem = static_cast<size_t>(C) * 36 * sizeof(float);
  int grid = ew_grid(total, 256);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  CUDA_TRY(launch_pdl(conv3x3_c4_kernel, dim3(grid), dim3(256), smem, ST(stream), in, B, H, W, C, CBF(w), bias, sgn, round_in,
                                                     BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_timestep_embed(const int64_t* t, int B, int C, void* out, void* stream) {
  const int n = B * (C / 2);
  CUDA_TRY(launch_pdl(timestep_embed_kernel, dim3((n + 127) / 128), dim3(128), 0, ST(stream), reinterpret_cast<const long long*>(t), B, C, BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_colsum(const void* x, int B, int HW, int C, void* out, void* stream) {
  if (C % 8) return set_error("colsum: C % 8 != 0");
  const int nvec = C / 8;
  CUDA_TRY(launch_pdl(colsum_kernel, dim3(dim3((nvec + 31) / 32, B)), dim3(dim3(32, 8)), 0, ST(stream), CBF(x), HW, C, BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_add_bf16(const void* a, const void* b, int64_t n, void* out, void* stream) {
  if (n % 8) return set_error("add: n % 8 != 0");
  CUDA_TRY(launch_pdl(add_bf16_kernel, dim3(ew_grid(n / 8, 256)), dim3(256), 0, ST(stream), CBF(a), CBF(b), n / 8, BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int pcm_cast_f32_bf16(const float* in, int64_t n, void* out, void* stream) {
  CUDA_TRY(launch_pdl(cast_f32_bf16_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, ST(stream), in, n, BF(out)));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
