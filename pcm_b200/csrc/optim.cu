// Optimiser side of the step on ONE flat fp32 LoRA buffer (67.25 M elements for SD1.5 r=64):
//   pcm_grad_sumsq    : global L2 norm of the (all-reduced) gradient
//   pcm_adamw_clip    : clip_grad_norm_(max_norm) folded into torch.optim.AdamW's update; the
//                       1/world average of the NCCL sum is folded in too; step counter and lr
//                       live in device memory so the whole step can sit in one CUDA graph
//   pcm_lora_refresh  : bf16 operand copies of the LoRA factors in the four layouts the tcgen05
//                       GEMMs consume (A, s*B, (s*B)^T, A^T), table driven, one launch
// Replaces accelerator.clip_grad_norm_ + optimizer.step + zero_grad
// (train_pcm_lora_sd15.py:1297-1301) and peft's per-op autocast casts of lora_A / lora_B.
#include "common.cuh"
#include "host_common.h"
#include "../../include/pcm_b200.h"

namespace pcm {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  __shared__ double s_part[32];
  double acc = 0.0;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += static_cast<double>(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) acc += static_cast<double>(g[i]) * g[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// state: [0] = lr, [1] = step (float, already incremented for this update)
__global__ void adamw_clip_kernel(float* __restrict__ p, float* __restrict__ g,
                                  float* __restrict__ m, float* __restrict__ v, long long n,
                                  const float* __restrict__ state, float beta1, float beta2,
                                  float eps, float wd, float max_norm, float inv_world,
                                  const double* __restrict__ sumsq, int zero_grad) {
  const float lr = state[0];
  const float step = state[1];
  const float norm = static_cast<float>(sqrt(*sumsq)) * inv_world;
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
  float coef = max_norm > 0.f ? fminf(max_norm / (norm + 1e-6f), 1.f) : 1.f;
  coef *= inv_world;
  const float bc1 = 1.f - powf(beta1, step);
  const float bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (zero_grad) g[i] = 0.f;
  }
}

__global__ void state_step_kernel(float* state) { state[1] += 1.f; }

// LoRA refresh table entry: masters A [r][taps][cin] and B [n][r] (fp32, offsets in elements)
struct RefreshEntry {
  long long a_off, b_off;          // into the fp32 master buffer
  long long a_fwd, sb_fwd, sb_t, a_t;  // into the bf16 operand buffer
  int cin, taps, n, r;
  long long work_begin;            // prefix sum of per-entry work items
};

// work item space per entry: [0, r*taps*cin) -> A element; then [.., + n*r) -> B element
__global__ void lora_refresh_kernel(const float* __restrict__ master,
                                    const RefreshEntry* __restrict__ tab, int num_entries,
                                    long long total_work, float scale, bf16* __restrict__ opnd) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total_work;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // binary search for the entry
    int lo = 0, hi = num_entries - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].work_begin <= i) lo = mid; else hi = mid - 1;
    }
    const RefreshEntry e = tab[lo];
    long long j = i - e.work_begin;
    const long long na = static_cast<long long>(e.r) * e.taps * e.cin;
    if (j < na) {
      // A[r][t][c]
      const float val = master[e.a_off + j];
      const bf16 h = __float2bfloat16_rn(val);
      opnd[e.a_fwd + j] = h;
      const int c = static_cast<int>(j % e.cin);
      const int t = static_cast<int>((j / e.cin) % e.taps);
      const int rr = static_cast<int>(j / (static_cast<long long>(e.cin) * e.taps));
      opnd[e.a_t + (static_cast<long long>(c) * e.taps + t) * e.r + rr] = h;  // A^T [c][t][r]
    } else {
      j -= na;
      const float val = master[e.b_off + j] * scale;
      const bf16 h = __float2bfloat16_rn(val);
      opnd[e.sb_fwd + j] = h;  // [n][r]
      const int rr = static_cast<int>(j % e.r);
      const int nn = static_cast<int>(j / e.r);
      opnd[e.sb_t + static_cast<long long>(rr) * e.n + nn] = h;  // [r][n]
    }
  }
}

}  // namespace pcm

using namespace pcm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int pcm_grad_sumsq(const float* g, int64_t n, double* out, void* stream) {
  CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double), ST(stream)));
  int grid = static_cast<int>((n / 4 + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  if (grid < 1) grid = 1;
  sumsq_kernel<<<grid, 256, 0, ST(stream)>>>(g, n, out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int pcm_adamw_clip(float* p, float* g, float* m, float* v, int64_t n, float* state,
                              float beta1, float beta2, float eps, float weight_decay,
                              float max_norm, float inv_world, const double* sumsq, int zero_grad,
                              void* stream) {
  state_step_kernel<<<1, 1, 0, ST(stream)>>>(state);
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  adamw_clip_kernel<<<grid, 256, 0, ST(stream)>>>(p, g, m, v, n, state, beta1, beta2, eps,
                                                  weight_decay, max_norm, inv_world, sumsq,
                                                  zero_grad);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int pcm_lora_refresh(const float* master, const void* table, int num_entries,
                                int64_t total_work, float scale, void* opnd, void* stream) {
  int grid = static_cast<int>((total_work + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  lora_refresh_kernel<<<grid, 256, 0, ST(stream)>>>(
      master, reinterpret_cast<const RefreshEntry*>(table), num_entries, total_work, scale,
      reinterpret_cast<bf16*>(opnd));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
