/* pcm_b200 -- C ABI of the B200-native PCM-LoRA distillation hot path.
 *
 * The reference (G-U-N/Phased-Consistency-Model) has no FFI/plugin layer: its hot path is Python
 * calling diffusers/peft/torch modules.  This header is the boundary a maintainer binds instead
 * (ctypes stub in INTEGRATION.md).  Every entry point cites the reference call site it replaces
 * (T15 = code/text_to_image_sd15/train_pcm_lora_sd15.py, S15 = scheduling_ddpm_modified.py).
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller (torch tensors); the callee
 * borrows them for the duration of the call; work is enqueued on `stream` (a cudaStream_t passed
 * as void*); return value 0 = ok, negative = error (see pcm_last_error()).  No allocation crosses
 * the ABI.  Activations are NHWC ("channels last") bf16 unless stated otherwise.
 */
#ifndef PCM_B200_H
#define PCM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCM_MAX_ASRC 4
#define PCM_MAX_BSRC 2
#define PCM_MAX_PROG 24

/* A-operand source: a bf16 NHWC tensor viewed as (C, W, H, B) with element strides. */
typedef struct {
  const void* ptr;
  int32_t C, W, H, B;
  int64_t sW, sH, sB; /* element strides of the W, H, B dims (C is contiguous) */
} pcm_asrc;

/* B-operand source: weights [N, K] row-major bf16 (K contiguous), leading dimension ld. */
typedef struct {
  const void* ptr;
  int32_t K, N;
  int64_t ld;
} pcm_bsrc;

/* One K-program entry: nchunks consecutive 64-wide K blocks read from a[a_src] at channel a_c0..,
 * spatially shifted by (dw, dh) (zero filled outside the image), against b[b_src] columns b_k0... */
typedef struct {
  int32_t a_src, b_src, dw, dh, nchunks, a_c0, b_k0, pad_;
} pcm_kentry;

/* Implicit-GEMM descriptor: out[m, n] = alpha * sum_k A[m, k] * Bw[n, k] (+bias +rowvec +residual).
 * Replaces nn.Conv2d / nn.Linear (+ peft LoRA branch) calls inside diffusers' UNet2DConditionModel
 * forward/backward issued by T15:1192-1198, 1219-1223, 1238-1244, 1263-1268, 1296. */
typedef struct {
  pcm_asrc a[PCM_MAX_ASRC];
  pcm_bsrc b[PCM_MAX_BSRC];
  pcm_kentry prog[PCM_MAX_PROG];
  int32_t num_a, num_b, num_prog;
  int32_t lin;        /* 1: A sources are plain [M, C] matrices (W = rows), no spatial taps */
  int32_t M, N;       /* output rows (B*H*W tokens) and columns */
  int32_t geoW, geoH; /* conv mode: output image width / height (tile -> TMA coordinates) */
  int32_t block_n;    /* N tile: multiple of 32, <= 256 */
  /* epilogue */
  void* out;          /* bf16 (or fp32 when out_fp32) */
  const float* bias;  /* [N] fp32 or NULL */
  const void* rowvec; /* bf16 [B, rowvec_ld] added per image (time embedding) or NULL */
  const void* residual; /* bf16, same row mapping as out, or NULL */
  int64_t osW, osH, osB; /* element strides of out/residual rows: off = b*osB + h*osH + w*osW */
  int64_t rowvec_ld;
  int32_t epiW, epiHW;   /* row m -> (b = m / epiHW, h = (m % epiHW) / epiW, w = m % epiW) */
  int32_t out_fp32, round_bf16;
  float alpha;
  int32_t act;           /* 0 none, 1 SiLU applied to the result */
} pcm_gemm_desc;

/* LoRA weight-gradient descriptor: out[ch, r] += alpha * sum_m P[m(+tap), ch] * Q[m, r], r < 64.
 * Replaces autograd's wgrad of the peft lora_A / lora_B modules (T15:1296). */
typedef struct {
  pcm_asrc p;          /* [tokens, Cp] activation (or grad) */
  pcm_asrc q;          /* [tokens, >=64] rank-side operand */
  int32_t q_c0;        /* first column of the 64-wide slice of q */
  int32_t lin;
  int32_t M;           /* tokens */
  int32_t geoW, geoH;
  int32_t num_taps;
  int32_t dw[9], dh[9];
  int64_t tap_off[9];  /* element offset into out per tap */
  float* out;          /* fp32, accumulated with atomics */
  int64_t os_row, os_col; /* out[tap_off + ch*os_row + r*os_col] */
  int32_t ksplit;      /* token-dimension splits (0 = auto) */
  float alpha;
} pcm_wgrad_desc;

const char* pcm_last_error(void);
int pcm_version(void);
int pcm_num_sms(void);

/* tcgen05 implicit GEMM / conv and LoRA wgrad */
int pcm_gemm(const pcm_gemm_desc* d, void* stream);
int pcm_wgrad(const pcm_wgrad_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCM_B200_H */
