// Host-side helpers shared by the C-ABI translation units (error string, TMA descriptor encode).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace pcm {

int set_error(const char* msg);
int set_cuda_error(cudaError_t e, const char* what);
int num_sms();
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda
// dependency); bf16 elements, SWIZZLE_128B, zero fill out of bounds.
int encode_tmap(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box, const cuuint32_t* estr,
                int swizzle_bytes = 128);

bool pdl_enabled();

// Launch with programmatic stream serialization (disable with PCM_NO_PDL=1).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define CUDA_TRY(expr)                                         \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::pcm::set_cuda_error(_e, #expr); \
  } while (0)

}  // namespace pcm
