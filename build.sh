#!/bin/bash
# Build libpcm_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
mkdir -p pcm_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -shared -Xcompiler -fPIC -o pcm_b200/lib/libpcm_b200.so pcm_b200/csrc/*.cu "$@"
