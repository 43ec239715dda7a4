#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --model sdxl --batch 4 --latent 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sdxl.log 2>&1
echo "rc=$?" >> gpurun_out/bench_sdxl.log
tail -3 gpurun_out/bench_sdxl.log | cut -c1-1500
nvidia-smi --query-gpu=memory.used --format=csv
