#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/norm_bench.py > gpurun_out/norm_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
for k in gn_stats_kernel gn_apply_kernel gn_bwd_stats_kernel gn_bwd_apply_kernel; do
  ITERS=1 timeout 200 $NCU -k regex:$k -s 2 -c 1 -o gpurun_out/r2_$k -f python tools/norm_bench.py gn > gpurun_out/ncu_$k.log 2>&1
done
ITERS=1 timeout 200 $NCU -k regex:ln_fwd_kernel -s 2 -c 1 -o gpurun_out/r2_ln_fwd_kernel -f python tools/norm_bench.py ln > gpurun_out/ncu_ln.log 2>&1
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py tests/test_unet_gpu.py tests/test_pcm_kernels_gpu.py -m gpu -q > gpurun_out/pytest_small.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_small.log
cat gpurun_out/norm_bench.log; tail -3 gpurun_out/pytest_small.log; ls -la gpurun_out/*.ncu-rep
