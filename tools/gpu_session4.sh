#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q > gpurun_out/pytest_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
PCM_EPI_V2=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q > gpurun_out/pytest_epi2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_epi2.log
timeout 120 python tools/gemm_bench.py 0 1 3 4 10 11 12 13 > gpurun_out/gemm_bench_v1.log 2>&1
PCM_EPI_V2=1 timeout 120 python tools/gemm_bench.py 0 1 3 4 10 11 12 13 > gpurun_out/gemm_bench_v2.log 2>&1
PCM_EPI_V2=1 timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_epi2.log 2>&1
echo "rc=$?" >> gpurun_out/bench_epi2.log
# ncu full captures (one kernel each) for profiles/
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:pcm_gemm_kernel -s 3 -c 1 -o gpurun_out/r2_gemm_conv320 -f python tools/gemm_bench.py 5 > gpurun_out/ncu1.log 2>&1
timeout 300 $NCU -k regex:pcm_gemm_kernel -s 3 -c 1 -o gpurun_out/r2_gemm_lin_k384_n320 -f python tools/gemm_bench.py 11 > gpurun_out/ncu2.log 2>&1
ITERS=2 timeout 300 $NCU -k regex:attn_fwd_tc2 -s 2 -c 1 -o gpurun_out/r2_attn_fwd -f python tools/attn_bench.py > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/pytest_gemm.log; tail -3 gpurun_out/pytest_epi2.log
cat gpurun_out/gemm_bench_v1.log gpurun_out/gemm_bench_v2.log
python - <<PY
import json
for l in open("gpurun_out/bench_epi2.log"):
    if l.startswith("{"):
        d=json.loads(l); print("epi2", d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["achieved"], d["loss"])
PY
ls -la gpurun_out/*.ncu-rep
