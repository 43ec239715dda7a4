#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
EXP=$PWD/pcm_b200/lib/libpcm_b200_attnexp.so
timeout 120 python tools/debug_splitk.py > gpurun_out/splitk.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
PCM_EPI_V2=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q > gpurun_out/pytest_epi2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_epi2.log
PCM_B200_LIB=$EXP timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention or attn" > gpurun_out/pytest_attnexp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_attnexp.log
timeout 120 python tools/attn_bench.py > gpurun_out/attn_bench_default.log 2>&1
PCM_B200_LIB=$EXP timeout 120 python tools/attn_bench.py > gpurun_out/attn_bench_exp.log 2>&1
timeout 120 python tools/gemm_bench.py 0 1 3 4 11 12 13 > gpurun_out/gemm_bench_v1.log 2>&1
PCM_EPI_V2=1 timeout 120 python tools/gemm_bench.py 0 1 3 4 11 12 13 > gpurun_out/gemm_bench_v2.log 2>&1
for v in default epi2 attnexp gn24; do
  case $v in
    default) envs="X=1" ;;
    epi2) envs="PCM_EPI_V2=1" ;;
    attnexp) envs="PCM_B200_LIB=$EXP" ;;
    gn24) envs="PCM_GN_CHUNK_MB=24" ;;
  esac
  env $envs timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_$v.log 2>&1
  echo "rc=$?" >> gpurun_out/bench_$v.log
done
tail -4 gpurun_out/pytest.log; tail -3 gpurun_out/pytest_epi2.log; tail -3 gpurun_out/pytest_attnexp.log; tail -12 gpurun_out/splitk.log
cat gpurun_out/attn_bench_default.log gpurun_out/attn_bench_exp.log gpurun_out/gemm_bench_v1.log gpurun_out/gemm_bench_v2.log
for v in default epi2 attnexp gn24; do python - <<PY
import json
for l in open("gpurun_out/bench_$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$v", d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["achieved"], d["loss"])
PY
done
