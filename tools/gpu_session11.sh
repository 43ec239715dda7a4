#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NOP=$PWD/pcm_b200/lib/libpcm_b200_nopoly.so
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q > gpurun_out/pytest_attn.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_attn.log
timeout 120 python tools/attn_bench.py > gpurun_out/attn_bench_poly.log 2>&1
PCM_B200_LIB=$NOP timeout 120 python tools/attn_bench.py > gpurun_out/attn_bench_nopoly.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_poly.log 2>&1
PCM_B200_LIB=$NOP timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_nopoly.log 2>&1
tail -3 gpurun_out/pytest_attn.log; cat gpurun_out/attn_bench_poly.log gpurun_out/attn_bench_nopoly.log
for v in poly nopoly; do grep -h '^{' gpurun_out/bench_$v.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('$v', d['ms_per_step'], d['loss'])
"; done
