#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/norm_bench.py > gpurun_out/norm_bench2.log 2>&1
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_unet_gpu.py tests/test_parity_gpu.py -m gpu -q -s > gpurun_out/pytest_norm.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_norm.log
timeout 400 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench8.log 2>&1
cat gpurun_out/norm_bench2.log; grep -n "passed\|failed\|parity config" gpurun_out/pytest_norm.log | tail -12
grep -h '^{' gpurun_out/bench8.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('bench', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['loss'])
"
