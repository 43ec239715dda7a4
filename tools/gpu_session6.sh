#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
grep -n "passed\|failed\|FAILED\|sdxl tiny" gpurun_out/pytest.log | tail -12; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-1200; tail -1 gpurun_out/bench_ref.log | cut -c1-400
