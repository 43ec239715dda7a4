#!/bin/bash
# One GPU session: parity tests, smoke, bench.  Everything bounded by `timeout`; logs -> gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log
