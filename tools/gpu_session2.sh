#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/debug_splitk.py > gpurun_out/splitk.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
for v in default gn0 late0; do
  case $v in
    default) envs="" ;;
    gn0) envs="PCM_GN_CHUNK_MB=0" ;;
    late0) envs="PCM_LATE_WAIT=0" ;;
  esac
  env $envs timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_$v.log 2>&1
  echo "rc=$?" >> gpurun_out/bench_$v.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --profile-only > gpurun_out/ncu_run.log 2>&1
tail -4 gpurun_out/pytest.log; tail -12 gpurun_out/splitk.log
for v in default gn0 late0; do python - <<PY
import json
for l in open("gpurun_out/bench_$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("$v", d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["achieved"], d["loss"])
PY
done
