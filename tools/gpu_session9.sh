#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default skipwgrad nostream; do
  case $v in
    default) envs="X=1" ;;
    skipwgrad) envs="PCM_DEBUG_SKIP_WGRAD=1" ;;
    nostream) envs="PCM_WGRAD_STREAM=0" ;;
  esac
  env $envs timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench9_$v.log 2>&1
  grep -h '^{' gpurun_out/bench9_$v.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('$v', d['ms_per_step'], d['roofline']['gemm_ms_per_step'])
"
done
