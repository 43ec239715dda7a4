#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/dp_smi.txt 2>&1
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -s > gpurun_out/pytest_dp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_dp.log
PORT=29511
for v in ingraph eager_ar nooverlap; do
  case $v in
    ingraph) envs="X=1" ;;
    eager_ar) envs="PCM_NCCL_IN_GRAPH=0" ;;
    nooverlap) envs="PCM_DP_OVERLAP=0 PCM_NCCL_IN_GRAPH=0" ;;
  esac
  PORT=$((PORT+1))
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2_$v.log 2>&1
  echo "rc=$?" >> gpurun_out/bench_dp2_$v.log
done
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_dp1.log 2>&1
tail -5 gpurun_out/pytest_dp.log
for v in ingraph eager_ar nooverlap; do grep -h '^{' gpurun_out/bench_dp2_$v.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('$v', d['n_gpus'], d['ms_per_step'], d['value'], d['loss'])
"; tail -2 gpurun_out/bench_dp2_$v.log | cut -c1-300; done
grep -h '^{' gpurun_out/bench_dp1.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('dp1', d['ms_per_step'], d['value'])
"
