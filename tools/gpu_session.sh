#!/bin/bash
# One GPU verification session (what the round-end driver does, plus the micro-benchmarks):
#   gpurun --timeout 3000 -- 'bash tools/gpu_session.sh'           (1 GPU)
#   gpurun --gpus 2 --timeout 2000 -- 'bash tools/gpu_session.sh dp' (2 GPUs: NCCL parity test + scaling)
# Everything is bounded by `timeout`; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" == "dp" ]; then
  timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -s > gpurun_out/pytest_dp.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_dp.log
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2.log 2>&1
  echo "rc=$?" >> gpurun_out/bench_dp2.log
  timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_dp1.log 2>&1
  tail -4 gpurun_out/pytest_dp.log; tail -2 gpurun_out/bench_dp2.log | cut -c1-400; tail -1 gpurun_out/bench_dp1.log | cut -c1-300
  exit 0
fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
PCM_CTX_GROUP=0 timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_noctx.log 2>&1
echo "ctx-group off: $(tail -1 gpurun_out/bench_noctx.log | cut -c1-160)"
timeout 120 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
timeout 120 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1
timeout 120 python tools/norm_bench.py > gpurun_out/norm_bench.log 2>&1
grep -n "passed\|failed\|FAILED\|parity config\|determinism" gpurun_out/pytest.log | tail -14
tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-1500
