#!/bin/bash
# round 2, visit B: parity of the restructured pipeline + zk_srs_open, bench, task-count sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_a.log 2>&1; echo "bench exit $?"; tail -c 1800 gpurun_out/bench_a.log
timeout 400 python tools/msm_tune.py > gpurun_out/msm_tune.log 2>&1; tail -40 gpurun_out/msm_tune.log
MSM_LOGS=8,11,12,14 timeout 200 python tools/msm_sizes.py > gpurun_out/msm_sizes.log 2>&1; tail -8 gpurun_out/msm_sizes.log
