#!/bin/bash
# usual GPU visit: tests, smoke, microbench, bench; logs under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/microbench.py > gpurun_out/microbench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/microbench.log; tail -2 gpurun_out/bench.log
