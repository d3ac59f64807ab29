#!/bin/bash
# round 2, visit I: full parity (no -x), TMA A/B, ncu of both accumulate kernels, open timing
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/msm_tma_ab.py 2>&1 | tail -12
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accumulate -c 4 -o gpurun_out/r02i_accumulate_ab -f python tools/prof_cmd.py msm_tma > gpurun_out/ncu_i.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_i.log
