#!/bin/bash
# round 2, visit Z: size / window sweeps for BASELINE.md
mkdir -p gpurun_out
CHUNKS=0 timeout 600 python tools/sweep.py msm > gpurun_out/sweep_msm.log 2>&1; echo "sweep msm exit $?"; tail -9 gpurun_out/sweep_msm.log | cut -c1-220
timeout 600 python tools/sweep.py ntt > gpurun_out/sweep_ntt.log 2>&1; echo "sweep ntt exit $?"; tail -13 gpurun_out/sweep_ntt.log
