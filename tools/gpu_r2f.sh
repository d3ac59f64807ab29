#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_lanes.py 2>&1 | tail -30
echo "--- blocking launches"
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/diag_lanes.py 2>&1 | tail -16
