#!/bin/bash
# round 2, visit Y: the evaluator's gate programs (four gates) on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_expr.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_y.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_y.log
