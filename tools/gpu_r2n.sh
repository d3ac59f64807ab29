#!/bin/bash
# round 2, visit N: the constraint evaluator — parity, timing, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_expr.py tests/test_gpu_d8_pipeline.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_n.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_n.log
timeout 600 python tools/expr_time.py 2>&1 | tail -12
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_expr_eval -c 2 -o gpurun_out/r02n_expr_eval -f python tools/expr_time.py > gpurun_out/ncu_n.log 2>&1; echo "ncu exit $?"
