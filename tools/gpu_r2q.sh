#!/bin/bash
# round 2, visit Q: NTT parity with the persistent grid, A/B timing, the C++ layer's open
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_cpp_layer.py tests/test_gpu_d8_pipeline.py tests/test_gpu_quotient_pipeline.py tests/test_gpu_lagrange.py -m gpu -q -x --timeout 1000 -p no:cacheprovider > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_q.log
# (the A/B of the persistent NTT grid ran here — commit 9c1289a; the experiment was reverted and its tool removed, result in profiles/r02_ntt_persistent.md)
