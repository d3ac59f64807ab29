#!/bin/bash
# round 2, visit A: parity of the restructured pipeline, bench, task-count sweep, then the field-product variants (HEAD pipeline)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.5 > gpurun_out/bench_a.log 2>&1; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_a.log
timeout 400 python tools/msm_tune.py > gpurun_out/msm_tune.log 2>&1; tail -40 gpurun_out/msm_tune.log
MSM_LOGS=8,11,12,14 timeout 200 python tools/msm_sizes.py > gpurun_out/msm_sizes.log 2>&1; tail -8 gpurun_out/msm_sizes.log
for K in k0 k2 k4 k5 k6; do
  export ZKB200_LIB=$PWD/proof_systems_b200/libzkb200_$K.so
  echo "=== $K" | tee -a gpurun_out/mul_variants.log
  timeout 200 python -m pytest tests/test_gpu_field.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1 | tee -a gpurun_out/mul_variants.log
  timeout 200 python tools/microbench.py 4,100 2>&1 | grep -E "warps/SM +(8|16|64) " | tee -a gpurun_out/mul_variants.log
  timeout 200 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['stage_ms']['accumulate'], d['ntt']['value'])" | tee -a gpurun_out/mul_variants.log
done
