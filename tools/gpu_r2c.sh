#!/bin/bash
# round 2, visit C: parity, the new bench line, K variants of the field product, ncu baselines with pipe metrics
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_c.log 2>gpurun_out/bench_c.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_c.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_c.log").read().strip().splitlines()[-1])
    print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"], d["clocks"])
    for k, v in d["extra"].items(): print(k, json.dumps(v)[:700])
except Exception as e: print("bench parse failed", e)
PY
WINDOWS=15,16 WAVES=512,1024 timeout 300 python tools/msm_tune.py > gpurun_out/msm_tune.log 2>&1; tail -12 gpurun_out/msm_tune.log
MSM_LOGS=8,11,12,14 timeout 200 python tools/msm_sizes.py > gpurun_out/msm_sizes.log 2>&1; tail -8 gpurun_out/msm_sizes.log
rm -f gpurun_out/mul_variants.log
for K in k2 k4 k5 k6; do
  export ZKB200_LIB=$PWD/proof_systems_b200/libzkb200_$K.so
  echo "=== $K" | tee -a gpurun_out/mul_variants.log
  timeout 200 python -m pytest tests/test_gpu_field.py tests/test_gpu_ntt.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1 | tee -a gpurun_out/mul_variants.log
  timeout 200 python tools/microbench.py 4,100 2>&1 | grep -E "warps/SM +(8|16|64) " | tee -a gpurun_out/mul_variants.log
  timeout 200 python bench.py --no-extra --steps 10 --warmup 3 --cpu-seconds 0.2 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['stage_ms']['accumulate'], d['ntt']['value'])" | tee -a gpurun_out/mul_variants.log
done
unset ZKB200_LIB
echo "=== k0 (shipped)" | tee -a gpurun_out/mul_variants.log
timeout 200 python tools/microbench.py 4,100 2>&1 | grep -E "warps/SM +(8|16|64) " | tee -a gpurun_out/mul_variants.log
PIPES=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --set full --metrics $PIPES --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o gpurun_out/r02a_prof_accumulate python tools/prof_cmd.py 16 2 > gpurun_out/ncu_acc_run.log 2>&1; echo "acc capture exit $?"
timeout 600 ncu --set full --metrics $PIPES --clock-control none --import-source on -k regex:k_ntt_pass -s 4 -c 4 -f -o gpurun_out/r02a_prof_ntt python tools/prof_cmd.py 16 2 > gpurun_out/ncu_ntt_run.log 2>&1; echo "ntt capture exit $?"
ls -la gpurun_out/*.ncu-rep
