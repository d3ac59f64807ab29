#!/bin/bash
# round 2, visit D: parity (new NTT schedule, three-pass plan, out-of-place), bench, MSM sweep, ncu of the new NTT pass
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_d.log 2>gpurun_out/bench_d.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_d.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_d.log").read().strip().splitlines()[-1])
    print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"])
    print("ntt16", d["ntt"]["ms_per_step"], d["ntt"]["roofline"]["kernel_ms"])
    for k, v in d["extra"].items(): print(k, json.dumps(v)[:900])
except Exception as e: print("bench parse failed", e)
PY
WINDOWS=15,16 WAVES=512 timeout 300 python tools/msm_tune.py > gpurun_out/msm_tune.log 2>&1; tail -8 gpurun_out/msm_tune.log
PIPES=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --set full --metrics $PIPES --clock-control none --import-source on -k regex:k_ntt_pass -s 4 -c 4 -f -o gpurun_out/r02b_prof_ntt python tools/prof_cmd.py 16 2 > gpurun_out/ncu_ntt_run.log 2>&1; echo "ntt capture exit $?"
ls -la gpurun_out/*.ncu-rep
