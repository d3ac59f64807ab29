#!/bin/bash
# round 2, visit E: parity (lane pool, library communicator, d8 pipeline), bench, ncu of the tuned NTT pass
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_e.log 2>gpurun_out/bench_e.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_e.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_e.log").read().strip().splitlines()[-1])
    print("headline", d["value"], d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"])
    print("ntt16", d["ntt"]["ms_per_step"], d["ntt"]["roofline"]["kernel_ms"])
    print("cfg3", json.dumps(d["extra"]["cfg3_fp_ntt_2^20"])[:500])
except Exception as e: print("bench parse failed", e)
PY
PIPES=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --set full --metrics $PIPES --clock-control none --import-source on -k regex:k_ntt_pass -s 4 -c 4 -f -o gpurun_out/r02c_prof_ntt python tools/prof_cmd.py 16 2 > gpurun_out/ncu_ntt_run.log 2>&1; echo "ntt capture exit $?"
