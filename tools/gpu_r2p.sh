#!/bin/bash
# round 2, visit P: the profiles of the final kernels — launch list (window 16), full captures of the accumulation, the four tail
# kernels, the 2^20 / 2^16 NTT passes and the constraint evaluator; then the bench once more
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_p.csv python tools/prof_cmd.py 16 3 > /dev/null 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_accumulate -s 2 -c 1 -f -o gpurun_out/r02p_accumulate python tools/prof_cmd.py 16 3 > gpurun_out/ncu_p1.log 2>&1; echo "acc exit $?"
timeout 600 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --clock-control none -k regex:"k_run_sum|k_bucket_finish_serial|k_gridsum" -s 8 -c 4 -f -o gpurun_out/r02p_tails python tools/prof_cmd.py 16 3 > gpurun_out/ncu_p2.log 2>&1; echo "tails exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ntt_pass -s 8 -c 4 -f -o gpurun_out/r02p_ntt python tools/prof_cmd.py 16 3 > gpurun_out/ncu_p3.log 2>&1; echo "ntt exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_p.log 2>gpurun_out/bench_p.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_p.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"], d["checks"], d["clocks"])
PY
ls -la gpurun_out/*.ncu-rep | tail -8
