#!/bin/bash
# round 2, visit S: evaluator with the top of the stack in registers, NTT register stage with its warp fence (racecheck again, MSM tails too)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 python -m pytest tests/test_gpu_expr.py tests/test_gpu_quotient_pipeline.py tests/test_gpu_ntt.py tests/test_gpu_d8_pipeline.py -m gpu -q -x --timeout 1000 -p no:cacheprovider > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s.log
timeout 600 python tools/expr_time.py 2>&1 | tail -12
timeout 900 $CS --tool racecheck --error-exitcode 86 --print-limit 10 python -m pytest tests/test_gpu_ntt.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "not 2_20 and not beyond and not 17 and not 18 and not 19" > gpurun_out/sanitize_race_ntt.log 2>&1; echo "racecheck (ntt) exit $?"; tail -3 gpurun_out/sanitize_race_ntt.log
timeout 900 $CS --tool racecheck --error-exitcode 86 --print-limit 10 python -m pytest tests/test_gpu_msm.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "edge or degenerate or kat or tma" > gpurun_out/sanitize_race_msm.log 2>&1; echo "racecheck (msm) exit $?"; tail -3 gpurun_out/sanitize_race_msm.log; grep -c "Race reported" gpurun_out/sanitize_race_msm.log; grep "Race reported" -A1 gpurun_out/sanitize_race_msm.log | grep -o "in [a-z_.]*cuh*:[0-9]*" | sort | uniq -c | head -20
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/bench_s.log 2>gpurun_out/bench_s.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_s.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["ntt"]["ms_per_step"], d["checks"])
PY
timeout 900 python tools/replay_kimchi.py > gpurun_out/replay.log 2>&1; echo "replay exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/replay_kimchi.json"))
    print("replay total ms", d["total_s"] * 1e3, {k: round(v * 1e3, 3) for k, v in d["stages_s"].items()})
    print("resident", {k: round(v * 1e3, 3) for k, v in d["resident_stages_s"].items()})
except Exception as e: print("replay parse failed", e)
PY
