#!/bin/bash
# round 2, visit O: full parity with the evaluator in, evaluator timing at 8 CTAs/SM, bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -10 gpurun_out/pytest_gpu.log
timeout 600 python tools/expr_time.py 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_o.log 2>gpurun_out/bench_o.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_o.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"], d["checks"], d["roofline"]["traffic"])
PY
