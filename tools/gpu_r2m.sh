#!/bin/bash
# round 2, visit M: full parity, bench (both arms), tail A/B numbers
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
WINDOWS=16,15 WAVES=512 CHUNKS=6,8 BATCHES=2,15 timeout 600 python tools/msm_tune.py 2>&1 | tail -12
timeout 300 python tools/open_time.py 2>&1 | tail -7
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_m.log 2>gpurun_out/bench_m.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_m.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"], d["checks"], d["clocks"])
print({k: (v.get("ms_per_step") or v.get("ms_per_round_trip")) for k, v in d["extra"].items() if isinstance(v, dict) and ("ms_per_step" in v or "ms_per_round_trip" in v)})
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_m_ref.log 2>&1; tail -c 600 gpurun_out/bench_m_ref.log
