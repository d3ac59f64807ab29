#!/bin/bash
# round 2, visit H: full parity, replay, open timing, bench
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -x --timeout 1200 -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
timeout 600 python tools/replay_kimchi.py > gpurun_out/replay.log 2>&1; echo "replay exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/replay_kimchi.json"))
    print("replay total ms", d["total_s"] * 1e3, {k: round(v * 1e3, 3) for k, v in d["stages_s"].items()})
except Exception as e: print("replay parse failed", e)
PY
timeout 300 python tools/open_time.py 2>&1 | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_h.log 2>gpurun_out/bench_h.err; echo "bench exit $?"; tail -c 400 gpurun_out/bench_h.err
