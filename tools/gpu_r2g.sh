#!/bin/bash
# round 2, visit G: full parity (lane pool, communicator, d8 pipeline, chunked Lagrange bases, 2^16 replay), bench
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -x --timeout 1200 -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/replay_kimchi.py > gpurun_out/replay.log 2>&1; echo "replay exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/replay_kimchi.json"))
    print("replay total", d["total_s"], {k: round(v * 1e3, 3) for k, v in d["stages_s"].items()})
    print("cpu", {k: round(v, 3) for k, v in d.get("cpu_oracle", {}).get("stages_s", {}).items()})
except Exception as e: print("replay parse failed", e)
PY
tail -3 gpurun_out/replay.log
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_g.log 2>gpurun_out/bench_g.err; echo "bench exit $?"; tail -c 400 gpurun_out/bench_g.err
