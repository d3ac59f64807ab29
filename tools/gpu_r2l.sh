#!/bin/bash
# round 2, visit L: tail kernels at 4 CTAs/SM + column trips, chunk sweep at windows 16/15, open with the phase trace, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_srs.py tests/test_gpu_ipa.py -m gpu -q -x --timeout 800 -p no:cacheprovider > gpurun_out/pytest_l.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_l.log
WINDOWS=16,15 WAVES=512 CHUNKS=3,4,5,6,8,12 BATCHES=2 timeout 600 python tools/msm_tune.py 2>&1 | tail -20
ZKB200_TRACE_OPEN=1 timeout 300 python tools/open_time.py 2>&1 | tail -14
timeout 400 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-extra > gpurun_out/bench_l.log 2>gpurun_out/bench_l.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_l.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"], d["checks"])
PY
