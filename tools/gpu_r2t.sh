#!/bin/bash
# round 2, visit T: the round-end sequence as the driver runs it — full GPU suite, smoke(), bench (both arms) — plus the launch list of
# the bench command itself for profiles/
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -9 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_t_ref.log 2>gpurun_out/bench_t_ref.err; echo "ref exit $?"
timeout 900 python bench.py > gpurun_out/bench_t.log 2>gpurun_out/bench_t.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_t.log").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/bench_t_ref.log").read().strip().splitlines()[-1])
print("ours", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["checks"], d["clocks"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["traffic"])
print("ref ", r["value"], r["cpu_baseline"]["cores"], "same config:", d["config"] == r["config"], "e2e ratio", d["e2e"]["value"] / r["value"])
print({k: (v.get("ms_per_step") or v.get("ms_per_round_trip")) for k, v in d["extra"].items() if isinstance(v, dict) and ("ms_per_step" in v or "ms_per_round_trip" in v)})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --cpu-seconds 0.3 > gpurun_out/ncu_bench.log 2>&1; echo "bench launch list exit $?"
