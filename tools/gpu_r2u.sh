#!/bin/bash
# round 2, visit U (2 GPUs): the multi-rank test and both bench arms under torchrun, as the driver launches them
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_u_ref.log 2> gpurun_out/bench_u_ref.err; echo "ref n2 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_u.log 2> gpurun_out/bench_u.err; echo "bench n2 exit $?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_u.log").read().strip().splitlines() if l.startswith("{")][-1])
r = json.loads([l for l in open("gpurun_out/bench_u_ref.log").read().strip().splitlines() if l.startswith("{")][-1])
print("ours", d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["checks"], d["clocks"])
print("ref ", r["n_gpus"], r["value"], "same config:", d["config"] == r["config"])
c4 = d["extra"]["cfg4_vesta_2^20_strong"]; print("cfg4 strong", c4["ms_per_step"], c4["value"], c4["result_matches_cpu_oracle"])
PY
