#!/bin/bash
# round 2, visit V (8 GPUs): the bench at N = 8 under torchrun (weak scaling of config 2, strong scaling of config 4)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_v.log 2> gpurun_out/bench_v.err; echo "bench n8 exit $?"; tail -c 500 gpurun_out/bench_v.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_v.log").read().strip().splitlines() if l.startswith("{")][-1])
print("ours", d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["checks"])
c4 = d["extra"]["cfg4_vesta_2^20_strong"]; print("cfg4 strong", c4["ms_per_step"], c4["value"], c4["result_matches_cpu_oracle"])
PY
