#!/bin/bash
# round 2, visit J (2 GPUs): the library-owned exchange at world size 2, bench at N=2 (both arms)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -15 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -c 600 gpurun_out/bench_n2.err; head -c 3000 gpurun_out/bench_n2.log
