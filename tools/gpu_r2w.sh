#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/diag_scale.py 2> gpurun_out/diag_scale.err | tail -12; tail -c 600 gpurun_out/diag_scale.err
