#!/bin/bash
# Next-round experiment (run under gpurun): the split-product variants of field.cuh (ZK_MUL_PLAIN_PER_ROW = K products per row as
# plain wide multiplies + carry adds on the ALU pipe, DESIGN.md "next levers").  For every K: rebuild the library, check parity on
# the GPU (field + MSM + NTT tests), then report the multiplication / mixed-addition rates and the headline bench.
# The tree is left on the shipped build (K = 0) at the end.
set -u
mkdir -p gpurun_out
for K in 0 2 4 5 6; do
  make -C proof_systems_b200/csrc clean > /dev/null
  if [ "$K" = 0 ]; then make -C proof_systems_b200/csrc -j8 > /dev/null; else make -C proof_systems_b200/csrc -j8 EXTRA=-DZK_MUL_PLAIN_PER_ROW=$K > /dev/null; fi
  echo "=== K=$K" | tee -a gpurun_out/mul_variants.log
  python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -1 | tee -a gpurun_out/mul_variants.log
  python tools/microbench.py 4,100 2>&1 | grep -E "ilp4|madd" | sort -t: -k3 | tail -4 | tee -a gpurun_out/mul_variants.log
  python bench.py --steps 20 --warmup 5 --cpu-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['stage_ms']['accumulate'], d['ntt']['value'])" | tee -a gpurun_out/mul_variants.log
done
make -C proof_systems_b200/csrc clean > /dev/null; make -C proof_systems_b200/csrc -j8 > /dev/null
