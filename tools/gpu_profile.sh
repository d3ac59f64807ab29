#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench command, (2) full captures of the two dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --cpu-seconds 0.3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch_run.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 4 -c 1 -f -o gpurun_out/prof_accumulate $CMD > gpurun_out/ncu_acc_run.log 2>&1
echo "accumulate capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 8 -c 2 -f -o gpurun_out/prof_ntt $CMD > gpurun_out/ncu_ntt_run.log 2>&1
echo "ntt capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gridsum -s 8 -c 1 -f -o gpurun_out/prof_gridsum $CMD > gpurun_out/ncu_grid_run.log 2>&1
echo "gridsum capture exit $?"
ls -la gpurun_out | head -30
