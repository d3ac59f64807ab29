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
# (gpurun copies back at most 64 MiB: the third capture keeps the summary sections only)
timeout 900 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --clock-control none -k regex:"k_gridsum|k_run_sum|k_bucket_finish" -s 16 -c 4 -f -o gpurun_out/prof_tails $CMD > gpurun_out/ncu_grid_run.log 2>&1
echo "gridsum capture exit $?"
ls -la gpurun_out | head -30
