#!/bin/bash
# round 2, visit K: fold building block parity, fold vs never-fold timing, ncu launch lists at table windows 16 and 15
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ipa.py tests/test_gpu_msm.py -m gpu -q --timeout 800 -p no:cacheprovider > gpurun_out/pytest_k.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_k.log
timeout 900 python tools/fold_vs_never_fold.py > gpurun_out/fold_vs_never_fold.log 2>&1; echo "fold exit $?"; tail -26 gpurun_out/fold_vs_never_fold.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_w16.csv python tools/prof_cmd.py 16 3 > /dev/null 2>&1; echo "ncu16 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_w15.csv python tools/prof_cmd.py 15 3 > /dev/null 2>&1; echo "ncu15 exit $?"
