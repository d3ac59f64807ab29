#!/bin/bash
# round 2, visit R: compute-sanitizer over the newer code paths (memcheck), shared-memory race check of the NTT and tail kernels,
# and the full suite twice more for flakiness
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_expr.py tests/test_gpu_quotient_pipeline.py tests/test_gpu_index_cache.py -m gpu -q -x --timeout 800 -p no:cacheprovider > gpurun_out/sanitize_mem1.log 2>&1; echo "memcheck (expr, quotient, index cache) exit $?"; tail -4 gpurun_out/sanitize_mem1.log
timeout 900 $CS --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_ipa.py "tests/test_gpu_srs.py" -m gpu -q -x --timeout 800 -p no:cacheprovider -k "not 2_16 and not 65536" > gpurun_out/sanitize_mem2.log 2>&1; echo "memcheck (ipa, srs/open) exit $?"; tail -4 gpurun_out/sanitize_mem2.log
timeout 900 $CS --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_msm.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "tma or edge or degenerate or kat or batch_shares or fold or synthetic" > gpurun_out/sanitize_mem3.log 2>&1; echo "memcheck (msm subset) exit $?"; tail -4 gpurun_out/sanitize_mem3.log
timeout 900 $CS --tool racecheck --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_ntt.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "not 2_20 and not beyond and not 17 and not 18 and not 19" > gpurun_out/sanitize_race.log 2>&1; echo "racecheck (ntt) exit $?"; tail -4 gpurun_out/sanitize_race.log
for i in 1 2; do timeout 1200 python -m pytest tests -m gpu -q --timeout 1000 -p no:cacheprovider > gpurun_out/pytest_soak_$i.log 2>&1; echo "soak $i exit $?"; tail -2 gpurun_out/pytest_soak_$i.log; done
