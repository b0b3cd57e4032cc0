#!/bin/bash
# compute-sanitizer over the CQT tests and one model test; logs -> gpurun_out/sanitizer/
# usage (GPU box): bash scripts/sanitize.sh [memcheck racecheck initcheck]
set -u
out=gpurun_out/sanitizer
mkdir -p $out
tools=${@:-memcheck racecheck initcheck}
CQT_TESTS="tests/test_cqt_gpu.py::test_magnitude_and_decibels tests/test_cqt_gpu.py::test_encode_complex_and_layout_helpers tests/test_cqt_gpu.py::test_forward_matches_oracle[cfg2-3-2] tests/test_cqt_gpu.py::test_inverse_matches_oracle[cfg1-3-2]"
MODEL_TESTS="tests/test_model_gpu.py::test_chunk_batching_is_invisible"
TMO=${SANITIZE_TIMEOUT:-240}
for tool in $tools; do
  for name in cqt model; do
    if [ $name = cqt ]; then sel="$CQT_TESTS"; else sel="$MODEL_TESTS"; fi
    if [ $name = model ] && [ $tool != memcheck ]; then continue; fi   # racecheck / initcheck of the tcgen05 kernels: opt-in (slow)
    log=$out/${tool}_${name}.txt
    echo "== compute-sanitizer --tool $tool : $sel" > $log
    timeout $TMO compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
        python -m pytest $sel -x -q -m gpu -p no:cacheprovider >> $log 2>&1
    echo "exit $?" >> $log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $log | tail -5
  done
done
