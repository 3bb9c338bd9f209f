#!/bin/bash
# compute-sanitizer over the kernel known-answer tests and one teacher-forced front-end test (run on the GPU box):
#   gpurun --timeout 1500 -- 'bash profiles/run_sanitizer.sh r2'
# Writes gpurun_out/sanitizer_<tag>_{memcheck,racecheck,initcheck}.txt; profiles/sanitizer_<tag>.txt is the committed summary.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=32
KTESTS="tests/test_kernels_gpu.py"
FTEST="tests/test_frontend_gpu.py::test_teacher_forced_1280x560"
GTEST="tests/test_group_gpu.py::test_group_equals_single_handles tests/test_group_gpu.py::test_group_masks_clahe_partial_ticks_and_reset"
for tool in memcheck racecheck; do
  lim=900
  [ $tool = racecheck ] && lim=1500
  log=$out/sanitizer_${tag}_${tool}.txt
  echo "== compute-sanitizer --tool $tool ($(date -u +%FT%TZ))" > $log
  timeout $lim compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
      python -m pytest $KTESTS "$FTEST" $GTEST -m gpu -x -q -p no:cacheprovider >> $log 2>&1
  echo "== exit code $?" >> $log
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|== exit\|passed\|failed\|== compute" $out/sanitizer_${tag}_*.txt > $out/sanitizer_${tag}_summary.txt
cat $out/sanitizer_${tag}_summary.txt
