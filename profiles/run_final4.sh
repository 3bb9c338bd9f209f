# last check of the round: full GPU suite on the final build, then compute-sanitizer memcheck over the line extractor with
# the thread walk forced and the group's teacher-forced test (new kernels: k_fld_walk_thread, dense tile labelling, LK staging)
mkdir -p gpurun_out
timeout 170 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_final_c.txt 2>&1; tail -2 gpurun_out/pytest_gpu_final_c.txt
log=gpurun_out/sanitizer_r2z_memcheck.txt
echo "== compute-sanitizer --tool memcheck ($(date -u +%FT%TZ))" > $log
timeout 90 compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20 python -m pytest "tests/test_kernels_gpu.py::test_fld_vs_restatement" "tests/test_group_gpu.py::test_group_equals_single_handles" -m gpu -x -q -p no:cacheprovider >> $log 2>&1
echo "== exit code $?" >> $log
grep -h "ERROR SUMMARY\|== exit\|passed\|failed" $log | tail -4
