mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_transforms_gpu.py tests/test_dataops_gpu.py -q -x -k "not expo_global300 and not spline_extra300" > gpurun_out/sanitizer_r02.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_r02.log
tail -15 gpurun_out/sanitizer_r02.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_transforms_gpu.py tests/test_dataops_gpu.py -q -x -k "op_matches or fps_subsample or co_unit" > gpurun_out/racecheck_r02.log 2>&1; echo "racecheck exit $?" >> gpurun_out/racecheck_r02.log
tail -8 gpurun_out/racecheck_r02.log
