mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 9"
(timeout 900 $S --tool memcheck python -m pytest tests/test_iter_gpu.py -m gpu -x -q -k "sample_count_is_exact or dynamic_schedule or edge_cases or test_float4_sums_are_exact") > gpurun_out/r02_sanitizer_memcheck_iter.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck_iter.log
(timeout 900 $S --tool racecheck python -m pytest tests/test_iter_gpu.py -m gpu -x -q -k "sample_count_is_exact or dynamic_schedule") > gpurun_out/r02_sanitizer_racecheck_iter.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck_iter.log
(timeout 900 $S --tool memcheck python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "not 1048576 and not 1060921") > gpurun_out/r02_sanitizer_memcheck_sort.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck_sort.log
(timeout 900 $S --tool racecheck python -m pytest tests/test_sort_gpu.py -m gpu -x -q -k "multisort or ignore_max") > gpurun_out/r02_sanitizer_racecheck_sort.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck_sort.log
