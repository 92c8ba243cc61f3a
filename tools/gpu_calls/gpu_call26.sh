S="compute-sanitizer --error-exitcode 9"
(timeout 900 $S --tool memcheck python -m pytest tests/test_iter_gpu.py -m gpu -x -q -k "sample_count_is_exact or dynamic_schedule or edge_cases or test_float4_sums_are_exact") > gpurun_out/r02_sanitizer_memcheck_iter.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck_iter.log
(timeout 600 python -m pytest tests/test_iter_gpu.py -m gpu -x -q -k "exact") 2>&1 | tail -3
