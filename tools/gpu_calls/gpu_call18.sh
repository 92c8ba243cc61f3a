mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_sort_gpu.py -m gpu -x -q) > gpurun_out/r02_gputests_sort.log 2>&1; tail -8 gpurun_out/r02_gputests_sort.log
timeout 300 python tools/deferred_bench.py 27 > gpurun_out/r02_deferred_bench.json 2>&1; cat gpurun_out/r02_deferred_bench.json | tail -40
