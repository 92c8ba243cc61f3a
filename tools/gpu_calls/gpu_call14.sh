mkdir -p gpurun_out
timeout 600 python tools/sched_bench.py G6F G3 G24H G2M > gpurun_out/r02_sched_bench.jsonl 2>&1; cut -c1-200 gpurun_out/r02_sched_bench.jsonl
W=3840 H=2160 SPP=500 timeout 300 python tools/sched_bench.py G6F > gpurun_out/r02_sched_bench_4k.jsonl 2>&1; cut -c1-200 gpurun_out/r02_sched_bench_4k.jsonl
timeout 300 python tools/spill_debug.py 2>&1 | tail -8
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gputests_12.log 2>&1; tail -8 gpurun_out/r02_gputests_12.log
