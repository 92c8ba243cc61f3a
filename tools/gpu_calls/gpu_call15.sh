mkdir -p gpurun_out
timeout 600 python tools/sched_bench.py G6F G24H > gpurun_out/r02_sched_bench2.jsonl 2>&1; cut -c1-200 gpurun_out/r02_sched_bench2.jsonl
W=3840 H=2160 SPP=500 timeout 300 python tools/sched_bench.py G6F > gpurun_out/r02_sched_bench2_4k.jsonl 2>&1; cut -c1-200 gpurun_out/r02_sched_bench2_4k.jsonl
(timeout 900 python -m pytest tests/test_iter_gpu.py -m gpu -x -q) > gpurun_out/r02_gputests_13.log 2>&1; tail -8 gpurun_out/r02_gputests_13.log
timeout 900 python bench.py > gpurun_out/r02_bench_dyn_n1.json 2> gpurun_out/r02_bench_dyn_n1.err; tail -3 gpurun_out/r02_bench_dyn_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_dyn_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e'], d['roofline']['l2_atomic'], d['roofline']['kernel_ms'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'), v.get('frames_per_second'))
P
