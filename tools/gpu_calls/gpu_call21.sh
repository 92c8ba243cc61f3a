mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gputests_15.log 2>&1; tail -12 gpurun_out/r02_gputests_15.log
CASES=dynamic:1 timeout 600 python tools/sched_bench.py G6F G3 G24H > gpurun_out/r02_sched_bench3.jsonl 2>&1; cut -c1-200 gpurun_out/r02_sched_bench3.jsonl
timeout 300 python tools/hot_bench.py G2M 2>&1 | cut -c1-200 | tail -6
timeout 900 python bench.py > gpurun_out/r02_bench_dyn2_n1.json 2> gpurun_out/r02_bench_dyn2_n1.err; tail -3 gpurun_out/r02_bench_dyn2_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_dyn2_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e'], d['roofline']['l2_atomic'], d['roofline']['kernel_ms'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'), v.get('frames_per_second'))
P
