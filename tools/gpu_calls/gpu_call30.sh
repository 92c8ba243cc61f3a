mkdir -p gpurun_out
(time timeout 900 python bench.py --no-cpu-baseline) > gpurun_out/r02_bench_pol_n1.json 2> gpurun_out/r02_bench_pol_n1.err; tail -3 gpurun_out/r02_bench_pol_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_pol_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e'], d['roofline']['l2_atomic'], d['roofline']['kernel_ms'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'), v.get('frames_per_second'))
P
(timeout 900 python -m pytest tests/test_render_gpu.py tests/test_iter_gpu.py -m gpu -x -q -k "4k or 8k or benchmark_sizes or conservation") 2>&1 | tail -3
