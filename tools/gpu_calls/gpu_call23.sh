mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_iter_gpu.py -m gpu -x -q -k "exact or dynamic or hot") > gpurun_out/r02_gputests_16.log 2>&1; tail -6 gpurun_out/r02_gputests_16.log
(time timeout 900 python bench.py --no-extras) > gpurun_out/r02_bench_dyn3_n1.json 2> gpurun_out/r02_bench_dyn3_n1.err; tail -5 gpurun_out/r02_bench_dyn3_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_dyn3_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e'], d['roofline']['l2_atomic'], d['roofline']['kernel_ms'])
print(d['cpu_baseline_filters'])
P
