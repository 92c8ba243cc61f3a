mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gputests_final.log 2>&1; tail -8 gpurun_out/r02_gputests_17.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke3.log
(time timeout 900 python bench.py) > gpurun_out/r02_bench_final2_n1.json 2> gpurun_out/r02_bench_final2_n1.err; tail -5 gpurun_out/r02_bench_final2_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_final2_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e'], d['roofline']['l2_atomic'], d['roofline']['kernel_ms'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e'), v.get('frames_per_second'))
P
