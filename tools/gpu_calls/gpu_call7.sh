mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/r02_gputests_6.log 2>&1; tail -12 gpurun_out/r02_gputests_6.log
CASES=G6F,G3,G24H timeout 900 python tools/iter_bench.py '' 'POINTS=2' 'POINTS=2,ITER_MIN_CTAS=6' 'STILL=0' 'STILL=0,POINTS=2' 'STILL=0,POINTS=2,ITER_MIN_CTAS=8' > gpurun_out/r02_iter_variants3.txt 2>&1; tail -20 gpurun_out/r02_iter_variants3.txt
timeout 200 python tools/host_profile.py > gpurun_out/r02_host_profile.txt 2>&1; head -50 gpurun_out/r02_host_profile.txt
(time timeout 400 python bench.py --no-cpu-baseline) > gpurun_out/r02_bench_pts2.json 2> gpurun_out/r02_bench_pts2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_pts2.json')); print('ms/step', d['ms_per_step'], 'iter', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('frames_per_second'), '%.4e'%v['value'])
PY
