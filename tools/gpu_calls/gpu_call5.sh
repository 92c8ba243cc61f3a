mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/r02_gputests_4.log 2>&1; tail -12 gpurun_out/r02_gputests_4.log
timeout 200 python tools/iter_timeline.py > gpurun_out/r02_iter_timeline.txt 2>&1; cat gpurun_out/r02_iter_timeline.txt
timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_1080b.txt 2>&1; cat gpurun_out/r02_bilat_1080b.txt
W=3840 H=2160 timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_4kb.txt 2>&1; cat gpurun_out/r02_bilat_4kb.txt
CASES=G6F,G3,G24H timeout 600 python tools/iter_bench.py '' 'STILL=0' 'STILL=0,ITER_MIN_CTAS=6' > gpurun_out/r02_iter_variants2.txt 2>&1; tail -12 gpurun_out/r02_iter_variants2.txt
(time timeout 300 python bench.py --no-extras --hot-bins off --no-cpu-baseline) > gpurun_out/r02_bench_hotoff.json 2> gpurun_out/r02_bench_hotoff.err
(time timeout 300 python bench.py --no-extras --no-cpu-baseline) > gpurun_out/r02_bench_hotauto.json 2> gpurun_out/r02_bench_hotauto.err
python - <<'PY'
import json
for f in ('hotoff','hotauto'):
    d=json.load(open('gpurun_out/r02_bench_%s.json'%f)); print(f, 'ms/step', d['ms_per_step'], 'iter', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])
PY
