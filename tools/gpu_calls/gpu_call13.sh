mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_iter_gpu.py tests/test_render_gpu.py -m gpu -x -q) > gpurun_out/r02_gputests_11.log 2>&1; tail -15 gpurun_out/r02_gputests_11.log
timeout 400 python tools/spill_bench.py G6F G3 G24H > gpurun_out/r02_spill_bench.jsonl 2>&1; cut -c1-220 gpurun_out/r02_spill_bench.jsonl
W=3840 H=2160 SPP=500 timeout 300 python tools/spill_bench.py G6F > gpurun_out/r02_spill_bench_4k.jsonl 2>&1; cut -c1-220 gpurun_out/r02_spill_bench_4k.jsonl
timeout 300 python tools/accuracy_probe.py > gpurun_out/r02_accuracy_g6f_v2.json 2>&1; python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/r02_accuracy_g6f_v2.json'))
    for m,rows in d['modes'].items():
        print(m, [(r['rank'], '%.2e'%r['mean_rel_err'], '%.2e'%r['max_rel_err']) for r in rows])
except Exception as e:
    print(open('gpurun_out/r02_accuracy_g6f_v2.json').read()[-2000:])
P
