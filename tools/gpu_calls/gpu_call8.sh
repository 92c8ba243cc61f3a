mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/r02_gputests_7.log 2>&1; tail -6 gpurun_out/r02_gputests_7.log
timeout 200 python tools/host_profile.py > gpurun_out/r02_host_profile2.txt 2>&1; head -4 gpurun_out/r02_host_profile2.txt
timeout 300 python tools/hot_bench.py G2M G6F G3 > gpurun_out/r02_hot_bench3.jsonl 2> gpurun_out/r02_hot_bench3.err
timeout 200 python tools/filter_bench.py > gpurun_out/r02_filter_1080.txt 2>&1; tail -14 gpurun_out/r02_filter_1080.txt
W=3840 H=2160 timeout 200 python tools/filter_bench.py > gpurun_out/r02_filter_4k.txt 2>&1; tail -14 gpurun_out/r02_filter_4k.txt
W=3840 H=2160 timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_4kc.txt 2>&1; cat gpurun_out/r02_bilat_4kc.txt
timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_1080c.txt 2>&1; cat gpurun_out/r02_bilat_1080c.txt
(time timeout 500 python bench.py) > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final_n1.json')); print('ms/step', d['ms_per_step'], 'iter', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'value %.4e'%d['value'])
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('frames_per_second'), '%.4e'%v['value'])
PY
(time timeout 300 python bench.py --impl reference --steps 2 --warmup 0) > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; tail -c 400 gpurun_out/r02_bench_ref_n1.json
