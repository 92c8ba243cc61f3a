mkdir -p gpurun_out
timeout 300 python tools/accuracy_probe.py > gpurun_out/r02_accuracy_g6f_v3.json 2>&1; python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/r02_accuracy_g6f_v3.json'))
    print('hottest share', d['hottest_bin_share'])
    for m,rows in d['modes'].items():
        print(m, [(r['rank'], '%.2e'%r['mean_rel_err'], '%.2e'%r['max_rel_err']) for r in rows])
except Exception as e:
    print(open('gpurun_out/r02_accuracy_g6f_v3.json').read()[-1500:])
P
(time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5) > gpurun_out/r02_gputests_14.log 2>&1; tail -12 gpurun_out/r02_gputests_14.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r02_smoke2.log 2>&1; tail -2 gpurun_out/r02_smoke2.log
