mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q --durations=4) > gpurun_out/r02_gputests_9.log 2>&1; tail -8 gpurun_out/r02_gputests_9.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 200 python tools/filter_bench.py > gpurun_out/r02_filter_1080b.txt 2>&1; tail -12 gpurun_out/r02_filter_1080b.txt
W=3840 H=2160 timeout 200 python tools/filter_bench.py > gpurun_out/r02_filter_4kb.txt 2>&1; tail -12 gpurun_out/r02_filter_4kb.txt
timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_1080d.txt 2>&1; cat gpurun_out/r02_bilat_1080d.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_launches2.out 2>&1
W=7680 H=4320 SPP=500 timeout 300 python tools/hot_bench.py G6F > gpurun_out/r02_g6f_8k.jsonl 2>&1; tail -3 gpurun_out/r02_g6f_8k.jsonl
