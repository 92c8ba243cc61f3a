mkdir -p gpurun_out
W=3840 H=2160 SPP=500 timeout 300 python tools/hot_bench.py G6F > gpurun_out/r02_g6f_4k_500.jsonl 2>&1; tail -6 gpurun_out/r02_g6f_4k_500.jsonl | cut -c1-200
W=3840 H=2160 SPP=4000 timeout 300 python tools/hot_bench.py G6F > gpurun_out/r02_g6f_4k_4000.jsonl 2>&1; tail -6 gpurun_out/r02_g6f_4k_4000.jsonl | cut -c1-200
timeout 300 python tools/accuracy_probe.py > gpurun_out/r02_accuracy_g6f.json 2>&1; cat gpurun_out/r02_accuracy_g6f.json | head -60
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bilateral_tile -s 2 -c 2 -o gpurun_out/r02_k_bilateral_tile -f env W=3840 H=2160 REPS=2 python tools/bilat_bench.py 0 4 > gpurun_out/r02_ncu_tile.out 2>&1; tail -2 gpurun_out/r02_ncu_tile.out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gputests_10.log 2>&1; tail -4 gpurun_out/r02_gputests_10.log
