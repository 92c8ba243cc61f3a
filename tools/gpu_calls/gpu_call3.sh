mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/r02_gputests_2.log 2>&1; tail -25 gpurun_out/r02_gputests_2.log
timeout 300 python tools/hot_bench.py G2M G6F > gpurun_out/r02_hot_bench2.jsonl 2> gpurun_out/r02_hot_bench2.err; tail -2 gpurun_out/r02_hot_bench2.err
(time timeout 600 python bench.py) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_launches.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 1 -c 1 -o gpurun_out/r02_cb_iter_1080 -f python tools/one_frame.py G6F 1920 1080 2000 hot=0 filters=0 > gpurun_out/r02_ncu1.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 1 -c 1 -o gpurun_out/r02_cb_iter_8k_packed -f python tools/one_frame.py G24H 7680 4320 500 filters=0 > gpurun_out/r02_ncu2.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 3 -c 1 -o gpurun_out/r02_cb_iter_g2m_hot -f python tools/one_frame.py G2M 1920 1080 1000 hot=1 filters=0 > gpurun_out/r02_ncu3.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 1 -c 1 -o gpurun_out/r02_cb_iter_1080_blur -f python tools/one_frame.py G6F 1920 1080 2000 hot=0 filters=0 blur=1 > gpurun_out/r02_ncu4.out 2>&1
ls -la gpurun_out | tail -20
