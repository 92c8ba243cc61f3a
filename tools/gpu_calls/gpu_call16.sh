mkdir -p gpurun_out
SCHED=static timeout 200 python tools/cta_timeline.py > gpurun_out/r02_cta_timeline.txt 2>&1
SCHED=dynamic timeout 200 python tools/cta_timeline.py >> gpurun_out/r02_cta_timeline.txt 2>&1; cat gpurun_out/r02_cta_timeline.txt
timeout 300 python tools/hot_bench.py G6F G2M > gpurun_out/r02_hot_bench_v4.jsonl 2>&1; cut -c1-210 gpurun_out/r02_hot_bench_v4.jsonl
timeout 300 python tools/accuracy_probe.py > gpurun_out/r02_accuracy_g6f_v3.json 2>&1; tail -5 gpurun_out/r02_accuracy_g6f_v3.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 1 -c 1 -o gpurun_out/r02_cb_iter_1080_dyn -f python tools/one_frame.py G6F 1920 1080 2000 hot=0 filters=0 > gpurun_out/r02_ncu_dyn.out 2>&1; tail -2 gpurun_out/r02_ncu_dyn.out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_launches3.out 2>&1; tail -2 gpurun_out/r02_launches3.out | cut -c1-300
