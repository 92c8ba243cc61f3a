mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/r02_gputests_3.log 2>&1; tail -25 gpurun_out/r02_gputests_3.log
CASES=G6F,G3,G24H timeout 600 python tools/iter_bench.py '' 'PAL_COMPACT=1' 'STILL=0' 'STILL=0,PAL_COMPACT=0' 'STILL=0,ITER_MIN_CTAS=6' 'STILL=0,ITER_MIN_CTAS=6,PAL_COMPACT=0' > gpurun_out/r02_iter_variants.txt 2>&1; tail -20 gpurun_out/r02_iter_variants.txt
timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_1080.txt 2>&1; cat gpurun_out/r02_bilat_1080.txt
W=3840 H=2160 timeout 200 python tools/bilat_bench.py > gpurun_out/r02_bilat_4k.txt 2>&1; cat gpurun_out/r02_bilat_4k.txt
(time timeout 600 python bench.py) > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; tail -3 gpurun_out/r02_bench_n1b.err
