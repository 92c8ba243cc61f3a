mkdir -p gpurun_out
# ncu --set full of the bilateral kernels (prologues + main pass): 1080p directions 1, 5, 0
# (window<1>, window<4>, fast<15>), then the TMA tile kernel at 4K (direction 0)
REPS=1 timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_bilat -c 9 -f -o gpurun_out/r02_bilateral_1080 python tools/bilat_bench.py 1 5 0 > gpurun_out/r02_ncu_bilat_1080.out 2>&1
tail -3 gpurun_out/r02_ncu_bilat_1080.out
W=3840 H=2160 REPS=1 timeout 45 ncu --set full --clock-control none --import-source on -k regex:k_bilateral -c 1 -f -o gpurun_out/r02_bilateral_4k python tools/bilat_bench.py 0 > gpurun_out/r02_ncu_bilat_4k.out 2>&1
tail -3 gpurun_out/r02_ncu_bilat_4k.out
