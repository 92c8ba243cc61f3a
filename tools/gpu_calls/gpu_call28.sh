mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 1 -c 1 -o gpurun_out/r02_cb_iter_8k_packed_final -f python tools/one_frame.py G24H 7680 4320 500 hot=0 filters=0 > gpurun_out/r02_ncu_8k_final.out 2>&1; tail -1 gpurun_out/r02_ncu_8k_final.out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cb_iter -s 1 -c 1 -o gpurun_out/r02_cb_iter_4k_final -f python tools/one_frame.py G6F 3840 2160 1000 hot=0 filters=0 > gpurun_out/r02_ncu_4k_final.out 2>&1; tail -1 gpurun_out/r02_ncu_4k_final.out
