mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $T tools/band_check.py > gpurun_out/r02_band_check_n2.json 2> gpurun_out/r02_band_check_n2.err; tail -2 gpurun_out/r02_band_check_n2.json; tail -3 gpurun_out/r02_band_check_n2.err
timeout 300 $T tools/shared_frame_check.py > gpurun_out/r02_shared_frame_n2.json 2> gpurun_out/r02_shared_frame_n2.err; tail -2 gpurun_out/r02_shared_frame_n2.json; tail -5 gpurun_out/r02_shared_frame_n2.err
(time timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err; tail -c 1500 gpurun_out/r02_bench_n2.json
(time timeout 300 $T bench.py --impl reference --gpus 2 --steps 1 --warmup 0) > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; tail -c 600 gpurun_out/r02_bench_ref_n2.json
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_5.log 2>&1; tail -4 gpurun_out/r02_gputests_5.log
