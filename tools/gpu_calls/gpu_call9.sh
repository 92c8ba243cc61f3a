mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
FRAME=3840,2160,100 timeout 300 $T tools/shared_frame_check.py > gpurun_out/r02_shared_frame_n8.json 2> gpurun_out/r02_shared_frame_n8.err; tail -1 gpurun_out/r02_shared_frame_n8.json; tail -3 gpurun_out/r02_shared_frame_n8.err
(time timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3) > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -4 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=8 ms/step', d['ms_per_step'], 'value %.4e'%d['value'], 'e2e', d['e2e']['ms_per_step'], d['config'].get('nccl_reduce_ms'))
        for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('frames_per_second'), '%.4e'%v['value'], v.get('nccl_reduce_ms'))
PY
(time timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3 --band-output gather --no-extras --workload still4k) > gpurun_out/r02_bench_n8_4k_gather.json 2> gpurun_out/r02_bench_n8_4k_gather.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n8_4k_gather.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=8 4k gather ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config'].get('nccl_reduce_ms'), d['config'].get('nccl_gather_ms'))
PY
