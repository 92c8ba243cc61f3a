mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
(time timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3) > gpurun_out/r02_bench_dyn_n8.json 2> gpurun_out/r02_bench_dyn_n8.err; tail -4 gpurun_out/r02_bench_dyn_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_dyn_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=8 ms/step', d['ms_per_step'], 'value %.4e'%d['value'], 'e2e', d['e2e']['ms_per_step'], d['config'].get('nccl_reduce_ms'))
        for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), (v.get('e2e') or {}).get('ms_per_step'), v.get('frames_per_second'), '%.4e'%v['value'], v.get('nccl_reduce_ms'))
PY
