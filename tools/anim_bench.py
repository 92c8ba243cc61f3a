"""
Config 4: 1080p animation with temporal motion blur, frames partitioned over the GPUs
of the job (frame k -> rank k mod N, no collective).  Renders through
RenderManager.queue_frame with the reference's software pipelining (queue frame k+1
before waiting on frame k, main.py:63-76) and reports steady-state frames/s.

    python tools/anim_bench.py [--frames 48] [--spp 2000]
    torchrun --nproc-per-node 8 tools/anim_bench.py --frames 96
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render, multigpu

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=48)
ap.add_argument('--spp', type=int, default=2000)
ap.add_argument('--genome', default='G6F')
ap.add_argument('--frame-width', type=float, default=1.0)
args = ap.parse_args()

rank, world, local = multigpu.env_rank_world()
dist = None
if world > 1:
    import torch, torch.distributed as dist
    multigpu.init_process_group('nccl')
N.init(local)
gnm = samples.GENOMES[args.genome](animated=True) if args.genome == 'G6F' else samples.GENOMES[args.genome]()
gprof = profile.wrap(dict(width=1920, height=1080, spp=args.spp, fps=24, duration=30,
                          frame_width=args.frame_width), gnm)
times = [t[0] for _, t in profile.enumerate_times(gprof)][:args.frames]
mine = multigpu.partition_frames(times, rank, world)
rmgr = render.RenderManager(seed=1 + rank)
rdr = render.Renderer(gnm, gprof)
# warm-up: compile, allocate
for t in mine[:2]:
    evt, buf = rmgr.queue_frame(rdr, gnm, gprof, t)
    evt.synchronize()
if dist is not None:
    dist.barrier()
t0 = time.perf_counter()
pending, gpu_ms, checksum = None, [], 0
for t in mine + [None]:
    nxt = rmgr.queue_frame(rdr, gnm, gprof, t) if t is not None else None
    if pending is not None:
        pending[0].synchronize()
        gpu_ms.append(pending[0].time())
        checksum += int(pending[1][::97, ::89, 0].sum())
    pending = nxt
wall = time.perf_counter() - t0
if dist is not None:
    tt = torch.tensor([wall], device='cuda')
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall = float(tt[0])
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps({'config': '1080p %s animation, %d spp, frame_width %g, %d frames over %d GPU(s)'
                                % (args.genome, args.spp, args.frame_width, len(times), world),
                      'frames_per_second': len(times) / wall, 'wall_s': wall,
                      'gpu_ms_per_frame_rank0': float(np.mean(gpu_ms)), 'checksum_rank0': checksum,
                      'iterations_per_second': len(times) * args.spp * 1920 * 1080 / wall}))
