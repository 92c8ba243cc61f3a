"""
Multi-GPU check of the sharded filter chain: every GPU runs its share of a still,
the histograms are all-reduced, and the chain runs (a) sharded -- BandFilter: each
GPU filters its band of rows plus halo, bands gathered on the root -- and (b) whole,
on the root, from a saved copy of the same combined histogram.  The two filtered
float4 frames must be equal bit for bit.

    torchrun --nproc-per-node 2 tools/band_check.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render, multigpu

rank, world, local = multigpu.env_rank_world()
import torch.distributed as dist
multigpu.init_process_group('nccl')
N.init(local)
gnm = samples.g6f()
w, h, spp = (int(x) for x in os.environ.get('FRAME', '1920,1080,200').split(','))
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=3, rank=rank, world=world)
rdr = render.Renderer(gnm, gprof)
s = rmgr.stream_a
dim = rmgr.fb.set_dim(w, h)
shape = (dim.ah, dim.astride, 4)
rmgr._copy(rdr, gnm)
rmgr._interp(rdr, gnm, dim, tc, 0.0)
rmgr._iter(rdr, gnm, gprof, dim, tc)
comm = multigpu.NativeComm(rank, world) if os.environ.get('COLLECTIVES') == 'native' else None
multigpu.HistReducer(root=None, comm=comm)(rmgr.fb, dim, s)
s.synchronize()
hist = N.from_device(rmgr.fb.d_front, shape, np.float32)
rmgr.band_filter = multigpu.BandFilter(rank, world, root=0, comm=comm or True)
rmgr._filter(rdr, gprof, dim, tc)
s.synchronize()
banded = N.from_device(rmgr.fb.d_front, shape, np.float32)
dist.barrier()
if rank == 0:
    N.memcpy_htod(rmgr.fb.d_front, hist, s)
    rmgr.band_filter = None
    rmgr._filter(rdr, gprof, dim, tc)
    s.synchronize()
    whole = N.from_device(rmgr.fb.d_front, shape, np.float32)
    same = banded.view(np.uint32) == whole.view(np.uint32)
    print(json.dumps({'world': world, 'frame': [w, h, spp],
                      'collectives': 'native (cb_hist_reduce, cb_band_gather)' if comm else 'torch.distributed',
                      'halo_rows': multigpu.chain_reach(rdr.filts, gprof, tc),
                      'bands': [multigpu.band_rows(dim.ah, r, world) for r in range(world)],
                      'samples_in_histogram': float(hist[..., 3].sum()),
                      'bit_identical': bool(same.all()),
                      'differing_words': int((~same).sum()),
                      'filtered_mean': float(whole.mean())}))
dist.barrier()
dist.destroy_process_group()
