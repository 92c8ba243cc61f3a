"""
Multi-GPU check of the banded OUTPUT path: every GPU renders its share of a still, the
histograms are all-reduced, every GPU filters its band, converts its own output rows
(cb_convert_rows) and copies them into one page-locked shared-memory frame.  The root then
renders the same combined histogram alone (whole-frame filter + convert) with the same
dither seeds; the two 8-bit frames must be byte-identical.

    torchrun --nproc-per-node 2 tools/shared_frame_check.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render, multigpu, mwc

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
shared = multigpu.SharedFrame(rdr.out.shape(dim), rdr.out.dtype, rank, world, barrier=dist.barrier)
reducer = multigpu.HistReducer(root=None)
INTEGER = os.environ.get('INTEGER', '1') == '1'      # reduce unscaled integer level sums
saved = {}


def hook(fb, dim_, stream):
    reducer(fb, dim_, stream)
    stream.synchronize()
    saved['hist'] = N.from_device(fb.d_front, shape, np.float32)
    # identical dither streams on every rank and for the reference frame below
    N.memcpy_htod(fb.d_seeds, mwc.make_seeds(fb.nstreams, host_seed=99), stream)


hook.integer_sums = INTEGER
rmgr.hist_hook = hook
rmgr.band_filter = multigpu.BandFilter(rank, world, shared=shared)
evt, out = rmgr.queue_frame(rdr, gnm, gprof, tc)
evt.synchronize()
dist.barrier()
if rank == 0:
    banded = np.array(shared.array)
    rmgr.band_filter = None
    def inject(fb, dim_, stream):
        N.memcpy_htod(fb.d_front, saved['hist'], stream)
        N.memcpy_htod(fb.d_seeds, mwc.make_seeds(fb.nstreams, host_seed=99), stream)
    inject.integer_sums = INTEGER
    rmgr.hist_hook = inject
    evt, whole = rmgr.queue_frame(rdr, gnm, gprof, tc)
    evt.synchronize()
    whole = np.array(whole)
    print(json.dumps({'world': world, 'frame': [w, h, spp],
                      'output_rows': [multigpu.BandFilter(r, world, comm=False).output_rows(dim)
                                      for r in range(world)],
                      'integer_sums': INTEGER, 'byte_identical': bool(np.array_equal(banded, whole)),
                      'differing_bytes': int((banded != whole).sum()),
                      'frame_mean': float(whole[..., :3].mean())}))
dist.barrier()
shared.close()
dist.destroy_process_group()
