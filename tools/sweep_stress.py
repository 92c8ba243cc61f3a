"""Does the spill sweep keep sums exact whatever the grid of CTAs (i.e. whatever the timing)?
python tools/sweep_stress.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from helpers import exact_level_sums
N.init(0)
gnm = samples.g3()
w, h, spp = 160, 90, 20000
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
for grid in (8, 64, 300, 1024):
    res = {}
    for mode in ('exact', 'swept'):
        rmgr = render.RenderManager(seed=17)
        rmgr.accumulate, rmgr.hot_bins, rmgr.schedule, rmgr.iter_grid = 'float4', False, 'static', grid
        rmgr.spill_interval = 1 << 21
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        rmgr._copy(rdr, gnm); rmgr._interp(rdr, gnm, dim, tc, 0.0)
        if mode == 'exact':
            res[mode] = exact_level_sums(N, rmgr, rdr, gnm, gprof, dim, tc, waves_per_chunk=max(1, 1024 // grid))
        else:
            rmgr._iter(rdr, gnm, gprof, dim, tc); rmgr.stream_a.synchronize()
            res[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
            fine = N.from_device(rmgr.fb.d_left, (dim.ah * dim.astride, 4), np.float32)
        rmgr.fb.free()
    exact, swept = res['exact'], res['swept']
    k = np.float32(1.0 / 255.0)
    small = (exact[..., :3] < 2 ** 24).all(axis=-1)
    ok_small = np.array_equal(swept[..., :3][small], (exact[..., :3].astype(np.float32) * k)[small])
    big = ~small
    ref = exact[..., :3][big] / 255.0
    rel = np.abs(swept[..., :3][big].astype(np.float64) - ref) / ref
    print('grid %4d: counts equal %s, small bins bit-exact %s, big bins %d max rel %.3g, fine max %.3g' % (
        grid, np.array_equal(swept[..., 3].astype(np.int64), exact[..., 3]), ok_small, big.sum(), rel.max(), fine.max()), flush=True)
