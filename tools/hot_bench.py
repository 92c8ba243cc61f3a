"""Time the iterate stage with and without the hot-bin variant (and the motion-blur
variant): python tools/hot_bench.py [GENOME ...]   -> JSON lines."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render

N.init(0)
names = sys.argv[1:] or ['G6F', 'G3', 'G2M', 'G24H']
W, H = int(os.environ.get('W', 1920)), int(os.environ.get('H', 1080))
for gname in names:
    spp = int(os.environ.get('SPP', 500 if gname == 'G24H' else 2000))
    gnm = samples.GENOMES[gname]()
    for fw in (0, 1e-9):
        gprof = profile.wrap(dict(width=W, height=H, spp=spp, frame_width=fw, start=1, end=2), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        for hot in (False, True, 'auto'):
            rmgr = render.RenderManager(seed=1)
            rmgr.hot_bins = hot
            rdr = render.Renderer(gnm, gprof)
            dim = rmgr.fb.set_dim(W, H)
            rmgr._copy(rdr, gnm)
            rmgr._interp(rdr, gnm, dim, tc, 0.0)
            ms = []
            for i in range(4):
                e0, e1 = N.Event(), N.Event()
                e0.record(rmgr.stream_a)
                rmgr._iter(rdr, gnm, gprof, dim, tc)
                e1.record(rmgr.stream_a)
                e1.synchronize()
                ms.append(e1.time_since(e0))
            n = rmgr.last_iter_samples
            best = min(ms[1:])
            print(json.dumps(dict(genome=gname, width=W, height=H, spp=spp, motion_blur=fw > 0,
                                  hot_bins=str(hot), used_hot=rmgr.last_iter_hot, ms=best,
                                  samples_per_s=n / best * 1e3, all_ms=ms)), flush=True)
            rmgr.fb.free()
