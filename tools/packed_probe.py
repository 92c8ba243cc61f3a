"""float4 vs packed-u64 accumulation at 4K and 8K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render
N.init(0)
for gname, w, h, spp in (('G6F', 7680, 4320, 500), ('G24H', 7680, 4320, 500), ('G6F', 3840, 2160, 1000), ('G6F', 1920, 1080, 2000)):
    for mode in ('float4', 'packed'):
        rmgr = render.RenderManager(seed=1)
        rmgr.accumulate = mode
        gnm = samples.GENOMES[gname]()
        gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        rmgr._copy(rdr, gnm); rmgr._interp(rdr, gnm, dim, tc, 0.0)
        ms = []
        for i in range(3):
            a, b = N.Event(), N.Event()
            a.record(rmgr.stream_a)
            rmgr._iter(rdr, gnm, gprof, dim, tc)
            b.record(rmgr.stream_a); b.synchronize()
            ms.append(b.time_since(a))
        n = rmgr.last_iter_samples
        print('%-5s %dx%d spp %-5d %-7s %9.2f ms  %.4g it/s' % (gname, w, h, spp, mode, min(ms), n / min(ms) * 1e3), flush=True)
        rmgr.fb.free()
