"""4K iterate stage by accumulation layout: python tools/grid4k_probe.py [GENOME SPP]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render
N.init(0)
gname = sys.argv[1] if len(sys.argv) > 1 else 'G6F'
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w, h = int(os.environ.get('W', 3840)), int(os.environ.get('H', 2160))
for mode, swz, spill in (('float4', True, True), ('float4', True, False), ('float4', False, True), ('float4', False, False), ('packed', False, False)):
    rmgr = render.RenderManager(seed=1)
    rmgr.accumulate, rmgr.swizzle, rmgr.spill, rmgr.hot_bins = mode, swz, spill, False
    gnm = samples.GENOMES[gname]()
    gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    ms = []
    for i in range(4):
        e0, e1 = N.Event(), N.Event()
        e0.record(rmgr.stream_a)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        e1.record(rmgr.stream_a); e1.synchronize()
        ms.append(e1.time_since(e0))
    n = rmgr.last_iter_samples
    print('%-5s %dx%d spp %d %-7s swizzle=%d spill=%d: %.2f ms  %.4g it/s' % (gname, w, h, spp, mode, swz, spill, min(ms[1:]), n / min(ms[1:]) * 1e3), flush=True)
    rmgr.fb.free()
