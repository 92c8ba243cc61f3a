"""Does cb_iter's speed depend on where the histogram sits (address hash vs hot bins)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render

class Ptr(object):
    def __init__(self, ptr): self.ptr = ptr
    def __int__(self): return self.ptr

N.init(0)
gname = os.environ.get('G', 'G6F')
w, h, spp = 1920, 1080, 1000
gnm = samples.GENOMES[gname]()
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=1); rmgr.swizzle = {'1': True, '0': False}.get(os.environ.get('SWZ', 'auto'), 'auto')
rdr = render.Renderer(gnm, gprof)
dim = rmgr.fb.set_dim(w, h)
rmgr._copy(rdr, gnm)
rmgr._interp(rdr, gnm, dim, tc, 0.0)
nb = 16 * dim.ah * dim.astride
big = N.DeviceBuffer(nb + (64 << 20))
for off in (0, 256, 512, 1024, 4096, 8192, 65536, 1 << 20, (1 << 20) + 256, 2 << 20, 3 << 20,
            (5 << 20) + 4096, 16 << 20, 32 << 20, 48 << 20):
    rmgr.fb.d_front = Ptr(big.ptr + off); rmgr.fb.d_left = Ptr(big.ptr + off)
    ms = []
    for i in range(3):
        e0, e1 = N.Event(), N.Event()
        e0.record(rmgr.stream_a)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        e1.record(rmgr.stream_a); e1.synchronize()
        ms.append(e1.time_since(e0))
    print('%s offset %9d: %.2f ms  (%s)' % (gname, off, min(ms), ' '.join('%.2f' % m for m in ms)), flush=True)
