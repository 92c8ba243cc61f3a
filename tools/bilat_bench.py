"""Per-direction timing of cb_bilateral_direction on a rendered histogram (1080p by
default; W/H override).  ncu: `ncu --set full -k regex:k_bilateral -c 4 python tools/bilat_bench.py 3`"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render, filters

N.init(0)
w, h = int(os.environ.get('W', 1920)), int(os.environ.get('H', 1080))
gnm = samples.g6f()
gprof = profile.wrap(dict(width=w, height=h, spp=500, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=1)
rdr = render.Renderer(gnm, gprof)
dim = rmgr.fb.set_dim(w, h)
rmgr._copy(rdr, gnm); rmgr._interp(rdr, gnm, dim, tc, 0.0); rmgr._iter(rdr, gnm, gprof, dim, tc)
fb, s, L = rmgr.fb, rmgr.stream_a, N.lib()
L.cb_yuv_to_rgb(fb.d_back.ptr, fb.d_front.ptr, N.byref(dim), s.handle)
s.synchronize()
f32 = np.float32
c1 = filters.gauss_coefs(1)
pats = [int(a) for a in sys.argv[1:]] or list(range(8))
reps = int(os.environ.get('REPS', 5))
for pat in pats:
    best = 1e9
    for _ in range(reps):
        e0, e1 = N.Event(), N.Event()
        e0.record(s)
        N.check(L.cb_bilateral_direction(fb.d_front.ptr, fb.d_back.ptr, fb.d_left.ptr, pat, 15, c1,
                                         f32(6 * w / 1920.), f32(0.05), f32(1.5), f32(0.8), f32(4),
                                         N.byref(dim), s.handle))
        e1.record(s); e1.synchronize()
        best = min(best, e1.time_since(e0))
    print('direction %d: %.1f us' % (pat, best * 1e3), flush=True)
