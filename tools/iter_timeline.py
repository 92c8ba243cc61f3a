"""Where does the time between the events around RenderManager._iter go?  Runs the bench's
device step with an event after every enqueue (L2 flushed before each step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render

N.init(0)
gnm = samples.g6f()
gprof = profile.wrap(dict(width=1920, height=1080, spp=2000, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
flush = N.DeviceBuffer(512 << 20)
for hot in (False, 'auto'):
    for do_flush in (0, 1):
        rmgr = render.RenderManager(seed=1)
        rmgr.hot_bins = hot
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(1920, 1080)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, tc, 0.0)
        s = rmgr.stream_a
        marks = []
        orig_launch, orig_scan = rmgr._launch_iter, N.lib().cb_hot_scan

        def launch(*a, **k):
            marks.append(('before launch', N.Event().record(s)))
            orig_launch(*a, **k)
            marks.append(('after launch', N.Event().record(s)))
        rmgr._launch_iter = launch
        res = []
        for rep in range(5):
            if do_flush:
                N.fill32(flush, (512 << 20) // 4, 0, s)
            s.synchronize()
            del marks[:]
            e0 = N.Event().record(s)
            rmgr._iter(rdr, gnm, gprof, dim, tc)
            e1 = N.Event().record(s)
            e1.synchronize()
            res.append((e1.time_since(e0), [(n, e.time_since(e0)) for n, e in marks]))
        tot, m = res[-1]
        print('hot=%s flush=%d total %.2f ms: %s   (all: %s)' % (
            hot, do_flush, tot, ' '.join('%s@%.2f' % (n.split()[0][0] + n.split()[1][0], t) for n, t in m),
            ' '.join('%.2f' % r[0] for r in res)), flush=True)
        rmgr.fb.free()
