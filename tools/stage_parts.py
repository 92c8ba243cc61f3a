"""Where the iterate stage's time goes (fills / cb_iter launches / finish), spill on and off.
   W=3840 H=2160 SPP=1000 python tools/stage_parts.py [GENOME]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render
N.init(0)
gname = sys.argv[1] if len(sys.argv) > 1 else 'G6F'
w, h, spp = int(os.environ.get('W', 3840)), int(os.environ.get('H', 2160)), int(os.environ.get('SPP', 1000))
flush = N.DeviceBuffer(512 << 20)
for spill in (False, True):
    rmgr = render.RenderManager(seed=1)
    rmgr.hot_bins, rmgr.spill = False, spill
    gnm = samples.GENOMES[gname]()
    gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    s = rmgr.stream_a
    marks = []
    orig = rmgr._launch_iter
    def launch(*a, **k):
        marks.append(N.Event().record(s))
        orig(*a, **k)
        marks.append(N.Event().record(s))
    rmgr._launch_iter = launch
    best = None
    for rep in range(4):
        N.fill32(flush, (512 << 20) // 4, 0, s)
        del marks[:]
        e0 = N.Event().record(s)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        e1 = N.Event().record(s)
        e1.synchronize()
        parts = (marks[0].time_since(e0), marks[1].time_since(marks[0]), e1.time_since(marks[1]), e1.time_since(e0))
        if rep and (best is None or parts[3] < best[3]):
            best = parts
    print('%s %dx%d spp %d spill=%d: fills %.3f ms, cb_iter %.3f ms, finish %.3f ms, total %.3f ms  (window %d)' % (
        (gname, w, h, spp, spill) + best + (rmgr._spill_window(dim.ah * dim.astride, rmgr.last_iter_samples),)), flush=True)
    rmgr.fb.free()
