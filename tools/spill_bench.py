"""A/B of the spill sweep of the float4 accumulation (render.RenderManager.spill): time of the
iterate stage (fills + cb_iter + cb_hist_finish) with sweeps off and on, still and
motion-blur variants.   python tools/spill_bench.py [GENOME ...]   (env W H SPP)  -> JSON lines"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render

N.init(0)
names = sys.argv[1:] or ['G6F', 'G3', 'G24H']
W, H = int(os.environ.get('W', 1920)), int(os.environ.get('H', 1080))
for gname in names:
    spp = int(os.environ.get('SPP', 500 if gname == 'G24H' else 2000))
    gnm = samples.GENOMES[gname]()
    for fw in (0, 1e-9):
        gprof = profile.wrap(dict(width=W, height=H, spp=spp, frame_width=fw, start=1, end=2), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        for spill, interval in ((False, 0), (True, 1 << 25), (True, 1 << 24), (True, 1 << 23)):
            rmgr = render.RenderManager(seed=1)
            rmgr.hot_bins, rmgr.spill = False, spill
            if spill:
                rmgr.spill_interval = interval
            rdr = render.Renderer(gnm, gprof)
            dim = rmgr.fb.set_dim(W, H)
            rmgr._copy(rdr, gnm)
            rmgr._interp(rdr, gnm, dim, tc, 0.0)
            ms = []
            for i in range(5):
                e0, e1 = N.Event(), N.Event()
                e0.record(rmgr.stream_a)
                rmgr._iter(rdr, gnm, gprof, dim, tc)
                e1.record(rmgr.stream_a)
                e1.synchronize()
                ms.append(e1.time_since(e0))
            n = rmgr.last_iter_samples
            best = min(ms[1:])
            print(json.dumps(dict(genome=gname, width=W, height=H, spp=spp, motion_blur=fw > 0,
                                  spill=spill, interval=interval,
                                  window=rmgr._spill_window(dim.ah * dim.astride, n), ms=best,
                                  samples_per_s=n / best * 1e3, all_ms=ms)), flush=True)
            rmgr.fb.free()
