"""A/B of the unit schedule (static ownership vs dynamic claiming) and of the spill sweep:
ms of the iterate stage.   python tools/sched_bench.py [GENOME ...]  (env W H SPP)  -> JSON lines"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render

N.init(0)
if os.environ.get('EXTRA'):             # extra -D defines for the iterate module
    from cuburn_b200.code import itergen
    _gen = itergen.generate_source
    def _with_defines(pk, params_const=False, extra_defines=None, **kw):
        d = dict(extra_defines or {})
        d.update(kv.split('=') for kv in os.environ['EXTRA'].split(','))
        return _gen(pk, params_const, extra_defines=d, **kw)
    itergen.generate_source = _with_defines
CASES = [c.split(':') for c in os.environ.get('CASES', 'static:0,static:1,dynamic:0,dynamic:1').split(',')]
names = sys.argv[1:] or ['G6F', 'G3', 'G24H']
W, H = int(os.environ.get('W', 1920)), int(os.environ.get('H', 1080))
flush = N.DeviceBuffer(512 << 20)
for gname in names:
    spp = int(os.environ.get('SPP', 500 if gname == 'G24H' else 2000))
    gnm = samples.GENOMES[gname]()
    for fw in [float(x) for x in os.environ.get('FW', '0,1e-9').split(',')]:
        gprof = profile.wrap(dict(width=W, height=H, spp=spp, frame_width=fw, start=1, end=2), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        for sched, spill in [(a, b == '1') for a, b in CASES]:
            rmgr = render.RenderManager(seed=1)
            rmgr.hot_bins, rmgr.spill, rmgr.schedule = False, spill, sched
            rdr = render.Renderer(gnm, gprof)
            dim = rmgr.fb.set_dim(W, H)
            rmgr._copy(rdr, gnm)
            rmgr._interp(rdr, gnm, dim, tc, 0.0)
            ms = []
            for i in range(6):
                N.fill32(flush, (512 << 20) // 4, 0, rmgr.stream_a)
                e0, e1 = N.Event(), N.Event()
                e0.record(rmgr.stream_a)
                rmgr._iter(rdr, gnm, gprof, dim, tc)
                e1.record(rmgr.stream_a)
                e1.synchronize()
                ms.append(e1.time_since(e0))
            n = rmgr.last_iter_samples
            best = min(ms[1:])
            print(json.dumps(dict(genome=gname, width=W, height=H, spp=spp, motion_blur=fw > 0,
                                  schedule=sched, spill=spill, ms=round(best, 3),
                                  samples_per_s=n / best * 1e3, all_ms=[round(m, 2) for m in ms])), flush=True)
            rmgr.fb.free()
