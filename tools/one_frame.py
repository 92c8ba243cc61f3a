"""Render frames of one workload (for ncu): python tools/one_frame.py GENOME W H SPP [key=value ...]
keys: accumulate=auto|float4|packed  hot=auto|0|1  frames=N  blur=0|1  filters=0|1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render

gname, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
opts = dict(kv.split('=') for kv in sys.argv[5:])
N.init(0)
gnm = samples.g6f(animated=True) if (gname == 'G6F' and opts.get('blur') == '1') else samples.GENOMES[gname]()
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=1.0 if opts.get('blur') == '1' else 0,
                          fps=24, duration=30.0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=1)
rmgr.accumulate = opts.get('accumulate', 'auto')
rmgr.hot_bins = {'auto': 'auto', '0': False, '1': True}[opts.get('hot', 'auto')]
rdr = render.Renderer(gnm, gprof)
for i in range(int(opts.get('frames', 2))):
    if opts.get('filters', '1') == '1':
        evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc)
        evt.synchronize()
    else:
        dim = rmgr.fb.set_dim(w, h)
        td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, tc - 0.5 * td, td)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
print('done', gname, w, h, spp, opts, 'samples', rmgr.last_iter_samples, 'hot', rmgr.last_iter_hot)
