import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from cuburn_b200.code import itergen
N.init(0)
w, h, spp = 1920, 1080, 1000
gnm = samples.GENOMES['G6F']()
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
orig_gen = itergen.generate_source
def gen(pk, params_const=False, extra_defines=None, **kw):
    d = dict(extra_defines or {}); d['SPILL_DEBUG'] = '1'
    return orig_gen(pk, params_const, extra_defines=d, **kw)
itergen.generate_source = gen
for watch in (541960, 101838, 372757):
    dbg = N.DeviceBuffer(4096 * 4)
    N.fill32(dbg, 4096, 0)
    init = np.zeros(4, np.int32); init[0] = watch
    N.memcpy_htod(dbg, init)
    orig_args = N.IterArgs
    def IterArgs(**kw):
        kw['hot_tags'] = dbg.ptr
        return orig_args(**kw)
    render.N.IterArgs = IterArgs
    rmgr = render.RenderManager(seed=17)
    rmgr.accumulate, rmgr.hot_bins = 'float4', False
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    rmgr.stream_a.synchronize()
    render.N.IterArgs = orig_args
    d = N.from_device(dbg, (4096,), np.int32)
    n = d[1]
    rec = d[4:4 + 3 * min(n, 300)].reshape(-1, 3)
    print('watch', watch, 'visits', n)
    print(' '.join('%d:%d@%d' % (r[0], int(np.int32(r[1]).view(np.float32)), r[2]) for r in rec[np.argsort(rec[:, 0])]))
    rmgr.fb.free()
