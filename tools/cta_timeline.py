"""Start / end time and SM of every persistent CTA of one cb_iter launch (debug build of the
module: -DCTA_TIMELINE writes %globaltimer / %smid per CTA).  python tools/cta_timeline.py [still|blur]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from cuburn_b200.code import itergen
N.init(0)
w, h, spp = 1920, 1080, int(os.environ.get('SPP', 1000))
gnm = samples.GENOMES[os.environ.get('GENOME', 'G6F')]()
orig_gen = itergen.generate_source
def gen(pk, params_const=False, extra_defines=None, **kw):
    d = dict(extra_defines or {}); d['CTA_TIMELINE'] = '1'
    return orig_gen(pk, params_const, extra_defines=d, **kw)
itergen.generate_source = gen
for fw in (0, 1e-9):
    gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=fw, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    dbg = N.DeviceBuffer(4096 * 4 + 1024 * 24 + 64)
    N.fill32(dbg, (4096 * 4 + 1024 * 24 + 64) // 4, 0)
    init = np.zeros(4, np.int32); init[0] = -1
    N.memcpy_htod(dbg, init)
    orig_args = N.IterArgs
    def IterArgs(**kw):
        kw['hot_tags'] = dbg.ptr
        return orig_args(**kw)
    render.N.IterArgs = IterArgs
    rmgr = render.RenderManager(seed=17)
    rmgr.accumulate, rmgr.hot_bins = 'float4', False
    rmgr.schedule = os.environ.get('SCHED', 'dynamic')
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    for rep in range(2):
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
    render.N.IterArgs = orig_args
    mod = rdr.variant(fw == 0)
    info = mod.kernel_info('cb_iter', 256)
    grid = rdr.grid_ctas(rmgr.fb.nstreams, mod)
    raw = N.from_device(dbg, (4096 * 4 + 1024 * 24 + 64,), np.uint8)
    tt = raw[4096:4096 + 1024 * 24].view(np.uint64).reshape(1024, 3)[:grid]
    t0 = tt[:, 0].min()
    st, en, sm = (tt[:, 0] - t0) / 1e6, (tt[:, 1] - t0) / 1e6, tt[:, 2].astype(int)
    print('schedule', rmgr.schedule, 'variant', 'still' if fw == 0 else 'blur', 'regs', info['num_regs'], 'ctas/sm', info['ctas_per_sm'], 'grid', grid)
    print(' kernel span %.2f ms; CTA start quantiles (ms) %s' % (en.max(), np.round(np.percentile(st, [0, 25, 50, 75, 90, 100]), 2).tolist()))
    print(' CTA end quantiles (ms) %s; run time quantiles %s' % (np.round(np.percentile(en, [0, 25, 50, 75, 90, 100]), 2).tolist(), np.round(np.percentile(en - st, [0, 25, 50, 75, 100]), 2).tolist()))
    late = st > 0.5
    print(' CTAs started later than 0.5 ms:', int(late.sum()), ' first such index', int(np.argmax(late)) if late.any() else None)
    per_sm = np.bincount(sm[~late], minlength=148)
    print(' initially resident per SM: min %d max %d, SMs used %d' % (per_sm[per_sm > 0].min(), per_sm.max(), (per_sm > 0).sum()))
    rmgr.fb.free()
