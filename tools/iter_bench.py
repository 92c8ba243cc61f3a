"""Time cb_iter variants on the GPU: python tools/iter_bench.py 'ITER_MIN_CTAS=5' 'ITER_MIN_CTAS=6,FOO=1' ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from cuburn_b200.code import itergen

N.init(0)
cases = [('G6F', 1920, 1080, 2000), ('G3', 1920, 1080, 2000), ('G24H', 1920, 1080, 500)]
if os.environ.get('CASES'):
    cases = [c for c in cases if c[0] in os.environ['CASES'].split(',')]
def _medium(nxf, keep_final):
    """G6F cut down to its first `nxf` xforms (for tuning the heavy / light boundary)."""
    def make():
        g = samples.g6f()
        g['xforms'] = dict((k, v) for k, v in g['xforms'].items() if int(k) < nxf)
        if not keep_final:
            g.pop('final_xform', None)
        return g
    return make
samples.GENOMES.update(G4M=_medium(4, False), G3F=_medium(3, True), G2M=_medium(2, False))
cases += [('G4M', 1920, 1080, 2000), ('G3F', 1920, 1080, 2000), ('G2M', 1920, 1080, 2000)]
if os.environ.get('CASES'):
    cases = [c for c in cases if c[0] in os.environ['CASES'].split(',')]
variants = sys.argv[1:] or ['']
rmgr = render.RenderManager(seed=1); rmgr.hot_bins = False; rmgr.swizzle = {'1': True, '0': False}.get(os.environ.get('SWZ', 'auto'), 'auto')
for gname, w, h, spp in cases:
    gnm = samples.GENOMES[gname]()
    gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    dim = rmgr.fb.set_dim(w, h)
    for var in variants:
        defs = dict(kv.split('=') for kv in var.split(',') if kv)
        still = defs.pop('STILL', '1') == '1'
        pk = itergen.GenomePacker(gnm)
        src = itergen.generate_source(pk, params_const=still, extra_defines=defs)
        names, hdrs = itergen.load_headers()
        mod = N.Module(src, 'iter.cu', hdrs, names, itergen.NVRTC_OPTIONS)
        rdr = render.Renderer(gnm, gprof)
        rdr._variants[(still, False, False)] = mod
        if not still:
            gprof2 = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=float(os.environ.get('FW', 1e-9)), start=1, end=2), gnm)
        else:
            gprof2 = gprof
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, tc, 0.0)
        ms = []
        flush = os.environ.get('FLUSH')
        if flush:
            fbuf = N.DeviceBuffer(512 << 20)
        for i in range(int(os.environ.get('REPS', 4))):
            if flush:
                N.fill32(fbuf, (512 << 20) // 4, 0, rmgr.stream_a)
                if flush == '2':
                    N.fill32(rmgr.fb.d_front, 4 * dim.ah * dim.astride, 0, rmgr.stream_a)
            e0, e1 = N.Event(), N.Event()
            e0.record(rmgr.stream_a)
            rmgr._iter(rdr, gnm, gprof2, dim, tc)
            e1.record(rmgr.stream_a)
            e1.synchronize()
            ms.append(e1.time_since(e0))
            if os.environ.get('FILTERS'):
                for filt in rdr.filts:
                    filt.apply(rmgr.fb, gprof, getattr(gprof.filters, filt.name), dim, tc, rmgr.stream_a)
                rmgr.stream_a.synchronize()
        n = rmgr.last_iter_samples
        info = mod.kernel_info('cb_iter', 256)
        print('%-5s %-40s regs %3d ctas/sm %d  %.2f ms  %.4g it/s  all: %s' % (
            gname, var or '(default)', info['num_regs'], info['ctas_per_sm'], min(ms[1:]),
            n / min(ms[1:]) * 1e3, ' '.join('%.2f' % m for m in ms)), flush=True)
