"""How exact are the colour sums of the hottest bins?  Renders one sample set (same seeds)
  * in chunks small enough that no float add can round, summed in int64 (the truth,
    tests/helpers.py::exact_level_sums),
  * with the default float4 accumulation (integer levels + spill sweep),
  * with the sweep off (plain float32 running sums: what round 1 shipped),
  * with the packed u64 cells (the reference's format, drained into floats),
and reports the relative error of each by heat of the bin.
    python tools/accuracy_probe.py [GENOME W H SPP]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from helpers import exact_level_sums

N.init(0)
gname, w, h, spp = (sys.argv[1:] + ['G6F', '1920', '1080', '2000'])[:4] if len(sys.argv) > 1 else ('G6F', 1920, 1080, 2000)
w, h, spp = int(w), int(h), int(spp)
gnm = samples.GENOMES[gname]()
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
out = {}
for mode in ('exact', 'float4', 'float4 unswept', 'packed'):
    rmgr = render.RenderManager(seed=17)
    rmgr.accumulate, rmgr.hot_bins = ('packed' if mode == 'packed' else 'float4'), False
    rmgr.spill = mode != 'float4 unswept'
    rmgr.schedule = 'static'           # the same sample set in every mode
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    if mode == 'exact':
        e = exact_level_sums(N, rmgr, rdr, gnm, gprof, dim, tc).astype(np.float64)
        e[..., :3] /= 255.0
        out[mode] = e
    else:
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        out[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32).astype(np.float64)
    rmgr.fb.free()
b = out['exact']
cnt = b[..., 3].ravel()
order = np.argsort(cnt)[::-1]
res = dict(genome=gname, width=w, height=h, spp=spp, hottest_bin_share=float(cnt.max() / (w * h * spp)), modes={})
for mode in ('float4', 'float4 unswept', 'packed'):
    a = out[mode]
    assert np.array_equal(a[..., 3], b[..., 3])
    rel = np.abs(a[..., :3] - b[..., :3]).reshape(-1, 3) / np.maximum(b[..., :3].reshape(-1, 3), 1e-3)
    rows = []
    for lo, hi in ((0, 1), (1, 10), (10, 100), (100, 1000), (1000, 10000), (10000, 100000)):
        sel = order[lo:hi]
        rows.append(dict(rank='%d-%d' % (lo + 1, hi), samples_per_bin=[float(cnt[sel].max()), float(cnt[sel].min())],
                         mean_rel_err=float(rel[sel].mean()), max_rel_err=float(rel[sel].max())))
    res['modes'][mode] = rows
print(json.dumps(res, indent=1))
