"""How far do the float32 running sums of the default (float4) accumulation drift in the
hottest bins?  Renders the same samples (same seeds, one launch each) with the float4 path
and with the packed path (integer level sums, exact) and compares the colour sums bin by bin.
    python tools/accuracy_probe.py [GENOME W H SPP]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render

N.init(0)
gname, w, h, spp = (sys.argv[1:] + ['G6F', '1920', '1080', '2000'])[:4] if len(sys.argv) > 1 else ('G6F', 1920, 1080, 2000)
w, h, spp = int(w), int(h), int(spp)
gnm = samples.GENOMES[gname]()
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
out = {}
for mode in ('float4', 'packed'):
    rmgr = render.RenderManager(seed=17)
    rmgr.accumulate, rmgr.hot_bins = mode, False
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    rmgr.stream_a.synchronize()
    out[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32).astype(np.float64)
    rmgr.fb.free()
a, b = out['float4'], out['packed']
assert np.array_equal(a[..., 3], b[..., 3])
cnt = b[..., 3].ravel()
order = np.argsort(cnt)[::-1]
rel = np.abs(a[..., :3] - b[..., :3]).reshape(-1, 3) / np.maximum(b[..., :3].reshape(-1, 3), 1e-3)
rows = []
for lo, hi in ((0, 1), (1, 10), (10, 100), (100, 1000), (1000, 10000), (10000, 100000)):
    sel = order[lo:hi]
    rows.append(dict(rank='%d-%d' % (lo + 1, hi), samples_per_bin=[float(cnt[sel].max()), float(cnt[sel].min())],
                     mean_rel_err=float(rel[sel].mean()), max_rel_err=float(rel[sel].max()),
                     bound_n_2pow_minus25=float(cnt[sel].max() * 2.0 ** -25)))
print(json.dumps(dict(genome=gname, width=w, height=h, spp=spp, bins_by_heat=rows), indent=1))
