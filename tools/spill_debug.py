import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from helpers import exact_level_sums
N.init(0)
gname, w, h, spp = 'G6F', int(os.environ.get('W', 1920)), int(os.environ.get('H', 1080)), int(os.environ.get('SPP', 1000))
gnm = samples.GENOMES[gname]()
gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
out = {}
for mode in ('exact', 'swept'):
    rmgr = render.RenderManager(seed=17)
    rmgr.accumulate, rmgr.hot_bins = 'float4', False
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    nbins = dim.ah * dim.astride
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    if mode == 'exact':
        out[mode] = exact_level_sums(N, rmgr, rdr, gnm, gprof, dim, tc)
    else:
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        out[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32).astype(np.float64)
        fine = N.from_device(rmgr.fb.d_left, (nbins, 4), np.float32).astype(np.float64)
        moved = N.from_device(rmgr.fb.d_right, (nbins, 4), np.float32).astype(np.float64)
    rmgr.fb.free()
swz = (nbins // 65536) * 65536
i = np.arange(nbins)
j = np.where(i < swz, (i & ~0xffff) | ((i * 40503) & 0xffff), i)
fine, moved = fine[j], moved[j]
ex = out['exact'].reshape(-1, 4).astype(np.float64)
got = out['swept'].reshape(-1, 4)
raw = fine + moved
print('count equal', np.array_equal(ex[:, 3], got[:, 3]), 'raw count equal', np.array_equal(raw[:, 3], ex[:, 3]))
err = np.abs(raw[:, :3] - ex[:, :3]).max(axis=1)
print('bins with raw != exact:', int((err > 0).sum()), 'of', int((ex[:, 3] > 0).sum()))
bad = np.argsort(err)[::-1][:12]
for b in bad:
    print(int(b), 'y,x', b // dim.astride, b % dim.astride, 'exact', ex[b].tolist(), 'fine', fine[b].tolist(), 'moved', moved[b].tolist(), 'diff', (raw[b] - ex[b]).tolist())
print('max fine', fine.max(axis=0).tolist(), 'max moved', moved.max(axis=0).tolist(), 'bins moved', int((moved[:, 3] > 0).sum()))
e2 = (err > 0)
print('erroneous bins: count quantiles', np.percentile(ex[e2, 3], [0, 10, 50, 90, 100]).tolist() if e2.any() else None)
print('moved sums > 2^24:', int((moved[:, :3].max(axis=1) >= 2 ** 24).sum()), ' erroneous with moved < 2^24:', int((e2 & (moved[:, :3].max(axis=1) < 2 ** 24)).sum()))
