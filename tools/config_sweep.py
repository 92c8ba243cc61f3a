"""All BASELINE.json configs that fit one GPU, through queue_frame (full frame, host frame out)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
N.init(0)
rows = []
for label, gname, w, h, spp in (
        ('config 1: 640x360 G3 256 spp', 'G3', 640, 360, 256),
        ('config 2: 1080p G6F 2000 spp', 'G6F', 1920, 1080, 2000),
        ('config 3 on one GPU: 4K G6F 4000 spp', 'G6F', 3840, 2160, 4000),
        ('config 5: 8K G24H 2000 spp', 'G24H', 7680, 4320, 2000),
        ('8K G6F 2000 spp', 'G6F', 7680, 4320, 2000)):
    gnm = samples.GENOMES[gname]()
    gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=1)
    rdr = render.Renderer(gnm, gprof)
    ms, it = [], []
    for rep in range(3):
        e0 = N.Event().record(rmgr.stream_a)
        dim = rmgr.fb.set_dim(w, h)
        evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc)
        evt.synchronize()
        ms.append(evt.time())
    # iterate alone
    dim = rmgr.fb.set_dim(w, h)
    for rep in range(2):
        a, b = N.Event(), N.Event()
        a.record(rmgr.stream_a)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        b.record(rmgr.stream_a); b.synchronize()
        it.append(b.time_since(a))
    n = w * h * spp
    row = dict(config=label, samples=n, frame_ms=min(ms[1:]), iter_ms=min(it),
               frame_its=n / min(ms[1:]) * 1e3, iter_its=n / min(it) * 1e3,
               hist_mib=16 * dim.ah * dim.astride / 2 ** 20)
    rows.append(row)
    print('%-40s hist %6.1f MiB  frame %8.2f ms  %.3g it/s   iterate %8.2f ms  %.3g it/s' % (
        label, row['hist_mib'], row['frame_ms'], row['frame_its'], row['iter_ms'], row['iter_its']), flush=True)
    rmgr.fb.free()
    del rmgr
json.dump(rows, open('gpurun_out/config_sweep.json', 'w'), indent=1)
