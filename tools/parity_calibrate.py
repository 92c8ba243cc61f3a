"""
Calibration of the density-parity test of the chaos game at the benchmarked sizes.

Both the oracle (oracle/chaos.c) and the device kernel are Monte-Carlo estimators of
the same measure with *different* sample sets, so bin counts can only be compared
statistically.  Counts of pooled bins are over-dispersed relative to Poisson on both
sides (consecutive iterations of one trajectory are correlated), by an amount that
depends on the genome -- so a fixed bound on the z statistics is either loose or
wrong.  This tool measures, per workload, the three numbers the tests then use:

  z_oo   oracle(seed A)  vs oracle(seed B)      -- the oracle's own dispersion
  z_gg   device(seed A)  vs device(seed B)      -- the device's own dispersion
  z_go   device(seed A)  vs oracle(seed A)

With dispersion indices D_o, D_g:  var z_oo = D_o, var z_gg = D_g and, if both sides
sample the same measure, var z_go = (D_o + D_g) / 2.  Any systematic difference between
the two measures inflates z_go beyond that (and shows in chi2 growing with spp).
Output: one JSON line per workload (run with --cpu-only for z_oo alone).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np

from parity_stats import density_z, colour_means, in_frame_fraction

WORKLOADS = [
    ('smoke', 'G3', 320, 180, 200, 'auto'),
    ('smoke-x4', 'G3', 320, 180, 800, 'auto'),
    ('smoke-x16', 'G3', 320, 180, 3200, 'auto'),
    ('g6f-small', 'G6F', 640, 360, 256, 'auto'),
    ('g6f-small-x8', 'G6F', 640, 360, 2048, 'auto'),
    ('g24h-small', 'G24H', 320, 180, 200, 'auto'),
    # name, genome, w, h, spp, accumulate
    ('config1', 'G3', 640, 360, 256, 'auto'),
    ('config2', 'G6F', 1920, 1080, 200, 'auto'),
    ('config3', 'G6F', 3840, 2160, 50, 'auto'),
    ('config5', 'G24H', 7680, 4320, 25, 'auto'),
    ('config5-float4', 'G24H', 7680, 4320, 25, 'float4'),
    ('config2-packed', 'G6F', 1920, 1080, 200, 'packed'),
]


def oracle_hist(gnm, w, h, spp, seed, tc):
    from cuburn_b200 import mwc
    from oracle import flame_ref as R
    ev = R.GenomeEval(gnm, w, h, tc, 0.0)
    seeds = mwc.make_seeds(32768, host_seed=seed)
    pal, seeds = R.palette_table(gnm, ev.ts, ev.td, seeds)
    hist, _ = R.iterate(ev, pal, seeds, w * h * spp)
    return hist


def device_hist(N, gnm, w, h, spp, seed, accumulate):
    from cuburn_b200 import render, profile
    prof = dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2)
    gprof = profile.wrap(prof, gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=seed)
    rmgr.accumulate = accumulate
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    rmgr.stream_a.synchronize()
    hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
    rmgr.fb.free()
    return hist


def main():
    cpu_only = '--cpu-only' in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith('--')]
    from cuburn_b200 import samples, profile
    from oracle.build import build
    build()
    N = None
    if not cpu_only:
        from cuburn_b200 import _native as N
        N.init(0)
    for name, gname, w, h, spp, acc in WORKLOADS:
        if only and name not in only:
            continue
        gnm = samples.GENOMES[gname]()
        gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        n = w * h * spp
        row = dict(workload=name, genome=gname, width=w, height=h, spp=spp, accumulate=acc,
                   samples=n)
        t = time.perf_counter()
        oa = oracle_hist(gnm, w, h, spp, 101, tc)
        row['oracle_s'] = time.perf_counter() - t
        ob = oracle_hist(gnm, w, h, spp, 202, tc)
        row['z_oo'] = density_z(oa, ob)
        row['in_frame_oracle'] = in_frame_fraction(oa, n)
        del ob
        if N is not None:
            ga = device_hist(N, gnm, w, h, spp, 101, acc)
            gb = device_hist(N, gnm, w, h, spp, 202, acc)
            row['z_gg'] = density_z(ga, gb)
            row['z_go'] = density_z(ga, oa)
            row['in_frame_device'] = in_frame_fraction(ga, n)
            row['colour_go'] = colour_means(ga, oa)
            row['colour_gg'] = colour_means(ga, gb)
            del ga, gb
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
