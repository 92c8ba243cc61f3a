"""Shared helpers for the parity tests."""
import ctypes

import numpy as np


def still_profile(gnm, w, h, spp, **extra):
    from cuburn_b200 import profile
    prof = dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2)
    prof.update(extra)
    gprof = profile.wrap(prof, gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    return gprof, tc


def frame_window(gprof, tc):
    td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
    return tc - 0.5 * td, td


def pool8(a):
    hh, ww = a.shape[0] // 8 * 8, a.shape[1] // 8 * 8
    return a[:hh, :ww].astype(np.float64).reshape(hh // 8, 8, ww // 8, 8).sum(axis=(1, 3))


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def upload_field(N, arr):
    return N.to_device(np.ascontiguousarray(arr, np.float32))


def single_xform_genome(var_name, params=None, weight=1.0, extra_vars=None):
    """One xform with identity-ish affine carrying one variation (plus extras)."""
    vs = {var_name: dict(weight=weight, **(params or {}))}
    for k, v in (extra_vars or {}).items():
        vs[k] = v
    return {
        'type': 'animation', 'time': {'duration': 1},
        'camera': {'scale': 0.25},
        'palette': [[0.0, 'rgb8'] + ['AAAA' * 16] * 16],
        'xforms': {'0': {
            'weight': 1, 'color': 0.7, 'color_speed': 0.25,
            'pre_affine': {'angle': 55, 'spread': 40,
                           'magnitude': {'x': 0.9, 'y': 1.1},
                           'offset': {'x': 0.31, 'y': -0.17}},
            'variations': vs}},
    }


def c_float_p(arr):
    return arr.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def exact_level_sums(N, rmgr, rdr, gnm, gprof, dim, tc, waves_per_chunk=1):
    """
    The integer sums (sum Y, sum U, sum V of 8-bit palette levels, count) per bin, in
    linear layout, of exactly the sample set ``rmgr._iter`` draws for this frame from the
    current seeds -- computed without trusting any long float32 running sum: the frame is
    launched in chunks of whole waves of the persistent grid (a frame split over several
    calls draws the samples one call would, cb_iter_args.first_sample / first_round), the
    raw grids are read back after every chunk, checked to be below 2^24 (every add was
    exact) and added up in int64 on the host.  int64 array [ah][astride][4].
    """
    from cuburn_b200 import render
    from cuburn_b200.code import itergen
    nbins = dim.ah * dim.astride
    acc = np.zeros((nbins, 4), np.int64)
    orig = rmgr._launch_iter

    def chunked(mod, rdr_, info, d_acc, swz, dim_, first, n, total, fuse, packed, hot, s,
                first_round=0):
        assert not packed and not hot
        grid = rmgr.iter_grid or rdr_.grid_ctas(rmgr.fb.nstreams, mod)
        ppt = itergen.points_per_thread(rdr_.packer, rdr_._points(rdr_.packer, True))
        step = grid * waves_per_chunk * render.UNIT_SAMPLES
        while n > 0:
            m = min(step, n)
            orig(mod, rdr_, info, d_acc, swz, dim_, first, m, total, fuse, packed, hot, s,
                 first_round)
            s.synchronize()
            for buf in (rmgr.fb.d_left, rmgr.fb.d_right):
                raw = N.from_device(buf, (nbins, 4), np.float32)
                assert raw.max() < 2.0 ** 24 and np.array_equal(raw, np.rint(raw))
                acc[:] += raw.astype(np.int64)
                N.fill32(buf, 4 * nbins, 0, s)
            first_round += fuse + waves_per_chunk * (render.UNIT_SAMPLES //
                                                     (render.ITER_THREADS * ppt))
            first, n, fuse = first + m, n - m, 0
    rmgr._launch_iter = chunked
    try:
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
    finally:
        rmgr._launch_iter = orig
    swz = (nbins // 65536) * 65536 if rmgr._use_swizzle(nbins) else 0
    i = np.arange(nbins)
    j = np.where(i < swz, (i & ~0xffff) | ((i * 40503) & 0xffff), i)
    return acc[j].reshape(dim.ah, dim.astride, 4)
