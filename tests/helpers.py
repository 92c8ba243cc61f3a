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
