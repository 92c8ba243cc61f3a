"""
Host side of the filter chain: a registry of named filters, each a short launch
recipe over the framebuffers.

API-compatible with the reference (cuburn/filters.py): ``Filter.register``,
``Filter.apply(fb, gprof, params, dim, tc, stream)`` with the rule that the
result is in ``fb.d_front`` when ``apply`` returns, and ``create(gprof)`` =
``['yuv'] + filter_order``.  Kernels are the C-ABI entry points of
csrc/cb_filters.cu; parameters are evaluated on the host at ``tc`` exactly as
in the reference recipes.
"""
import ctypes

import numpy as np
from numpy import float32 as f32

from . import _native as N


_gauss_cache = {}


def gauss_coefs(stdev=1):
    """7-tap normalised Gaussian (set_blur_width, filters.py:11-16).  The returned ctypes
    array is shared between calls with the same width (read-only by convention)."""
    key = float(stdev)
    out = _gauss_cache.get(key)
    if out is None:
        coefs = np.exp(np.float32(np.arange(-3, 4)) ** 2 / (-2 * stdev ** 2)).astype(np.float32)
        coefs /= np.sum(coefs)
        if len(_gauss_cache) > 256:         # an animated width: do not grow without bound
            _gauss_cache.clear()
        out = _gauss_cache[key] = (ctypes.c_float * 7)(*coefs.astype(np.float32))
    return out


# The first eight of the device's 16 filter directions (code/filters.py:8-17), for
# working out how far a stencil reaches.
DIRECTIONS = [(1.0, 0.0), (0.0, 1.0), (1.0, 1.0), (-1.0, 1.0),
              (1.0, 0.5), (-0.5, 1.0), (1.0, -0.5), (0.5, 1.0)]


def reach_rows(pattern, *steps):
    """Rows a chain of stencils along direction ``pattern`` can reach, each stage
    ``steps`` taps out (tap offsets are rounded per stage, tex_shear)."""
    dy = abs(DIRECTIONS[pattern][1])
    return int(sum(np.ceil(dy * n) for n in steps))


def _h(stream):
    return stream.handle if stream is not None else None


class Filter(object):
    filter_map = {}
    name = ''
    # True if the filter needs a full 4-channel side buffer
    full_side = False

    def apply(self, fb, gprof, params, dim, tc, stream=None):
        raise NotImplementedError()

    def reach(self, gprof, params, tc):
        """Rows above / below a bin that its result depends on (0: pointwise).  An
        addition: lets a multi-GPU still filter row bands with exact halos."""
        return 0

    @classmethod
    def register(cls, name):
        def register_(subcls):
            cls.filter_map[name] = subcls
            subcls.name = name
            return subcls
        return register_


@Filter.register('yuv')
class YuvFilterLib(Filter):
    def apply(self, fb, gprof, params, dim, tc, stream=None):
        N.check(N.lib().cb_yuv_to_rgb(fb.d_back.ptr, fb.d_front.ptr, N.byref(dim), _h(stream)))
        fb.flip()


@Filter.register('bilateral')
class Bilateral(Filter):
    radius = 15
    directions = 8

    def apply(self, fb, gprof, params, dim, tc, stream=None):
        L, s = N.lib(), _h(stream)
        coefs = gauss_coefs(1)
        # a "pixel" of spatial_std means a 1080p pixel (filters.py:74-76)
        sstd = f32(params.spatial_std(tc) * dim.w / 1920.)
        cstd, dstd = f32(params.color_std(tc)), f32(params.density_std(tc))
        dpow, grad = f32(params.density_pow(tc)), f32(params.gradient(tc))
        for pattern in range(self.directions):
            # den_blur -> den_blur_1c -> bilateral of the reference recipe
            # (filters.py:80-94), fused into one restructured direction pass
            N.check(L.cb_bilateral_direction(
                fb.d_back.ptr, fb.d_front.ptr, fb.d_left.ptr, pattern, self.radius, coefs,
                sstd, cstd, dstd, dpow, grad, N.byref(dim), s))
            fb.flip()

    def reach(self, gprof, params, tc):
        # per direction: taps to radius + 1 (the gradient reads one further), whose
        # two-octave blur reaches 6 steps, whose density blur reaches 3
        return sum(reach_rows(p, self.radius + 1, 6, 3) for p in range(self.directions))


@Filter.register('logscale')
class Logscale(Filter):
    def apply(self, fb, gprof, params, dim, tc, stream=None):
        k1 = f32(params.brightness(tc) * 268 / 256)
        # area of the frame in IFS units: h / (scale^2 * w)  (filters.py:103-105)
        area = dim.h / (params.scale(tc) ** 2 * dim.w)
        k2 = f32(1.0 / (area * gprof.spp(tc)))
        N.check(N.lib().cb_logscale(fb.d_front.ptr, fb.d_front.ptr, k1, k2,
                                    N.byref(dim), _h(stream)))


@Filter.register('haloclip')
class HaloClip(Filter):
    def apply(self, fb, gprof, params, dim, tc, stream=None):
        L, s = N.lib(), _h(stream)
        gam = f32(1 / gprof.filters.colorclip.gamma(tc) - 1)
        coefs = gauss_coefs(1)
        N.check(L.cb_apply_gamma(fb.d_left.ptr, fb.d_front.ptr, f32(0.1), N.byref(dim), s))
        N.check(L.cb_den_blur_1c(fb.d_back.ptr, fb.d_left.ptr, 2, 0, coefs, N.byref(dim), s))
        N.check(L.cb_den_blur_1c(fb.d_left.ptr, fb.d_back.ptr, 3, 0, coefs, N.byref(dim), s))
        N.check(L.cb_haloclip(fb.d_front.ptr, fb.d_left.ptr, gam, N.byref(dim), s))

    def reach(self, gprof, params, tc):
        return reach_rows(2, 3) + reach_rows(3, 3)


def calc_lingam(params, tc):
    gam = f32(1 / params.gamma(tc))
    lin = f32(params.gamma_threshold(tc))
    lingam = f32(lin ** (gam - 1.0) if lin > 0 else 0)
    return gam, lin, lingam


@Filter.register('smearclip')
class SmearClip(Filter):
    full_side = True

    def apply(self, fb, gprof, params, dim, tc, stream=None):
        L, s = N.lib(), _h(stream)
        gam, lin, lingam = calc_lingam(gprof.filters.colorclip, tc)
        coefs = gauss_coefs(params.width(tc))
        N.check(L.cb_apply_gamma_full_hi(fb.d_left.ptr, fb.d_front.ptr, f32(gam - 1),
                                         N.byref(dim), s))
        N.check(L.cb_full_blur(fb.d_back.ptr, fb.d_left.ptr, 2, 0, coefs, N.byref(dim), s))
        N.check(L.cb_full_blur(fb.d_left.ptr, fb.d_back.ptr, 3, 0, coefs, N.byref(dim), s))
        N.check(L.cb_full_blur(fb.d_back.ptr, fb.d_left.ptr, 0, 0, coefs, N.byref(dim), s))
        N.check(L.cb_full_blur(fb.d_left.ptr, fb.d_back.ptr, 1, 0, coefs, N.byref(dim), s))
        N.check(L.cb_smearclip(fb.d_front.ptr, fb.d_left.ptr, f32(gam - 1), lin, lingam,
                               N.byref(dim), s))

    def reach(self, gprof, params, tc):
        return sum(reach_rows(p, 3) for p in (2, 3, 0, 1))


@Filter.register('colorclip')
class ColorClip(Filter):
    def apply(self, fb, gprof, params, dim, tc, stream=None):
        vib = f32(params.vibrance(tc))
        hipow = f32(params.highlight_power(tc))
        gam, lin, lingam = calc_lingam(params, tc)
        N.check(N.lib().cb_colorclip(fb.d_front.ptr, vib, hipow, gam, lin, lingam,
                                     N.byref(dim), _h(stream)))


@Filter.register('plainclip')
class PlainClip(Filter):
    def apply(self, fb, gprof, params, dim, tc, stream=None):
        gam, lin, lingam = calc_lingam(gprof.filters.colorclip, tc)
        N.check(N.lib().cb_plainclip(
            fb.d_front.ptr, f32(gam - 1), lin, lingam,
            f32(gprof.filters.plainclip.brightness(tc)), N.byref(dim), _h(stream)))


@Filter.register('logencode')
class LogEncode(Filter):
    def apply(self, fb, gprof, params, dim, tc, stream=None):
        N.check(N.lib().cb_logencode(fb.d_back.ptr, fb.d_front.ptr,
                                     f32(params.degamma(tc)), N.byref(dim), _h(stream)))
        fb.flip()


def create(gprof):
    order = ['yuv'] + list(gprof.filter_order)
    return [Filter.filter_map[f]() for f in order]
