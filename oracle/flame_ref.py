"""
ORACLE (test infrastructure only) -- genome interpolation, precalc, palette and
the chaos game, restated on the CPU.

Follows, in order: spline knots  cuburn/genome/use.py:129-158; packing defaults
cuburn/code/interp.py:207-232; knot search  cuburn/code/util.py:220-229;
Catmull-Rom (plain and magnitude domain)  cuburn/code/interp.py:295-366; precalc
cuburn/code/iter.py:12-30,56-95 and the five variation precalcs in
cuburn/code/variations.py; palette  cuburn/code/interp.py:372-433 with
cuburn/genome/util.py:75-87; iteration  oracle/chaos.c.

float32 arithmetic is done with numpy float32 scalars/arrays (every operation
rounds once, no fused multiply-add), transcendental steps with the same fixed
float64 operation sequences as csrc/device/det_math.cuh, so the packed
parameters can be compared bit for bit with the device.

The genome *schema* (defaults, which parameters interpolate in the magnitude
domain) is read from cuburn_b200.genome.specs, which tests pin against the
reference schema; no other product code is used.
"""
import base64
import ctypes
import math
import os

import numpy as np

from cuburn_b200.genome import specs as _specs
from cuburn_b200.genome.spectypes import Map as _Map
from cuburn_b200.genome.variations import var_names as _var_names, VAR_TABLE as _VAR_TABLE

f32 = np.float32
NTS = 1024
PAL_ROWS = 64
GUTTER = 12

_VAR_NUM = {name: num for num, name in _var_names.items()}
_VAR_PARAMS = {name: [p for p, _ in params] for _, name, params in _VAR_TABLE}


# ---- dims / sample counts (render.py:80-89, 330-332) ----------------------------
def calc_dim(w, h):
    aw = w + 2 * GUTTER
    ah = 16 * int(math.ceil((h + 2 * GUTTER) / 16.0))
    astride = 32 * int(math.ceil(aw / 32.0))
    return dict(w=w, h=h, aw=aw, ah=ah, astride=astride)


# ---- MWC, vectorised over streams (code/mwc.py:56-77) ---------------------------
class MwcStreams(object):
    def __init__(self, seeds):
        seeds = np.asarray(seeds, dtype=np.uint32)
        self.mul = seeds[:, 0].astype(np.uint64)
        self.state = seeds[:, 1].astype(np.uint64)
        self.carry = seeds[:, 2].astype(np.uint64)

    def next_u32(self, mask=None):
        t = self.mul * self.state + self.carry
        ns, nc = t & np.uint64(0xffffffff), t >> np.uint64(32)
        if mask is None:
            self.state, self.carry = ns, nc
            return ns.astype(np.uint32)
        self.state = np.where(mask, ns, self.state)
        self.carry = np.where(mask, nc, self.carry)
        return ns.astype(np.uint32)

    def next_01(self, mask=None):
        return self.next_u32(mask).astype(f32) * f32(1.0 / 4294967296.0)

    def next_11(self, mask=None):
        return self.next_u32(mask).view(np.int32).astype(f32) * f32(1.0 / 2147483648.0)

    def seeds(self):
        out = np.empty((self.mul.size, 3), np.uint32)
        out[:, 0], out[:, 1], out[:, 2] = self.mul, self.state, self.carry
        return out


# ---- deterministic float64 elementary functions (twin of det_math.cuh) ---------
def det_log2f(x):
    x = np.asarray(x, f32)
    bits = x.view(np.uint32)
    e = (bits >> np.uint32(23)).astype(np.int32) - 127
    m = ((bits & np.uint32(0x007fffff)) | np.uint32(0x3f800000)).view(f32)
    big = m > f32(1.41421354)
    m = np.where(big, m * f32(0.5), m)
    e = np.where(big, e + 1, e)
    md = m.astype(np.float64)
    s = (md - 1.0) / (md + 1.0)
    z = s * s
    p = np.full_like(z, 1.0 / 23.0)
    for k in (21, 19, 17, 15, 13, 11, 9, 7, 5, 3):
        p = p * z + 1.0 / k
    p = p * z + 1.0
    ln_m = (2.0 * s) * p
    r = e.astype(np.float64) + ln_m * 1.4426950408889634
    return r.astype(f32)


def det_exp2f(v):
    vd = np.asarray(v, f32).astype(np.float64)
    n = np.floor(vd + 0.5)
    f = vd - n
    t = f * 0.6931471805599453
    p = np.full_like(t, 1.0 / 6227020800.0)
    for d in (479001600.0, 39916800.0, 3628800.0, 362880.0, 40320.0, 5040.0,
              720.0, 120.0, 24.0, 6.0):
        p = p * t + 1.0 / d
    p = p * t + 0.5
    p = p * t + 1.0
    p = p * t + 1.0
    nc = np.clip(n, -1000.0, 1000.0)
    scale = np.ldexp(1.0, np.nan_to_num(nc).astype(np.int64).astype(np.int32))
    with np.errstate(over='ignore'):
        r = (p * scale).astype(f32)
    return np.where(np.isnan(n), f32(np.nan), r)


def det_sincosf(x):
    xd = np.asarray(x, f32).astype(np.float64)
    k = np.floor(xd * 0.6366197723675814 + 0.5)
    r = (xd - k * 1.57079632673412561417e+00) - k * 6.07710050650619224932e-11
    z = r * r
    ps = np.full_like(z, 1.0 / 355687428096000.0)
    ps = ps * z - 1.0 / 1307674368000.0
    ps = ps * z + 1.0 / 6227020800.0
    ps = ps * z - 1.0 / 39916800.0
    ps = ps * z + 1.0 / 362880.0
    ps = ps * z - 1.0 / 5040.0
    ps = ps * z + 1.0 / 120.0
    ps = ps * z - 1.0 / 6.0
    ps = ps * z + 1.0
    s = r * ps
    pc = np.full_like(z, 1.0 / 20922789888000.0)
    pc = pc * z - 1.0 / 87178291200.0
    pc = pc * z + 1.0 / 479001600.0
    pc = pc * z - 1.0 / 3628800.0
    pc = pc * z + 1.0 / 40320.0
    pc = pc * z - 1.0 / 720.0
    pc = pc * z + 1.0 / 24.0
    pc = pc * z - 0.5
    pc = pc * z + 1.0
    c = pc
    q = k.astype(np.int64) & 3
    so = np.select([q == 0, q == 1, q == 2], [s, c, -s], -c)
    co = np.select([q == 0, q == 1, q == 2], [c, -s, -c], s)
    return so.astype(f32), co.astype(f32)


# ---- splines ------------------------------------------------------------------------
def normalize_spline(val, scale):
    """Any JSON spelling of a spline -> sorted (times, values) with guard knots."""
    if isinstance(val, (int, float)):
        v0 = v1 = 0.0
        pts = [(0.0, float(val)), (1.0, float(val))]
    else:
        if len(val) % 2:
            raise ValueError('odd-length spline')
        if len(val) == 2:
            v0 = v1 = 0.0
            pts = [(0.0, val[0]), (1.0, val[1])]
        else:
            v0, v1 = val[1], val[3]
            pts = [(0.0, val[0]), (1.0, val[2])]
            pts += [(val[i], val[i + 1]) for i in range(4, len(val), 2)]
    v0 *= scale
    v1 *= scale
    pts.sort()
    if pts[0][0] >= 0:
        pts = [(-2.0, pts[1][1] - (pts[1][0] + 2.0) * v0)] + pts
    if pts[-1][0] <= 1:
        pts = pts + [(3.0, pts[-2][1] + (3.0 - pts[-2][0]) * v1)]
    t = np.full(32, 1e9, f32)
    k = np.zeros(32, f32)
    t[:len(pts)] = [p[0] for p in pts]
    k[:len(pts)] = [p[1] for p in pts]
    return t, k


def knot_search(times, t):
    """Rightmost index with times[i] < t, 5 halving steps; t is an array."""
    lo = np.zeros(t.shape, np.int64)
    for step in (16, 8, 4, 2, 1):
        lo = np.where(t > times[lo + step], lo + step, lo)
    return lo


_ELBOW, _EOFF = f32(0.0625), f32(5.0)


def _mag_fwd(x):
    x = np.asarray(x, f32)
    pos, neg = x > _ELBOW, x < -_ELBOW
    safe = np.where(pos, x, np.where(neg, -x, f32(1.0)))
    lg = det_log2f(safe) + _EOFF
    return np.where(pos, lg, np.where(neg, -lg, x / _ELBOW)).astype(f32)


def _mag_inv(v):
    v = np.asarray(v, f32)
    pos, neg = v >= f32(1.0), v <= f32(-1.0)
    arg = np.where(pos, v - _EOFF, np.where(neg, -v - _EOFF, f32(0.0))).astype(f32)
    ex = det_exp2f(arg)
    return np.where(pos, ex, np.where(neg, -ex, v * _ELBOW)).astype(f32)


def _mag_slope(x, m):
    x, m = np.asarray(x, f32), np.asarray(m, f32)
    return np.where(x >= _ELBOW, m / x,
                    np.where(x <= -_ELBOW, m / -x, m / _ELBOW)).astype(f32)


def catmull_rom(times, knots, t, mag=False):
    """interp.py:318-355 in float32; t is a float32 array."""
    with np.errstate(all='ignore'):
        t = np.asarray(t, f32)
        idx = np.maximum(knot_search(times, t), 1)
        i3 = np.minimum(idx + 2, 31)
        t1 = times[idx]
        t2 = times[idx + 1] - t1
        rt2 = f32(1.0) / t2
        t0 = (times[idx - 1] - t1) * rt2
        t3 = (times[i3] - t1) * rt2
        u = (t - t1) * rt2
        k0, k1, k2, k3 = knots[idx - 1], knots[idx], knots[idx + 1], knots[i3]
        m1 = (k2 - k0) / (f32(1.0) - t0)
        m2 = (k3 - k1) / t3
        if mag:
            m1 = _mag_slope(k1, m1)
            m2 = _mag_slope(k2, m2)
            k1 = _mag_fwd(k1)
            k2 = _mag_fwd(k2)
        uu = u * u
        uuu = uu * u
        b1 = (uuu - f32(2.0) * uu) + u
        b2 = (f32(2.0) * uuu - f32(3.0) * uu) + f32(1.0)
        b3 = uuu - uu
        b4 = f32(-2.0) * uuu + f32(3.0) * uu
        r = ((m1 * b1 + k1 * b2) + m2 * b3) + k2 * b4
        r = r.astype(f32)
        if mag:
            r = _mag_inv(r)
        return r


def sample_times(tstart, tstep, n):
    """time_i = fma(i, tstep, tstart) in float32 (exact product in float64)."""
    i = np.arange(n, dtype=np.float64)
    return (i * np.float64(f32(tstep)) + np.float64(f32(tstart))).astype(f32)


# ---- genome access -------------------------------------------------------------------
def _spec_at(path):
    sp = _specs.anim
    for name in path:
        sp = sp.type if isinstance(sp, _Map) else sp[name]
    return sp


class GenomeEval(object):
    """Evaluates every parameter the hot path needs at the frame's temporal samples."""

    def __init__(self, gnm, w, h, tc, td, nts=NTS):
        self.gnm, self.nts = gnm, nts
        self.dim = calc_dim(w, h)
        self.scale = gnm.get('time', {}).get('duration', 1)
        ts = tc - 0.5 * td
        self.ts, self.td = ts, td
        self.times = sample_times(ts, td / nts, nts)
        self.xform_ids = sorted(str(k) for k in gnm['xforms'])
        self.has_final = 'final_xform' in gnm
        self.xaos = any('chaos' in gnm['xforms'][k] for k in gnm['xforms'])
        self.values = {}
        self._build()

    def spline(self, path):
        """float32 [nts] values of the animated parameter at `path`."""
        attr = self.gnm
        for name in path:
            if not isinstance(attr, dict) or name not in attr:
                attr = _spec_at(path).default
                break
            attr = attr[name]
        t, k = normalize_spline(attr, self.scale)
        return catmull_rom(t, k, self.times, mag=(_spec_at(path).interp == 'mag'))

    def _set(self, path, val):
        self.values['.'.join(path)] = np.asarray(val, f32)

    def _affine(self, apath):
        d2r = lambda a: (a * f32(3.14159274101257)) / f32(180.0)
        pri, spr = d2r(self.spline(apath + ('angle',))), d2r(self.spline(apath + ('spread',)))
        magx = self.spline(apath + ('magnitude', 'x'))
        magy = self.spline(apath + ('magnitude', 'y'))
        sm, cm = det_sincosf(pri - spr)
        sp, cp = det_sincosf(pri + spr)
        self._set(apath + ('xx',), magx * cm)
        self._set(apath + ('yx',), -magx * sm)
        self._set(apath + ('xy',), -magy * cp)
        self._set(apath + ('yy',), magy * sp)
        self._set(apath + ('xo',), self.spline(apath + ('offset', 'x')))
        self._set(apath + ('yo',), -self.spline(apath + ('offset', 'y')))

    def _xform(self, xpath, xf):
        self._affine(xpath + ('pre_affine',))
        if 'post_affine' in xf:
            self._affine(xpath + ('post_affine',))
        self._set(xpath + ('color',), self.spline(xpath + ('color',)))
        self._set(xpath + ('color_speed',), self.spline(xpath + ('color_speed',)))
        if 'opacity' in xf:
            self._set(xpath + ('opacity',), self.spline(xpath + ('opacity',)))
        for v in sorted(xf.get('variations', {})):
            vp = xpath + ('variations', v)
            self._set(vp + ('weight',), self.spline(vp + ('weight',)))
            for p in _VAR_PARAMS[v]:
                self._set(vp + (p,), self.spline(vp + (p,)))
            with np.errstate(all='ignore'):
                if v == 'waves':
                    dx = self.spline(xpath + ('pre_affine', 'offset', 'x'))
                    dy = self.spline(xpath + ('pre_affine', 'offset', 'y'))
                    self._set(vp + ('dx2',), f32(1.0) / (dx * dx + f32(1.0e-20)))
                    self._set(vp + ('dy2',), f32(1.0) / (dy * dy + f32(1.0e-20)))
                elif v == 'perspective':
                    ang = self.spline(vp + ('angle',)) * f32(1.57079637050629)
                    pdist = np.maximum(f32(1e-9), self.spline(vp + ('dist',)))
                    sn, cs = det_sincosf(ang)
                    self._set(vp + ('mdist',), pdist)
                    self._set(vp + ('sin',), sn)
                    self._set(vp + ('cos',), pdist * cs)
                elif v in ('julian', 'juliascope'):
                    self._set(vp + ('cn',), self.spline(vp + ('dist',))
                              / (f32(2.0) * self.spline(vp + ('power',))))
                elif v == 'curve':
                    xl, yl = self.spline(vp + ('xlength',)), self.spline(vp + ('ylength',))
                    self._set(vp + ('x2',), f32(1.0) / np.maximum(f32(1e-20), xl * xl))
                    self._set(vp + ('y2',), f32(1.0) / np.maximum(f32(1e-20), yl * yl))

    def _build(self):
        g, dim = self.gnm, self.dim
        for xid in self.xform_ids:
            self._xform(('xforms', xid), g['xforms'][xid])
        if self.has_final:
            self._xform(('final_xform',), g['final_xform'])
        # cumulative normalised densities (iter.py:12-30)
        with np.errstate(all='ignore'):
            ws = [self.spline(('xforms', xid, 'weight')) for xid in self.xform_ids]
            total = np.zeros(self.nts, f32)
            for w_ in ws:
                total = total + w_
            rsum = f32(1.0) / total
            run = np.zeros(self.nts, f32)
            if not self.xaos:
                for xid, w_ in list(zip(self.xform_ids, ws))[:-1]:
                    run = run + w_ * rsum
                    self._set(('xforms', xid, 'density'), run)
            else:
                # precalc_chaos (code/iter.py:32-54): per previous xform p, the cumulative
                # normalised products weight[n] * chaos[p][n]
                for pid in self.xform_ids:
                    den = [w_ * self.spline(('xforms', pid, 'chaos', nid))
                           for nid, w_ in zip(self.xform_ids, ws)]
                    total = np.zeros(self.nts, f32)
                    for d in den:
                        total = total + d
                    rsum = f32(1.0) / total
                    run = np.zeros(self.nts, f32)
                    for nid, d in list(zip(self.xform_ids, den))[:-1]:
                        run = run + d * rsum
                        self._set(('xforms', pid, 'chaos_den', nid), run)
        # camera (iter.py:56-79)
        rot = (self.spline(('camera', 'rotation')) * f32(3.14159274101257)) / f32(180.0)
        rs, rc = det_sincosf(rot)
        cenx, ceny = self.spline(('camera', 'center', 'x')), self.spline(('camera', 'center', 'y'))
        scale = self.spline(('camera', 'scale')) * f32(dim['w'])
        self._set(('camera', 'xx'), scale * rc)
        self._set(('camera', 'xy'), scale * -rs)
        self._set(('camera', 'xo'), scale * (rs * ceny - rc * cenx) + f32(0.5) * f32(dim['aw']))
        self._set(('camera', 'yx'), scale * rs)
        self._set(('camera', 'yy'), scale * rc)
        self._set(('camera', 'yo'), scale * -(rs * cenx + rc * ceny) + f32(0.5) * f32(dim['ah']))

    # ---- flat records for chaos.c --------------------------------------------------
    def _var_args(self, xpath, v):
        g = lambda *p: self.values['.'.join(xpath + p)]
        vp = ('variations', v)
        if v == 'waves':
            return [g('pre_affine', 'xy'), g('pre_affine', 'yy'), g(*vp, 'dx2'), g(*vp, 'dy2')]
        if v == 'popcorn':
            return [g('pre_affine', 'xo'), g('pre_affine', 'yo')]
        if v == 'rings':
            return [g('pre_affine', 'xo')]
        if v == 'fan':
            return [g('pre_affine', 'xo'), g('pre_affine', 'yo')]
        if v == 'perspective':
            return [g(*vp, 'mdist'), g(*vp, 'sin'), g(*vp, 'cos')]
        if v in ('julian', 'juliascope'):
            return [g(*vp, 'power'), g(*vp, 'cn')]
        if v == 'curve':
            return [g(*vp, 'xamp'), g(*vp, 'yamp'), g(*vp, 'x2'), g(*vp, 'y2')]
        return [g(*vp, p) for p in _VAR_PARAMS[v]]

    def xform_record(self, xpath, xf, lib):
        n = lib.oracle_xf_floats()
        rec = np.zeros((self.nts, n), f32)
        g = lambda *p: self.values['.'.join(xpath + p)]
        for i, c in enumerate(('xx', 'xy', 'xo', 'yx', 'yy', 'yo')):
            rec[:, i] = g('pre_affine', c)
            if 'post_affine' in xf:
                rec[:, 6 + i] = g('post_affine', c)
        rec[:, 12] = 1.0 if 'post_affine' in xf else 0.0
        rec[:, 13] = g('color')
        rec[:, 14] = g('color_speed')
        names = sorted(xf.get('variations', {}))
        if len(names) > 16:
            raise ValueError('oracle supports at most 16 variations per xform')
        rec[:, 15] = len(names)
        if 'opacity' in xf:
            rec[:, n - 2] = 1.0
            rec[:, n - 1] = g('opacity')
        for vi, v in enumerate(names):
            base = 16 + vi * 12
            rec[:, base] = _VAR_NUM[v]
            rec[:, base + 1] = g('variations', v, 'weight')
            for ai, arr in enumerate(self._var_args(xpath, v)):
                rec[:, base + 2 + ai] = arr
        return rec

    def frame_records(self, lib):
        nxf = len(self.xform_ids)
        total = nxf + (1 if self.has_final else 0)
        stride = lib.oracle_frame_floats(total)
        fr = np.zeros((self.nts, stride), f32)
        for i, c in enumerate(('xx', 'xy', 'xo', 'yx', 'yy', 'yo')):
            fr[:, i] = self.values['camera.' + c]
        if not self.xaos:
            for i, xid in enumerate(self.xform_ids[:-1]):
                fr[:, 6 + i] = self.values['xforms.%s.density' % xid]
        xfn = lib.oracle_xf_floats()
        off = 6 + 64
        for i, xid in enumerate(self.xform_ids):
            fr[:, off + i * xfn: off + (i + 1) * xfn] = self.xform_record(
                ('xforms', xid), self.gnm['xforms'][xid], lib)
        if self.has_final:
            fr[:, off + nxf * xfn: off + (nxf + 1) * xfn] = self.xform_record(
                ('final_xform',), self.gnm['final_xform'], lib)
        return fr


# ---- palette (interp.py:372-433; genome/util.py:75-87) --------------------------------
def decode_palette(entry):
    assert entry[0] == 'rgb8'
    raw = base64.b64decode(''.join(entry[1:]))
    rgb = np.frombuffer(raw, np.uint8).reshape(256, 3)
    out = np.ones((256, 4), f32)
    out[:, :3] = rgb / 255.0
    return out


def palette_table(gnm, ts, td, seeds, rows=PAL_ROWS):
    """
    float32 [rows][256][4] = dithered 8-bit (Y,U,V)/255 and 1.  Entry (r, c) uses
    RNG stream r*256+c (three draws).  Returns (table, updated seeds).
    """
    pals = sorted((float(p[0]), decode_palette(p[1:])) for p in gnm['palette'])
    ptimes = np.full(32, 1e9, f32)
    ptimes[:len(pals)] = [p[0] for p in pals]
    src = np.stack([p[1] for p in pals] + [np.zeros((256, 4), f32)])
    tt = sample_times(ts, td / rows, rows)
    idx = np.maximum(knot_search(ptimes, tt) + 1, 1)
    tr = ptimes[idx]
    with np.errstate(all='ignore'):
        lf = (tr - tt) / (tr - ptimes[idx - 1])
    single = tr > f32(1.0)
    lf = np.where(single, f32(1.0), lf).astype(f32)
    rf = np.where(single, f32(0.0), f32(1.0) - lf).astype(f32)
    left = src[idx - 1]                                       # [rows][256][4]
    right = np.where(single[:, None, None], left, src[np.minimum(idx, len(pals))])
    lf, rf = lf[:, None], rf[:, None]

    def yuv(p):
        r, g, b = p[..., 0], p[..., 1], p[..., 2]
        y = (f32(0.299) * r + f32(0.587) * g) + f32(0.114) * b
        u = (f32(-0.168736) * r - f32(0.331264) * g) + f32(0.5) * b
        v = (f32(0.5) * r - f32(0.418688) * g) - f32(0.081312) * b
        return y, u, v
    ly, lu, lv = yuv(left)
    ry, ru, rv = yuv(right)
    y = ly * lf + ry * rf
    u = (lu * lf + ru * rf) + f32(0.5)
    v = (lv * lf + rv * rf) + f32(0.5)

    rng = MwcStreams(seeds[:rows * 256])
    out = np.ones((rows, 256, 4), f32)
    for ch, val in enumerate((y, u, v)):
        d = rng.next_11().reshape(rows, 256)
        q = val * f32(255.0) + f32(0.49) * d
        q = np.minimum(f32(255.0), np.maximum(f32(0.0), np.trunc(q))) + f32(0.0)   # no -0
        out[:, :, ch] = q * f32(1.0 / 255.0)
    new_seeds = np.array(seeds, copy=True)
    new_seeds[:rows * 256] = rng.seeds()
    return out, new_seeds


# ---- the C part ---------------------------------------------------------------------------
_lib = None


def chaos_lib():
    global _lib
    if _lib is None:
        from .build import build
        L = ctypes.CDLL(build())
        L.oracle_xf_floats.restype = ctypes.c_int
        L.oracle_frame_floats.restype = ctypes.c_int
        L.oracle_frame_floats.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _p(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


def mwc_sums(seeds, rounds):
    seeds = np.ascontiguousarray(seeds, np.uint32).copy()
    sums = np.zeros(seeds.shape[0], np.uint64)
    chaos_lib().oracle_mwc_sums(_p(seeds), ctypes.c_int(seeds.shape[0]),
                                ctypes.c_int(rounds), _p(sums))
    return sums, seeds


def variation(var_id, args, w, txs, tys, seeds):
    """One variation (flam3 number) on arrays: returns (tx, ty, ox, oy, seeds)."""
    txs, tys = np.array(txs, f32), np.array(tys, f32)
    oxs, oys = np.zeros_like(txs), np.zeros_like(txs)
    seeds = np.array(seeds, np.uint32)
    a = np.zeros(12, f32)
    a[:len(args)] = args
    chaos_lib().oracle_variation(ctypes.c_int(var_id), _p(a), ctypes.c_float(w), _p(txs), _p(tys),
                                 _p(oxs), _p(oys), _p(seeds), ctypes.c_int(txs.size))
    return txs, tys, oxs, oys, seeds


def var_number(name):
    return _VAR_NUM[name]


def apply_xform(xf_record, xs, ys, cs, seeds):
    xs, ys, cs = (np.ascontiguousarray(a, f32).copy() for a in (xs, ys, cs))
    seeds = np.ascontiguousarray(seeds, np.uint32).copy()
    rec = np.ascontiguousarray(xf_record, f32)
    chaos_lib().oracle_apply_xform(_p(rec), _p(xs), _p(ys), _p(cs), _p(seeds),
                                   ctypes.c_int(xs.size))
    return xs, ys, cs, seeds


def point_to_bin(cam, xs, ys, cs, dithers, astride, aheight):
    xs, ys, cs, dithers = (np.ascontiguousarray(a, f32) for a in (xs, ys, cs, dithers))
    cam = np.ascontiguousarray(cam, f32)
    bins = np.empty(xs.size, np.int32)
    cidx = np.empty(xs.size, np.int32)
    chaos_lib().oracle_point_to_bin(_p(cam), _p(xs), _p(ys), _p(cs), _p(dithers),
                                    ctypes.c_int(xs.size), ctypes.c_int(astride),
                                    ctypes.c_int(aheight), _p(bins), _p(cidx))
    return bins, cidx


def iterate(ev, palette, seeds, nsamples, ntraj=4096, fuse=32, nthreads=0):
    """Run the chaos game; returns the float32 [ah][astride][4] histogram."""
    L = chaos_lib()
    fr = np.ascontiguousarray(ev.frame_records(L))
    dim = ev.dim
    hist = np.zeros((dim['ah'], dim['astride'], 4), f32)
    seeds = np.ascontiguousarray(seeds, np.uint32).copy()
    pal = np.ascontiguousarray(palette, f32)
    xaos = None
    if ev.xaos and len(ev.xform_ids) > 1:
        ids = ev.xform_ids
        xaos = np.ascontiguousarray(np.stack(
            [np.stack([ev.values['xforms.%s.chaos_den.%s' % (p, n)] for n in ids[:-1]], axis=1)
             for p in ids], axis=1), f32)                     # [nts][nxf][nxf-1]
    L.oracle_iterate_ex(_p(fr), ctypes.c_int(fr.shape[1]), ctypes.c_int(ev.nts),
                        ctypes.c_int(len(ev.xform_ids)), ctypes.c_int(1 if ev.has_final else 0),
                        _p(pal), ctypes.c_int(pal.shape[0]), _p(hist),
                        ctypes.c_int(dim['astride']), ctypes.c_int(dim['ah']), _p(seeds),
                        ctypes.c_int(ntraj), ctypes.c_uint64(int(nsamples)),
                        ctypes.c_int(fuse), ctypes.c_int(nthreads),
                        _p(xaos) if xaos is not None else None)
    return hist, seeds
