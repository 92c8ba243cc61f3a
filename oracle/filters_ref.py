"""
ORACLE (test infrastructure only) -- the filter chain in numpy float32.

Restates the reference filter kernels (cuburn/code/filters.py:4-413,
cuburn/code/color.py:33-40) and their host recipes (cuburn/filters.py:46-198)
on [ah][astride][4] float32 arrays.  Neighbour taps use clamp-to-edge
addressing (the only mode CUDA honours for unnormalised texture coordinates,
SURVEY.md Q16).  libm-accurate math; the device kernels use fast-math
intrinsics, so comparisons carry a tolerance.
"""
import numpy as np

f32 = np.float32

DIRS = np.array([
    (1.0, 0.0), (0.0, 1.0), (1.0, 1.0), (-1.0, 1.0),
    (1.0, 0.5), (-0.5, 1.0), (1.0, -0.5), (0.5, 1.0),
    (1.0, 0.666667), (-0.666667, 1.0), (1.0, -0.666667), (0.666667, 1.0),
    (1.0, 0.333333), (-0.333333, 1.0), (1.0, -0.333333), (0.333333, 1.0),
], dtype=f32)


def ftz(a):
    """Flush subnormals to zero: the reference's modules are built with
    -use_fast_math (code/util.py:96), i.e. every float op runs in FTZ mode."""
    a = np.asarray(a, f32)
    return np.where(np.abs(a) < f32(1.17549435e-38), f32(0), a).astype(f32)


def shear_offset(pattern, radius):
    """Round-half-even of each direction component times radius (filters.py:22-33)."""
    d = DIRS[pattern]
    return int(np.rint(f32(d[0]) * f32(radius))), int(np.rint(f32(d[1]) * f32(radius)))


def tap(arr, pattern, radius):
    """arr sampled at (x + i, y + j) with clamp-to-edge, for every pixel."""
    i, j = shear_offset(pattern, radius)
    h, w = arr.shape[:2]
    rows = np.clip(np.arange(h) + j, 0, h - 1)
    cols = np.clip(np.arange(w) + i, 0, w - 1)
    return arr[rows][:, cols]


def gauss_coefs(stdev=1):
    c = np.exp(np.float32(np.arange(-3, 4)) ** 2 / (-2 * stdev ** 2)).astype(f32)
    return (c / np.sum(c)).astype(f32)


def yuv_to_rgb(pix):
    y, u, v, w = (pix[..., k] for k in range(4))
    u = u - f32(0.5) * w
    v = v - f32(0.5) * w
    out = np.empty_like(pix)
    out[..., 0] = np.maximum(0, y + f32(1.402) * v)
    out[..., 1] = np.maximum(0, y - f32(0.34414) * u - f32(0.71414) * v)
    out[..., 2] = np.maximum(0, y + f32(1.772) * u)
    out[..., 3] = w
    return out


def blur7(arr, pattern, upsample, coefs):
    acc = np.zeros_like(arr)
    for i in range(7):
        acc = acc + tap(arr, pattern, (i - 3) * (1 << upsample)) * coefs[i]
    return ftz(acc)


def bilateral_pass(pix, pattern, radius, sstd, cstd, dstd, dpow, gspeed):
    """One direction of the bilateral filter (code/filters.py:166-264)."""
    with np.errstate(all='ignore'):
        pix = ftz(pix)
        coefs = gauss_coefs(1)
        den1 = blur7(pix[..., 3], pattern, 0, coefs)
        blur = blur7(den1, pattern, 1, coefs)
        sq2 = f32(1.41421353816986)
        spa = np.exp(np.arange(32, dtype=f32) ** 2 / (-sq2 * f32(sstd))).astype(f32)
        cscale = f32(1.0) / (-sq2 * f32(3.0) * f32(cstd))
        dscale = f32(-0.5) / f32(dstd)
        cw = pix[..., 3]
        cdrcp = f32(1.0) / (cw + f32(1.0e-6))
        cen = pix[..., :3] * cdrcp[..., None]
        cpow = np.power(cw, f32(dpow))
        acc = np.zeros_like(pix)
        wsum = np.zeros_like(cw)
        for r in range(-radius, radius + 1):
            p = tap(pix, pattern, r)
            prev = tap(pix[..., 3], pattern, r - 1)
            nxt = tap(pix[..., 3], pattern, r + 1)
            pw = p[..., 3]
            both = (pw > 0) & (cw > 0)
            pdrcp = f32(1.0) / np.where(pw > 0, pw, f32(1.0))
            diff = p[..., :3] * pdrcp[..., None] - cen
            cdiff = np.where(both, (diff * diff).sum(axis=-1), f32(0.5)).astype(f32)
            dfact = np.exp2(dscale * np.abs(cpow - np.power(pw, f32(dpow))))
            avg = tap(blur, pattern, r)
            grad = (nxt - prev) / (avg + f32(1.0e-6))
            if r < 0:
                grad = -grad
            gfact = np.exp2(-np.exp2(f32(gspeed) * grad))
            fac = spa[abs(r)] * np.exp(cscale * cdiff) * dfact
            if r != 0:
                fac = fac * gfact
            fac = ftz(fac)
            wsum = wsum + fac
            acc = acc + fac[..., None] * p
        rcp = f32(1.0) / (wsum + f32(1e-10))
        return ftz(acc * rcp[..., None])


def _in_strips(fn, pix, halo, threads):
    """Run ``fn`` (a stencil whose result at a row depends on at most ``halo`` rows either
    side) on horizontal strips in parallel threads; identical to ``fn(pix)``: a strip is
    cut with its halo, real image borders stay borders (so edge clamping is unchanged)."""
    from concurrent.futures import ThreadPoolExecutor
    rows = pix.shape[0]
    n = max(1, min(threads, rows // (4 * halo)))
    if n == 1:
        return fn(pix)
    edges = [rows * k // n for k in range(n + 1)]

    def job(k):
        a, b = edges[k], edges[k + 1]
        lo, hi = max(a - halo, 0), min(b + halo, rows)
        return fn(pix[lo:hi])[a - lo:a - lo + (b - a)]
    with ThreadPoolExecutor(n) as pool:
        return np.concatenate(list(pool.map(job, range(n))), axis=0)


def bilateral(pix, w, spatial_std=6, color_std=0.05, density_std=1.5, density_pow=0.8,
              gradient=4.0, radius=15, directions=8, threads=1):
    """``threads`` > 1: each direction pass runs on row strips in parallel (same result;
    a pass reaches radius + 1 taps plus the 6 + 3 rows of the two density blurs)."""
    sstd = spatial_std * w / 1920.
    for pattern in range(directions):
        fn = lambda p: bilateral_pass(p, pattern, radius, sstd, color_std, density_std,
                                      density_pow, gradient)
        pix = _in_strips(fn, pix, radius + 1 + 9 + 6, threads)
    return pix


def logscale(pix, k1, k2):
    with np.errstate(all='ignore'):
        w = pix[..., 3]
        ls = np.maximum(0, f32(k1) * np.log(f32(1.0) + w * f32(k2)) / w)
        ls = np.where(np.isnan(ls), f32(0), ls)     # fmaxf(0, NaN) == 0
        return (pix * ls[..., None]).astype(f32)


def logscale_consts(brightness, scale, w, h, spp):
    k1 = f32(brightness * 268 / 256)
    area = h / (scale ** 2 * w)
    return k1, f32(1.0 / (area * spp))


def calc_lingam(gamma, threshold):
    gam = f32(1 / gamma)
    lin = f32(threshold)
    lingam = f32(lin ** (gam - 1.0) if lin > 0 else 0)
    return gam, lin, lingam


def _gamma_toe(w, gm1, lin, lingam):
    with np.errstate(all='ignore'):
        ls = np.power(w, f32(gm1))
        frac = w / f32(lin) if lin > 0 else np.zeros_like(w)
        toe = (f32(1.0) - frac) * f32(lingam) + frac * ls
        return np.where(w < lin, toe, ls).astype(f32)


def smearclip(pix, width=0.7, gamma=4, threshold=0.01):
    gam, lin, lingam = calc_lingam(gamma, threshold)
    with np.errstate(all='ignore'):
        w = pix[..., 3]
        ls = np.where(w > 0, np.maximum(0, w - f32(1.0)) / np.where(w > 0, w, f32(1)), f32(0))
        hi = (pix * ls[..., None]).astype(f32)
        coefs = gauss_coefs(width)
        for pattern in (2, 3, 0, 1):
            hi = blur7(hi, pattern, 0, coefs)
        p = pix + hi
        w = p[..., 3]
        ls = _gamma_toe(w, gam - f32(1), lin, lingam)
        out = p * ls[..., None]
        out[w <= 0] = 0
        return out.astype(f32)


def plainclip(pix, brightness=1.0, gamma=4, threshold=0.01):
    gam, lin, lingam = calc_lingam(gamma, threshold)
    w = pix[..., 3]
    ls = _gamma_toe(w, gam - f32(1), lin, lingam) * f32(brightness)
    out = pix * ls[..., None]
    out[w <= 0] = 0
    return out.astype(f32)


def haloclip(pix, gamma=4):
    gm1 = f32(1 / gamma - 1)
    with np.errstate(all='ignore'):
        coefs = gauss_coefs(1)
        plane = np.power(pix[..., 0], f32(0.1)).astype(f32)
        plane = blur7(plane, 2, 0, coefs)
        plane = blur7(plane, 3, 0, coefs)
        w = pix[..., 3]
        ls = np.power(w, gm1) / np.maximum(f32(1.0), plane)
        out = pix * ls[..., None]
        out[w <= 0] = 0
        return out.astype(f32)


def colorclip(pix, vibrance=1, highlight_power=-1, gamma=4, threshold=0.01):
    gam, lin, lingam = calc_lingam(gamma, threshold)
    vib, hp = f32(vibrance), f32(highlight_power)
    with np.errstate(all='ignore'):
        w = pix[..., 3]
        rgb = pix[..., :3]
        alpha = np.power(w, gam)
        if lin > 0:
            frac = w / lin
            alpha = np.where(w < lin, (f32(1) - frac) * w * lingam + frac * alpha, alpha)
        ls = vib * alpha / w
        alpha = np.clip(alpha, 0, 1)
        maxc = rgb.max(axis=-1)
        maxa = maxc * ls
        newls = f32(1.0) / maxc
        hi = (maxa > 1) & (hp >= 0)
        lsratio = np.power(newls / ls, hp)
        a = maxc[..., None] - (maxc[..., None] - rgb * newls[..., None]) * lsratio[..., None]
        adjhlp = np.where((-hp > 1) | (maxa <= 1), f32(1.0), -hp)
        adj = (f32(1.0) - adjhlp) * newls + adjhlp * ls
        b = np.where((maxc > 0)[..., None], rgb * adj[..., None], rgb)
        out = np.where(hi[..., None], a, b)
        out = out + (f32(1.0) - vib) * np.power(rgb, gam)
        res = np.empty_like(pix)
        res[..., :3] = np.minimum(1, out)
        res[..., 3] = alpha
        res[w <= 0] = 0
        return res.astype(f32)


def logencode(pix, degamma=2.2):
    with np.errstate(all='ignore'):
        return (np.log2(np.power(pix, f32(degamma))) / f32(12.0) + f32(1.0)).astype(f32)


def default_chain(hist, w, h, scale, spp, brightness=4, gamma=4, threshold=0.01,
                  smear_width=0.7, threads=1):
    """yuv -> bilateral -> logscale -> smearclip, with schema defaults."""
    pix = yuv_to_rgb(hist)
    pix = bilateral(pix, w, threads=threads)
    k1, k2 = logscale_consts(brightness, scale, w, h, spp)
    pix = logscale(pix, k1, k2)
    return smearclip(pix, smear_width, gamma, threshold)
