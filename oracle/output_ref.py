"""
ORACLE (test infrastructure only) -- pixel-format conversion in numpy.

Restates cuburn/code/output.py:7-225 with the deterministic RNG assignment the
C ABI documents for cb_convert: pixel i (row-major over w x h) is produced by
MWC stream i % nstreams, streams walk their pixels in increasing order, and a
random number is drawn for every dithered channel whether or not it is used.
Arithmetic is float32 with one rounding per operation, so the result is
comparable bit for bit with the device.
"""
import numpy as np

from .flame_ref import MwcStreams

f32 = np.float32
GUTTER = 12


def _dclamp(rnd, peak, v):
    q = np.minimum(f32(peak), v * f32(peak) + f32(0.99) * rnd)
    return np.where(v > 0, q, f32(0)).astype(f32)


def _jpeg(p):
    r, g, b = p[..., 0], p[..., 1], p[..., 2]
    y = (f32(0.299) * r + f32(0.587) * g) + f32(0.114) * b
    cb = (f32(-0.168736) * r - f32(0.331264) * g) + f32(0.5) * b
    cr = (f32(0.5) * r - f32(0.418688) * g) - f32(0.081312) * b
    return y.astype(f32), cb.astype(f32), cr.astype(f32)


def convert(fmt, src, w, h, seeds, nstreams=None):
    """
    fmt in {'rgba_u8','rgba_u16','yuv444p','yuv444p10','yuv420p10','yuv444p12'}.
    src: float32 [ah][astride][4].  Returns (flat output array, updated seeds).
    """
    seeds = np.asarray(seeds, np.uint32)
    nstreams = nstreams or seeds.shape[0]
    rng = MwcStreams(seeds[:nstreams])
    crop = src[GUTTER:GUTTER + h, GUTTER:GUTTER + w].reshape(-1, 4).astype(f32)
    npix = w * h
    rounds = (npix + nstreams - 1) // nstreams

    if fmt in ('rgba_u8', 'rgba_u16'):
        peak = 255.0 if fmt == 'rgba_u8' else 65535.0
        out = np.zeros((npix, 4), np.uint8 if fmt == 'rgba_u8' else np.uint16)
        nchan = 4
    elif fmt == 'yuv420p10':
        out = np.zeros(npix * 3 // 2, np.uint16)
        nchan = 1
    else:
        out = np.zeros(3 * npix, np.uint8 if fmt == 'yuv444p' else np.uint16)
        nchan = 3

    for j in range(rounds):
        lo = j * nstreams
        n = min(nstreams, npix - lo)
        active = np.arange(nstreams) < n
        idx = lo + np.arange(n)
        p = crop[idx]
        draw = lambda m=active: rng.next_01(m)[:n]
        if fmt in ('rgba_u8', 'rgba_u16'):
            for ch in range(4):
                out[idx, ch] = _dclamp(draw(), peak, p[:, ch]).astype(out.dtype)
        elif fmt in ('yuv444p', 'yuv444p10'):
            peak = 255.0 if fmt == 'yuv444p' else 1023.0
            y, cb, cr = _jpeg(p)
            cb, cr = cb + f32(0.5), cr + f32(0.5)
            out[idx] = _dclamp(draw(), peak, y).astype(out.dtype)
            qu = _dclamp(draw(), peak, cb)
            if fmt == 'yuv444p10':
                qu = np.minimum(f32(1023.0), np.maximum(f32(0), f32(1023.0) * cb))
            out[idx + npix] = qu.astype(out.dtype)
            out[idx + 2 * npix] = _dclamp(draw(), peak, cr).astype(out.dtype)
        elif fmt == 'yuv444p12':
            q = np.clip(p[:, :3], 0, 1).astype(f32)
            r, g, b = q[:, 0], q[:, 1], q[:, 2]
            yy = (f32(0.2126) * r + f32(0.7152) * g) + f32(0.0722) * b
            cb = ((f32(-0.11457) * r - f32(0.38543) * g) + f32(0.5) * b) + f32(0.5)
            cr = ((f32(0.5) * r - f32(0.45416) * g) - f32(0.04585) * b) + f32(0.5)
            out[idx] = (_dclamp(draw(), 3504.0, yy) + f32(256.0)).astype(out.dtype)
            out[idx + npix] = (_dclamp(draw(), 3584.0, cb) + f32(256.0)).astype(out.dtype)
            out[idx + 2 * npix] = (_dclamp(draw(), 3584.0, cr) + f32(256.0)).astype(out.dtype)
        else:   # yuv420p10
            y, _, _ = _jpeg(p)
            out[idx] = _dclamp(draw(), 1023.0, y).astype(out.dtype)
            py, px = idx // w, idx % w
            site = (2 * px < w) & (2 * py < h)
            smask = np.zeros(nstreams, bool)
            smask[:n] = site
            r1 = rng.next_01(smask)[:n]
            r2 = rng.next_01(smask)[:n]
            if site.any():
                sy, sx = py[site], px[site]
                blk = src[GUTTER + 2 * sy[:, None] + np.array([0, 0, 1, 1]),
                          GUTTER + 2 * sx[:, None] + np.array([0, 1, 0, 1])].astype(f32)
                wts = blk[..., 3]
                _, bcb, bcr = _jpeg(blk)
                tot = (wts[:, 0].astype(np.float64) + 1e-12).astype(f32)
                cbs = wts[:, 0] * bcb[:, 0]
                crs = wts[:, 0] * bcr[:, 0]
                for k in (1, 2, 3):
                    tot = tot + wts[:, k]
                    cbs = cbs + wts[:, k] * bcb[:, k]
                    crs = crs + wts[:, k] * bcr[:, k]
                ci = npix + (w // 2) * sy + sx
                out[ci] = _dclamp(r1[site], 1023.0, cbs / tot + f32(0.5)).astype(out.dtype)
                out[ci + npix // 4] = _dclamp(r2[site], 1023.0, crs / tot + f32(0.5)).astype(out.dtype)

    new_seeds = np.array(seeds, copy=True)
    new_seeds[:nstreams] = rng.seeds()
    return out, new_seeds
