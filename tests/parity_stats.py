"""
Statistics for comparing two Monte-Carlo histograms of the same flame
(float32 [ah][astride][4]: sum Y, sum U, sum V, count -- SURVEY a17).

Used by the density-parity tests and by tools/parity_calibrate.py.
"""
import numpy as np


def pool(a, k=8):
    hh, ww = a.shape[0] // k * k, a.shape[1] // k * k
    return a[:hh, :ww].astype(np.float64).reshape(hh // k, k, ww // k, k).sum(axis=(1, 3))


def density_z(ha, hb, k=8, min_count=400):
    """
    z statistics of pooled k x k bin counts: z = (a - b) / sqrt(a + b) over pooled bins
    holding more than ``min_count`` samples in total.  For two Poisson fields of equal
    mean z is ~N(0,1); over-dispersed counts (index D on each side) give var z = D.
    """
    pa, pb = pool(ha[..., 3], k), pool(hb[..., 3], k)
    m = (pa + pb) > min_count
    z = (pa - pb)[m] / np.sqrt((pa + pb)[m])
    return dict(mean=float(z.mean()), std=float(z.std()), max=float(np.abs(z).max()),
                cells=int(m.sum()), mass=float((pa + pb)[m].sum() / max((pa + pb).sum(), 1)))


def colour_means(ha, hb, k=8, min_count=400):
    """Mean absolute difference of the density-normalised channel means (Y, U, V in
    [0, 1] of full scale) over the well-populated pooled bins, and the largest one."""
    pa, pb = pool(ha[..., 3], k), pool(hb[..., 3], k)
    m = (pa > min_count / 2) & (pb > min_count / 2)
    out = {}
    for ch, name in enumerate('YUV'):
        ca, cb = pool(ha[..., ch], k), pool(hb[..., ch], k)
        d = np.abs(ca[m] / pa[m] - cb[m] / pb[m])
        out[name] = dict(mean=float(d.mean()), max=float(d.max()))
    return out


def in_frame_fraction(h, n):
    """Share of the launched samples that landed inside the accumulation grid."""
    return float(h[..., 3].astype(np.float64).sum() / n)
