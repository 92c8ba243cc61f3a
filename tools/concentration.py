"""
How concentrated are a flame's samples?  For each sample genome: the share of all samples
that the hottest 16x16-bin blocks (as many as fit a shared-memory budget) and the hottest
single bins capture, from oracle histograms (CPU only).  Input to the accumulation design
(profiles/r02_smem_atomics.md):  python tools/concentration.py [GENOME ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cuburn_b200 import samples, profile
from parity_calibrate import oracle_hist

W, H, SPP = 1920, 1080, 40
for gname in (sys.argv[1:] or ['G6F', 'G3', 'G24H', 'G2M']):
    gnm = samples.GENOMES[gname]()
    gprof = profile.wrap(dict(width=W, height=H, spp=SPP, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    hst = oracle_hist(gnm, W, H, SPP, 5, tc)[..., 3].astype(np.float64)
    tot = float(W * H * SPP)
    ah, ast = hst.shape
    blk = hst[:ah // 16 * 16, :ast // 16 * 16].reshape(ah // 16, 16, ast // 16, 16).sum((1, 3)).ravel()
    cs = np.cumsum(np.sort(blk)[::-1]) / tot
    flat = np.sort(hst.ravel())[::-1]
    print('%s %dx%d: hottest bin %.4f %% of all samples' % (gname, W, H, 100 * flat[0] / tot))
    for kb in (28, 56, 112, 224):
        nblk = kb * 1024 // 8 // 256
        print('   %3d KB of 8-byte cells = %3d blocks of 16x16 bins: %.1f %%' % (kb, nblk, 100 * cs[nblk - 1]))
    for share in (1 / 512., 1 / 2048.):
        sel = flat >= share * tot
        print('   bins holding >= 1/%d of the samples: %d, together %.1f %%'
              % (round(1 / share), sel.sum(), 100 * flat[sel].sum() / tot))
