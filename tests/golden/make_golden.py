#!/usr/bin/env python
"""
Regenerate tests/golden/oracle_golden.json.

The reference (Python 2 + PyCUDA) cannot be imported or run in this
environment, so these are *oracle* outputs frozen at the commit that first
passed bit-exact parity against the device on a B200 -- a regression pin for the
oracle, not reference-generated vectors.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cuburn_b200 import samples, mwc          # noqa: E402
from oracle import flame_ref as R              # noqa: E402


def main():
    out = {'params': []}
    cases = [dict(genome='G3', w=640, h=360, tc=1.5 / 720, td=0.0),
             dict(genome='G6F', kwargs={'animated': True}, w=1920, h=1080, tc=0.37, td=1.0 / 720),
             dict(genome='G24H', w=3840, h=2160, tc=0.5, td=0.0)]
    for case in cases:
        g = samples.GENOMES[case['genome']](**case.get('kwargs', {}))
        ev = R.GenomeEval(g, case['w'], case['h'], case['tc'], case['td'])
        names = sorted(ev.values)
        pick = names[::max(1, len(names) // 24)]
        case['values'] = {n: [[i, int(ev.values[n][i].view(np.uint32))] for i in (0, 511, 1023)]
                          for n in pick}
        out['params'].append(case)
    g = samples.g6f()
    pal, _ = R.palette_table(g, 0.1, 0.2, mwc.make_seeds(16384, host_seed=5))
    out['palette'] = dict(genome='G6F', ts=0.1, td=0.2, seed=5, entries=[
        [r, c, [int(x) for x in np.round(pal[r, c, :3] * 255)]]
        for r in (0, 31, 63) for c in (0, 1, 100, 255)])
    ev = R.GenomeEval(samples.g3(), 160, 90, 0.5, 0.0)
    seeds = mwc.make_seeds(16384 + 64, host_seed=9)
    pal, seeds = R.palette_table(samples.g3(), ev.ts, 0.0, seeds)
    hist, _ = R.iterate(ev, pal, seeds, 300000, ntraj=64, nthreads=1)
    out['chaos'] = dict(genome='G3', w=160, h=90, tc=0.5, seed=9, nsamples=300000,
                        inside=int(hist[..., 3].sum()), argmax=int(np.argmax(hist[..., 3])),
                        max=int(hist[..., 3].max()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_golden.json')
    with open(path, 'w') as fp:
        json.dump(out, fp, indent=1)
    print('wrote', path)


if __name__ == '__main__':
    main()
