#!/usr/bin/env python
"""
Generate tests/golden/host_golden.json by EXECUTING reference host modules
(cuburn/genome/use.py, cuburn/profile.py, cuburn/code/mwc.py make_seeds,
cuburn/filters.py calc_lingam / blur coefficients) under Python 3 with minimal
in-memory syntax patches.  Only runs where /root/reference is mounted.

    python tests/golden/make_host_golden.py
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference/cuburn/'

from cuburn_b200.genome import spectypes, specs      # noqa: E402


def exec_module(name, src, inject):
    saved = {k: sys.modules.get(k) for k in inject}
    sys.modules.update(inject)
    try:
        mod = types.ModuleType(name)
        exec(compile(src, name, 'exec'), mod.__dict__)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_use():
    src = open(REF + 'genome/use.py').read()
    src = src.replace('knots = [(0, p0), (1, p1)] + zip(knots[4::2], knots[5::2])',
                      'knots = [(0, p0), (1, p1)] + list(zip(knots[4::2], knots[5::2]))')
    return exec_module('ref_use', src, {'spectypes': spectypes, 'specs': specs})


def load_profile(use_mod):
    src = open(REF + 'profile.py').read()
    src = src.replace('from genome.specs import toplevels', 'from specs import toplevels')
    src = src.replace('from genome.use import RefWrapper, SplineWrapper',
                      'from ref_use import RefWrapper, SplineWrapper')
    src = src.replace('import output\n', '')
    src = src.replace('choices=BUILTIN.keys()', 'choices=list(BUILTIN.keys())')
    return exec_module('ref_profile', src, {'specs': specs, 'ref_use': use_mod})


SPLINES = [
    1.5, [2.0, 4.0], [45, -360, -135, 0, 0.3, 150], [0.1, 2.0, 0.9, -1.0],
    [1.0, 0.0, 3.0, 0.5, 0.25, 2.0, 0.5, -1.0, 0.75, 4.0],
    [5, 1, 5, 1, 0.9, 7, 0.1, 3],
]


def main():
    use = load_use()
    out = {'splines': []}
    for sp in SPLINES:
        for scale in (1.0, 2.5):
            se = use.SplineEval(sp, scale)
            ts = [0.0, 0.013, 0.25, 0.5, 0.77, 1.0]
            out['splines'].append({
                'value': sp, 'scale': scale,
                'knots': se.knots.tolist(),
                'at': [[t, se(t), se(t, 1)] for t in ts]})
    prof = load_profile(use)
    import copy
    pristine = copy.deepcopy(prof.BUILTIN)
    out['profile'] = []
    for argv in ([], ['-P', '1080p', '--still'], ['-P', '720p', '--fps=1', '--duration=5', '--shard=5'],
                 ['-P', 'preview', '--start', '3', '--end', '40', '--skip', '2'],
                 ['--duration', '2', '--fps', '12', '--end', '-3']):
        # the reference updates its BUILTIN table in place (profile.py:84-88), which
        # leaks overrides from one call into the next; start every case clean
        prof.BUILTIN.clear()
        prof.BUILTIN.update(copy.deepcopy(pristine))
        args = prof.add_args().parse_args(argv)
        name, p = prof.get_from_args(args)
        gprof = prof.wrap(dict(p), {'type': 'animation', 'camera': {'spp': 1.5},
                                    'time': {'duration': 2, 'frame_width': [1.0, 0.5]}})
        times = prof.enumerate_times(gprof)
        out['profile'].append({
            'argv': argv, 'name': name, 'profile': p,
            'times': [[i, [float(x) for x in t]] for i, t in times][:50], 'ntimes': len(times),
            'spp': gprof.spp(0.5), 'frame_width': gprof.frame_width(0.25),
            'duration': gprof.duration, 'size': [gprof.width, gprof.height]})
    # seeds (code/mwc.py:30-47); the table itself is compared byte for byte elsewhere
    mults = np.fromfile(REF + 'code/primes.bin', dtype='<u4')
    src = open(REF + 'code/mwc.py').read()
    src = src[src.index('def make_seeds'):src.index('mwclib = devlib')]
    ns = {'np': np, 'mults': mults, 'load_mults': lambda: mults}
    exec(compile(src, 'ref_mwc', 'exec'), ns)
    seeds = ns['make_seeds'](4096, host_seed=42)
    out['seeds'] = {'host_seed': 42, 'n': 4096, 'first': seeds[:4].tolist(),
                    'last': seeds[-2:].tolist(),
                    'xor': [int(np.bitwise_xor.reduce(seeds[:, k])) for k in range(3)]}
    # filter host constants (cuburn/filters.py:11-16,132-136)
    f32 = np.float32
    out['filters'] = {'gauss': {}, 'lingam': []}
    for stdev in (1, 0.7, 2.5):
        c = np.exp(np.float32(np.arange(-3, 4)) ** 2 / (-2 * stdev ** 2))
        c = (c / np.sum(c)).astype(np.float32)
        out['filters']['gauss'][str(stdev)] = [float(x) for x in c]
    for gamma, thr in ((4, 0.01), (2.2, 0.0), (3.0, 0.05)):
        gam = f32(1 / gamma)
        lin = f32(thr)
        lingam = f32(lin ** (gam - 1.0) if lin > 0 else 0)
        out['filters']['lingam'].append([gamma, thr, float(gam), float(lin), float(lingam)])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_golden.json')
    with open(path, 'w') as fp:
        json.dump(out, fp, indent=1)
    print('wrote', path)


if __name__ == '__main__':
    main()
