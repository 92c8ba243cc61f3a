"""
Per-genome source generation for the iterate module.

The reference renders tempita templates into one CUDA source per genome
structure (cuburn/code/iter.py:121-155, 547-575).  This generator emits a much
smaller translation unit: one ``apply_xf_<n>`` function per xform that calls
the inline variation library with packed-parameter slots, the weighted choice
(``chaos_step``), an optional ``final_step``, a handful of ``#define``s, and
then includes the fixed kernel skeleton (csrc/device/iter_kernel.cuh).
Numeric values never appear in the source -- they stay data -- so genomes with
the same structure share one compiled module (genome.util.hash).
"""
import os

from . import varlib
from .packer import GenomePacker, AFFINE_COEFS

DEVICE_DIR = os.path.join(os.path.dirname(os.path.dirname(__file__)),
                          'csrc', 'device')
HEADER_NAMES = ('mwc.cuh', 'variations.cuh', 'iter_params.cuh', 'iter_kernel.cuh')

NVRTC_OPTIONS = (
    '--gpu-architecture=sm_100a',
    '--use_fast_math',          # the reference builds with -use_fast_math (code/util.py:96)
    '--std=c++17',
    '-lineinfo',
    '--extra-device-vectorization',
)


def load_headers():
    out = []
    for n in HEADER_NAMES:
        with open(os.path.join(DEVICE_DIR, n)) as fp:
            out.append(fp.read())
    return list(HEADER_NAMES), out


class _Params(object):
    """
    How generated code reads parameter slots.  Stills: ``P[slot]`` with a constant index
    is a constant-bank operand.  Motion blur: the block lives in shared memory, where
    every scalar read is an LDS through the pipe the histogram reductions use
    (profiles/r01_iter_variants.md: +20 % frame time); the slots of an xform are
    contiguous, so they are fetched as aligned float4s -- a quarter of the loads --
    declared right before their first use so the compiler sees short live ranges.
    """
    def __init__(self, vector):
        self.vector = vector
        self.loaded = set()

    def ref(self, slot, lines):
        if not self.vector:
            return 'P[%d]' % slot
        k = slot // 4
        if k not in self.loaded:
            self.loaded.add(k)
            lines.append('    const float4 q%d = P4[%d];' % (k, k))
        return 'q%d.%s' % (k, 'xyzw'[slot % 4])


def _affine_lines(packer, prm, decl, apath, src_x, src_y, dst_x, dst_y):
    s = {c: prm.ref(packer.slot(apath + (c,)), decl) for c in AFFINE_COEFS}
    return [
        '        %s = %s * %s + %s * %s + %s;' % (dst_x, s['xx'], src_x, s['xy'], src_y, s['xo']),
        '        %s = %s * %s + %s * %s + %s;' % (dst_y, s['yx'], src_x, s['yy'], src_y, s['yo']),
    ]


def _xform_function(packer, fname, xpath, variations, has_post, vector=False):
    """``void fname(point_set &pt, mwc_st &rng)``: the xform applied to every point of the
    thread.  Parameter fetches (``decl``) come first and are shared by all points."""
    prm = _Params(vector)
    decl, body = [], ['        float x = pt.x[p], y = pt.y[p], color = pt.c[p];',
                      '        float tx, ty;']
    body += _affine_lines(packer, prm, decl, xpath + ('pre_affine',), 'x', 'y', 'tx', 'ty')
    body.append('        float ox = 0.0f, oy = 0.0f;')
    # variations in sorted-name order (use.py:90-91, iter.py:132-137)
    for v in variations:
        vpath = xpath + ('variations', v)
        args = ['tx', 'ty', prm.ref(packer.slot(vpath + ('weight',)), decl), 'ox', 'oy']
        if varlib.uses_rng(v):
            args.append('rng')
        for kind, name in varlib.var_args(v):
            if kind == 'pre':
                args.append(prm.ref(packer.slot(xpath + ('pre_affine', name)), decl))
            else:
                args.append(prm.ref(packer.slot(vpath + (name,)), decl))
        body.append('        var_%s(%s);' % (v, ', '.join(args)))
    if has_post:
        body.append('        tx = ox; ty = oy;')
        body += _affine_lines(packer, prm, decl, xpath + ('post_affine',), 'tx', 'ty', 'ox', 'oy')
    body.append('        const float csp = %s;' % prm.ref(packer.slot(xpath + ('color_speed',)), decl))
    body.append('        pt.x[p] = ox; pt.y[p] = oy;')
    body.append('        pt.c[p] = color * (1.0f - csp) + %s * csp;'
                % prm.ref(packer.slot(xpath + ('color',)), decl))
    L = ['__device__ __forceinline__ void %s(point_set &pt, mwc_st &rng) {' % fname]
    L += decl
    L += ['#pragma unroll', '    for (int p = 0; p < POINTS; p++) {'] + body + ['    }', '}']
    return '\n'.join(L)


def _choice_chain(packer, prm, names, den_slot, indent, xaos):
    """The cumulative-density if-chain over the xforms (iter.py:263-272; with xaos one
    chain per previous xform, iter.py:236-257).  ``den_slot(i)`` is the slot of the
    cumulative density that ends xform i's interval."""
    L = []
    refs = [prm.ref(den_slot(i), L) for i in range(len(names) - 1)]
    for i, (fname, xpath) in enumerate(names):
        if i < len(names) - 1:
            head = '%sif (sel <= %s) {' % ('else ' if i else '', refs[i])
        else:
            head = 'else {' if len(names) > 1 else '{'
        body = ' %s(pt, rng);' % fname
        if xpath in packer.opacity:
            OL = []
            op = _Params(prm.vector).ref(packer.opacity[xpath], OL)
            body += ''.join(' ' + l.strip() for l in OL)
            body += (' for (int p = 0; p < POINTS; p++) vis[p] = opacity_visible(%s, rng);' % op)
        if xaos:
            body += ' pt.last[0] = %d;' % i
        L.append('%s%s%s }' % (indent, head, body))
    return L


def is_heavy(packer):
    """
    Genomes whose warps drift apart within a round (many variation evaluations, a final
    xform) run faster with the histogram reduction issued before the exchange barrier;
    light genomes, whose warps stay in lockstep, with it after (iter_kernel.cuh).
    Measured on B200 (variation uses over all xforms): G3 (5) and a four-xform cut of
    G6F (8) after; G6F (16 incl. final xform) and G24H (72) before.
    """
    return sum(len(variations) for _, variations, _ in packer.xforms) >= 12


def generate_source(packer, params_const=False, extra_defines=None, acc_packed=False,
                    hot_bins=False, points=1):
    """
    CUDA source of the iterate module for ``packer``'s genome structure.
    ``params_const`` selects the variant whose parameter block lives in
    __constant__ memory (one block per launch: stills) instead of shared memory;
    ``acc_packed`` the packed-u64 accumulation; ``hot_bins`` the variant that keeps
    private shared-memory cells for the bins listed in ``iter_args::hot_tags``;
    ``points`` the number of trajectories a thread carries (xaos genomes: always 1).
    """
    extra_defines = dict(extra_defines or {})
    vector = (not params_const) and extra_defines.pop('PARAMS_VECTOR', '1') != '0'
    points = 1 if packer.xaos else int(extra_defines.pop('POINTS', points))
    out = ['// generated by cuburn_b200.code.itergen -- do not edit']
    for k, v in extra_defines.items():
        out.append('#define %s %s' % (k, v))
    # Eight CTAs of 32 registers for every variant.  (Under the static unit schedule six
    # CTAs of 40 registers were faster for the motion-blur variant -- fewer, older warps
    # suffered less from the schedulers' bias; with units claimed dynamically eight win:
    # G6F 22.8 -> 22.3 ms, G24H 12.0 -> 11.4, profiles/r02_iter_variants_dyn.txt.)
    out += ['#include "mwc.cuh"', '#include "variations.cuh"', '']
    out.append('#define NSLOTS %d' % packer.nslots)
    out.append('#define POINTS %d' % points)
    out.append('#define PARAMS_CONST %d' % (1 if params_const else 0))
    out.append('#define ACC_PACKED %d' % (1 if acc_packed else 0))
    out.append('#define HOT_BINS %d' % (1 if hot_bins else 0))
    out.append('#define XAOS %d' % (1 if packer.xaos else 0))
    out.append('#define HAS_FINAL %d' % (1 if packer.has_final else 0))
    if 'RED_BEFORE_PULL' not in extra_defines:
        out.append('#define RED_BEFORE_PULL %d' % (1 if is_heavy(packer) else 0))
    out.append('#include "iter_params.cuh"')
    out.append('')

    names = []
    for xpath, variations, has_post in packer.xforms:
        if xpath[0] == 'final_xform':
            fname = 'apply_xf_final'
        else:
            fname = 'apply_xf_%d' % len(names)
            names.append((fname, xpath))
        out.append(_xform_function(packer, fname, xpath, variations, has_post, vector))
        out.append('')

    L = ['__device__ __forceinline__ void chaos_step(float sel, point_set &pt, mwc_st &rng, '
         'bool (&vis)[POINTS]) {',
         '    for (int p = 0; p < POINTS; p++) vis[p] = true;']
    ids = [xpath[1] for _, xpath in names]
    if packer.xaos:
        L.append('    switch (pt.last[0]) {')
        for p, pid in enumerate(ids):
            prm = _Params(vector)
            L.append('    %s: {' % ('default' if p == len(ids) - 1 else 'case %d' % p))
            L += _choice_chain(packer, prm, names,
                               lambda i: packer.slot(('xforms', pid, 'chaos_den', ids[i])),
                               '        ', True)
            L.append('        break; }')
        L.append('    }')
    else:
        prm = _Params(vector)
        L += _choice_chain(packer, prm, names,
                           lambda i: packer.slot(('xforms', ids[i], 'density')), '    ', False)
    L.append('}')
    out.append('\n'.join(L))
    out.append('')
    if packer.has_final:
        out.append('__device__ __forceinline__ void final_step(point_set &pt, mwc_st &rng) {\n'
                   '    apply_xf_final(pt, rng);\n}')
        out.append('')
    # camera affine (iter.py:302-311) as six named coefficients
    prm, CL = _Params(vector), []
    cam = [prm.ref(packer.slot('camera', c), CL) for c in AFFINE_COEFS]
    out.append('__device__ __forceinline__ void camera_coefs(float &xx, float &xy, float &xo, '
               'float &yx, float &yy, float &yo) {\n%s    xx = %s; xy = %s; xo = %s; yx = %s; '
               'yy = %s; yo = %s;\n}\n' % (''.join(l + '\n' for l in CL), *cam))
    out.append('#include "iter_kernel.cuh"')
    return '\n'.join(out) + '\n'


def mkiterlib(gnm, params_const=False, acc_packed=False, hot_bins=False, points=1):
    """``(packer, source)`` for a genome (mirrors iter.mkiterlib, iter.py:559-575)."""
    packer = GenomePacker(gnm)
    return packer, generate_source(packer, params_const, acc_packed=acc_packed,
                                   hot_bins=hot_bins, points=points)


def points_per_thread(packer, points):
    """Trajectories per thread a module built with ``points`` really uses."""
    return 1 if packer.xaos else points


def best_points(packer, params_const):
    """
    Trajectories per thread of the production module.  Two points per thread share the
    xform choice, the parameter fetches and the loop overhead.  Measured on B200
    (profiles/r02_iter_variants.md, 1080p, ms per frame, one / two points):

                          G3 (5 variation uses)   G6F (16)        G24H (72)
        still             20.9 / 24.2             24.0 / 23.6     11.7 / 15.8
        motion blur       21.6 / 24.5             27.1 / 24.7     12.4 / 16.5

    That table is from the static unit schedule.  With units claimed dynamically
    (profiles/r02_iter_variants_dyn.txt; every variant at eight CTAs per SM):

        still             20.0 / 19.95            21.6 / 21.0     11.55 / 15.4
        motion blur       20.3 / 19.9             25.4 / 22.3     11.4  / 15.3

    It pays for mid-sized genomes in both variants (G6F: -2.6 % still, -12 % blurred,
    which brings the blurred frame within 3 % of the still), is neutral for light genomes
    and hurts heavy ones (twice the code no longer fits the instruction cache).
    """
    uses = sum(len(variations) for _, variations, _ in packer.xforms)
    return 2 if (not packer.xaos and 12 <= uses <= 32) else 1
