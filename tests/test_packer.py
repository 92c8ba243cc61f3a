"""Genome packing: rows, slots, precalc program, generated source."""
import numpy as np
import pytest

from cuburn_b200 import samples
from cuburn_b200.code import itergen, packer, varlib
from cuburn_b200.genome.use import SplineEval
from cuburn_b200.genome.variations import var_params


def test_structure_g6f():
    g = samples.g6f()
    pk = packer.GenomePacker(g)
    assert pk.xform_ids == ['0', '1', '2', '3', '4', '5'] and pk.has_final
    assert pk.row_paths[:6] == [('xforms', str(i), 'weight') for i in range(6)]
    assert pk.nslots == len(set(pk.slot_names))
    # every xform has its affine, colour and variation slots
    for xid in pk.xform_ids:
        for c in packer.AFFINE_COEFS:
            pk.slot('xforms', xid, 'pre_affine', c)
        pk.slot('xforms', xid, 'color')
    pk.slot('xforms', '2', 'post_affine', 'yo')
    with pytest.raises(KeyError):
        pk.slot('xforms', '1', 'post_affine', 'xx')
    pk.slot('xforms', '1', 'variations', 'julian', 'cn')
    pk.slot('final_xform', 'variations', 'eyefish', 'weight')
    with pytest.raises(KeyError):
        pk.slot('xforms', '5', 'density')        # last xform takes the remainder
    assert pk.param_stride % 4 == 0 and pk.param_stride >= pk.nslots
    prog = pk.program_array()
    assert prog.shape[1] == packer.PROG_WIDTH and prog.dtype == np.int32
    # every slot is written exactly once
    written = []
    for w in prog:
        n = {varlib.OP_AFFINE: 6, varlib.OP_CAMERA: 6, varlib.OP_DENSITY: w[10] - 1,
             varlib.OP_WAVES: 2, varlib.OP_PERSPECTIVE: 3, varlib.OP_CURVE: 2}.get(w[0], 1)
        written += list(range(w[1], w[1] + n))
    # (alignment padding between blocks is never written)
    assert sorted(written) == [i for i, _ in pk.named_slots()]
    assert all(pk.slot(*n.split('.')) == i for i, n in pk.named_slots())
    assert pk.slot('xforms', '3', 'pre_affine', 'xx') % 4 == 0 and pk.slot('camera', 'xx') % 4 == 0


def test_mag_rows_follow_schema():
    pk = packer.GenomePacker(samples.g6f())
    mags = {'.'.join(p): m for p, m in zip(pk.row_paths, pk.row_mag)}
    assert mags['camera.scale'] == 1 and mags['camera.rotation'] == 0
    assert mags['xforms.0.pre_affine.magnitude.x'] == 1
    assert mags['xforms.0.pre_affine.angle'] == 0
    assert mags['xforms.1.variations.julian.power'] == 1
    assert mags['xforms.3.variations.curl.c1'] == 0


def test_pack_matches_normalize_and_defaults():
    g = samples.g6f(animated=True)
    pk = packer.GenomePacker(g)
    times, knots = pk.pack(g)
    assert times.shape == (pk.nrows, 32) and times.dtype == np.float32
    row = pk.row_paths.index(('xforms', '0', 'pre_affine', 'angle'))
    kn = SplineEval.normalize(g['xforms']['0']['pre_affine']['angle'], 1)
    n = kn.shape[1]
    assert np.array_equal(times[row, :n], kn[0].astype(np.float32))
    assert np.array_equal(knots[row, :n], kn[1].astype(np.float32))
    assert np.all(times[row, n:] == np.float32(1e9))
    # a key missing from the genome packs its schema default
    row = pk.row_paths.index(('xforms', '0', 'pre_affine', 'spread'))
    g2 = samples.g6f()
    del g2['xforms']['0']['pre_affine']['spread']
    t2, k2 = packer.GenomePacker(g2).pack(g2)
    assert np.all(k2[row, :4] == 45)
    row = pk.row_paths.index(('final_xform', 'color_speed'))
    assert np.all(knots[row, :4] == 0.0)


def test_too_many_knots_rejected():
    g = samples.g3()
    g['camera']['rotation'] = [0, 0, 10, 0] + [v for i in range(40) for v in (i / 41.0 + 0.01, i)]
    pk = packer.GenomePacker(g)
    with pytest.raises(ValueError):
        pk.pack(g)


def test_unknown_variation_rejected():
    g = samples.g3()
    g['xforms']['0']['variations']['twintrian'] = {'weight': 1}
    with pytest.raises(KeyError):
        packer.GenomePacker(g)


def test_generated_source_shape():
    pk, src = itergen.mkiterlib(samples.g6f())
    assert src.count('__device__ __forceinline__ void apply_xf_') == 7
    assert '#define HAS_FINAL 1' in src and '#include "iter_kernel.cuh"' in src
    # variations are applied in sorted-name order: pre_blur after gaussian_blur,
    # before waves2
    body = src[src.index('void apply_xf_4'):src.index('void apply_xf_5')]
    order = [body.index('var_' + v + '(') for v in ('cylinder', 'gaussian_blur', 'pre_blur', 'waves2')]
    assert order == sorted(order)
    # structure, not numbers, decides the source
    g2 = samples.g6f()
    g2['xforms']['0']['weight'] = 7.5
    assert itergen.mkiterlib(g2)[1] == src
    g3 = samples.g6f()
    g3['xforms']['0']['variations']['swirl'] = {'weight': 0.1}
    assert itergen.mkiterlib(g3)[1] != src


def test_var_args_cover_device_signatures():
    """Argument lists agree with the device library's function signatures."""
    import os, re
    hdr = open(os.path.join(itergen.DEVICE_DIR, 'variations.cuh')).read()
    sigs = dict(re.findall(r'VFN var_(\w+)\(([^)]*)\)', hdr))
    assert set(sigs) == set(var_params)
    for name, sig in sigs.items():
        args = [a.strip() for a in sig.split(',')]
        assert [a.split()[-1].lstrip('&') for a in args[:5]] == ['tx', 'ty', 'w', 'ox', 'oy'], name
        rest = args[5:]
        has_rng = bool(rest) and 'mwc_st' in rest[0]
        assert has_rng == varlib.uses_rng(name), name
        if has_rng:
            rest = rest[1:]
        assert len(rest) == len(varlib.var_args(name)), name


def test_pack_constant_row_fast_path_equals_normalize():
    """Constant rows are written in one vectorised store; the result must be the same
    bits as normalising every row on its own (interp.py:207-232)."""
    from cuburn_b200 import samples
    from cuburn_b200.code import itergen, packer as P
    from cuburn_b200.genome.use import SplineEval
    for gnm in (samples.g3(), samples.g6f(), samples.g6f(animated=True), samples.g24h()):
        pk = itergen.GenomePacker(gnm)
        times, knots = pk.pack(gnm)
        want_t = np.full_like(times, 1e9)
        want_k = np.zeros_like(knots)
        scale = gnm.get('time', {}).get('duration', 1)
        nconst = 0
        for idx, path in enumerate(pk.row_paths):
            attr = gnm
            for name in path:
                if not isinstance(attr, dict) or name not in attr:
                    attr = P.resolve_spec(pk.spec, path).default
                    break
                attr = attr[name]
            nconst += type(attr) in (int, float)
            kn = SplineEval.normalize(attr, scale)
            want_t[idx, :kn.shape[1]] = kn[0]
            want_k[idx, :kn.shape[1]] = kn[1]
        assert nconst > 0
        assert np.array_equal(times.view(np.uint32), want_t.view(np.uint32))
        assert np.array_equal(knots.view(np.uint32), want_k.view(np.uint32))


def test_reduction_placement_follows_the_genome_weight():
    """itergen.is_heavy picks where the histogram reduction sits relative to the
    exchange barrier (measured per genome class, profiles/r01_iter_variants.md)."""
    from cuburn_b200 import samples
    from cuburn_b200.code import itergen
    heavy = {name: itergen.is_heavy(itergen.GenomePacker(make()))
             for name, make in samples.GENOMES.items()}
    assert heavy == {'G3': False, 'G6F': True, 'G24H': True, 'G2M': False}
    for name, make in samples.GENOMES.items():
        src = itergen.generate_source(itergen.GenomePacker(make()), 1)
        assert ('#define RED_BEFORE_PULL %d' % heavy[name]) in src
        forced = itergen.generate_source(itergen.GenomePacker(make()), 1,
                                         extra_defines={'RED_BEFORE_PULL': 0})
        assert forced.count('#define RED_BEFORE_PULL') == 1
