"""
The C-ABI library loads without a GPU and exports every symbol that
include/cuburn_b200.h declares; the per-genome module compiles for sm_100a with
NVRTC (a compiler: no GPU needed) for every variation.  No compute calls here.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'cuburn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(built):
    from cuburn_b200 import _native as N
    syms = _header_symbols()
    assert len(syms) >= 45
    lib = N.lib()
    for s in syms:
        assert hasattr(lib, s), s
    # the ctypes table binds exactly the header's functions
    assert sorted(N.EXPORTS) == syms
    out = subprocess.run(['nm', '-D', '--defined-only', N.LIB_PATH], stdout=subprocess.PIPE,
                         text=True).stdout
    exported = set(re.findall(r' T (cb_\w+)', out))
    assert set(syms) <= exported


def test_library_is_sm100a_only(built):
    from cuburn_b200 import _native as N
    out = subprocess.run(['cuobjdump', '-lelf', N.LIB_PATH], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_no_gpu_calls_fail_loudly(built):
    """Without a device the library reports an error instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from cuburn_b200 import _native as N
    with pytest.raises(N.NativeError):
        N.init(0)
    assert N.calc_dim(1920, 1080).astride == 1952      # pure host helper still works
    d = N.calc_dim(1280, 720)
    assert (d.aw, d.ah, d.astride) == (1304, 752, 1312)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cuburn_b200/ may reference it."""
    pkg = os.path.join(ROOT, 'cuburn_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f
                assert 'liboracle' not in text, f


def _variation_names():
    from cuburn_b200.genome.variations import VAR_TABLE
    return [n for _, n, _ in VAR_TABLE]


def test_nvrtc_compiles_every_variation(built):
    """All 95 variations in 5 generated modules (19 xforms each, plus final/post)."""
    from cuburn_b200 import _native as N
    from cuburn_b200.code import itergen
    names = _variation_names()
    hn, hs = itergen.load_headers()
    for k in range(0, len(names), 19):
        chunk = names[k:k + 19]
        g = {'type': 'animation', 'camera': {'scale': 0.3},
             'xforms': {str(i): {'weight': 1, 'variations': {v: {'weight': 0.5},
                                                             'linear': {'weight': 0.5}}}
                        for i, v in enumerate(chunk)}}
        g['xforms']['0']['post_affine'] = {'angle': 10}
        g['final_xform'] = {'variations': {chunk[0]: {'weight': 1}}}
        pk, src = itergen.mkiterlib(g)
        mod = N.Module(src, 'vars_%d.cu' % k, hs, hn, itergen.NVRTC_OPTIONS)
        cubin = mod.cubin
        assert cubin[:4] == b'\x7fELF' and len(cubin) > 4096


def test_sample_genome_modules_use_vector_red(built, tmp_path):
    """SASS evidence: 16-byte float4 reductions and SFU intrinsics; the 32-register
    cap of the still variant (64 warps/SM, measured fastest) may cost a few bytes of
    spill, no more; the motion-blur variant runs six CTAs of at most 40 registers."""
    from cuburn_b200 import _native as N, samples
    from cuburn_b200.code import itergen
    hn, hs = itergen.load_headers()
    pk, src = itergen.mkiterlib(samples.g6f())
    blur = N.Module(src, 'g6f_blur.cu', hs, hn, itergen.NVRTC_OPTIONS)
    pb = tmp_path / 'g6f_blur.cubin'
    pb.write_bytes(blur.cubin)
    usage = subprocess.run(['cuobjdump', '-res-usage', str(pb)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True).stdout
    m = re.search(r'Function cb_iter:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', usage)
    assert m and int(m.group(1)) <= 40 and int(m.group(2)) == 0 and int(m.group(4)) == 0
    pk, src = itergen.mkiterlib(samples.g6f(), params_const=True)
    mod = N.Module(src, 'g6f.cu', hs, hn, itergen.NVRTC_OPTIONS)
    p = tmp_path / 'g6f.cubin'
    p.write_bytes(mod.cubin)
    sass = subprocess.run(['cuobjdump', '-sass', str(p)], stdout=subprocess.PIPE, text=True).stdout
    assert 'REDG.E.ADD.F32x4' in sass
    assert 'MUFU.SIN' in sass and 'MUFU.RCP' in sass
    usage = subprocess.run(['cuobjdump', '-res-usage', str(p)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True).stdout
    m = re.search(r'Function cb_iter:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', usage)
    assert m and int(m.group(2)) <= 64 and int(m.group(4)) == 0 and int(m.group(1)) <= 32


def test_every_module_variant_compiles(built, tmp_path):
    """The iterate module exists in variants (still / motion blur) x (float4 / packed cells)
    x (hot-bin cells) x (evict-last reductions) that RenderManager._iter picks per frame;
    the rarely taken ones must build too -- here for a plain genome and for one with xaos
    and opacity -- without local-memory arrays; register spill stays small under the
    32-register cap (0 B for the plain genome; the xaos motion-blur variant, with its
    per-previous-xform chains, spills 72 B)."""
    import itertools
    from cuburn_b200 import samples, render
    xaos = samples.g3()
    xaos['xforms']['0']['opacity'] = 0.35
    xaos['xforms']['0']['chaos'] = {'0': 0.25, '1': 2.0}
    xaos['xforms']['2']['chaos'] = {'1': 0.5, '2': 0.1}
    seen = set()
    for gnm in (samples.g3(), xaos):
        for const, packed, hot, big in itertools.product((False, True), repeat=4):
            if packed and hot:
                continue                    # never combined (_hot_decision)
            pk, src, mod = render.Renderer.compile(gnm, params_const=const, acc_packed=packed,
                                                   hot_bins=hot, big_grid=big)
            assert mod.cubin[:4] == b'\x7fELF'
            seen.add(src)
            p = tmp_path / 'variant.cubin'
            p.write_bytes(mod.cubin)
            usage = subprocess.run(['cuobjdump', '-res-usage', str(p)], stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True).stdout
            m = re.search(r'Function cb_iter:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)',
                          usage)
            assert m and int(m.group(4)) == 0 and int(m.group(2)) <= 96, (const, packed, hot, big,
                                                                         m and m.groups())
    assert len(seen) == 24                  # every combination is its own source


def test_compile_error_carries_log(built):
    from cuburn_b200 import _native as N
    with pytest.raises(N.CompileError) as e:
        N.Module('__global__ void k() { undeclared(); }', 'oops.cu', [], [],
                 ['--gpu-architecture=sm_100a'])
    assert 'undeclared' in str(e.value)
