"""
ORACLE (test infrastructure only) -- compile the reference's OWN device code for
the CPU, straight from the sources under /root/reference.

cuburn's kernels are CUDA C held in Python strings.  The per-variation bodies
(cuburn/code/variations.py), the Catmull-Rom evaluator (cuburn/code/interp.py:284-367
with the knot search of cuburn/code/util.py:210-231), the YUV helpers
(cuburn/code/color.py) and the dither clamp (cuburn/code/output.py:7-13) are plain
C once the template placeholders are replaced by variables and a 20-line shim
supplies `mwc_st`, `mwc_next*` and the float constants.  This script extracts those
strings by executing the reference modules with stub `util` / `tempita`, writes a
generated translation unit to oracle/_ref/ref_kernels.cpp and compiles it with g++
into oracle/_ref/libref_kernels.so (+ ref_kernels.json describing the argument
order of every variation).  Nothing is copied into the repository: oracle/_ref/ is
git-ignored, built artefacts only.  Tests use the library to pin oracle/chaos.c
and oracle/flame_ref.py against the reference's code; on the GPU box (no
/root/reference) the prebuilt library that travelled with the snapshot is used.
"""
import json
import os
import re
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_ref')
LIB = os.path.join(OUT_DIR, 'libref_kernels.so')
META = os.path.join(OUT_DIR, 'ref_kernels.json')
REF = '/root/reference/cuburn/code/'


def available():
    return os.path.exists(LIB) and os.path.exists(META)


class _Template(object):
    def __init__(self, content, name=None, namespace=None, **kw):
        self.content, self.name = content, name

    def substitute(self, *a, **kw):
        raise RuntimeError('templates are not rendered here')


def _exec_reference(fname, stubs):
    src = open(REF + fname).read()
    if 'if __name__ ==' in src:            # Python 2 self-test blocks do not parse
        src = src[:src.index('if __name__ ==')]
    mod = types.ModuleType('ref_' + fname[:-3])
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        exec(compile(src, REF + fname, 'exec'), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _stub_util():
    u = types.ModuleType('util')
    u.Template = _Template

    def devlib(deps=(), headers='', decls='', defs=''):
        return types.SimpleNamespace(deps=deps, headers=headers, decls=decls, defs=defs)
    u.devlib = devlib
    u.assemble_code = lambda *a: ''
    u.binsearchlib = u.ringbuflib = None
    u.snd = lambda ab: ab[1]
    u.DEFAULT_SEARCH_ROUNDS = 5
    return u


SHIM = r'''
#include <math.h>
#include <stdint.h>
#include <string.h>
typedef struct { uint32_t mul, state, carry; } mwc_st;
static inline uint32_t mwc_next(mwc_st &st) {
    uint64_t t = (uint64_t)st.mul * st.state + st.carry;
    st.state = (uint32_t)t; st.carry = (uint32_t)(t >> 32); return st.state;
}
static inline float mwc_next_01(mwc_st &st) { return mwc_next(st) * (1.0f / 4294967296.0f); }
static inline float mwc_next_11(mwc_st &st) { return (float)(int32_t)mwc_next(st) * (1.0f / 2147483648.0f); }
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
static inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
#define __device__
#define __noinline__
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float max(float a, float b) { return a > b ? a : b; }
'''


def _constants_block():
    """The #undef/#define M_* block of the reference stdlib (code/util.py:145-172)."""
    src = open(REF + 'util.py').read()
    start = src.index('#undef M_E')
    end = src.index('#define bfe(')
    return src[start:end]


def generate():
    util_stub = _stub_util()
    varmod = _exec_reference('variations.py', {'util': util_stub})
    meta = {'variations': {}}
    parts = [SHIM, _constants_block()]

    for name, tmpl in sorted(varmod.var_code.items()):
        code = tmpl.content.replace('{{precalc_fun(pv, px)}}', '')
        pv, pa = [], []

        def repl(m):
            expr = m.group(1).strip()
            if expr.startswith('pv.'):
                n = expr[3:]
                if n not in pv:
                    pv.append(n)
                return 'pv_' + n
            if expr.startswith('px.pre_affine.'):
                n = expr[len('px.pre_affine.'):]
                if n not in pa:
                    pa.append(n)
                return 'pa_' + n
            raise ValueError('unexpected placeholder %r in %s' % (expr, name))
        code = re.sub(r'\{\{(.*?)\}\}', repl, code)
        decl = ''.join('    float pv_%s = pv[%d];\n' % (n, i) for i, n in enumerate(pv))
        decl += ''.join('    float pa_%s = pa[%d];\n' % (n, i) for i, n in enumerate(pa))
        parts.append('''
extern "C" void ref_var_%s(float *txs, float *tys, float w, float *oxs, float *oys,
                           uint32_t *seeds, int n, const float *pv, const float *pa) {
%s    for (int i__ = 0; i__ < n; i__++) {
        mwc_st rctx = {seeds[3*i__], seeds[3*i__+1], seeds[3*i__+2]};
        float tx = txs[i__], ty = tys[i__], ox = oxs[i__], oy = oys[i__];
        {
%s
        }
        txs[i__] = tx; tys[i__] = ty; oxs[i__] = ox; oys[i__] = oy;
        seeds[3*i__+1] = rctx.state; seeds[3*i__+2] = rctx.carry;
    }
}
''' % (name, decl, code))
        meta['variations'][name] = {'pv': pv, 'pa': pa}

    # knot search + Catmull-Rom (code/util.py:210-231, code/interp.py:284-367)
    rounds = 5
    search = ['static int bitwise_binsearch(const float *hay, float needle) {', '    int lo = 0;']
    for i in range(rounds - 1, -1, -1):
        search.append('    if (needle > hay[lo + %d]) lo += %d;' % (1 << i, 1 << i))
    search += ['    return lo;', '}']
    parts.append('\n'.join(search))
    interp = _exec_reference('interp.py', {
        'util': util_stub, 'numpy': __import__('numpy'),
        'cuburn': types.ModuleType('cuburn'), 'cuburn.genome': types.ModuleType('cuburn.genome'),
        'cuburn.genome.specs': types.ModuleType('specs'),
        'cuburn.genome.util': types.SimpleNamespace(resolve_spec=None),
        'cuburn.genome.use': types.SimpleNamespace(Wrapper=object, SplineEval=None),
        'color': types.SimpleNamespace(yuvlib=None), 'mwc': types.SimpleNamespace(mwclib=None)})
    parts.append(interp.catmullromlib.decls)
    parts.append(interp.catmullromlib.defs)
    parts.append('''
extern "C" void ref_catmull_rom(const float *times, const float *knots, const float *ts,
                                float *out, int n, int mag) {
    for (int i = 0; i < n; i++)
        out[i] = mag ? catmull_rom_mag(times, knots, ts[i]) : catmull_rom(times, knots, ts[i]);
}
''')
    # colour helpers (code/color.py:12-42)
    color = _exec_reference('color.py', {'util': util_stub, 'numpy': __import__('numpy')})
    parts.append(color.yuvlib.decls)
    parts.append('''
extern "C" void ref_rgb2yuv(const float *rgb, float *yuv, int n) {
    for (int i = 0; i < n; i++) {
        float3 r = rgb2yuv(make_float3(rgb[3*i], rgb[3*i+1], rgb[3*i+2]));
        yuv[3*i] = r.x; yuv[3*i+1] = r.y; yuv[3*i+2] = r.z;
    }
}
extern "C" void ref_yuvo2rgb(float *pix, int n) {
    for (int i = 0; i < n; i++) {
        float4 p = make_float4(pix[4*i], pix[4*i+1], pix[4*i+2], pix[4*i+3]);
        yuvo2rgb(p);
        pix[4*i] = p.x; pix[4*i+1] = p.y; pix[4*i+2] = p.z; pix[4*i+3] = p.w;
    }
}
''')
    return '\n'.join(parts), meta


def build(force=False):
    """Build oracle/_ref/libref_kernels.so if the reference tree is mounted."""
    if not os.path.isdir(REF):
        return LIB if available() else None
    if available() and not force and \
            os.path.getmtime(LIB) >= os.path.getmtime(os.path.abspath(__file__)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    src, meta = generate()
    cpp = os.path.join(OUT_DIR, 'ref_kernels.cpp')
    with open(cpp, 'w') as fp:
        fp.write(src)
    cmd = ['g++', '-O1', '-ffp-contract=off', '-fno-fast-math', '-w', '-shared', '-fPIC',
           '-o', LIB, cpp, '-lm']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('reference kernels did not compile:\n' + r.stdout[-4000:])
    os.remove(cpp)          # keep only the built artefacts
    with open(META, 'w') as fp:
        json.dump(meta, fp, indent=1)
    return LIB


_lib = None


def lib():
    """ctypes handle of the reference-code library (None when it cannot be had)."""
    global _lib
    if _lib is None:
        import ctypes
        path = build()
        if path is None:
            return None
        _lib = ctypes.CDLL(path)
    return _lib


def meta():
    with open(META) as fp:
        return json.load(fp)


def ref_variation(name, txs, tys, w, seeds, pv, pa):
    """Run the reference's body of variation `name` on arrays of input points.
    Returns (tx, ty, ox, oy, seeds) after the call (ox, oy start at zero)."""
    import ctypes
    import numpy as np
    L = lib()
    f32 = np.float32
    txs, tys = np.array(txs, f32), np.array(tys, f32)
    oxs, oys = np.zeros_like(txs), np.zeros_like(txs)
    seeds = np.array(seeds, np.uint32)
    pv = np.array(list(pv) + [0.0], f32)
    pa = np.array(list(pa) + [0.0], f32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    getattr(L, 'ref_var_' + name)(p(txs), p(tys), ctypes.c_float(w), p(oxs), p(oys), p(seeds),
                                  ctypes.c_int(txs.size), p(pv), p(pa))
    return txs, tys, oxs, oys, seeds


def ref_catmull_rom(times, knots, ts, mag=False):
    import ctypes
    import numpy as np
    L = lib()
    times, knots, ts = (np.ascontiguousarray(a, np.float32) for a in (times, knots, ts))
    out = np.empty_like(ts)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.ref_catmull_rom(p(times), p(knots), p(ts), p(out), ctypes.c_int(ts.size),
                      ctypes.c_int(1 if mag else 0))
    return out


def ref_rgb2yuv(rgb):
    import ctypes
    import numpy as np
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.empty_like(rgb)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib().ref_rgb2yuv(p(rgb), p(out), ctypes.c_int(rgb.shape[0]))
    return out


def ref_yuvo2rgb(pix):
    import ctypes
    import numpy as np
    pix = np.array(pix, np.float32)
    lib().ref_yuvo2rgb(pix.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(pix.shape[0]))
    return pix


if __name__ == '__main__':
    print(build(force=True))
