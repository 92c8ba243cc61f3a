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


def _sample_genomes():
    """Genome structures the iterate functions are rendered for (the synthetic test
    genomes: BASELINE configs 1-5 plus one with a post-affine and no final xform)."""
    import sys
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from cuburn_b200 import samples
    return {'g3': samples.g3(), 'g6f': samples.g6f(), 'g24h': samples.g24h()}


class _Template(object):
    """Stand-in for tempita.Template, enough of it for the reference's device code:
    ``{{expr}}`` (``None`` renders as nothing), ``{{for a, b in expr}} ... {{endfor}}``,
    ``{{if expr}} ... {{elif expr}} ... {{else}} ... {{endif}}`` and ``{{py: stmt}}``,
    nested to any depth.  Expressions see the substitution arguments, the template's
    namespace and the variables bound by enclosing loops; ``locals()`` inside an
    expression is that combined scope (iter.py renders the variation bodies with it)."""
    _TOKEN = re.compile(r'\{\{(.*?)\}\}', re.S)

    def __init__(self, content, name=None, namespace=None, **kw):
        self.content, self.name, self.namespace = content, name, namespace or {}

    def _parse(self):
        pos, stack = 0, [[]]         # stack of open blocks; a block is a list of nodes
        heads = []
        for m in self._TOKEN.finditer(self.content):
            if m.start() > pos:
                stack[-1].append(('text', self.content[pos:m.start()]))
            pos = m.end()
            tok = m.group(1).strip()
            if tok.startswith('for '):
                var, expr = tok[4:].split(' in ', 1)
                node = ('for', [v.strip() for v in var.split(',')], expr, [])
                stack[-1].append(node)
                stack.append(node[3])
                heads.append(node)
            elif tok.startswith('if '):
                node = ('if', [(tok[3:], [])])
                stack[-1].append(node)
                stack.append(node[1][0][1])
                heads.append(node)
            elif tok.startswith('elif ') or tok == 'else':
                node = heads[-1]
                assert node[0] == 'if', 'stray %s in template %s' % (tok, self.name)
                stack.pop()
                node[1].append((tok[5:] if tok != 'else' else 'True', []))
                stack.append(node[1][-1][1])
            elif tok in ('endfor', 'endif'):
                assert heads and heads[-1][0] == tok[3:], 'unbalanced %s in %s' % (tok, self.name)
                heads.pop()
                stack.pop()
            elif tok.startswith('py:'):
                stack[-1].append(('py', tok[3:].strip()))
            else:
                stack[-1].append(('expr', tok))
        if pos < len(self.content):
            stack[-1].append(('text', self.content[pos:]))
        assert len(stack) == 1, 'unclosed block in template %s' % self.name
        return stack[0]

    def _render(self, nodes, ns):
        out = []
        for node in nodes:
            kind = node[0]
            if kind == 'text':
                out.append(node[1])
            elif kind == 'expr':
                val = eval(node[1], ns)
                out.append('' if val is None else str(val))
            elif kind == 'py':
                exec(node[1], ns)
            elif kind == 'for':
                _, names, expr, body = node
                for item in list(eval(expr, ns)):
                    if len(names) == 1:
                        ns[names[0]] = item
                    else:
                        for n, v in zip(names, item):
                            ns[n] = v
                    out.append(self._render(body, ns))
            elif kind == 'if':
                for cond, body in node[1]:
                    if eval(cond, ns):
                        out.append(self._render(body, ns))
                        break
        return ''.join(out)

    def substitute(self, *a, **kw):
        ns = dict(self.namespace)
        for d in a:
            if isinstance(d, dict):
                ns.update(d)
        ns.update(kw)
        return self._render(self._parse(), ns)


class _FakeNode(object):
    """Stands in for the packer views the precalc templates are rendered against:
    reading ``node.a.b`` yields the C identifier ``in_<prefix>_a_b``, ``_set('x')``
    yields ``out_<prefix>_x``; both are recorded in order of first use."""
    def __init__(self, prefix, reg=None, code=None):
        self._prefix = prefix
        self._reg = reg if reg is not None else {'in': [], 'out': []}
        self._codes = code if code is not None else []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _FakeNode(self._prefix + '_' + name, self._reg, self._codes)

    def _precalc(self):
        return self

    def _set(self, name):
        ident = 'out_%s_%s' % (self._prefix, name)
        if ident not in self._reg['out']:
            self._reg['out'].append(ident)
        return ident

    def _code(self, code):
        self._codes.append(code)

    def __str__(self):
        ident = 'in_' + self._prefix
        if ident not in self._reg['in']:
            self._reg['in'].append(ident)
        return ident


class _FakeXforms(object):
    """``cp.xforms`` of the packer view for a genome with xforms named `names`, in
    the reference's iteration order: a Python 2 dict-like whose ``keys()`` slices."""
    def __init__(self, node, names):
        self._node, self._names = node, list(names)

    def __iter__(self):
        return iter(self._names)

    def keys(self):
        return list(self._names)

    def __getitem__(self, name):
        return getattr(self._node, name)


class _View(object):
    """
    Stand-in for the packer view (PackerWrapper / PrecalcWrapper, code/interp.py:28-123) that
    the iterate templates are rendered against.  A node is a genome path; coercing it to a
    string yields the C identifier of that parameter: ``out_<path>`` when a precalc hunk
    has produced it with ``_set`` (rendered earlier in the same template, as in the
    reference), else ``in_<path>``, an input of the generated function.  Container
    emulation follows genome/use.py:90-99 (sorted keys).
    """
    def __init__(self, path, reg, present):
        self.__dict__.update(_path=tuple(path), _reg=reg, _present=present)

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _View(self._path + (name,), self._reg, self._present)

    def __getitem__(self, name):
        return getattr(self, str(name))

    def __contains__(self, name):
        return self._path + (str(name),) in self._present

    def keys(self):
        n = len(self._path)
        return sorted(set(p[n] for p in self._present if len(p) > n and p[:n] == self._path))

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def _precalc(self):
        return self

    def _set(self, name):
        path = self._path + (name,)
        if path not in self._reg['out']:
            self._reg['out'].append(path)
        return 'out_' + '_'.join(path)

    def _code(self, code):
        self._reg['codes'].append(code)

    def __str__(self):
        if self._path in self._reg['out']:
            return 'out_' + '_'.join(self._path)
        if self._path not in self._reg['in']:
            self._reg['in'].append(self._path)
        return 'in_' + '_'.join(self._path)


def _present_paths(root, doc):
    out = set()

    def walk(path, d):
        out.add(path)
        if isinstance(d, dict):
            for k, v in d.items():
                walk(path + (str(k),), v)
    walk((root,), doc)
    return out


def _iterate_functions(itermod, tag, gnm):
    """
    The reference's own text of the per-xform functions (``apply_xf_<id>``,
    code/iter.py:121-149, with the variation bodies and affine precalcs it pulls in), the
    xform choice chain (iter.py:263-272) and the final xform + camera + `trunca` + bounds
    test (iter.py:302-317), rendered for genome ``gnm`` and wrapped for the CPU:

      ref_<tag>_setup(in[])                 load the genome's direct parameters, run the precalcs
      ref_<tag>_apply(xf, x[], y[], c[], seeds[], n)      xf = index in sorted order, -1 = final
      ref_<tag>_choose(sel, x[], y[], c[], seeds[], last[], n)
      ref_<tag>_bin(x[], y[], c[], seeds[], idx[], cc[], n)   idx = -1 when rejected
    Returns (C++ source, [input paths in order]).
    """
    reg = {'in': [], 'out': [], 'codes': []}
    present = _present_paths('cp', gnm)
    cp = _View(('cp',), reg, present)
    xids = cp.xforms.keys()
    bodies = [itermod.iter_xf_body(cp, xid, cp.xforms[xid]) for xid in xids]
    has_final = 'final_xform' in gnm
    if has_final:
        bodies.append(itermod.iter_xf_body(cp, 'final', cp.final_xform))
    itermod.precalc_camera(cp.camera)
    itermod.precalc_densities(cp)
    body = itermod.iter_body_code
    a0 = body.index("float xfsel = cosel[threadIdx.y];")
    a1 = body.index("// Rotate points between threads.")
    chain = '{{py:xk = cp.xforms.keys()}}' + body[a0 + len("float xfsel = cosel[threadIdx.y];"):a1]
    b0 = body.index("float cx, cy, cc;")
    b1 = body.index("uint32_t hotspot_i = (")
    binseg = body[b0:b1].replace('continue;', 'return -1;')
    ns = dict(itermod.__dict__, cp=cp)
    chain_c = _Template(chain, tag + '_chain').substitute(ns)
    bin_c = _Template(binseg, tag + '_bin').substitute(ns)
    ins, outs = list(reg['in']), list(reg['out'])
    L = ['namespace ref_%s {' % tag]
    L += ['static float in_%s;' % '_'.join(p) for p in ins]
    L += ['static float out_%s;' % '_'.join(p) for p in outs]
    L += [b.replace('__device__', 'static') for b in bodies]
    L.append('static int choose(float xfsel, float &x, float &y, float &color, mwc_st &rctx) {\n'
             '    int last_xf_used = 0;\n%s\n    return last_xf_used;\n}' % chain_c)
    L.append('static int bin(float x, float y, float color, mwc_st &rctx, float &cc_out) {\n%s\n'
             '    cc_out = cc;\n    return (int)(iy * acc_size.astride + ix);\n}' % bin_c)
    L.append('static void setup(const float *in) {')
    L += ['    in_%s = in[%d];' % ('_'.join(p), i) for i, p in enumerate(ins)]
    L += ['    {\n%s\n    }' % c for c in reg['codes']]
    L.append('}')
    L.append('}  // namespace')
    calls = ''.join('        case %d: ref_%s::apply_xf_%s(x[i], y[i], c[i], r); break;\n' % (k, tag, xid)
                    for k, xid in enumerate(xids))
    if has_final:
        calls += '        case -1: ref_%s::apply_xf_final(x[i], y[i], c[i], r); break;\n' % tag
    L.append("""
extern "C" void ref_%(t)s_setup(const float *in, int width, int awidth, int aheight, int astride) {
    acc_size.width = width; acc_size.awidth = awidth; acc_size.aheight = aheight;
    acc_size.astride = astride;
    ref_%(t)s::setup(in);
}
extern "C" void ref_%(t)s_apply(int xf, float *x, float *y, float *c, uint32_t *seeds, int n) {
    for (int i = 0; i < n; i++) {
        mwc_st r = {seeds[3*i], seeds[3*i+1], seeds[3*i+2]};
        switch (xf) {
%(calls)s        }
        seeds[3*i+1] = r.state; seeds[3*i+2] = r.carry;
    }
}
extern "C" void ref_%(t)s_choose(float sel, float *x, float *y, float *c, uint32_t *seeds,
                                 int *last, int n) {
    for (int i = 0; i < n; i++) {
        mwc_st r = {seeds[3*i], seeds[3*i+1], seeds[3*i+2]};
        last[i] = ref_%(t)s::choose(sel, x[i], y[i], c[i], r);
        seeds[3*i+1] = r.state; seeds[3*i+2] = r.carry;
    }
}
extern "C" void ref_%(t)s_bin(const float *x, const float *y, const float *c, uint32_t *seeds,
                              int *idx, float *cc, int n) {
    for (int i = 0; i < n; i++) {
        mwc_st r = {seeds[3*i], seeds[3*i+1], seeds[3*i+2]};
        idx[i] = ref_%(t)s::bin(x[i], y[i], c[i], r, cc[i]);
        seeds[3*i+1] = r.state; seeds[3*i+2] = r.carry;
    }
}
""" % dict(t=tag, calls=calls))
    return '\n'.join(L), ['.'.join(p[1:]) for p in ins]


def _precalc_function(cname, codes, reg, extra_args=''):
    """Wrap rendered precalc hunks into  void cname(const float *in, float *out, ...)."""
    body = ''.join('    float %s = in[%d];\n' % (n, i) for i, n in enumerate(reg['in']))
    body += ''.join('    float %s = 0.0f;\n' % n for n in reg['out'])
    body += ''.join('    {\n%s\n    }\n' % c for c in codes)
    body += ''.join('    out[%d] = %s;\n' % (i, n) for i, n in enumerate(reg['out']))
    return 'extern "C" void %s(const float *in, float *out%s) {\n%s}\n' % (cname, extra_args, body)


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


CUDA_SHIM = r'''
struct uint2 { uint32_t x, y; };
struct float2 { float x, y; };
struct uchar3 { unsigned char x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort3 { unsigned short x, y, z; };
struct ushort4 { unsigned short x, y, z, w; };
static inline uchar3 make_uchar3(float x, float y, float z) { uchar3 r = {(unsigned char)x, (unsigned char)y, (unsigned char)z}; return r; }
static inline uchar4 make_uchar4(float x, float y, float z, float w) { uchar4 r = {(unsigned char)x, (unsigned char)y, (unsigned char)z, (unsigned char)w}; return r; }
static inline ushort3 make_ushort3(float x, float y, float z) { ushort3 r = {(unsigned short)x, (unsigned short)y, (unsigned short)z}; return r; }
static inline ushort4 make_ushort4(float x, float y, float z, float w) { ushort4 r = {(unsigned short)x, (unsigned short)y, (unsigned short)z, (unsigned short)w}; return r; }
struct idx3 { int x, y, z; };
static idx3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __shared__ static
#define __constant__ static
#define __restrict__
static inline void __syncthreads() {}
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline void scale_float4(float4 &pix, float scale) { pix.x *= scale; pix.y *= scale; pix.z *= scale; pix.w *= scale; }
// point-sampled 2-D texture reference with unnormalised coordinates: clamp to edge
enum { cudaTextureType2D = 2, cudaSurfaceType2D = 2 };
template <typename T, int D> struct texture { const T *ptr; int w, h; };
template <typename T> static inline T tex2D(texture<T, cudaTextureType2D> &t, float x, float y) {
    int xi = (int)floorf(x), yi = (int)floorf(y);
    xi = xi < 0 ? 0 : (xi >= t.w ? t.w - 1 : xi);
    yi = yi < 0 ? 0 : (yi >= t.h ? t.h - 1 : yi);
    return t.ptr[yi * t.w + xi];
}
template <typename T, int D> struct surface { uint2 *ptr; int wbytes; };
static inline void surf2Dwrite(uint2 v, surface<void, cudaSurfaceType2D> &s, int xbytes, int y) {
    s.ptr[(y * s.wbytes + xbytes) / 8] = v;
}
// ring buffer slot claim (code/util.py:329-335) for a serial emulation in which
// thread 0 of a block runs first
typedef struct { uint32_t head; uint32_t tail; } ringbuf;
static uint32_t rb_idx;
static uint32_t rb_incr(uint32_t &rb_base, int tidx) {
    if (threadIdx.y == 0 && threadIdx.x == 0) rb_idx = 256 * ((rb_base++) & 1023);
    return rb_idx + tidx;
}
// run a kernel body for every thread of a (bx, by) grid of (tx, ty) blocks, serially
template <typename F> static void run_grid(int gx, int gy, int tx, int ty, int passes, F f) {
    gridDim.x = gx; gridDim.y = gy; gridDim.z = 1; blockDim.x = tx; blockDim.y = ty; blockDim.z = 1;
    for (int by = 0; by < gy; by++) for (int bx = 0; bx < gx; bx++)
        for (int pass = 0; pass < passes; pass++)        // pass 0 fills __shared__ tables
            for (int y = 0; y < ty; y++) for (int x = 0; x < tx; x++) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0;
                threadIdx.x = x; threadIdx.y = y; threadIdx.z = 0;
                f();
            }
}
'''

KERNEL_ENTRIES = r'''
static void bind4(const float *p, int w, int h) { chan4_src.ptr = (const float4 *)p; chan4_src.w = w; chan4_src.h = h; }
static void bind1(const float *p, int w, int h) { chan1_src.ptr = p; chan1_src.w = w; chan1_src.h = h; }
#define GRID2(W, H, PASSES, CALL) run_grid((W) / 32, (H) / 8, 32, 8, PASSES, [&]() { CALL; })
extern "C" {
void ref_set_gauss(const float *c) { for (int i = 0; i < 7; i++) gauss_coefs[i] = c[i]; }
void ref_yuv_to_rgb(float *dst, const float *src, int w, int h) { GRID2(w, h, 1, yuv_to_rgb((float4 *)dst, (const float4 *)src)); }
void ref_logscale(float *dst, const float *src, float k1, float k2, int w, int h) { GRID2(w, h, 1, logscale((float4 *)dst, (const float4 *)src, k1, k2)); }
void ref_logencode(float *dst, const float *src, float degamma, int w, int h) { GRID2(w, h, 1, logencode((float4 *)dst, (const float4 *)src, degamma)); }
void ref_den_blur(float *dst, const float *src4, int pattern, int up, int w, int h) { bind4(src4, w, h); GRID2(w, h, 1, den_blur(dst, pattern, up)); }
void ref_den_blur_1c(float *dst, const float *src1, int pattern, int up, int w, int h) { bind1(src1, w, h); GRID2(w, h, 1, den_blur_1c(dst, pattern, up)); }
void ref_full_blur(float *dst, const float *src4, int pattern, int up, int w, int h) { bind4(src4, w, h); GRID2(w, h, 1, full_blur((float4 *)dst, pattern, up)); }
void ref_bilateral(float *dst, const float *src4, const float *blur1, int pattern, int radius, float sstd, float cstd,
                   float dstd, float dpow, float gspeed, int w, int h) {
    bind4(src4, w, h); bind1(blur1, w, h);
    GRID2(w, h, 2, bilateral((float4 *)dst, pattern, radius, sstd, cstd, dstd, dpow, gspeed));
}
void ref_apply_gamma(float *dst, float *src4, float gamma, int w, int h) { GRID2(w, h, 1, apply_gamma(dst, (float4 *)src4, gamma)); }
void ref_haloclip(float *pix, const float *den, float gm1, int w, int h) { GRID2(w, h, 1, haloclip((float4 *)pix, den, gm1)); }
void ref_apply_gamma_full_hi(float *dst, float *src, float gm1, int w, int h) { GRID2(w, h, 1, apply_gamma_full_hi((float4 *)dst, (float4 *)src, gm1)); }
void ref_smearclip(float *pix, const float *smear, float gm1, float lin, float lingam, int w, int h) { GRID2(w, h, 1, smearclip((float4 *)pix, (const float4 *)smear, gm1, lin, lingam)); }
void ref_plainclip(float *pix, float gm1, float lin, float lingam, float brightness, int w, int h) { GRID2(w, h, 1, plainclip((float4 *)pix, gm1, lin, lingam, brightness)); }
void ref_colorclip(float *pix, float vib, float hipow, float gamma, float lin, float lingam, int w, int h) { GRID2(w, h, 1, colorclip((float4 *)pix, vib, hipow, gamma, lin, lingam)); }

// pixel formats: launchC geometry (cuburn/output.py:21-26); rctxs must hold 1024*256 streams
#define OUT_KERNEL(NAME, T) \
void ref_##NAME(void *dst, const float *src, int gutter, int w, int sstride, int h, uint32_t *seeds) { \
    ringbuf rb = {0, 0}; \
    run_grid((w + 31) / 32, (h + 7) / 8, 32, 8, 1, [&]() { \
        if ((int)(blockIdx.x * blockDim.x + threadIdx.x) >= w || (int)(blockIdx.y * blockDim.y + threadIdx.y) >= h) { \
            /* keep the slot bookkeeping of out-of-frame threads without touching memory */ \
            int tid = blockDim.x * threadIdx.y + threadIdx.x; rb_incr(rb.head, tid); rb_incr(rb.tail, tid); return; } \
        NAME((T *)dst, (const float4 *)src, gutter, w, sstride, h, &rb, (mwc_st *)seeds); }); }
OUT_KERNEL(f32_to_rgba_u8, uchar4)
OUT_KERNEL(f32_to_rgba_u16, ushort4)
OUT_KERNEL(f32_to_yuv444p, char)
OUT_KERNEL(f32_to_yuv444p10, uint16_t)
OUT_KERNEL(f32_to_yuv420p10, uint16_t)
OUT_KERNEL(f32_to_yuv444p12, uint16_t)

// palette: 64 blocks x 256 threads (cuburn/render.py:295-301)
void ref_interp_palette(uint32_t *out /* [rows][256][2] */, uint32_t *seeds, const float *times,
                        const float *sources, float tstart, float tstep, int rows) {
    ringbuf rb = {0, 0};
    flatpal.ptr = (uint2 *)out; flatpal.wbytes = 256 * 8;
    run_grid(rows, 1, 256, 1, 1, [&]() {
        interp_palette_flat(&rb, (mwc_st *)seeds, times, (const float4 *)sources, tstart, tstep); });
}
}
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
    # precalc hunks: camera and affine (code/iter.py:56-95) and the variation
    # precalcs (code/variations.py), rendered against fake packer views
    mwc_stub = types.SimpleNamespace(mwclib=None)
    cuburn_pkg = types.ModuleType('cuburn')
    cuburn_pkg.genome = types.ModuleType('cuburn.genome')
    cuburn_pkg.genome.specs = types.ModuleType('cuburn.genome.specs')
    itermod = _exec_reference('iter.py', {
        'variations': varmod, 'interp': interp, 'util': util_stub, 'mwc': mwc_stub,
        'cuburn': cuburn_pkg, 'cuburn.genome': cuburn_pkg.genome,
        'cuburn.genome.specs': cuburn_pkg.genome.specs})
    meta['precalc'] = {}
    parts.append('typedef struct { uint32_t width, height, awidth, aheight, astride; } acc_size_t;\n'
                 'static acc_size_t acc_size;')
    cam = _FakeNode('cam')
    itermod.precalc_camera(cam)
    parts.append(_precalc_function('ref_precalc_camera_', cam._codes, cam._reg))
    parts.append('extern "C" void ref_precalc_camera(const float *in, float *out, int width, '
                 'int awidth, int aheight) {\n    acc_size.width = width; acc_size.awidth = awidth; '
                 'acc_size.aheight = aheight;\n    ref_precalc_camera_(in, out);\n}')
    meta['precalc']['camera'] = cam._reg
    px = _FakeNode('px')
    itermod.precalc_xf_affine(px)
    parts.append(_precalc_function('ref_precalc_affine', px._codes, px._reg))
    meta['precalc']['affine'] = px._reg
    # cumulative xform densities (code/iter.py:12-30) for 3- and 6-xform genomes
    for names in (('0', '1', '2'), ('0', '1', '2', '3', '4', '5')):
        cp = _FakeNode('cp')
        cp.__dict__['xforms'] = _FakeXforms(_FakeNode('xf', cp._reg, cp._codes), names)
        itermod.precalc_densities(cp)
        kind = 'densities%d' % len(names)
        parts.append(_precalc_function('ref_precalc_' + kind, cp._codes, cp._reg))
        meta['precalc'][kind] = cp._reg
    for vname in ('waves', 'perspective', 'julian', 'juliascope', 'curve'):
        reg, codes = {'in': [], 'out': []}, []
        pvn, pxn = _FakeNode('pv', reg, codes), _FakeNode('px', reg, codes)
        varmod.var_code[vname].namespace['precalc_fun'](pvn, pxn)
        parts.append(_precalc_function('ref_precalc_' + vname, codes, reg))
        meta['precalc'][vname] = reg

    # the iterate kernel's own text: apply_xf_<id>, the choice chain, camera + trunca
    # + bounds (code/iter.py:121-149, 263-272, 302-317), rendered for the sample genomes
    parts.append('static inline uint32_t trunca(float f) {   // cvt.rni.s32.f32 (util.py:194-200;\n'
                 '    // the instruction itself is executed by oracle/ptx_emu.py)\n'
                 '    if (f != f) return 0u;\n'
                 '    if (f >= 2147483648.0f) return 0x7fffffffu;\n'
                 '    if (f <= -2147483648.0f) return 0x80000000u;\n'
                 '    return (uint32_t)(int32_t)nearbyintf(f);\n}')
    meta['iterate'] = {}
    for tag, gnm in _sample_genomes().items():
        code, inputs = _iterate_functions(itermod, tag, gnm)
        parts.append(code)
        meta['iterate'][tag] = inputs
    # inline PTX kept as text: executed by oracle/ptx_emu.py
    blocks = re.findall(r'\{\{crep\("""(.*?)"""\)\}\}', itermod.iter_body_code, re.S)
    usrc_ = open(REF + 'util.py').read()
    meta['ptx'] = {
        'iter_accumulate': blocks[0],       # code/iter.py:332-407; %0..%7 = cc, color_dither,
                                            # time, i, atom_ptr, cosel, out_ptr, hotspot_mult
        'flush_atom': blocks[1],            # code/iter.py:429-540; %0..%6 = gi, hoti, atom_ptr,
                                            # out_ptr, hotspot_ptr, xi, yi
        'trunca': re.search(r'asm\("(cvt\.rni\.s32\.f32[^"]*)"', usrc_).group(1),
    }

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
    # ---- kernels: filters, pixel formats, palette --------------------------------
    parts.append(CUDA_SHIM)
    usrc = open(REF + 'util.py').read()
    parts.append(usrc[usrc.index('#define GET_IDX_2'):usrc.index('""", defs=r\'\'\'\n__device__ uint32_t gtid')])
    filt = _exec_reference('filters.py', {'util': util_stub, 'color': color})
    shear = filt.texshearlib.defs
    # the only inline PTX in the filter code: round-to-nearest-even of i and j
    a0 = shear.index('asm("{')
    a1 = shear.index(': "+f"(i), "+f"(j));') + len(': "+f"(i), "+f"(j));')
    shear = shear[:a0] + 'i = nearbyintf(i); j = nearbyintf(j);' + shear[a1:]
    parts.append('extern "C" {')
    parts.append(filt.denblurlib.decls)
    parts.append(shear)
    parts.append('}')
    for lib_ in (filt.logscalelib, filt.yuvfilterlib, filt.logencodelib, filt.denblurlib,
                 filt.fullblurlib, filt.bilaterallib, filt.halocliplib, filt.smearcliplib,
                 filt.plaincliplib, filt.colorcliplib):
        parts.append(lib_.defs)
    outmod = _exec_reference('output.py', {'util': util_stub, 'mwc': mwc_stub})
    parts.append(outmod.pixfmtlib.defs)
    parts.append(interp.palintlib.decls)
    parts.append(interp.palintlib.defs)
    parts.append(KERNEL_ENTRIES)
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


def _fp(a):
    import ctypes
    return a.ctypes.data_as(ctypes.c_void_p)


def ref_precalc(kind, inputs, width=0, awidth=0, aheight=0):
    """Run a reference precalc hunk; returns {output identifier: value}."""
    import ctypes
    import numpy as np
    reg = meta()['precalc'][kind]
    inp = np.array([inputs[n] for n in reg['in']], np.float32)
    out = np.zeros(len(reg['out']), np.float32)
    if kind == 'camera':
        lib().ref_precalc_camera(_fp(inp), _fp(out), ctypes.c_int(width), ctypes.c_int(awidth),
                                 ctypes.c_int(aheight))
    else:
        getattr(lib(), 'ref_precalc_' + kind)(_fp(inp), _fp(out))
    return dict(zip(reg['out'], out))


def ref_filter(name, *arrays_and_scalars, shape=None):
    """Call ref_<name>(array..., scalar..., astride, aheight); arrays are float32 and
    modified in place where the reference kernel writes them."""
    import ctypes
    import numpy as np
    args = []
    for a in arrays_and_scalars:
        if isinstance(a, np.ndarray):
            args.append(_fp(a))
        elif isinstance(a, (int, np.integer)):
            args.append(ctypes.c_int(int(a)))
        else:
            args.append(ctypes.c_float(float(a)))
    ah, astride = shape
    getattr(lib(), 'ref_' + name)(*args, ctypes.c_int(astride), ctypes.c_int(ah))


def ref_set_gauss(coefs):
    import numpy as np
    lib().ref_set_gauss(_fp(np.ascontiguousarray(coefs, np.float32)))


def ref_convert(fmt, src, w, h, seeds, gutter=12):
    """Reference pixel-format kernel `f32_to_<fmt>`; seeds: uint32 [262144][3]."""
    import ctypes
    import numpy as np
    src = np.ascontiguousarray(src, np.float32)
    seeds = np.array(seeds, np.uint32)
    assert seeds.shape[0] >= 262144
    nbytes = {'rgba_u8': 4, 'rgba_u16': 8, 'yuv444p': 3, 'yuv444p10': 6, 'yuv420p10': 3,
              'yuv444p12': 6}[fmt] * w * h
    # slack: the reference's `>` bounds tests let the x = w/2, y = h/2 threads of the
    # 4:2:0 kernel write past the chroma planes (code/output.py:164,186-189)
    out = np.zeros(nbytes + 8 * w + 64, np.uint8)
    getattr(lib(), 'ref_f32_to_' + fmt)(_fp(out), _fp(src), ctypes.c_int(gutter), ctypes.c_int(w),
                                        ctypes.c_int(src.shape[1]), ctypes.c_int(h), _fp(seeds))
    return out[:nbytes].view(np.uint8 if fmt in ('rgba_u8', 'yuv444p') else np.uint16)


def ref_palette(ptimes, pals, seeds, tstart, tstep, rows=64):
    """Reference interp_palette_flat: returns (uint32 [rows][256][2] packed, seeds)."""
    import ctypes
    import numpy as np
    ptimes = np.ascontiguousarray(ptimes, np.float32)
    pals = np.ascontiguousarray(pals, np.float32)
    seeds = np.array(seeds, np.uint32)
    out = np.zeros((rows, 256, 2), np.uint32)
    lib().ref_interp_palette(_fp(out), _fp(seeds), _fp(ptimes), _fp(pals), ctypes.c_float(tstart),
                             ctypes.c_float(tstep), ctypes.c_int(rows))
    return out, seeds


if __name__ == '__main__':
    print(build(force=True))


class RefIterate(object):
    """The reference's iterate-kernel text rendered for one of the sample genomes
    (`tag` in ref_kernels.json['iterate']): apply_xf_<id>, the xform choice chain and
    final xform + camera + trunca + bounds test, on arrays."""
    def __init__(self, tag, ev):
        """``ev``: oracle.flame_ref.GenomeEval of the same genome; the direct parameters
        are read from its splines at the first temporal sample."""
        import ctypes
        import numpy as np
        self.tag, self.L = tag, lib()
        self.inputs = meta()['iterate'][tag]
        vals = np.array([ev.spline(tuple(p.split('.')))[0] for p in self.inputs], np.float32)
        d = ev.dim
        getattr(self.L, 'ref_%s_setup' % tag)(_fp(vals), ctypes.c_int(d['w']), ctypes.c_int(d['aw']),
                                              ctypes.c_int(d['ah']), ctypes.c_int(d['astride']))

    def _arrays(self, xs, ys, cs, seeds):
        import numpy as np
        return (np.array(xs, np.float32), np.array(ys, np.float32), np.array(cs, np.float32),
                np.array(seeds, np.uint32))

    def apply(self, xf, xs, ys, cs, seeds):
        """xf: index of the xform in sorted-key order, -1 for the final xform."""
        import ctypes
        xs, ys, cs, seeds = self._arrays(xs, ys, cs, seeds)
        getattr(self.L, 'ref_%s_apply' % self.tag)(ctypes.c_int(xf), _fp(xs), _fp(ys), _fp(cs),
                                                   _fp(seeds), ctypes.c_int(xs.size))
        return xs, ys, cs, seeds

    def choose(self, sel, xs, ys, cs, seeds):
        import ctypes
        import numpy as np
        xs, ys, cs, seeds = self._arrays(xs, ys, cs, seeds)
        last = np.zeros(xs.size, np.int32)
        getattr(self.L, 'ref_%s_choose' % self.tag)(ctypes.c_float(sel), _fp(xs), _fp(ys), _fp(cs),
                                                    _fp(seeds), _fp(last), ctypes.c_int(xs.size))
        return last, xs, ys, cs, seeds

    def bin(self, xs, ys, cs, seeds):
        """(bin index or -1, colour after the final xform, seeds)."""
        import ctypes
        import numpy as np
        xs, ys, cs, seeds = self._arrays(xs, ys, cs, seeds)
        idx, cc = np.zeros(xs.size, np.int32), np.zeros(xs.size, np.float32)
        getattr(self.L, 'ref_%s_bin' % self.tag)(_fp(xs), _fp(ys), _fp(cs), _fp(seeds), _fp(idx),
                                                 _fp(cc), ctypes.c_int(xs.size))
        return idx, cc, seeds


def ref_ptx(name):
    """Inline-PTX text of the reference: 'iter_accumulate', 'flush_atom', 'trunca'."""
    return meta()['ptx'][name]
