"""
In-tree build of the native pieces:

  csrc/libcuburn_b200.so       the C-ABI library (nvcc, sm_100a, links NVRTC)
  data/mwc_multipliers.bin     the 262144 MWC multipliers (tools/gen_mwc_multipliers.c)

``python -m cuburn_b200.build`` builds both; ``__graft_entry__.build()`` calls
``build_all()``.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libcuburn_b200.so')
MULT_PATH = os.path.join(HERE, 'data', 'mwc_multipliers.bin')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC']
# (source, extra flags).  Only the filter kernels use fast-math (like the
# reference); interpolation / output are bit-exact stages and must not.
SOURCES = [
    ('cb_core.cu', []),
    ('cb_interp.cu', ['-fmad=false']),
    ('cb_module.cu', []),
    ('cb_filters.cu', ['-use_fast_math']),
    ('cb_output.cu', ['-fmad=false']),
    ('cb_comm.cu', []),
    ('cb_sort.cu', []),
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('command failed: %s\n%s' % (' '.join(cmd), r.stdout))
    return r.stdout


def build_library(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', 'nvcc')
    hdrs = [os.path.join(CSRC, 'cb_common.h'),
            os.path.join(ROOT, 'include', 'cuburn_b200.h')]
    dev = os.path.join(CSRC, 'device')
    hdrs += [os.path.join(dev, f) for f in os.listdir(dev)]
    objs = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace('.cu', '.o'))
        if force or _newer(o, [s] + hdrs):
            out = _run([nvcc] + ARCH + COMMON + extra + ['-c', s, '-o', o])
            if verbose and out.strip():
                print(out)
        objs.append(o)
    if force or _newer(LIB_PATH, objs):
        _run([nvcc] + ARCH + ['-shared', '-o', LIB_PATH] + objs +
             ['-lnvrtc', '-Xlinker', '-rpath,/usr/local/cuda/lib64'])
    return LIB_PATH


def build_multipliers(force=False):
    if os.path.exists(MULT_PATH) and not force and os.path.getsize(MULT_PATH) == 4 * 262144:
        return MULT_PATH
    os.makedirs(os.path.dirname(MULT_PATH), exist_ok=True)
    exe = os.path.join(ROOT, 'tools', 'gen_mwc_multipliers')
    src = exe + '.c'
    _run(['gcc', '-O2', '-fopenmp', '-o', exe, src])
    _run([exe, MULT_PATH, '262144'])
    return MULT_PATH


def build_all(force=False, verbose=False):
    return build_library(force, verbose), build_multipliers(force)


if __name__ == '__main__':
    lib, mult = build_all(force='--force' in sys.argv, verbose=True)
    print(lib)
    print(mult)
