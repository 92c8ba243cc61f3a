"""Does scattered-RED throughput depend on which allocation holds the histogram?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, mwc
from cuburn_b200.code import itergen
sys.argv = [sys.argv[0]]
src = open(os.path.join(os.path.dirname(__file__), 'red_microbench.py')).read()
SRC = src[src.index("SRC = r'''") + 10: src.index("'''\n\nN.init(0)")]
N.init(0)
names, hdrs = itergen.load_headers()
mod = N.Module(SRC, 'red_bench.cu', hdrs, names, ['--gpu-architecture=sm_100a', '--std=c++17'])
seeds = N.to_device(mwc.make_seeds(262144, host_seed=3))
dim = N.calc_dim(1920, 1080)
nbins = dim.ah * dim.astride
bufs = [N.DeviceBuffer(16 * nbins) for _ in range(10)]
big = N.DeviceBuffer(16 * nbins * 10)
cands = [(b.ptr, 'alloc%d' % i) for i, b in enumerate(bufs)] + \
        [(big.ptr + k * 16 * nbins, 'big+%d' % k) for k in range(10)]
for ptr, name in cands:
    N.fill32(ptr, 4 * nbins, 0)
    best = 1e9
    for rep in range(3):
        e0, e1 = N.Event(), N.Event()
        e0.record(None)
        mod.launch('red_bench', (148 * 4,), (256,), [C.c_uint64(ptr), C.c_uint64(seeds.ptr),
                                                      C.c_uint(nbins), C.c_int(4096), C.c_int(0)])
        e1.record(None); e1.synchronize()
        best = min(best, e1.time_since(e0))
    print('%-8s 0x%x  %.3f ms  %.4g red/s' % (name, ptr, best, 148 * 4 * 256 * 4096 / best * 1e3))
