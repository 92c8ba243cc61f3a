"""Time cb_fill32 on histogram-sized buffers: python tools/fill_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N
N.init(0)
s = N.Stream()
for mib in (32.9, 129.5, 512.1):
    n = int(mib * 2 ** 20) // 16 * 16
    buf = N.DeviceBuffer(n)
    best = 1e9
    for i in range(6):
        e0, e1 = N.Event(), N.Event()
        e0.record(s); N.fill32(buf, n // 4, 0, s); e1.record(s); e1.synchronize()
        best = min(best, e1.time_since(e0))
    got = N.from_device(buf, (n // 4,), np.uint32)
    N.fill32(buf, n // 4 - 3, 0x7fc00000, s); s.synchronize()
    g2 = N.from_device(buf, (n // 4,), np.uint32)
    assert (got == 0).all() and (g2[:-3] == 0x7fc00000).all() and (g2[-3:] == 0).all()
    print('%7.1f MiB  %7.1f us  %6.0f GB/s' % (mib, best * 1e3, n / best / 1e6))
