"""Does the L2 float reduction (red.global.add.f32 / .v4.f32) round to nearest?
Adds the same value n times to one address and compares with exact and with host
float32 running sums under round-to-nearest and round-toward-zero."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, ctypes
from cuburn_b200 import _native as N
N.init(0)
src = r'''
extern "C" __global__ void k(float4 *p, float v, int per_thread) {
    for (int i = 0; i < per_thread; i++)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v), "f"(v * 0.5f), "f"(v * 0.01f), "f"(1.0f) : "memory");
}
'''
mod = N.Module(src, 'r.cu', [], [], ['--gpu-architecture=sm_100a'])
buf = N.DeviceBuffer(16)
for v in (0.94, 0.6196078431, 0.0039215686):
    N.fill32(buf, 4, 0)
    n_threads, per = 256 * 148, 16
    mod.launch('k', (148,), (256,), [ctypes.c_uint64(buf.ptr), ctypes.c_float(v), ctypes.c_int(per)])
    N.check(N.lib().cb_device_sync())
    got = N.from_device(buf, (4,), np.float32)
    n = n_threads * per
    vals = np.float32([v, np.float32(v) * np.float32(0.5), np.float32(v) * np.float32(0.01), 1.0])
    exact = vals.astype(np.float64) * n
    # host float32 running sums: nearest
    rn = np.zeros(4, np.float32)
    for i in range(n):
        rn = rn + vals
    print('v=%g n=%d' % (v, n))
    print('  gpu   ', got)
    print('  exact ', exact)
    print('  host f32 round-to-nearest sequential', rn)
    print('  gpu/exact - 1', got / exact - 1)
