"""
Per-kernel timing of the filter chain (CUDA events).  W/H select the frame size:
at 1080p one float4 plane (33 MiB) stays L2-resident between kernels, at 4K (130 MiB
per plane) every kernel streams from and to HBM.  FLUSH=1 writes 512 MiB before each
kernel (which also charges the kernel with writing those dirty lines back).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render, filters

N.init(0)
w, h = int(os.environ.get('W', 1920)), int(os.environ.get('H', 1080))
gnm = samples.g6f()
gprof = profile.wrap(dict(width=w, height=h, spp=500, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=1)
rdr = render.Renderer(gnm, gprof)
dim = rmgr.fb.set_dim(w, h)
rmgr._copy(rdr, gnm); rmgr._interp(rdr, gnm, dim, tc, 0.0); rmgr._iter(rdr, gnm, gprof, dim, tc)
rmgr.stream_a.synchronize()
nbins = dim.ah * dim.astride
hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
flush = N.DeviceBuffer(512 << 20)
L, s = N.lib(), rmgr.stream_a
peak = 6451.2
def timed(name, bytes_per_bin, fn, reps=5):
    best = 1e9
    for _ in range(reps):
        if os.environ.get('FLUSH', '0') == '1':
            N.fill32(flush, (512 << 20) // 4, 0, s)
        e0, e1 = N.Event(), N.Event()
        e0.record(s); fn(); e1.record(s); e1.synchronize()
        best = min(best, e1.time_since(e0))
    gbs = bytes_per_bin * nbins / best / 1e6
    print('%-28s %8.1f us  %7.1f GB/s  %5.1f%% of %.0f' % (name, best * 1e3, gbs, 100 * gbs / peak, peak), flush=True)
    return best
fb = rmgr.fb
c1 = filters.gauss_coefs(1)
D = N.byref(dim)
f32 = np.float32
total = 0
total += timed('yuv_to_rgb', 32, lambda: L.cb_yuv_to_rgb(fb.d_back.ptr, fb.d_front.ptr, D, s.handle))
total += 8 * timed('bilateral_direction (x8)', 64, lambda: L.cb_bilateral_direction(
    fb.d_front.ptr, fb.d_back.ptr, fb.d_left.ptr, 3, 15, c1, f32(6), f32(0.05), f32(1.5), f32(0.8), f32(4), D, s.handle))
total += timed('logscale', 32, lambda: L.cb_logscale(fb.d_front.ptr, fb.d_front.ptr, f32(4.2), f32(1e-4), D, s.handle))
total += timed('apply_gamma_full_hi', 32, lambda: L.cb_apply_gamma_full_hi(fb.d_left.ptr, fb.d_front.ptr, f32(-0.75), D, s.handle))
total += 4 * timed('full_blur (x4)', 32, lambda: L.cb_full_blur(fb.d_back.ptr, fb.d_left.ptr, 2, 0, c1, D, s.handle))
total += timed('smearclip', 48, lambda: L.cb_smearclip(fb.d_front.ptr, fb.d_left.ptr, f32(-0.75), f32(0.01), f32(31.6), D, s.handle))
total += timed('colorclip', 32, lambda: L.cb_colorclip(fb.d_front.ptr, f32(1), f32(-1), f32(0.25), f32(0.01), f32(31.6), D, s.handle))
total += timed('convert rgba8', 20, lambda: L.cb_convert(0, fb.d_back.ptr, fb.d_front.ptr, 12, D, fb.d_seeds.ptr, fb.nstreams, s.handle))
total += timed('hist_unswizzle', 32, lambda: L.cb_hist_unswizzle(fb.d_front.ptr, fb.d_left.ptr, (nbins // 65536) * 65536, D, s.handle))
print('default chain (yuv + 8 bilateral + logscale + smearclip + rgba8): %.1f us' % ((total) * 1e3))
