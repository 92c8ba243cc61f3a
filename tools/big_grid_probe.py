"""cb_iter at 4K / 8K (grids beyond L2) and u64 atom/red rates."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render, mwc
from cuburn_b200.code import itergen
N.init(0)
rmgr = render.RenderManager(seed=1)
for gname, w, h, spp in (('G6F', 3840, 2160, 1000), ('G24H', 7680, 4320, 250), ('G6F', 7680, 4320, 250)):
    for swz in (True, False):
        rmgr.swizzle = swz
        gnm = samples.GENOMES[gname]()
        gprof = profile.wrap(dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, tc, 0.0)
        ms = []
        for i in range(3):
            e0, e1 = N.Event(), N.Event()
            e0.record(rmgr.stream_a)
            rmgr._iter(rdr, gnm, gprof, dim, tc)
            e1.record(rmgr.stream_a); e1.synchronize()
            ms.append(e1.time_since(e0))
        n = rmgr.last_iter_samples
        print('%-5s %dx%d spp %d swizzle=%d: %.1f ms  %.4g it/s' % (gname, w, h, spp, swz, min(ms), n / min(ms) * 1e3), flush=True)

SRC = r'''
#include "mwc.cuh"
extern "C" __global__ void __launch_bounds__(256)
u64_bench(unsigned long long *hist, mwc_st *seeds, unsigned int nbins, int rounds, int mode, unsigned long long *sink) {
    int g = blockIdx.x * 256 + threadIdx.x;
    mwc_st rng = seeds[g];
    unsigned long long acc = 0;
    for (int r = 0; r < rounds; r++) {
        unsigned int u = mwc_next(rng);
        unsigned int bin = __umulhi(u, nbins);
        unsigned long long val = (1ull << 54) | ((unsigned long long)(u & 255) << 36) | (77ull << 18) | 99ull;
        bool check = mode == 1 || (mode == 2 && (r & 31) == (threadIdx.x >> 5 & 31) % 32 && (r % 32 == 0));
        if (mode == 2) check = ((r + (threadIdx.x >> 5)) & 31) == 0;      // one round in 32 per warp
        if (check) {
            unsigned long long old;
            asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(hist + bin), "l"(val) : "memory");
            if ((old >> 54) >= 512) acc += old;
        } else {
            asm volatile("red.global.add.u64 [%0], %1;" :: "l"(hist + bin), "l"(val) : "memory");
        }
    }
    if (acc == 0x1234567) sink[0] = acc;
    seeds[g] = rng;
}
'''
names, hdrs = itergen.load_headers()
mod = N.Module(SRC, 'u64.cu', hdrs, names, ['--gpu-architecture=sm_100a', '--std=c++17'])
seeds = N.to_device(mwc.make_seeds(262144, host_seed=3))
sink = N.DeviceBuffer(64)
for label, w, h in (('1080p', 1920, 1080), ('4K', 3840, 2160), ('8K', 7680, 4320)):
    dim = N.calc_dim(w, h)
    nbins = dim.ah * dim.astride
    hist = N.DeviceBuffer(8 * nbins)
    for mode, mname in ((0, 'red.u64'), (1, 'atom.u64 always'), (2, 'atom 1 round in 32')):
        best = 1e9
        for rep in range(3):
            N.fill32(hist, 2 * nbins, 0)
            e0, e1 = N.Event(), N.Event()
            e0.record(None)
            mod.launch('u64_bench', (148 * 6,), (256,), [C.c_uint64(hist.ptr), C.c_uint64(seeds.ptr), C.c_uint(nbins),
                       C.c_int(1024), C.c_int(mode), C.c_uint64(sink.ptr)])
            e1.record(None); e1.synchronize()
            best = min(best, e1.time_since(e0))
        print('%-6s u64 grid %.1f MiB %-20s %.4g /s' % (label, 8 * nbins / 2**20, mname, 148 * 6 * 256 * 1024 / best * 1e3), flush=True)
    hist.free()
