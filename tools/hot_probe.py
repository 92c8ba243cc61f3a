"""How do hot bins / hot regions limit scattered float4 REDs at L2?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, mwc
from cuburn_b200.code import itergen
SRC = r'''
#include "mwc.cuh"
// with probability phot a sample goes to one of K hot bins laid out from `hot0`
// with stride `hstride` bins; otherwise uniform over nbins
extern "C" __global__ void __launch_bounds__(256)
hot_bench(float4 *hist, mwc_st *seeds, unsigned int nbins, int rounds, float phot,
          unsigned int K, unsigned int hot0, unsigned int hstride, int dedup) {
    int g = blockIdx.x * 256 + threadIdx.x;
    mwc_st rng = seeds[g];
    float4 v = make_float4(0.25f, 0.5f, 0.75f, 1.0f);
    for (int r = 0; r < rounds; r++) {
        unsigned int u = mwc_next(rng);
        unsigned int bin = __umulhi(u, nbins);
        if (mwc_next_01(rng) < phot) bin = hot0 + (u % K) * hstride;
        float4 o = v;
        bool lead = true;
        if (dedup) {
            unsigned int m = __match_any_sync(0xffffffffu, bin);
            int cnt = __popc(m);
            lead = (__ffs(m) - 1) == (threadIdx.x & 31);
            o.x *= cnt; o.y *= cnt; o.z *= cnt; o.w *= cnt;
        }
        if (lead)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(hist + bin), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
    }
    seeds[g] = rng;
}
'''
N.init(0)
names, hdrs = itergen.load_headers()
mod = N.Module(SRC, 'hot.cu', hdrs, names, ['--gpu-architecture=sm_100a', '--std=c++17'])
seeds = N.to_device(mwc.make_seeds(262144, host_seed=3))
dim = N.calc_dim(1920, 1080)
nbins = dim.ah * dim.astride
hist = N.DeviceBuffer(16 * nbins)
N.fill32(hist, 4 * nbins, 0)
grid, rounds = 148 * 4, 2048
def run(phot, K, hstride, dedup=0):
    best = 1e9
    for rep in range(3):
        e0, e1 = N.Event(), N.Event()
        e0.record(None)
        mod.launch('hot_bench', (grid,), (256,), [C.c_uint64(hist.ptr), C.c_uint64(seeds.ptr), C.c_uint(nbins),
                   C.c_int(rounds), C.c_float(phot), C.c_uint(K), C.c_uint(1000000), C.c_uint(hstride), C.c_int(dedup)])
        e1.record(None); e1.synchronize()
        best = min(best, e1.time_since(e0))
    return grid * 256 * rounds / best * 1e3
print('uniform            : %.4g /s' % run(0.0, 1, 1))
for phot in (1e-4, 1e-3, 3e-3, 1e-2, 3e-2, 0.1):
    for K, hs, label in ((1, 1, '1 bin'), (16, 1, '16 adjacent bins (one 256B chunk)'), (16, 1952, '16 bins in a column'),
                         (16, 4099, '16 scattered bins'), (256, 1, '256 adjacent bins'), (256, 4099, '256 scattered bins')):
        a = run(phot, K, hs)
        b = run(phot, K, hs, 1)
        print('phot %-6g %-34s: %.4g /s   with warp dedup: %.4g /s' % (phot, label, a, b), flush=True)
