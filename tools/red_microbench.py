"""
Microbenchmark: peak scattered-reduction rate into an L2-resident histogram on
this GPU -- the roofline the chaos-game kernel's accumulation runs against.
Each thread draws MWC random bins and issues one reduction per sample; no other
work.  Variants: red.v4.f32 (16 B), 2 x red.v2.f32, 4 x red.f32, red.u64 (8 B),
red.u32 (4 B); histogram sizes 1080p / 4K / 8K float4 grids.
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, mwc
from cuburn_b200.code import itergen

SRC = r'''
#include "mwc.cuh"
extern "C" __global__ void __launch_bounds__(256)
red_bench(float4 *hist, mwc_st *seeds, unsigned int nbins, int rounds, int mode) {
    int g = blockIdx.x * 256 + threadIdx.x;
    mwc_st rng = seeds[g];
    float4 v = make_float4(0.25f, 0.5f, 0.75f, 1.0f);
    for (int r = 0; r < rounds; r++) {
        unsigned int u = mwc_next(rng);
        unsigned int bin = __umulhi(u, nbins);
        float4 *p = hist + bin;
        if (mode == 0) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else if (mode == 1) {
            asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
            asm volatile("red.global.add.v2.f32 [%0+8], {%1,%2};" :: "l"(p), "f"(v.z), "f"(v.w) : "memory");
        } else if (mode == 2) {
            asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v.x) : "memory");
            asm volatile("red.global.add.f32 [%0+4], %1;" :: "l"(p), "f"(v.y) : "memory");
            asm volatile("red.global.add.f32 [%0+8], %1;" :: "l"(p), "f"(v.z) : "memory");
            asm volatile("red.global.add.f32 [%0+12], %1;" :: "l"(p), "f"(v.w) : "memory");
        } else if (mode == 3) {
            unsigned long long val = (1ull << 54) | ((unsigned long long)(u & 255) << 36) | (77ull << 18) | 99ull;
            asm volatile("red.global.add.u64 [%0], %1;" :: "l"((unsigned long long *)hist + bin), "l"(val) : "memory");
        } else if (mode == 4) {
            asm volatile("red.global.add.u32 [%0], %1;" :: "l"((unsigned int *)hist + bin), "r"(1u) : "memory");
        } else if (mode == 5) {
            asm volatile("red.global.add.f32 [%0], %1;" :: "l"((float *)hist + bin), "f"(v.x) : "memory");
        }
    }
    seeds[g] = rng;
}
'''

N.init(0)
names, hdrs = itergen.load_headers()
mod = N.Module(SRC, 'red_bench.cu', hdrs, names, ['--gpu-architecture=sm_100a', '--std=c++17', '-lineinfo'])
sms = N.device_info(0)['sm_count']
seeds = N.to_device(mwc.make_seeds(262144, host_seed=3))
out = []
modes = ['red.v4.f32', '2x red.v2.f32', '4x red.f32', 'red.u64', 'red.u32', 'red.f32']
for (label, w, h) in (('1080p', 1920, 1080), ('4K', 3840, 2160), ('8K', 7680, 4320)):
    dim = N.calc_dim(w, h)
    nbins = dim.ah * dim.astride
    hist = N.DeviceBuffer(16 * nbins)
    N.fill32(hist, 4 * nbins, 0)
    for ctas_per_sm in (2, 4, 6):
        grid = sms * ctas_per_sm
        rounds = 4096
        for mode, mname in enumerate(modes):
            best = 1e9
            for rep in range(3):
                e0, e1 = N.Event(), N.Event()
                e0.record(None)
                mod.launch('red_bench', (grid,), (256,),
                           [C.c_uint64(hist.ptr), C.c_uint64(seeds.ptr), C.c_uint(nbins),
                            C.c_int(rounds), C.c_int(mode)])
                e1.record(None)
                e1.synchronize()
                best = min(best, e1.time_since(e0))
            n = grid * 256 * rounds
            rate = n / best * 1e3
            out.append(dict(grid=label, hist_mib=16 * nbins / 2 ** 20, ctas_per_sm=ctas_per_sm,
                            mode=mname, samples_per_s=rate, ms=best))
            print('%-6s %6.1f MiB  %d CTA/SM  %-14s %8.3f ms  %.4g samples/s' % (
                label, 16 * nbins / 2 ** 20, ctas_per_sm, mname, best, rate), flush=True)
    hist.free()
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/red_microbench.json', 'w'), indent=1)
