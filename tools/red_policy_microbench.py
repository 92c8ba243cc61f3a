"""Scattered red.global.add.v4.f32 into grids around the size of L2, with and without L2
eviction-priority hints on the reduction (does a 'persisting' hint stop the write-backs that
the 4K histogram suffers?).   python tools/red_policy_microbench.py  -> JSON lines"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, mwc
from cuburn_b200.code import itergen

SRC = r'''
#include "mwc.cuh"
extern "C" __global__ void __launch_bounds__(256)
red_policy(float4 *hist, mwc_st *seeds, unsigned int nbins, int rounds) {
    int g = blockIdx.x * 256 + threadIdx.x;
    mwc_st rng = seeds[g];
#if POLICY
    unsigned long long pol;
#if POLICY == 1
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#elif POLICY == 4
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.5;" : "=l"(pol));
#elif POLICY == 5
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.25;" : "=l"(pol));
#elif POLICY == 6
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.5;" : "=l"(pol));
#elif POLICY == 2
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#else
    asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
#endif
#endif
    for (int r = 0; r < rounds; r++) {
        unsigned int bin = __umulhi(mwc_next(rng), nbins);
#if POLICY
        asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                     :: "l"(hist + bin), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(1.0f), "l"(pol) : "memory");
#else
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(hist + bin), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(1.0f) : "memory");
#endif
    }
    seeds[g] = rng;
}
'''
N.init(0)
sms = N.device_info(0)['sm_count']
names, hdrs = itergen.load_headers()
seeds = N.to_device(mwc.make_seeds(262144, host_seed=5))
grid, rounds = sms * 6, 8192
big = N.DeviceBuffer(320 << 20)
POLICIES = ((0, 'none'), (1, 'evict_last'), (2, 'evict_first'), (3, 'evict_unchanged'),
            (4, 'evict_last 0.5 / evict_first'), (5, 'evict_last 0.25 / evict_first'), (6, 'evict_last 0.5 / normal'))
SIZES = [int(x) for x in os.environ.get('SIZES', '33,66,100,120,130,160').split(',')]
if os.environ.get('POLICIES'):
    POLICIES = [p for p in POLICIES if str(p[0]) in os.environ['POLICIES'].split(',')]
for policy, pname in POLICIES:
    try:
        mod = N.Module(SRC, 'redpol.cu', hdrs, names, ['--gpu-architecture=sm_100a', '--std=c++17', '-DPOLICY=%d' % policy])
    except Exception as e:
        print(json.dumps(dict(policy=pname, error=str(e)[:200]))); continue
    for mb in SIZES:
        nbins = (mb << 20) // 16
        best = 1e9
        for rep in range(3):
            N.fill32(big, (320 << 20) // 4, 0)
            e0, e1 = N.Event(), N.Event()
            e0.record(None)
            mod.launch('red_policy', (grid,), (256,), [C.c_uint64(big.ptr), C.c_uint64(seeds.ptr), C.c_uint(nbins), C.c_int(rounds)])
            e1.record(None); e1.synchronize()
            best = min(best, e1.time_since(e0))
        print(json.dumps(dict(policy=pname, grid_mib=mb, ms=round(best, 3), reds_per_s=grid * 256 * rounds / best * 1e3)), flush=True)
