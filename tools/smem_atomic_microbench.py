"""
Microbenchmark: shared-memory atomic rates on sm_100a, for the accumulation design
BASELINE.json's north star names ("accumulation in the 228 KB-per-SM shared memory ...
shared atomics where bins fit, L2 red.global spill only where they do not").

Every thread draws MWC random bins inside its CTA's shared-memory tile and issues the
accumulation of one sample per round; nothing else.  What one sample costs depends on the
cell format, because sm_100a has exactly one native shared atomic add, ATOMS.ADD (32-bit
integer): `red.shared.add.u64` and `red.shared.add.f32` compile to ATOMS.CAST.SPIN
compare-and-swap loops (cuobjdump of this file's kernel, see profiles/r02_smem_atomics.md).

  mode u32        1 x red.shared.add.u32              (a count-only cell: lower bound)
  mode 2xu32      2 x red.shared.add.u32              (the reference's packed cell, a15, split
                                                       in two carry-free words: count:14|Y:18,
                                                       U:16|V:16)
  mode u64cas     1 x red.shared.add.u64              (the packed u64 cell as the reference
                                                       defines it; CAS loop in SASS)
  mode 4xf32      4 x red.shared.add.f32              (a float4 cell; CAS loops)
  mode match2xu32 match.any on the bin, leader adds popc * value with 2 x u32
  mode sts        1 x st.shared.u64 (no atomicity)    (the LSU floor)
  mode redg_v4    1 x red.global.add.v4.f32 into an L2-resident 33 MiB grid (what cb_iter
                                                       does today), in the same harness

Address streams: `uniform` over the tile; `flame` -- bin = tile * u^4 (a heavy head: half
of the samples fall into 6 % of the bins, the shape of a flame's bright core inside a
tile); `hot` -- 25 % of the samples to one bin, the rest uniform.
Tile sizes: 2048 / 8192 / 28672 cells of 8 bytes per CTA (16 / 64 / 224 KB), 8 / 3 / 1
CTAs per SM.  Output: JSON lines with samples/s for the whole GPU.
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
extern __shared__ unsigned long long cells[];

__device__ __forceinline__ unsigned int pick(unsigned int u, unsigned int ncells, int shape) {
    if (shape == 0) return __umulhi(u, ncells);
    if (shape == 1) {
        float f = (float)u * 2.3283064365386962890625e-10f;
        f = f * f; f = f * f;
        unsigned int b = (unsigned int)(f * (float)ncells);
        return b < ncells ? b : ncells - 1;
    }
    return (u & 3u) == 0u ? 7u : __umulhi(u, ncells);
}

extern "C" __global__ void __launch_bounds__(256)
smem_bench(unsigned long long *out, float4 *hist, mwc_st *seeds, unsigned int ncells,
           unsigned int nbins, int rounds, int mode, int shape) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    for (unsigned int i = threadIdx.x; i < ncells; i += 256) cells[i] = 0ull;
    __syncthreads();
    mwc_st rng = seeds[g];
    const unsigned int base = (unsigned int)__cvta_generic_to_shared(cells);
    for (int r = 0; r < rounds; r++) {
        unsigned int u = mwc_next(rng);
        unsigned int bin = pick(u, ncells, shape);
        unsigned int a = base + bin * 8u;
        unsigned int lvl = u & 255u;
        if (mode == 0) {
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(1u) : "memory");
        } else if (mode == 1) {
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"((1u << 18) | lvl) : "memory");
            asm volatile("red.shared.add.u32 [%0+4], %1;" :: "r"(a), "r"((lvl << 16) | 99u) : "memory");
        } else if (mode == 2) {
            unsigned long long v = (1ull << 54) | ((unsigned long long)lvl << 36) | (77ull << 18) | 99ull;
            asm volatile("red.shared.add.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory");
        } else if (mode == 3) {
            // float4 cells: 16 bytes, half as many fit
            unsigned int a4 = base + (bin >> 1) * 16u;
            asm volatile("red.shared.add.f32 [%0], %1;" :: "r"(a4), "f"(0.25f) : "memory");
            asm volatile("red.shared.add.f32 [%0+4], %1;" :: "r"(a4), "f"(0.5f) : "memory");
            asm volatile("red.shared.add.f32 [%0+8], %1;" :: "r"(a4), "f"(0.75f) : "memory");
            asm volatile("red.shared.add.f32 [%0+12], %1;" :: "r"(a4), "f"(1.0f) : "memory");
        } else if (mode == 4) {
            unsigned int peers = __match_any_sync(0xffffffffu, bin);
            unsigned int n = __popc(peers);
            if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) {
                asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(n * ((1u << 18) | 128u)) : "memory");
                asm volatile("red.shared.add.u32 [%0+4], %1;" :: "r"(a), "r"(n * ((128u << 16) | 99u)) : "memory");
            }
        } else if (mode == 5) {
            asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"((unsigned long long)u) : "memory");
        } else if (mode == 6) {
            unsigned int gb = __umulhi(u, nbins);
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                         :: "l"(hist + gb), "f"(0.25f), "f"(0.5f), "f"(0.75f), "f"(1.0f) : "memory");
        }
    }
    __syncthreads();
    // keep the tile observable: one checksum per CTA
    unsigned long long acc = 0ull;
    for (unsigned int i = threadIdx.x; i < ncells; i += 256) acc += cells[i];
    atomicAdd(out + (blockIdx.x & 1023), acc);
    seeds[g] = rng;
}
'''

MODES = ['u32', '2xu32', 'u64cas', '4xf32', 'match2xu32', 'sts', 'redg_v4']
SHAPES = ['uniform', 'flame', 'hot']


def main():
    N.init(0)
    names, hdrs = itergen.load_headers()
    mod = N.Module(SRC, 'smem_bench.cu', hdrs, names,
                   ['--gpu-architecture=sm_100a', '--std=c++17', '-lineinfo'])
    if '--cubin' in sys.argv:
        with open(sys.argv[sys.argv.index('--cubin') + 1], 'wb') as fp:
            fp.write(mod.cubin)
    sms = N.device_info(0)['sm_count']
    seeds = N.to_device(mwc.make_seeds(262144, host_seed=3))
    dim = N.calc_dim(1920, 1080)
    nbins = dim.ah * dim.astride
    hist = N.DeviceBuffer(16 * nbins)
    N.fill32(hist, 4 * nbins, 0)
    out = N.DeviceBuffer(8 * 1024)
    N.fill32(out, 2 * 1024, 0)
    rows = []
    for ncells, ctas_per_sm in ((2048, 8), (8192, 3), (28672, 1)):
        grid = min(sms * ctas_per_sm, 1024)
        for shape, sname in enumerate(SHAPES):
            for mode, mname in enumerate(MODES):
                if mname == 'redg_v4' and shape != 0:
                    continue
                rounds = 512 if mname in ('u64cas', '4xf32') else 2048
                best = 1e9
                for rep in range(3):
                    e0, e1 = N.Event(), N.Event()
                    e0.record(None)
                    mod.launch('smem_bench', (grid,), (256,),
                               [C.c_uint64(out.ptr), C.c_uint64(hist.ptr), C.c_uint64(seeds.ptr),
                                C.c_uint(ncells), C.c_uint(nbins), C.c_int(rounds),
                                C.c_int(mode), C.c_int(shape)], dyn_smem=8 * ncells)
                    e1.record(None)
                    e1.synchronize()
                    best = min(best, e1.time_since(e0))
                n = grid * 256 * rounds
                row = dict(mode=mname, shape=sname, cells_per_cta=ncells,
                           tile_kb=8 * ncells // 1024, ctas_per_sm=ctas_per_sm, grid=grid,
                           samples=n, ms=best, samples_per_s=n / best * 1e3)
                rows.append(row)
                print(json.dumps(row), flush=True)
    return rows


if __name__ == '__main__':
    main()
