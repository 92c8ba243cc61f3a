"""
Microbenchmark: atomics into the shared memory of *other* CTAs of a thread-block cluster
(distributed shared memory) on sm_100a.  Question: could the histogram's bright region live
in the pooled shared memory of a cluster (8 or 16 SMs x ~100 KB), so that an ordinary flame
sends a third of its samples there instead of to L2?  Each sample = CELL_WORDS x
red.shared::cluster.add.u32 into a cell of a window that is spread over the cluster's CTAs
(owner = cell / cells_per_cta, address through mapa), cells drawn uniformly from the window.

  python tools/dsmem_atomic_microbench.py   -> JSON lines: samples/s for the whole GPU by
  cluster size, words per cell, CTAs per SM and share of remote samples
"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, mwc
from cuburn_b200.code import itergen

SRC = r'''
#include "mwc.cuh"
#ifndef CLUSTER
#define CLUSTER 8
#endif
#ifndef WORDS
#define WORDS 2
#endif
extern __shared__ unsigned int cells[];

__device__ __forceinline__ unsigned int cluster_rank() {
    unsigned int r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void red_cluster(unsigned int local_addr, unsigned int rank, unsigned int v) {
    unsigned int remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("red.shared::cluster.add.u32 [%0], %1;" :: "r"(remote), "r"(v) : "memory");
}

extern "C" __global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(256)
dsmem_bench(unsigned long long *out, mwc_st *seeds, unsigned int cells_per_cta, int rounds,
            int local_only) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    for (unsigned int i = threadIdx.x; i < cells_per_cta * WORDS; i += 256) cells[i] = 0u;
    cluster_sync();
    mwc_st rng = seeds[g];
    const unsigned int me = cluster_rank();
    const unsigned int base = (unsigned int)__cvta_generic_to_shared(cells);
    const unsigned int window = cells_per_cta * CLUSTER;
    for (int r = 0; r < rounds; r++) {
        const unsigned int u = mwc_next(rng);
        const unsigned int c = __umulhi(u, window);
        const unsigned int owner = local_only ? me : c / cells_per_cta;
        const unsigned int off = (c % cells_per_cta) * (WORDS * 4);
#pragma unroll
        for (int w = 0; w < WORDS; w++) red_cluster(base + off + 4 * w, owner, 1u + (u & 255u));
    }
    cluster_sync();
    unsigned long long s = 0;
    for (unsigned int i = threadIdx.x; i < cells_per_cta * WORDS; i += 256) s += cells[i];
    if (s == 0xffffffffffffffffull) out[0] = s;         // keep the cells alive
    seeds[g] = rng;
}
'''

N.init(0)
sms = N.device_info(0)['sm_count']
names, hdrs = itergen.load_headers()
seeds = N.to_device(mwc.make_seeds(262144, host_seed=5))
out = N.DeviceBuffer(64)
rounds = 4096
for cluster in (1, 2, 4, 8, 16):
    for words in (2, 4):
        mod = N.Module(SRC, 'dsmem.cu', hdrs, names,
                       ['--gpu-architecture=sm_100a', '--std=c++17', '-DCLUSTER=%d' % cluster, '-DWORDS=%d' % words])
        for ctas_per_sm, kb in ((4, 24), (2, 96)):
            cells = kb * 1024 // (4 * words)
            grid = (sms * ctas_per_sm) // cluster * cluster
            if grid * 256 > 262144:
                grid = 262144 // 256 // cluster * cluster
            for local_only in (0, 1):
                best = 1e9
                try:
                    for _ in range(3):
                        e0, e1 = N.Event(), N.Event()
                        e0.record(None)
                        mod.launch('dsmem_bench', (grid,), (256,),
                                   [C.c_uint64(out.ptr), C.c_uint64(seeds.ptr), C.c_uint(cells), C.c_int(rounds),
                                    C.c_int(local_only)], dyn_smem=kb * 1024)
                        e1.record(None); e1.synchronize()
                        best = min(best, e1.time_since(e0))
                    rate = grid * 256 * rounds / (best * 1e-3)
                    err = None
                except Exception as e:
                    rate, err = None, str(e)[:120]
                print(json.dumps(dict(cluster=cluster, words_per_cell=words, kb_per_cta=kb, ctas=grid,
                                      local_only=bool(local_only), ms=best if rate else None,
                                      samples_per_s=rate, error=err)), flush=True)
