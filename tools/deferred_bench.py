"""What would tile-binned deferred accumulation cost?  (BASELINE north star: "tiled or
deferred sorted writes"; reference design evidence: cuburn/code/sort.py, helpers/sortbench.cu.)

The scheme: the chaos game appends one 32-bit record (bin << 8 | palette column) per sample
instead of a reduction; records are partitioned by tile (one stable radix pass on the bin's
high bits: code/sort.py = cb_sort_pass); CTAs replay the records of a tile into shared memory
with native shared atomics and add the tile to the histogram.

Measured here, every stage alone and at its best case, on records drawn from a real flame
(a G6F 1080p histogram rendered by cb_iter is the address distribution):
  append   coalesced 4-byte stores of ready-made records (the floor for what would replace
           the reduction in cb_iter)
  sort     Sorter.sort by the 8 bits above bit 21 (256 tiles of 8192 bins)
  replay   tiles cut into chunks of <= 131072 records, one CTA (1024 threads, 64 KB of
           cells) per chunk: 2 x ATOMS.ADD per record (the optimistic packed-cell count; an
           exact 4-channel cell needs 4), non-empty cells added to the float4 grid with one
           reduction each
  direct   what cb_iter does: one red.global.add.v4.f32 per record into the same grid.
    python tools/deferred_bench.py [log2 records]   -> JSON"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render
from cuburn_b200.code import itergen
from cuburn_b200.code.sort import Sorter

SRC = r'''
#define TILE_BINS 8192
extern "C" __global__ void __launch_bounds__(256)
k_append(unsigned int *dst, const unsigned int *src, unsigned int n) {
    for (unsigned int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        dst[i] = src[i] ^ 1u;          // a record per thread and round, stored coalesced
}
extern "C" __global__ void __launch_bounds__(256)
k_direct(float4 *hist, const unsigned int *records, unsigned int n) {
    for (unsigned int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const unsigned int rec = records[i];
        const float c = (float)(rec & 255u);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(hist + (rec >> 8)), "f"(c), "f"(c), "f"(c), "f"(1.0f) : "memory");
    }
}
// chunk = (tile, first record, end record): all its records have bin >> 13 == tile
extern "C" __global__ void __launch_bounds__(1024)
k_replay(float4 *hist, const unsigned int *records, const int *chunks) {
    extern __shared__ unsigned int cell[];           // [TILE_BINS][2]: count, level sum
    for (int i = threadIdx.x; i < 2 * TILE_BINS; i += 1024) cell[i] = 0u;
    __syncthreads();
    const int tile = chunks[3 * blockIdx.x];
    const unsigned int lo = chunks[3 * blockIdx.x + 1], hi = chunks[3 * blockIdx.x + 2];
    for (unsigned int i = lo + threadIdx.x; i < hi; i += 1024) {
        const unsigned int rec = records[i];
        const unsigned int b = (rec >> 8) & (TILE_BINS - 1u);
        atomicAdd(&cell[2 * b], 1u);
        atomicAdd(&cell[2 * b + 1], rec & 255u);
    }
    __syncthreads();
    float4 *out = hist + (size_t)tile * TILE_BINS;
    for (int b = threadIdx.x; b < TILE_BINS; b += 1024) {
        const float n = (float)cell[2 * b], s = (float)cell[2 * b + 1];
        if (n > 0.0f)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                         :: "l"(out + b), "f"(s), "f"(s), "f"(s), "f"(n) : "memory");
    }
}
'''

N.init(0)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << logn
nbins = 256 * 8192

# address distribution: a real flame
gnm = samples.g6f()
gprof = profile.wrap(dict(width=1920, height=1080, spp=200, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=3)
rdr = render.Renderer(gnm, gprof)
dim = rmgr.fb.set_dim(1920, 1080)
rmgr._copy(rdr, gnm); rmgr._interp(rdr, gnm, dim, tc, 0.0); rmgr._iter(rdr, gnm, gprof, dim, tc)
rmgr.stream_a.synchronize()
count = N.from_device(rmgr.fb.d_front, (dim.ah * dim.astride, 4), np.float32)[:nbins, 3].astype(np.float64)
rmgr.fb.free()
cdf = np.cumsum(count); cdf /= cdf[-1]
rs = np.random.RandomState(1)
bins = np.searchsorted(cdf, rs.rand(n)).astype(np.uint32).clip(0, nbins - 1)
records = (bins << np.uint32(8)) | rs.randint(0, 256, n).astype(np.uint32)

names, hdrs = itergen.load_headers()
mod = N.Module(SRC, 'deferred.cu', hdrs, names, ['--gpu-architecture=sm_100a', '--std=c++17'])
sms = N.device_info(0)['sm_count']
hist = N.DeviceBuffer(16 * nbins)
rec_a, rec_b, rec_c = N.to_device(records), N.DeviceBuffer(4 * n), N.DeviceBuffer(4 * n)
srt = Sorter(n)


def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        e0, e1 = N.Event(), N.Event()
        e0.record(None); fn(); e1.record(None); e1.synchronize()
        best = min(best, e1.time_since(e0))
    return best


u = lambda b: C.c_uint64(b.ptr)
res = dict(records=n, nbins=nbins, tiles=256, address_distribution='G6F 1080p histogram (cb_iter, 200 spp)',
           hottest_bin_share=float(count.max() / count.sum()))
N.fill32(hist, 4 * nbins, 0)
t = timed(lambda: mod.launch('k_direct', (sms * 8,), (256,), [u(hist), u(rec_a), C.c_uint(n)]))
res['direct_red_v4'] = dict(ms=t, per_s=n / t * 1e3)
direct = N.from_device(hist, (nbins, 4), np.float32)[:, 3].astype(np.float64) / 4      # 4 timed runs
t = timed(lambda: mod.launch('k_append', (sms * 8,), (256,), [u(rec_c), u(rec_a), C.c_uint(n)]))
res['append'] = dict(ms=t, per_s=n / t * 1e3)
t = timed(lambda: srt.sort(rec_b, rec_a, n, lo_bit=21))
res['sort_pass'] = dict(ms=t, per_s=n / t * 1e3)
starts = srt.digit_starts().astype(np.int64)
assert starts[-1] == n
CH = 131072
chunks = []
for tile in range(256):
    for lo in range(starts[tile], starts[tile + 1], CH):
        chunks.append((tile, lo, min(lo + CH, starts[tile + 1])))
chunks = np.array(chunks, np.int32)
d_chunks = N.to_device(chunks)
launch_replay = lambda: mod.launch('k_replay', (len(chunks),), (1024,), [u(hist), u(rec_b), u(d_chunks)], dyn_smem=65536)
N.fill32(hist, 4 * nbins, 0)
t = timed(launch_replay, reps=1)
got = N.from_device(hist, (nbins, 4), np.float32)[:, 3].astype(np.float64)
assert np.array_equal(got, direct), 'replayed histogram differs from the direct one'
t = min(t, timed(launch_replay))
res['replay'] = dict(ms=t, per_s=n / t * 1e3, chunks=len(chunks), records_per_chunk=CH,
                     largest_tile_share=float(np.diff(starts).max()) / n)
tot = res['append']['ms'] + res['sort_pass']['ms'] + res['replay']['ms']
res['deferred_total'] = dict(ms=tot, per_s=n / tot * 1e3,
                             vs_direct=res['direct_red_v4']['ms'] / tot)
print(json.dumps(res, indent=1))
