#!/usr/bin/env python
"""
bench.py -- the reference's headline metric on its headline config.

Metric: IFS iterations/s (BASELINE.json).  Workload (N=1): configs[1], the
1080p still of the 6-xform + final-xform flame G6F at 2000 spp, synthetic
genome (cuburn_b200/samples.py).  One *step* renders one frame on the device:
interpolate the packed genome, run the chaos game, run the default filter
chain (yuv, bilateral, logscale, smearclip) and convert to RGBA8.

  value   samples / device time of the step, inputs resident in HBM
  e2e     the same through RenderManager.queue_frame -- H2D of the packed
          genome + palettes from pinned memory, D2H of the finished frame
  N > 1   one process per GPU (torchrun); every GPU runs the config's sample
          count with its own RNG streams (weak scaling: the still gets N x the
          samples), the float4 histograms are summed onto rank 0 with one NCCL
          reduce, rank 0 filters; time = max over ranks, device events.

  --impl reference   the CPU oracle's chaos game (oracle/chaos.c, OpenMP on all
          host cores) on a bounded sample of the same workload.  cuburn itself
          is GPU-only Python 2 + PyCUDA and cannot run here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # BASELINE.json configs[1]: the default; weak-scaled (every GPU runs 2000 spp)
    'still1080': dict(genome='G6F', width=1920, height=1080, spp=2000, scaling='weak',
                      label='1080p still, G6F (6 xforms + final xform, 12 variation types), '
                            '2000 spp per GPU'),
    # BASELINE.json configs[2]: one 4K frame at 4000 spp split over the GPUs (strong)
    'still4k': dict(genome='G6F', width=3840, height=2160, spp=4000, scaling='strong',
                    label='3840x2160 still, G6F, 4000 spp in total, samples split over the GPUs'),
    # BASELINE.json configs[4]: 8K, 24 heavy xforms; the 512 MiB histogram does not fit L2
    'still8k': dict(genome='G24H', width=7680, height=4320, spp=2000, scaling='strong',
                    label='7680x4320 still, G24H (24 xforms, heavy variations), 2000 spp in '
                          'total, samples split over the GPUs'),
}
CONFIG = dict(WORKLOADS['still1080'])
# profiles/r01_final_cb_iter.md (ncu --set full, still1080 workload): 37.69 MB read +
# 0.20 MB written per cb_iter launch -- the first touch of the histogram; the 66 GB of
# atomic payload never leaves L2
NCU_DRAM_BYTES_PER_LAUNCH = 37.9e6
UNIT = 65536


def read_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            return float(json.load(fp)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device=0):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summarise the samples taken in [t0, t1] (all samples if none fall inside)."""
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        inside = [ln for t, ln in self.lines
                  if t0 is None or (t0 - 0.02 <= t <= t1 + 0.12)]
        window = 'timed region'
        if not inside:
            inside, window = [ln for _, ln in self.lines], 'warm-up + timed region'
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown',
                                  'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm), window=window)


RED_SRC = r'''
#include "mwc.cuh"
// uniformly scattered 16-byte reductions into an L2-resident float4 grid, nothing else
extern "C" __global__ void __launch_bounds__(256)
red_peak(float4 *hist, mwc_st *seeds, unsigned int nbins, int rounds) {
    int g = blockIdx.x * 256 + threadIdx.x;
    mwc_st rng = seeds[g];
    for (int r = 0; r < rounds; r++) {
        unsigned int bin = __umulhi(mwc_next(rng), nbins);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(hist + bin), "f"(0.25f), "f"(0.5f), "f"(0.75f), "f"(1.0f) : "memory");
    }
    seeds[g] = rng;
}
'''


def measure_red_peak(N, hist_ptr, nbins, seeds_ptr, sms):
    """Live microbenchmark: the scattered float4-reduction rate this GPU sustains into
    a grid of the workload's size -- the roofline the accumulation runs against."""
    import ctypes as C
    from cuburn_b200.code import itergen
    names, hdrs = itergen.load_headers()
    mod = N.Module(RED_SRC, 'red_peak.cu', hdrs, names,
                   ['--gpu-architecture=sm_100a', '--std=c++17'])
    grid, rounds, best = sms * 4, 2048, 1e9
    for _ in range(4):
        e0, e1 = N.Event(), N.Event()
        e0.record(None)
        mod.launch('red_peak', (grid,), (256,),
                   [C.c_uint64(hist_ptr), C.c_uint64(seeds_ptr), C.c_uint(nbins), C.c_int(rounds)])
        e1.record(None)
        e1.synchronize()
        best = min(best, e1.time_since(e0))
    return grid * 256 * rounds / (best * 1e-3)


def frame_setup(n_gpus, rank, seed=1):
    from cuburn_b200 import _native as N, samples, profile, render
    gnm = samples.GENOMES[CONFIG['genome']]()
    spp = CONFIG['spp'] * (n_gpus if CONFIG['scaling'] == 'weak' else 1)
    prof = dict(width=CONFIG['width'], height=CONFIG['height'],
                spp=spp, frame_width=0, start=1, end=2)
    gprof = profile.wrap(prof, gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=seed, rank=rank, world=n_gpus)
    rdr = render.Renderer(gnm, gprof)
    return N, gnm, gprof, tc, rmgr, rdr


def cpu_chaos_rate(nsamples, nthreads=0, seed=1):
    """Oracle chaos game on the host cores: (iterations/s, threads used)."""
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    gnm = samples.GENOMES[CONFIG['genome']]()
    w, h = CONFIG['width'], CONFIG['height']
    ev = R.GenomeEval(gnm, w, h, 1.5 / 720, 0.0)
    seeds = mwc.make_seeds(32768, host_seed=seed)
    pal, seeds = R.palette_table(gnm, ev.ts, ev.td, seeds)
    cores = nthreads or (os.cpu_count() or 1)
    R.iterate(ev, pal, seeds, 2 ** 20, nthreads=cores)          # warm
    t = time.perf_counter()
    R.iterate(ev, pal, seeds, nsamples, nthreads=cores)
    return nsamples / (time.perf_counter() - t), cores


def cpu_filter_rates(w=320, h=180):
    """The numpy restatement of the default filter chain on a small synthetic histogram:
    algorithmic GB/s per filter (SURVEY 8(d) byte counts), one host core."""
    from cuburn_b200 import _native as N
    from oracle import filters_ref as F
    dim = N.calc_dim(w, h)
    rs = np.random.RandomState(5)
    dens = rs.gamma(0.3, 40.0, size=(dim.ah, dim.astride)).astype(np.float32)
    dens[rs.rand(dim.ah, dim.astride) < 0.3] = 0
    hist = np.empty((dim.ah, dim.astride, 4), np.float32)
    for ch in range(3):
        hist[..., ch] = dens * rs.rand(dim.ah, dim.astride).astype(np.float32)
    hist[..., 3] = dens
    nbins = dim.ah * dim.astride
    k1, k2 = F.logscale_consts(4, 0.5, w, h, 256)
    stages = (('yuv_to_rgb', 32, lambda p: F.yuv_to_rgb(p)),
              ('bilateral (8 directions)', 512, lambda p: F.bilateral(p, 1920)),
              ('logscale', 32, lambda p: F.logscale(p, k1, k2)),
              ('smearclip', 208, lambda p: F.smearclip(p, 0.7, 4, 0.01)))
    out, pix, total = {}, hist, 0.0
    for name, bytes_per_bin, fn in stages:
        t = time.perf_counter()
        pix = fn(pix)
        dt = time.perf_counter() - t
        total += dt
        out[name] = bytes_per_bin * nbins / dt / 1e9
    out['chain'] = 784.0 * nbins / total / 1e9
    return {'unit': 'GB/s (algorithmic bytes)', 'cores': 1, 'kind': 'port', 'value': out,
            'sample': 'default chain without output conversion on a %dx%d synthetic histogram '
                      '(%d bins)' % (w, h, nbins)}


def run_reference(args):
    """--impl reference: the CPU restatement, all host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    sample_spp = 100                       # 2.07e8 samples per step, ~1-3 s on 8 cores
    n = CONFIG['width'] * CONFIG['height'] * sample_spp
    rates = []
    for i in range(args.warmup + args.steps):
        r, cores = cpu_chaos_rate(n)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    line = {
        'impl': 'reference', 'metric': 'ifs_iterations_per_second', 'value': value,
        'unit': 'iterations/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * n / value, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '1080p still, G6F (6 xforms + final, 12 variation types), '
                               '2000 spp; CPU arm runs a %d-spp sample per step' % sample_spp},
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': cores,
                         'kind': 'port',
                         'sample': '%d samples (1080p x %d spp) of the chaos game per step'
                                   % (n, sample_spp)},
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--filter-shard', default='band', choices=['band', 'root'],
                    help='multi-GPU: filter row bands on every GPU (all-reduce + gather) '
                         'or the whole frame on the root (reduce)')
    ap.add_argument('--collectives', default='torch', choices=['torch', 'native'],
                    help="multi-GPU exchange through torch.distributed or through the "
                         "library's own NCCL communicator (cb_hist_reduce / cb_band_gather)")
    ap.add_argument('--workload', default='still1080', choices=sorted(WORKLOADS),
                    help='still1080 = BASELINE configs[1] (default); still4k = configs[2]')
    args = ap.parse_args()
    CONFIG.clear()
    CONFIG.update(WORKLOADS[args.workload])
    args.warmup = max(args.warmup, 3) if args.impl == 'native' else args.warmup

    if args.impl == 'reference':
        return run_reference(args)

    from cuburn_b200 import multigpu
    rank, world, local = multigpu.env_rank_world()
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    n_gpus = world
    dist = None
    if n_gpus > 1:
        import torch
        import torch.distributed as dist
        multigpu.init_process_group('nccl')

    from cuburn_b200 import _native as N
    N.init(local)
    N_, gnm, gprof, tc, rmgr, rdr = frame_setup(n_gpus, rank)
    reducer = None
    if n_gpus > 1:
        banded = args.filter_shard == 'band'
        comm = multigpu.NativeComm(rank, n_gpus) if args.collectives == 'native' else None
        reducer = multigpu.HistReducer(root=None if banded else 0, comm=comm)
        rmgr.hist_hook = reducer
        if banded:
            rmgr.band_filter = multigpu.BandFilter(rank, n_gpus, root=0, comm=comm or True)
    dim = rmgr.fb.set_dim(gprof.width, gprof.height)
    td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
    ts = tc - 0.5 * td
    total, first, mine = rmgr.frame_samples(gprof, dim, tc)

    # L2 flush buffer: 512 MiB > 126 MiB L2, written between timed steps
    flush = N.DeviceBuffer(512 << 20)

    def l2_flush():
        N.fill32(flush, (512 << 20) // 4, 0, rmgr.stream_a)

    def barrier():
        rmgr.stream_a.synchronize()
        rmgr.stream_b.synchronize()
        N.check(N.lib().cb_device_sync())
        if dist is not None:
            dist.barrier()

    def device_step(ev0, ev_iter0, ev_iter1, ev1):
        """One frame with inputs resident: interp + iterate (+reduce) + filters + convert."""
        s = rmgr.stream_a
        ev0.record(s)
        rmgr._interp(rdr, gnm, dim, ts, td)
        ev_iter0.record(s)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        ev_iter1.record(s)
        if reducer is not None:
            reducer(rmgr.fb, dim, s)
        if rmgr.band_filter is not None or rank == 0:
            rmgr._filter(rdr, gprof, dim, tc)
        if rank == 0:
            rdr.out.convert(rmgr.fb, gprof, dim, s)
        ev1.record(s)

    # ---- device-resident timing ----------------------------------------------------
    rmgr._copy(rdr, gnm)
    evs = [[N.Event() for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        device_step(*[N.Event() for _ in range(4)])
        l2_flush()
    barrier()
    wall0 = time.perf_counter()
    step_ms, iter_ms = [], []
    for k in range(args.steps):
        l2_flush()
        barrier()
        device_step(*evs[k])
        barrier()
        step_ms.append(evs[k][3].time_since(evs[k][0]))
        iter_ms.append(evs[k][2].time_since(evs[k][1]))
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    my_ms = float(np.sum(step_ms))
    if dist is not None:
        import torch
        t = torch.tensor([my_ms, float(np.sum(iter_ms))], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        my_ms, iter_total = float(t[0]), float(t[1])
    else:
        iter_total = float(np.sum(iter_ms))
    ms_per_step = my_ms / args.steps
    value = total / (ms_per_step * 1e-3)

    # ---- end to end through queue_frame (host buffers in, host frame out) --------------
    e2e_ms = []
    for k in range(args.warmup + args.steps):
        l2_flush()
        barrier()
        evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc, copy=True)
        evt.synchronize()
        barrier()
        if k >= args.warmup:
            e2e_ms.append(evt.time())
    e2e_my = float(np.sum(e2e_ms))
    if dist is not None:
        import torch
        t = torch.tensor([e2e_my], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_my = float(t[0])
    e2e_value = total / (e2e_my / args.steps * 1e-3)
    pk = rdr.packer
    h2d = 2 * pk.nrows * 32 * 4 + len(gnm['palette']) * 256 * 16 + 32 * 4 \
        + pk.nrows * 4 + pk.program_array().nbytes
    d2h = int(buf.nbytes)

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel (cb_iter) -----------------------------------
    peak, peak_src = read_peaks()
    iter_ms_mean = iter_total / args.steps
    packed = rmgr._use_packed(dim.ah * dim.astride)
    # one 16-byte float4 accumulate per sample (8-byte packed cell beyond 1.5 x L2)
    algo_bytes = (8.0 if packed else 16.0) * mine
    achieved = algo_bytes / (iter_ms_mean * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'cb_iter', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': peak_src,
                'note': 'algorithmic bytes = 16 B float4 accumulate per sample; the 33 MiB '
                        'histogram is L2-resident so the true bound is L2 atomic / issue rate, '
                        'see profiles/',
                'kernel_ms': iter_ms_mean,
                'samples_per_second_kernel': mine / (iter_ms_mean * 1e-3)}
    nbins = dim.ah * dim.astride
    red_peak = measure_red_peak(N, rmgr.fb.d_left.ptr, nbins, rmgr.fb.d_seeds.ptr,
                                N.device_info(local)['sm_count'])
    kernel_rate = mine / (iter_ms_mean * 1e-3)
    roofline['atomic'] = {
        'what': 'scattered red.global.add.v4.f32 into an L2-resident float4 grid of the '
                'same size, measured in this run (the binding resource: L2 slice atomic rate)',
        'achieved': kernel_rate, 'peak': red_peak, 'unit': 'reductions/s',
        'frac': kernel_rate / red_peak}
    # dram__bytes_read.sum + dram__bytes_write.sum of one cb_iter launch, from the
    # committed ncu capture (profiles/); None until a capture exists for this kernel
    roofline['traffic'] = NCU_DRAM_BYTES_PER_LAUNCH if args.workload == 'still1080' else None
    if packed:
        roofline['note'] = ('algorithmic bytes = 8 B packed-u64 accumulate per sample; the grid is '
                            'far larger than L2, so the bound is HBM sector read-modify-write')
    filt_ms = ms_per_step - iter_ms_mean
    roofline_filters = {'bound': 'hbm', 'stage': 'interp + filter chain + convert',
                        'algorithmic_bytes_per_bin': 804,
                        'achieved': 804.0 * nbins / (filt_ms * 1e-3) / 1e9, 'peak': peak,
                        'unit': 'GB/s', 'frac': 804.0 * nbins / (filt_ms * 1e-3) / 1e9 / peak,
                        'ms': filt_ms}

    cpu = None
    if not args.no_cpu_baseline and n_gpus == 1:      # rank 0 at N = 1 only
        n_cpu = 1920 * 1080 * 500                               # ~10-20 s of CPU work
        rate, cores = cpu_chaos_rate(n_cpu)
        cpu = {'value': rate, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
               'sample': '%d samples (500 spp worth of a 1080p frame) of the same genome, chaos game only'
                         % n_cpu}

    cpu_filters = cpu_filter_rates() if cpu is not None else None

    # 3 interp + fill + iter + unswizzle + yuv + 8 x 3 bilateral + logscale + 6 smearclip + convert
    launches_per_step = 3 + 1 + 1 + 1 + 1 + 8 * 3 + 1 + 6 + 1
    line = {
        'metric': 'ifs_iterations_per_second', 'value': value, 'unit': 'iterations/s',
        'n_gpus': n_gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': CONFIG['scaling'],
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': CONFIG['label'] + ', default filter chain, RGBA8 out',
                   'samples_per_step': total, 'frames_per_second': 1e3 / ms_per_step,
                   'l2': 'L2 flushed between timed steps (512 MiB fill)',
                   'timing': 'CUDA events on the launching stream, per step, max over ranks',
                   'collectives': args.collectives if n_gpus > 1 else None,
                   'parallelism': ('single GPU' if n_gpus == 1 else
                                   'independent RNG streams per GPU + NCCL all-reduce; filter '
                                   'chain sharded by row bands, gathered on the root'
                                   if rmgr.band_filter is not None else
                                   'independent RNG streams per GPU + NCCL reduce; root filters')},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'iterations/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_my / args.steps},
        'gpu_launches': launches_per_step * args.steps,
        'roofline': roofline, 'roofline_filters': roofline_filters,
        'cpu_baseline': cpu, 'cpu_baseline_filters': cpu_filters,
        'wall_s_timed_region': wall,
    }
    if reducer is not None:
        line['config']['nccl_reduce_ms'] = reducer.mean_reduce_ms(args.steps)
        if rmgr.band_filter is not None:
            line['config']['nccl_gather_ms'] = rmgr.band_filter.mean_gather_ms(args.steps)
    print(json.dumps(line))


if __name__ == '__main__':
    main()
