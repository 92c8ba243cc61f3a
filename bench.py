#!/usr/bin/env python
"""
bench.py -- the reference's headline metric on its headline config, plus the other
BASELINE.json configurations as `extra` records of the same JSON line.

Metric: IFS iterations/s (BASELINE.json).  Headline workload (`config.workload`):
configs[1], the 1080p still of the 6-xform + final-xform flame G6F at 2000 spp, synthetic
genome (cuburn_b200/samples.py).  One *step* renders one frame on the device: interpolate
the packed genome, run the chaos game, run the default filter chain (yuv, bilateral,
logscale, smearclip) and convert to RGBA8.

  value   samples / device time of the step, inputs resident in HBM
  e2e     the same through RenderManager.queue_frame -- H2D of the packed genome +
          palettes from pinned memory, D2H of the finished frame
  N > 1   one process per GPU (torchrun).  still1080 is weak-scaled (every GPU runs the
          config's sample count with its own RNG streams, so the still gets N x the
          samples); the float4 histograms are combined with one NCCL all-reduce, every
          GPU filters, converts and copies out one band of rows (exact halos); time = max
          over ranks, device events.
  extra   configs[2] (`still4k`, strong-scaled: 4000 spp in total split over the GPUs),
          configs[4] (`still8k`, N = 1 only) and configs[3] (`anim1080`, motion-blurred
          frames round-robin over the GPUs, frames/s), each with its own value / e2e.

  --impl reference   the CPU oracle's chaos game (oracle/chaos.c, OpenMP on all host
          cores) on the same workload; the sample per step is sized so that the whole run
          stays within a few minutes (the full 2000 spp when --steps + --warmup <= 4).
          cuburn itself is GPU-only Python 2 + PyCUDA and cannot run here.
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
                            '2000 spp per GPU, default filter chain, RGBA8 out'),
    # BASELINE.json configs[2]: one 4K frame at 4000 spp split over the GPUs (strong)
    'still4k': dict(genome='G6F', width=3840, height=2160, spp=4000, scaling='strong',
                    label='3840x2160 still, G6F, 4000 spp in total, samples split over the '
                          'GPUs, default filter chain, RGBA8 out'),
    # BASELINE.json configs[4]: 8K, 24 heavy xforms; the 512 MiB histogram does not fit L2
    'still8k': dict(genome='G24H', width=7680, height=4320, spp=2000, scaling='strong',
                    label='7680x4320 still, G24H (24 xforms, heavy variations), 2000 spp in '
                          'total, samples split over the GPUs, default filter chain, RGBA8 out'),
    # BASELINE.json configs[3]: 1080p animation with motion blur, frames over the GPUs
    'anim1080': dict(genome='G6F', width=1920, height=1080, spp=2000, scaling='weak',
                     animated=True,
                     label='1080p 24 fps animation of G6F (rotating pre-affines), 2000 spp, '
                           'motion blur over 1024 temporal samples, frames round-robin over '
                           'the GPUs'),
}
UNIT = 65536


def read_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            return float(json.load(fp)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def read_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel,
    from the committed ncu capture of this workload (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(path):
        with open(path) as fp:
            rec = json.load(fp).get(workload)
        if rec:
            return rec['dram_bytes_per_launch'], rec['source']
    return None, None


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


def workload_genome(cfg):
    from cuburn_b200 import samples
    if cfg.get('animated'):
        return samples.g6f(animated=True)
    return samples.GENOMES[cfg['genome']]()


def cpu_chaos_rate(cfg, nsamples, nthreads=0, seed=1):
    """Oracle chaos game on the host cores: (iterations/s, threads used)."""
    from cuburn_b200 import mwc
    from oracle import flame_ref as R
    gnm = workload_genome(dict(cfg, animated=False))
    w, h = cfg['width'], cfg['height']
    ev = R.GenomeEval(gnm, w, h, 1.5 / 720, 0.0)
    seeds = mwc.make_seeds(32768, host_seed=seed)
    pal, seeds = R.palette_table(gnm, ev.ts, ev.td, seeds)
    cores = nthreads or (os.cpu_count() or 1)
    R.iterate(ev, pal, seeds, 2 ** 20, nthreads=cores)          # warm
    t = time.perf_counter()
    R.iterate(ev, pal, seeds, nsamples, nthreads=cores)
    return nsamples / (time.perf_counter() - t), cores


def cpu_filter_rates(w=1920, h=1080):
    """The numpy restatement of the default filter chain on a synthetic 1080p histogram, the
    bilateral passes (97 % of the time) on row strips in all host threads: algorithmic GB/s
    per filter (SURVEY 8(d) byte counts)."""
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
    cores = os.cpu_count() or 1
    k1, k2 = F.logscale_consts(4, 0.5, w, h, 256)
    stages = (('yuv_to_rgb', 32, lambda p: F.yuv_to_rgb(p)),
              ('bilateral (8 directions)', 512, lambda p: F.bilateral(p, w, threads=cores)),
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
    return {'unit': 'GB/s (algorithmic bytes)', 'cores': cores, 'kind': 'port', 'value': out,
            'seconds': total,
            'sample': 'default chain without output conversion on a %dx%d synthetic histogram '
                      '(%d bins); numpy restatement, bilateral passes on row strips in %d '
                      'threads, pointwise stages on one' % (w, h, nbins, cores)}


def run_reference(args):
    """--impl reference: the CPU restatement of the chaos game, all host cores, on the
    headline workload; per-step sample sized to a ~150 s total budget."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    w, h, spp = cfg['width'], cfg['height'], cfg['spp']
    probe_rate, cores = cpu_chaos_rate(cfg, 2 ** 24)
    nsteps = max(1, args.warmup + args.steps)
    budget_s = 150.0 / nsteps
    sample_spp = int(max(25, min(spp, budget_s * probe_rate / (w * h))))
    n = w * h * sample_spp
    rates = []
    for i in range(args.warmup + args.steps):
        r, cores = cpu_chaos_rate(cfg, n)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    same = sample_spp == spp
    sample = ('the whole workload: %d samples (%dx%d x %d spp) of the chaos game per step'
              % (n, w, h, spp)) if same else \
        ('%d samples (%dx%d x %d of the %d spp) of the chaos game per step, sized so that '
         '%d steps fit ~150 s' % (n, w, h, sample_spp, spp, nsteps))
    line = {
        'impl': 'reference', 'metric': 'ifs_iterations_per_second', 'value': value,
        'unit': 'iterations/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * n / value, 'higher_is_better': True,
        'scaling': cfg['scaling'], 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['label']},
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': cores,
                         'kind': 'port', 'sample': sample,
                         'what': 'oracle/chaos.c (scalar C + OpenMP): xform choice, variations, '
                                 'final xform, camera, palette, float4 accumulation; no filters'},
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


class Bench(object):
    """One process's view of the job: device, ranks, timing helpers."""
    def __init__(self, args):
        from cuburn_b200 import multigpu
        self.args = args
        self.rank, self.world, self.local = multigpu.env_rank_world()
        if self.world != args.gpus and self.world > 1:
            raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, self.world))
        self.dist = self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            multigpu.init_process_group('nccl')
            self.dist, self.torch = dist, torch
        from cuburn_b200 import _native as N
        N.init(self.local)
        self.N = N
        self.comm = multigpu.NativeComm(self.rank, self.world) \
            if (self.world > 1 and args.collectives == 'native') else None
        # L2 flush buffer: 512 MiB > 126 MiB L2, written between timed steps
        self.flush = N.DeviceBuffer(512 << 20)

    def barrier(self, rmgr=None):
        if rmgr is not None:
            rmgr.stream_a.synchronize()
            rmgr.stream_b.synchronize()
        self.N.check(self.N.lib().cb_device_sync())
        if self.dist is not None:
            self.dist.barrier()

    def l2_flush(self, stream):
        self.N.fill32(self.flush, (512 << 20) // 4, 0, stream)

    def max_over_ranks(self, *values):
        if self.dist is None:
            return [float(v) for v in values]
        t = self.torch.tensor([float(v) for v in values], device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, value):
        if self.dist is None:
            return float(value)
        t = self.torch.tensor([float(value)], device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])


def run_still(b, name, steps, warmup, sampler=None):
    """Device-resident and end-to-end timing of one still workload on all ranks."""
    from cuburn_b200 import multigpu, profile, render
    N, args, cfg = b.N, b.args, WORKLOADS[name]
    gnm = workload_genome(cfg)
    spp = cfg['spp'] * (b.world if cfg['scaling'] == 'weak' else 1)
    gprof = profile.wrap(dict(width=cfg['width'], height=cfg['height'], spp=spp,
                              frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=1, rank=b.rank, world=b.world)
    rmgr.hot_bins = {'auto': 'auto', 'off': False, 'on': True}[args.hot_bins]
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(gprof.width, gprof.height)
    reducer = shared = None
    if b.world > 1:
        banded = args.filter_shard == 'band'
        reducer = multigpu.HistReducer(root=None if banded else 0, comm=b.comm,
                                       integer_sums=True)
        rmgr.hist_hook = reducer
        if banded:
            if args.band_output == 'shared':
                shared = multigpu.SharedFrame(rdr.out.shape(dim), rdr.out.dtype, b.rank, b.world,
                                              barrier=b.dist.barrier)
            rmgr.band_filter = multigpu.BandFilter(b.rank, b.world, root=0, comm=b.comm or True,
                                                   shared=shared)
    td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
    ts = tc - 0.5 * td
    total, first, mine = rmgr.frame_samples(gprof, dim, tc)

    def device_step(ev0, ev_iter0, ev_iter1, ev1):
        """One frame with inputs resident: interp + iterate (+reduce) + filters + convert."""
        s = rmgr.stream_a
        ev0.record(s)
        rmgr._interp(rdr, gnm, dim, ts, td)
        ev_iter0.record(s)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        ev_iter1.record(s)
        if reducer is not None:
            rmgr._combine(dim)              # the reducer (rmgr.hist_hook) + the 1/255 scale
        if rmgr.band_filter is not None or b.rank == 0:
            rmgr._filter(rdr, gprof, dim, tc)
        if shared is not None:
            rdr.out.convert(rmgr.fb, gprof, dim, s,
                            rows=rmgr.band_filter.output_rows(dim, rmgr.fb.gutter))
        elif b.rank == 0:
            rdr.out.convert(rmgr.fb, gprof, dim, s)
        ev1.record(s)

    rmgr._copy(rdr, gnm)
    evs = [[N.Event() for _ in range(4)] for _ in range(steps)]
    for _ in range(warmup):
        device_step(*[N.Event() for _ in range(4)])
        b.l2_flush(rmgr.stream_a)
    b.barrier(rmgr)
    if sampler is not None:
        sampler.start()
    wall0 = time.perf_counter()
    launches0 = N.launch_count()
    step_ms, iter_ms = [], []
    for k in range(steps):
        b.l2_flush(rmgr.stream_a)
        b.barrier(rmgr)
        device_step(*evs[k])
        b.barrier(rmgr)
        step_ms.append(evs[k][3].time_since(evs[k][0]))
        iter_ms.append(evs[k][2].time_since(evs[k][1]))
    launches = N.launch_count() - launches0 - steps          # minus the L2 flush fills
    wall1 = time.perf_counter()
    clocks = sampler.stop(wall0, wall1) if sampler is not None else None
    my_ms, iter_total = b.max_over_ranks(np.sum(step_ms), np.sum(iter_ms))
    ms_per_step = my_ms / steps
    value = total / (ms_per_step * 1e-3)

    # ---- end to end through queue_frame (host buffers in, host frame out) --------------
    e2e_ms = []
    for k in range(warmup + steps):
        b.l2_flush(rmgr.stream_a)
        b.barrier(rmgr)
        evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc, copy=True)
        evt.synchronize()
        b.barrier(rmgr)
        if k >= warmup:
            e2e_ms.append(evt.time())
    e2e_my, = b.max_over_ranks(np.sum(e2e_ms))
    e2e_value = total / (e2e_my / steps * 1e-3)
    pk = rdr.packer
    h2d = 2 * pk.nrows * 32 * 4 + len(gnm['palette']) * 256 * 16 + 32 * 4 \
        + pk.nrows * 4 + pk.program_array().nbytes
    if shared is not None:
        r0, r1 = rmgr.band_filter.output_rows(dim, rmgr.fb.gutter)
        d2h = int(b.sum_over_ranks((r1 - r0) * buf.strides[0]))
    else:
        d2h = int(buf.nbytes)

    nbins = dim.ah * dim.astride
    res = dict(
        workload=name, label=cfg['label'], scaling=cfg['scaling'], value=value,
        ms_per_step=ms_per_step, frames_per_second=1e3 / ms_per_step,
        samples_per_step=total, samples_this_gpu=mine, iter_ms=iter_total / steps,
        e2e=dict(value=e2e_value, unit='iterations/s', h2d_bytes_per_step=h2d,
                 d2h_bytes_per_step=d2h, ms_per_step=e2e_my / steps),
        launches_per_step=launches / float(steps), wall_s=wall1 - wall0, clocks=clocks,
        accumulate='packed u64 cells' if rmgr._use_packed(nbins) else 'float4 reductions',
        hot_bins=bool(getattr(rmgr, 'last_iter_hot', False)), nbins=nbins,
        parallelism=('single GPU' if b.world == 1 else
                     'independent RNG streams per GPU + NCCL all-reduce; every GPU filters, '
                     'converts and copies out one band of rows' if shared is not None else
                     'independent RNG streams per GPU + NCCL all-reduce; filter chain sharded '
                     'by row bands, gathered on the root' if rmgr.band_filter is not None else
                     'independent RNG streams per GPU + NCCL reduce; root filters'))
    if reducer is not None:
        res['nccl_reduce_ms'] = reducer.mean_reduce_ms(steps)
        if rmgr.band_filter is not None and shared is None:
            res['nccl_gather_ms'] = rmgr.band_filter.mean_gather_ms(steps)
    res['_state'] = (rmgr, rdr, dim)        # for the roofline microbenchmark; dropped later
    return res


def run_anim(b, name, frames_per_gpu, warmup):
    """configs[3]: motion-blurred 1080p frames, frame k on GPU k mod N, through
    queue_frame (double-buffered: frame k+1's upload and interpolation overlap frame k)."""
    from cuburn_b200 import profile, render
    N, cfg = b.N, WORKLOADS[name]
    gnm = workload_genome(cfg)
    prof = dict(width=cfg['width'], height=cfg['height'], spp=cfg['spp'], fps=24, duration=30.0,
                frame_width=1.0)
    gprof = profile.wrap(prof, gnm)
    times = [t for _, ts in profile.enumerate_times(gprof) for t in ts]
    mine = times[b.rank::b.world][:warmup + frames_per_gpu]
    rmgr = render.RenderManager(seed=1)
    rdr = render.Renderer(gnm, gprof)
    pending = []
    for k, tc in enumerate(mine):
        if k == warmup:
            for evt, _ in pending:
                evt.synchronize()
            pending = []
            b.barrier(rmgr)
            e0 = N.Event().record(rmgr.stream_a)
            t0 = time.perf_counter()
        pending.append(rmgr.queue_frame(rdr, gnm, gprof, tc, frame_seed=1000 + k))
        if len(pending) > 2:
            pending.pop(0)[0].synchronize()
    for evt, _ in pending:
        evt.synchronize()
    wall = time.perf_counter() - t0
    wall, = b.max_over_ranks(wall)
    nframes = (len(mine) - warmup) * b.world
    fps = nframes / wall
    dim = rmgr.fb.calc_dim(cfg['width'], cfg['height'])
    samples = cfg['spp'] * cfg['width'] * cfg['height']
    rmgr.fb.free()
    return dict(workload=name, label=cfg['label'], scaling='weak', frames=nframes,
                frames_per_second=fps, ms_per_frame_per_gpu=1e3 * wall / (len(mine) - warmup),
                value=fps * samples, unit='iterations/s',
                timing='host wall clock over queue_frame calls (H2D of every genome, D2H of '
                       'every RGBA8 frame inside), max over ranks',
                d2h_bytes_per_frame=cfg['width'] * cfg['height'] * 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true',
                    help='only the headline workload (no still4k / still8k / anim1080 records)')
    ap.add_argument('--filter-shard', default='band', choices=['band', 'root'],
                    help='multi-GPU: filter row bands on every GPU or the whole frame on the root')
    ap.add_argument('--band-output', default='shared', choices=['shared', 'gather'],
                    help='multi-GPU, banded filtering: every GPU converts its band and copies it '
                         'into one shared pinned host frame, or bands are gathered on the root')
    ap.add_argument('--collectives', default='torch', choices=['torch', 'native'],
                    help="multi-GPU exchange through torch.distributed or through the "
                         "library's own NCCL communicator (cb_hist_reduce / cb_band_gather)")
    ap.add_argument('--hot-bins', default='auto', choices=['auto', 'off', 'on'],
                    help='hot-bin privatisation: probe per genome (default), never, always')
    ap.add_argument('--workload', default='still1080',
                    choices=['still1080', 'still4k', 'still8k'],
                    help='headline workload: still1080 = BASELINE configs[1] (default)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'native' else args.warmup

    if args.impl == 'reference':
        return run_reference(args)

    b = Bench(args)
    N = b.N
    sampler = ClockSampler(b.local) if b.rank == 0 else None
    head = run_still(b, args.workload, args.steps, args.warmup, sampler)
    rmgr, rdr, dim = head.pop('_state')

    # ---- roofline of the dominant kernel (cb_iter), rank 0 ----------------------------
    roofline = roofline_filters = None
    if b.rank == 0:
        peak, peak_src = read_peaks()
        nbins, mine, iter_ms = head['nbins'], head['samples_this_gpu'], head['iter_ms']
        packed = head['accumulate'].startswith('packed')
        kernel_rate = mine / (iter_ms * 1e-3)
        red_peak = measure_red_peak(N, rmgr.fb.d_left.ptr, nbins, rmgr.fb.d_seeds.ptr,
                                    N.device_info(b.local)['sm_count'])
        algo = 8.0 if packed else 16.0
        gbs = algo * kernel_rate / 1e9
        traffic, traffic_src = read_traffic(args.workload)
        roofline = {
            'kernel': 'cb_iter', 'bound': 'hbm' if packed else 'l2_atomic',
            'achieved': gbs if packed else kernel_rate,
            'peak': peak if packed else red_peak,
            'unit': 'GB/s' if packed else 'reductions/s',
            'frac': gbs / peak if packed else kernel_rate / red_peak,
            'traffic': traffic, 'traffic_source': traffic_src,
            'kernel_ms': iter_ms, 'samples_per_second_kernel': kernel_rate,
            'what': ('8 B packed-u64 accumulate per sample into a grid far larger than L2: '
                     'HBM sector read-modify-write' if packed else
                     'one red.global.add.v4.f32 per sample into an L2-resident float4 grid; peak '
                     '= uniformly scattered reductions of the same size into the same grid, '
                     'measured in this run (L2 request rate; the histogram never leaves L2, so '
                     'DRAM traffic is ~0 and HBM does not bound the kernel).  A flame\'s '
                     'addresses are more concentrated than a uniform scatter, so the kernel can '
                     'sit a few per cent above this figure'),
            'hbm': {'algorithmic_bytes_per_sample': algo, 'achieved': gbs, 'peak': peak,
                    'unit': 'GB/s', 'frac': gbs / peak, 'peak_source': peak_src},
            'l2_atomic': {'achieved': kernel_rate, 'peak': red_peak, 'unit': 'reductions/s',
                          'frac': kernel_rate / red_peak},
        }
        filt_ms = head['ms_per_step'] - iter_ms
        roofline_filters = {
            'bound': 'hbm for the pointwise and blur kernels; FMA / SFU issue for the 31-tap '
                     'bilateral passes (profiles/)',
            'stage': 'everything of the step but cb_iter: interpolation, filter chain, RGBA8 '
                     'conversion' + (', NCCL all-reduce' if b.world > 1 else ''),
            'algorithmic_bytes_per_bin': 804,
            'achieved': 804.0 * nbins / (filt_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
            'frac': 804.0 * nbins / (filt_ms * 1e-3) / 1e9 / peak, 'ms': filt_ms,
            # not measured live: the ncu capture that says what the stage's dominant kernels
            # are bound by (three quarters of the stage is the 8 x 31-tap bilateral passes)
            'bilateral_main_pass': {
                'bound': 'fma_pipe', 'frac': 0.61,
                'source': 'profiles/r02_bilateral_kernels_ncu.md (ncu --set full, 1080p): '
                          'sm__pipe_fma_cycles_active 60-61 %, XU pipe 45 %, issue 61 %, largest '
                          'stall math-pipe throttle; DRAM 9-12 % of peak, bytes moved below the '
                          'algorithmic 40 B/bin'}}
    if getattr(rmgr.band_filter, 'shared', None) is not None:
        rmgr.band_filter.shared.close()
    rmgr.fb.free()
    del rmgr, rdr

    # ---- the other BASELINE configurations -----------------------------------------------
    extra = {}
    if not args.no_extras:
        names = [n for n in ('still4k', 'still8k') if n != args.workload]
        if b.world > 1:
            names.remove('still8k')          # 4 x 512 MiB planes per GPU: single-GPU record only
        for name in names:
            r = run_still(b, name, steps=max(3, min(5, args.steps)), warmup=3)
            st = r.pop('_state')
            if getattr(st[0].band_filter, 'shared', None) is not None:
                st[0].band_filter.shared.close()
            st[0].fb.free()
            del st
            r.pop('clocks', None)
            extra[name] = r
        extra['anim1080'] = run_anim(b, 'anim1080', frames_per_gpu=24,
                                     warmup=3)

    cpu = cpu_filters = None
    if not args.no_cpu_baseline and b.world == 1:      # rank 0 at N = 1 only
        cfg = WORKLOADS[args.workload]
        n_cpu = cfg['width'] * cfg['height'] * (500 if args.workload == 'still1080' else 60)
        rate, cores = cpu_chaos_rate(cfg, n_cpu)
        cpu = {'value': rate, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
               'sample': '%d samples of the same genome at the same resolution, chaos game only '
                         '(oracle/chaos.c, OpenMP)' % n_cpu}
        cpu_filters = cpu_filter_rates()

    if b.dist is not None:
        b.dist.barrier()
        b.dist.destroy_process_group()
    if b.rank != 0:
        return

    cfg = WORKLOADS[args.workload]
    line = {
        'metric': 'ifs_iterations_per_second', 'value': head['value'], 'unit': 'iterations/s',
        'n_gpus': b.world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
        'scaling': cfg['scaling'], 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['label'],
                   'samples_per_step': head['samples_per_step'],
                   'frames_per_second': head['frames_per_second'],
                   'accumulate': head['accumulate'], 'hot_bins': head['hot_bins'],
                   'l2': 'L2 flushed between timed steps (512 MiB fill)',
                   'timing': 'CUDA events on the launching stream, per step, max over ranks',
                   'collectives': args.collectives if b.world > 1 else None,
                   'parallelism': head['parallelism']},
        'clocks': head['clocks'],
        'e2e': head['e2e'],
        'gpu_launches': int(round(head['launches_per_step'] * args.steps)),
        'gpu_launches_per_step': head['launches_per_step'],
        'roofline': roofline, 'roofline_filters': roofline_filters,
        'cpu_baseline': cpu, 'cpu_baseline_filters': cpu_filters,
        'wall_s_timed_region': head['wall_s'],
        'extra': extra,
    }
    for k in ('nccl_reduce_ms', 'nccl_gather_ms'):
        if k in head:
            line['config'][k] = head[k]
    print(json.dumps(line))


if __name__ == '__main__':
    main()
