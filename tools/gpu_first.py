"""First contact with the GPU: run the whole path once, compare with the oracle."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cuburn_b200 import _native as N, samples, profile, render, mwc
from oracle import flame_ref as R, filters_ref as F

out_dir = 'gpurun_out'
os.makedirs(out_dir, exist_ok=True)
N.init(0)
print(N.device_info(0))

def run(gname, w, h, spp, seed=7, save=None, oracle=True):
    gnm = samples.GENOMES[gname]()
    prof = dict(width=w, height=h, spp=spp, frame_width=0, start=1, end=2)
    gprof = profile.wrap(prof, gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=seed)
    t0 = time.time()
    rdr = render.Renderer(gnm, gprof)
    print(gname, 'compile+load %.2fs' % (time.time() - t0), 'grid', rdr.grid_ctas(rmgr.fb.nstreams), rdr.kernel_info)
    # stage by stage for inspection
    dim = rmgr.fb.set_dim(w, h)
    td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
    ts = tc - 0.5 * td
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, ts, td)
    rmgr.stream_a.synchronize()
    pk = rdr.packer
    params = N.from_device(rmgr.info_a.d_params, (1024, pk.param_stride), np.float32)
    if oracle:
        ev = R.GenomeEval(gnm, w, h, tc, td)
        bad = 0
        for i, name in pk.named_slots():
            a, b = params[:, i], ev.values[name]
            if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                bad += 1
                print('  MISMATCH', name, a[:2], b[:2])
        print(gname, 'packed params bit-exact:', bad == 0, '(%d slots)' % len(pk.slot_names))
        seeds0 = mwc.make_seeds(rmgr.fb.nstreams, host_seed=seed)
        opal, _ = R.palette_table(gnm, ts, td, seeds0)
        dpal = N.from_device(rmgr.info_a.d_palette, (64, 256, 4), np.float32)
        print(gname, 'palette bit-exact:', np.array_equal(opal.view(np.uint32), dpal.view(np.uint32)))
    e0, e1 = N.Event(), N.Event()
    e0.record(rmgr.stream_a)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    e1.record(rmgr.stream_a)
    e1.synchronize()
    ms = e1.time_since(e0)
    n = rmgr.last_iter_samples
    print(gname, '%dx%d spp=%d: %d samples in %.2f ms -> %.3g it/s' % (w, h, spp, n, ms, n / ms * 1e3))
    hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
    print('  hist sum count', hist[..., 3].sum(), 'frac', hist[..., 3].sum() / n, 'max', hist[..., 3].max(), 'nan', np.isnan(hist).sum())
    if oracle:
        t = time.time()
        ohist, _ = R.iterate(ev, opal, seeds0, n)
        print('  oracle iterate %.2fs -> %.3g it/s' % (time.time() - t, n / (time.time() - t)))
        print('  oracle frac', ohist[..., 3].sum() / n, 'max', ohist[..., 3].max())
        # pooled density comparison 8x8
        def pool(a):
            hh, ww = a.shape[0] // 8 * 8, a.shape[1] // 8 * 8
            return a[:hh, :ww].reshape(hh // 8, 8, ww // 8, 8).sum(axis=(1, 3))
        pa, pb = pool(hist[..., 3].astype(np.float64)), pool(ohist[..., 3].astype(np.float64))
        m = (pa + pb) > 200
        z = (pa - pb)[m] / np.sqrt((pa + pb)[m])
        print('  pooled z: mean %.3f std %.3f max %.2f n=%d' % (z.mean(), z.std(), np.abs(z).max(), m.sum()))
        for ch in range(3):
            ca, cb = pool(hist[..., ch].astype(np.float64)), pool(ohist[..., ch].astype(np.float64))
            print('  channel %d mean color diff %.5f' % (ch, np.abs(ca[m] / pa[m] - cb[m] / pb[m]).mean()))
    # full frame through the public API
    evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc)
    evt.synchronize()
    print(gname, 'queue_frame %.2f ms' % evt.time(), buf.shape, buf.dtype, buf.mean(axis=(0, 1)))
    if save:
        from PIL import Image
        Image.fromarray(np.ascontiguousarray(buf[:, :, :3])).save(os.path.join(out_dir, save))
    if oracle and w * h <= 640 * 360:
        pix = F.default_chain(ohist, w, h, gnm['camera']['scale'], spp)
        from oracle import output_ref as O
        o8, _ = O.convert('rgba_u8', pix, w, h, seeds0)
        o8 = o8.reshape(h, w, 4)
        mse = np.mean((o8[..., :3].astype(np.float64) - buf[..., :3].astype(np.float64)) ** 2)
        print(gname, 'PSNR vs oracle frame: %.2f dB' % (10 * np.log10(255 ** 2 / mse)))
        Image.fromarray(o8[:, :, :3]).save(os.path.join(out_dir, 'oracle_' + save))
    return rmgr

run('G3', 640, 360, 256, save='g3.png')
run('G6F', 640, 360, 256, save='g6f_small.png')
run('G6F', 1920, 1080, 2000, save='g6f_1080p.png', oracle=False)
run('G24H', 1920, 1080, 500, save='g24h_1080p.png', oracle=False)
