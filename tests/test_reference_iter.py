"""
The iterate kernel's OWN text from the reference, executed on the CPU:

  * `apply_xf_<id>` (code/iter.py:121-149), the xform choice chain (iter.py:263-272) and
    final xform + camera + trunca + bounds test (iter.py:302-317): the tempita templates
    rendered by oracle/build_ref.py for the sample genomes and compiled with g++;
  * `trunca` (code/util.py:194-200), the packed-cell add / overflow spill
    (iter.py:332-407) and flush_atom's unpack (iter.py:429-479): inline PTX, executed
    verbatim by oracle/ptx_emu.py.

CPU tests pin the oracle (oracle/chaos.c, oracle/flame_ref.py, oracle/packed_ref.py)
against them; GPU tests pin the device kernels against the same executions.
"""
import ctypes

import numpy as np
import pytest


def _ref():
    from oracle import build_ref
    if build_ref.lib() is None or 'iterate' not in build_ref.meta():
        pytest.skip('reference kernels library not available')
    return build_ref


def _points(n, seed):
    rs = np.random.RandomState(seed)
    return (rs.uniform(-1.6, 1.6, n).astype(np.float32), rs.uniform(-1.4, 1.4, n).astype(np.float32),
            rs.uniform(0, 1, n).astype(np.float32))


def _setup(tag, w=640, h=360):
    B = _ref()
    from cuburn_b200 import samples
    from oracle import flame_ref as R
    gnm = {'g3': samples.g3, 'g6f': samples.g6f, 'g24h': samples.g24h}[tag]()
    ev = R.GenomeEval(gnm, w, h, 0.5, 0.0)
    return B, R, gnm, ev, B.RefIterate(tag, ev)


def _close(a, b, tol=2e-5):
    ok = np.isfinite(a) & np.isfinite(b) & (np.abs(b) < 1e4)
    err = np.abs(a[ok] - b[ok]) / (1.0 + np.abs(b[ok]))
    return ok.mean() > 0.9 and (err < tol).mean() > 0.999, float(err.max())


@pytest.mark.parametrize('tag', ['g3', 'g6f', 'g24h'])
def test_oracle_apply_xform_equals_reference_apply_xf(built, tag):
    """Every xform of the sample genomes (pre-affine precalc, the variations in sorted
    order, post-affine, colour blend): oracle/chaos.c against the reference's rendered
    apply_xf_<id>; identical RNG consumption."""
    from cuburn_b200 import mwc
    B, R, gnm, ev, ref = _setup(tag)
    xs, ys, cs = _points(3000, 1)
    seeds = mwc.make_seeds(xs.size, host_seed=9)
    ids = ev.xform_ids
    todo = [(k, ('xforms', xid), gnm['xforms'][xid]) for k, xid in enumerate(ids)]
    if ev.has_final:
        todo.append((-1, ('final_xform',), gnm['final_xform']))
    for k, xpath, xf in todo:
        rec = ev.xform_record(xpath, xf, R.chaos_lib())[0]
        ox, oy, oc, oseeds = R.apply_xform(rec, xs, ys, cs, seeds)
        rx, ry, rc, rseeds = ref.apply(k, xs, ys, cs, seeds)
        assert np.array_equal(oseeds, rseeds), (tag, xpath, 'RNG draws differ')
        for a, b in ((ox, rx), (oy, ry)):
            ok, worst = _close(a, b)
            assert ok, (tag, xpath, worst)
        assert np.abs(oc - rc).max() < 1e-6


@pytest.mark.parametrize('tag', ['g3', 'g6f'])
def test_oracle_choice_chain_equals_reference(built, tag):
    """`if (xfsel <= den_0) ... else if ... else last`: the chosen xform for a sweep of
    selectors including the exact boundaries, and the point it produces."""
    from cuburn_b200 import mwc
    B, R, gnm, ev, ref = _setup(tag)
    ids = ev.xform_ids
    den = [ev.values['xforms.%s.density' % i][0] for i in ids[:-1]]
    xs, ys, cs = _points(64, 2)
    seeds = mwc.make_seeds(64, host_seed=4)
    sels = [0.0, 1.0, 0.5]
    for d in den:
        sels += [float(d), float(np.nextafter(d, np.float32(2))), float(np.nextafter(d, np.float32(-1)))]
    for sel in sels:
        pick = len(ids) - 1
        for i, d in enumerate(den):
            if np.float32(sel) <= d:
                pick = i
                break
        last, rx, ry, rc, rseeds = ref.choose(sel, xs, ys, cs, seeds)
        assert (last == pick).all(), (sel, pick, last[:4])
        ax, ay, ac, aseeds = ref.apply(pick, xs, ys, cs, seeds)
        assert np.array_equal(rx, ax, equal_nan=True) and np.array_equal(rseeds, aseeds)
    # the cumulative densities themselves (precalc_densities) are pinned in
    # test_reference_code.py::test_precalc_densities


@pytest.mark.parametrize('tag', ['g3', 'g6f'])
def test_oracle_point_to_bin_equals_reference(built, tag):
    """Final xform on a copy, camera affine, `trunca`, unsigned bounds test against
    (astride, aheight), bin index (iter.py:302-317,331)."""
    from cuburn_b200 import mwc
    B, R, gnm, ev, ref = _setup(tag)
    n = 20000
    xs, ys, cs = _points(n, 3)
    xs[:8] = [np.nan, np.inf, -np.inf, 1e30, -1e30, 0, 0, 0]
    seeds = mwc.make_seeds(n, host_seed=6)
    ridx, rcc, rseeds = ref.bin(xs, ys, cs, seeds)
    fx, fy, fc, oseeds = xs, ys, cs, seeds
    if ev.has_final:
        rec = ev.xform_record(('final_xform',), gnm['final_xform'], R.chaos_lib())[0]
        fx, fy, fc, oseeds = R.apply_xform(rec, xs, ys, cs, seeds)
    cam = np.array([ev.values['camera.' + c][0] for c in ('xx', 'xy', 'xo', 'yx', 'yy', 'yo')], np.float32)
    obins, _ = R.point_to_bin(cam, fx, fy, fc, np.zeros(n, np.float32), ev.dim['astride'], ev.dim['ah'])
    assert np.array_equal(oseeds, rseeds)
    # the reference computes the camera affine as mul/add (no fma): a point within one
    # float32 ulp of a .5 boundary may round to the neighbouring bin
    assert (obins != ridx).mean() < 2e-3, (obins != ridx).mean()
    assert ((obins >= 0) == (ridx >= 0)).mean() > 0.999
    both = (obins >= 0) & (ridx >= 0)
    d = np.abs(obins[both] - ridx[both])
    assert np.isin(d, [0, 1, ev.dim['astride'], ev.dim['astride'] - 1, ev.dim['astride'] + 1]).all()
    assert (ridx >= 0).sum() > 1000 and (ridx < 0).sum() > 100
    acc = ridx >= 0
    assert np.abs(rcc[acc] - fc[acc]).max() < 1e-6


def test_trunca_ptx_equals_oracle_rounding(built):
    """`cvt.rni.s32.f32` as written in code/util.py:198, executed: round half to even,
    saturating, NaN -> 0 -- what the oracle's binning and the device's
    __float2int_rn compute."""
    B = _ref()
    from oracle import ptx_emu as E, flame_ref as R
    v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 3.49999, 1e10, -1e10, np.nan, np.inf, -np.inf,
                  2147483520.0, -2147483648.0, 7.0, 1951.5, 1952.5], np.float32)
    v = np.concatenate([v, np.random.RandomState(1).uniform(-3000, 3000, 500).astype(np.float32)])
    text = '.reg .u32 r; ' + B.ref_ptx('trunca').replace('%0', 'r')
    regs = E.run(text, v.size, [None, v], E.Memory())
    got = regs['r'].view(np.int32)
    want = np.where(np.isnan(v), 0, np.clip(np.rint(v.astype(np.float64)), -2 ** 31, 2 ** 31 - 1)).astype(np.int64)
    assert np.array_equal(got.astype(np.int64), want)
    # through the oracle's point_to_bin with an identity camera: ix = rni(x), iy = 0
    cam = np.array([1, 0, 0, 0, 0, 0], np.float32)
    ok = (v >= -0.5) & (v < 4000)
    bins, _ = R.point_to_bin(cam, v, np.zeros_like(v), np.zeros_like(v), np.zeros_like(v), 4096, 16)
    assert np.array_equal(bins[ok], got[ok])


def _surface(rs):
    """A packed palette surface as interp_palette_flat writes it (interp.py:428-429)."""
    Y, U, V = (rs.randint(0, 256, (64, 256)).astype(np.uint32) for _ in range(3))
    pal = np.zeros((64, 256, 2), np.uint32)
    pal[..., 1] = (1 << 22) + (Y << 4)
    pal[..., 0] = (U << 18) + V

    def suld(xbytes, y):
        x, y = np.minimum(xbytes // 8, 255), np.minimum(y, 63)
        return pal[y, x, 0], pal[y, x, 1]
    return Y, U, V, suld


def test_packed_cell_add_and_spill_ptx_equals_oracle(built):
    """The reference's accumulation PTX (iter.py:332-407), executed: palette column,
    packed entry, 64-bit add, the 3 % check with its 512 threshold, the drain into the
    float4 histogram -- against oracle/packed_ref.py, bit for bit."""
    B = _ref()
    from oracle import ptx_emu as E, packed_ref as P
    rs = np.random.RandomState(11)
    Y, U, V, suld = _surface(rs)
    nb, n = 24, 512
    cells, hist = np.zeros(nb, np.uint64), np.zeros((nb, 4), np.float32)
    # some cells start close to / beyond the spill threshold
    for b, cnt in ((3, 511), (4, 512), (5, 700), (6, 900)):
        cells[b] = P.pack_entry(0, 0, 0) * np.uint64(0) + ((np.uint64(cnt) << np.uint64(54)) |
                                                         (np.uint64(cnt * 90) << np.uint64(36)) |
                                                         (np.uint64(cnt * 40) << np.uint64(18)) |
                                                         np.uint64(cnt * 200))
    ocells, ohist = cells.copy(), hist.copy()
    mem = E.Memory()
    A, O = mem.add(0x10000000, cells), mem.add(0x20000000, hist)
    cc = rs.rand(n).astype(np.float32)
    cc[:4] = [0.0, 1.0, 0.5 / 255, 254.5 / 255]
    dith = (0.49 * rs.uniform(-1, 1, n)).astype(np.float32)
    dith[:4] = [-0.49, 0.49, 0.0, 0.0]
    time = rs.randint(0, 64, n).astype(np.uint32)
    bins = rs.randint(0, nb, n).astype(np.uint32)
    cosel = np.where(rs.rand(n) < 0.3, 0.99, 0.5).astype(np.float32)
    E.run(B.ref_ptx('iter_accumulate'), n,
          [cc, dith, time, bins, np.uint64(A), cosel, np.uint64(O), np.float32(1.0)], mem, suld=suld)
    col = np.minimum(P.palette_column(cc, dith), 255)
    entries = P.pack_entry(Y[time, col], U[time, col], V[time, col])
    P.accumulate(ocells, ohist, bins, entries, cosel > np.float32(0.97))
    assert np.array_equal(cells, ocells)
    assert np.array_equal(hist.view(np.uint32), ohist.view(np.uint32))
    assert hist[:, 3].sum() > 1500 and (cells >> np.uint64(54)).sum() > 100      # both paths ran
    total = (cells >> np.uint64(54)).sum() + hist[:, 3].sum()
    assert total == n + 511 + 512 + 700 + 900


def test_flush_atom_ptx_equals_oracle(built):
    """flush_atom (iter.py:429-540), executed for a 32 x 8 block: the unpack + fma into
    the float4 histogram bit for bit against oracle/packed_ref.flush, the cells zeroed,
    and the hotspot ballots consistent with the 128 / 512 / 2048 thresholds."""
    B = _ref()
    from oracle import ptx_emu as E, packed_ref as P
    rs = np.random.RandomState(12)
    astride, rows = 32, 16
    nb = astride * rows
    cnt = rs.choice([0, 1, 7, 100, 129, 500, 513, 1023], nb).astype(np.uint64)
    cells = (cnt << np.uint64(54)) | ((cnt * np.uint64(200)) << np.uint64(36)) | \
            ((cnt * np.uint64(128)) << np.uint64(18)) | (cnt * np.uint64(33))
    hist = (rs.rand(nb, 4) * rs.choice([0, 10, 3000], (nb, 1))).astype(np.float32)
    hot = np.zeros(nb, np.uint32)
    ocells, ohist = cells.copy(), hist.copy()
    mem = E.Memory()
    A, O, H = mem.add(0x10000000, cells), mem.add(0x20000000, hist), mem.add(0x30000000, hot)
    for y0 in range(0, rows, 8):                   # blocks of (32, 8) threads, warp = one row
        ty, tx = np.meshgrid(np.arange(8), np.arange(32), indexing='ij')
        tx, ty = tx.ravel().astype(np.uint32), ty.ravel().astype(np.uint32)
        xi, yi = tx, ty + np.uint32(y0)
        gi = yi * np.uint32(astride) + xi
        hoti = (yi >> 4) * np.uint32(astride) + (xi & np.uint32(0xfffffff0)) + (yi & np.uint32(15))
        E.run(B.ref_ptx('flush_atom'), 256,
              [gi, hoti, np.uint64(A), np.uint64(O), np.uint64(H), xi, yi], mem,
              special={'%tid.x': tx, '%tid.y': ty, '%laneid': tx})
    P.flush(ocells, ohist)
    assert not cells.any() and not ocells.any()
    assert np.array_equal(hist.view(np.uint32), ohist.view(np.uint32))
    # flag words: thread x = 0 of each row writes 16 two-bit flags; per SURVEY Q21 the two
    # bits of a flag are swapped (even lane = low bit carries the "> 512" test)
    lvl = P.hot_levels(hist[:, 3]).reshape(rows, astride)
    for y in range(rows):
        word = int(hot[(y >> 4) * astride + (y & 15)])
        half = lvl[y, 16:] if (y & 1) else lvl[y, :16]      # odd rows store the high half
        for k in range(16):
            lo, hi = (word >> (2 * k)) & 1, (word >> (2 * k + 1)) & 1
            assert lo == int(half[k] >= 2) and hi == int(half[k] in (1, 3)), (y, k, half[k], lo, hi)


# ---- the device against the same executions ----------------------------------------
@pytest.mark.gpu
def test_device_flush_and_palette_pack_equal_reference_ptx(native, built):
    """cb_palette_pack against the reference's surface words (interp.py:428-429) and
    cb_flush_packed against flush_atom's PTX, bit for bit."""
    B = _ref()
    N = native
    from oracle import ptx_emu as E, packed_ref as P
    rs = np.random.RandomState(13)
    lev = rs.randint(0, 256, (64, 256, 3))
    pal4 = np.ones((64, 256, 4), np.float32)
    pal4[..., :3] = (lev / 255.0).astype(np.float32)
    d_p4, d_pp = N.to_device(pal4), N.DeviceBuffer(64 * 256 * 8)
    N.check(N.lib().cb_palette_pack(d_pp.ptr, d_p4.ptr, 64, None))
    got = N.from_device(d_pp, (64, 256), np.uint64)
    assert np.array_equal(got, P.pack_entry(lev[..., 0], lev[..., 1], lev[..., 2]))

    dim = N.calc_dim(40, 8)                       # 64 x 32 bins
    nb = dim.ah * dim.astride
    cnt = rs.choice([0, 1, 7, 100, 129, 500, 513, 1023], nb).astype(np.uint64)
    cells = (cnt << np.uint64(54)) | ((cnt * np.uint64(200)) << np.uint64(36)) | \
            ((cnt * np.uint64(128)) << np.uint64(18)) | (cnt * np.uint64(33))
    hist = (rs.rand(nb, 4) * rs.choice([0, 10, 3000], (nb, 1))).astype(np.float32)
    d_c, d_h = N.to_device(cells), N.to_device(hist)
    N.check(N.lib().cb_flush_packed(d_h.ptr, d_c.ptr, N.byref(dim), None))
    N.check(N.lib().cb_device_sync())
    dev = N.from_device(d_h, (nb, 4), np.float32)
    hot = np.zeros(nb, np.uint32)
    mem = E.Memory()
    A, O, H = mem.add(0x10000000, cells), mem.add(0x20000000, hist), mem.add(0x30000000, hot)
    m = 512
    gi = np.arange(m, dtype=np.uint32)
    E.run(B.ref_ptx('flush_atom'), m, [gi, gi * 0, np.uint64(A), np.uint64(O), np.uint64(H), gi, gi * 0],
          mem, special={'%tid.x': gi & 31, '%tid.y': gi * 0, '%laneid': gi & 31})
    assert np.array_equal(dev[:m].view(np.uint32), hist[:m].view(np.uint32))


@pytest.mark.gpu
def test_device_packed_accumulate_equals_reference_ptx(native, built):
    """The device's accumulate_packed (red.global.add.u64, the unconditional drain of one
    lane per warp-round) on an explicit sample list, followed by cb_flush_packed, against
    the reference's accumulation PTX + flush_atom on the same list: cell contents are
    identical where nothing was drained; after the flush the density is identical
    everywhere and the colour sums agree to float32 rounding of the partial sums."""
    B = _ref()
    N = native
    from cuburn_b200 import samples
    from cuburn_b200.code import itergen
    from oracle import ptx_emu as E, packed_ref as P
    rs = np.random.RandomState(14)
    Y, U, V, suld = _surface(rs)
    pk, src = itergen.mkiterlib(samples.g3(), params_const=False, acc_packed=True)
    names, hdrs = itergen.load_headers()
    mod = N.Module(src, 'probe_packed.cu', hdrs, names, itergen.NVRTC_OPTIONS)
    dim = N.calc_dim(40, 8)
    nb = dim.ah * dim.astride
    n = 4096
    bins = rs.randint(0, 64, n).astype(np.int32)            # few bins: cells fill up
    row = 5
    col = rs.randint(0, 256, n).astype(np.int32)
    palp = P.pack_entry(Y, U, V)                            # [64][256]
    c = ctypes
    for drain_share in (0.0, 1.0 / 32):
        drain = (rs.rand(n) < drain_share).astype(np.int32)
        d_cells, d_hist = N.DeviceBuffer(8 * nb), N.DeviceBuffer(16 * nb)
        N.fill32(d_cells, 2 * nb, 0)
        N.fill32(d_hist, 4 * nb, 0)
        d_b, d_c, d_d = N.to_device(bins), N.to_device(col), N.to_device(drain)
        d_pal = N.to_device(np.ascontiguousarray(palp[row]))
        mod.launch('cb_probe_accumulate', ((n + 255) // 256,), (256,),
                   [c.c_uint64(d_cells.ptr), c.c_uint64(d_hist.ptr), c.c_uint64(d_b.ptr),
                    c.c_uint64(d_c.ptr), c.c_uint64(d_d.ptr), c.c_uint64(d_pal.ptr), c.c_int(n)])
        N.check(N.lib().cb_device_sync())
        dcells = N.from_device(d_cells, (nb,), np.uint64)
        # reference: the same samples through its PTX (no thread checks: cosel = 0.5)
        cells, hist = np.zeros(nb, np.uint64), np.zeros((nb, 4), np.float32)
        hot = np.zeros(nb, np.uint32)
        mem = E.Memory()
        A, O, H = mem.add(0x10000000, cells), mem.add(0x20000000, hist), mem.add(0x30000000, hot)
        cc = (col / 255.0).astype(np.float32)
        if drain_share == 0.0:
            cnt = np.bincount(bins, minlength=nb)
            assert cnt.max() < 1024
            E.run(B.ref_ptx('iter_accumulate'), n,
                  [cc, np.float32(0.0), np.uint32(row), bins.astype(np.uint32), np.uint64(A),
                   np.float32(0.5), np.uint64(O), np.float32(1.0)], mem, suld=suld)
            assert np.array_equal(dcells, cells)             # same 64-bit sums
        else:
            # many samples per bin: let the reference spill where it checks
            E.run(B.ref_ptx('iter_accumulate'), n,
                  [cc, np.float32(0.0), np.uint32(row), bins.astype(np.uint32), np.uint64(A),
                   np.where(rs.rand(n) < 0.5, 0.99, 0.5).astype(np.float32), np.uint64(O),
                   np.float32(1.0)], mem, suld=suld)
        N.check(N.lib().cb_flush_packed(d_hist.ptr, d_cells.ptr, N.byref(dim), None))
        N.check(N.lib().cb_device_sync())
        dev = N.from_device(d_hist, (nb, 4), np.float32)
        m = 64
        gi = np.arange(m, dtype=np.uint32)
        E.run(B.ref_ptx('flush_atom'), m, [gi, gi * 0, np.uint64(A), np.uint64(O), np.uint64(H), gi, gi * 0],
              mem, special={'%tid.x': gi & 31, '%tid.y': gi * 0, '%laneid': gi & 31})
        assert np.array_equal(dev[:, 3], hist[:, 3])          # density: exact
        assert dev[:, 3].sum() == n
        if drain_share == 0.0:
            assert np.array_equal(dev.view(np.uint32), hist.view(np.uint32))
        else:
            assert np.allclose(dev[:, :3], hist[:, :3], rtol=3e-6, atol=1e-5)
