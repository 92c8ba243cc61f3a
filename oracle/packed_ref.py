"""
ORACLE (test infrastructure only) -- numpy restatement of the reference's packed
accumulator.

  cell format   cuburn/code/interp.py:428-429, iter.py:334-407: one u64 per bin,
                count:10 | sum Y:18 | sum U:18 | sum V:18 of 8-bit palette levels; the
                palette entry is pre-packed as hi = (1 << 22) + (Y << 4), lo = (U << 18) + V
  add / spill   iter.py:359-407: 64-bit add; a checking thread (3 % of the warps) that sees
                a count >= 512 swaps the cell for zero and adds its contents, scaled, to the
                float4 histogram with four scalar reductions
  flush         iter.py:454-479 (`flush_atom`): hist += unpack(cell) * mult, cell = 0

Pinned against the reference's own PTX, executed by oracle/ptx_emu.py
(tests/test_reference_code.py).
"""
import numpy as np

f32 = np.float32
K255 = f32(1.0) / f32(255.0)        # the PTX immediate (1.0/255.0), rounded to float32


def pack_entry(y, u, v):
    """8-bit levels -> the u64 addend ((hi << 32) | lo)."""
    y, u, v = (np.asarray(a, np.uint64) for a in (y, u, v))
    hi = (np.uint64(1) << np.uint64(22)) + (y << np.uint64(4))
    lo = (u << np.uint64(18)) + v
    return (hi << np.uint64(32)) | lo


def unpack_cell(cell):
    """u64 cells -> (count, sum Y, sum U, sum V) as uint32 arrays."""
    cell = np.asarray(cell, np.uint64)
    count = (cell >> np.uint64(54)).astype(np.uint32)
    y = ((cell >> np.uint64(36)) & np.uint64(0x3ffff)).astype(np.uint32)
    u = ((cell >> np.uint64(18)) & np.uint64(0x3ffff)).astype(np.uint32)
    v = (cell & np.uint64(0x3ffff)).astype(np.uint32)
    return count, y, u, v


def palette_column(color, dither):
    """fma.rn(color, 255, dither) -> cvt.rni.u32 (iter.py:346-347; no clamp: the surface
    load clamps the coordinate instead)."""
    x = (np.asarray(color, f32).astype(np.float64) * 255.0
         + np.asarray(dither, f32).astype(np.float64)).astype(f32)
    r = np.rint(x.astype(np.float64))
    return np.where(np.isnan(r), 0, np.clip(r, 0, 4294967295.0)).astype(np.uint32)


def accumulate(cells, hist, bins, entries, check, mult=None):
    """
    One instruction stream over ``len(bins)`` lanes, as the hardware runs the block
    (iter.py:359-406): first the non-checking lanes add their entries (`red`), then the
    checking lanes (`atom`, keeping the value the cell held BEFORE their add), lanes in
    order within each instruction; then every checking lane that saw a
    count >= 512 swaps the cell for zero and, if it got a non-empty cell, adds its
    contents -- scaled by its hotspot multiplier -- to the float4 histogram with four
    scalar float reductions.  cells: uint64 [nbins]; hist: float32 [nbins][4]; in place.
    """
    mult = np.ones(len(bins), f32) if mult is None else np.asarray(mult, f32)
    seen = np.zeros(len(bins), np.uint64)
    order = [k for k in range(len(bins)) if not check[k]] + \
            [k for k in range(len(bins)) if check[k]]
    for k in order:
        b = int(bins[k])
        seen[k] = cells[b]
        with np.errstate(over='ignore'):
            cells[b] = cells[b] + entries[k]
    for k in range(len(bins)):
        if not check[k] or int(seen[k] >> np.uint64(32)) < (256 << 23):
            continue
        b = int(bins[k])
        old = cells[b]
        cells[b] = 0
        if int(old >> np.uint64(32)) == 0:
            continue
        d, y, u, v = unpack_cell(old)
        m = f32(mult[k])
        scale = f32(m * K255)
        hist[b, 0] += f32(f32(y) * scale)
        hist[b, 1] += f32(f32(u) * scale)
        hist[b, 2] += f32(f32(v) * scale)
        hist[b, 3] += f32(f32(d) * m)


def _fma(a, b, c):
    return (np.asarray(a, f32).astype(np.float64) * np.asarray(b, f32).astype(np.float64)
            + np.asarray(c, f32).astype(np.float64)).astype(f32)


def flush(cells, hist, mult=None):
    """flush_atom's accumulate half (iter.py:454-479): hist = fma(unpack(cell), scale, hist),
    cells = 0.  ``mult``: per-bin hotspot multiplier (1 when thinning is off)."""
    mult = np.ones(len(cells), f32) if mult is None else np.asarray(mult, f32)
    d, y, u, v = unpack_cell(cells)
    hist[:, 3] = _fma(d.astype(f32), mult, hist[:, 3])
    scale = (mult * K255).astype(f32)
    hist[:, 0] = _fma(y.astype(f32), scale, hist[:, 0])
    hist[:, 1] = _fma(u.astype(f32), scale, hist[:, 1])
    hist[:, 2] = _fma(v.astype(f32), scale, hist[:, 2])
    cells[:] = 0


def hot_levels(density):
    """Thresholds of the hotspot flags (iter.py:483-488): 0 / 1 / 2 / 3 for a flushed
    density above 128 / 512 / 2048."""
    d = np.asarray(density, f32)
    return (d > 128).astype(np.int32) + (d > 512) + (d > 2048)
