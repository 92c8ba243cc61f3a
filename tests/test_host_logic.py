"""Host-side decisions of the render manager that need no device: how often the accumulation
grid is swept, when the hot-bin pilot runs, how much scratch a sort pass needs."""
import ctypes
import types

import numpy as np


def _manager(l2=126 * 1024 * 1024, **attrs):
    from cuburn_b200 import render
    m = types.SimpleNamespace(spill=True, spill_interval=render.RenderManager.spill_interval,
                              spill_max_window=render.RenderManager.spill_max_window,
                              hot_bins='auto', hot_min_waves=render.RenderManager.hot_min_waves,
                              hot_recheck=render.RenderManager.hot_recheck, _hot_probe=None)
    m._l2 = lambda: l2
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def test_spill_window_sizes():
    """One sweep of the grid per 2^26 samples while the grid is L2-resident (capped by the
    kernel's window buffer), 0.002 x samples-per-bin sweeps beyond 0.6 x L2, none when
    switched off."""
    from cuburn_b200 import render
    f = render.RenderManager._spill_window
    unit = render.UNIT_SAMPLES
    # 1080p, 2000 spp: 62 sweeps wanted -> 1056 bins per unit, capped at the kernel's 1024
    nb, n = 2155008, 1920 * 1080 * 2000
    assert f(_manager(), nb, n) == 1024
    # 720p, 500 spp: 7 sweeps over 14063 units, below the cap
    nb7, n7 = 752 * 1312, 1280 * 720 * 500
    assert f(_manager(), nb7, n7) == -(-nb7 * 7 // (-(-n7 // unit))) < 1024
    # 4K, 4000 spp: beyond 0.6 x L2 -> int(0.002 * 3909) = 7 sweeps
    nb4, n4 = 8487424, 3840 * 2160 * 4000
    assert f(_manager(), nb4, n4) == -(-nb4 * 7 // (-(-n4 // unit)))
    # at least one sweep, never more bins than the grid
    assert f(_manager(), 24576, unit) == 1024
    assert f(_manager(), 512, 10 * unit) <= 512
    assert f(_manager(spill=False), nb, n) == 0 and f(_manager(), nb, 0) == 0


def test_hot_pilot_cadence():
    """'auto': probe a genome on its first frame; afterwards a genome without hot bins is
    looked at every hot_recheck-th frame only, a hot one on every frame."""
    from cuburn_b200 import render
    f = render.RenderManager._hot_decision
    m = _manager()
    rdr = types.SimpleNamespace(hot=None)
    big = 10 ** 6
    assert f(m, rdr, big, False, 1024) == (True, None)
    rdr.hot = False
    runs = [f(m, rdr, big, False, 1024)[0] for _ in range(32)]
    assert sum(runs) == 32 // m.hot_recheck and all(f(m, rdr, big, False, 1024)[1] is False for _ in range(3))
    rdr.hot = True
    assert all(f(m, rdr, big, False, 1024) == (True, True) for _ in range(10))
    # never for packed grids, short frames, or when switched off
    assert f(m, rdr, big, True, 1024) == (False, False)
    assert f(m, rdr, 3 * 1024, False, 1024) == (False, False)
    assert f(_manager(hot_bins=False), rdr, big, False, 1024) == (False, False)
    assert f(_manager(hot_bins=True), types.SimpleNamespace(hot=None), big, False, 1024) == (True, True)


def test_sort_scratch_words():
    """cb_sort_scratch_words: 2^bits counters per group of 8192 keys + one word per 1024
    counters + 8; argument errors are reported, not ignored."""
    from cuburn_b200 import _native as N
    L = N.lib()
    words = ctypes.c_uint64()
    for n, bits in ((0, 8), (1, 8), (8192, 8), (8193, 8), (1 << 26, 8), (100000, 4)):
        N.check(L.cb_sort_scratch_words(n, bits, ctypes.byref(words)))
        groups = max(1, -(-n // 8192))
        counters = groups << bits
        assert words.value == counters + -(-counters // 1024) + 8, (n, bits)
    assert L.cb_sort_scratch_words(100, 9, ctypes.byref(words)) != 0
    assert L.cb_sort_scratch_words(100, 0, ctypes.byref(words)) != 0
    # a pass on null buffers is refused before anything is launched
    assert L.cb_sort_pass(0, 0, 10, 0, 8, 0, 0, None) != 0
