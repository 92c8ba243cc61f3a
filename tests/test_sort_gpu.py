"""The radix-sort primitive (cuburn_b200/code/sort.py, csrc/cb_sort.cu) against numpy.
Reference behaviour: cuburn/code/sort.py:443-523 (Sorter.sort / multisort) and its own
self-test (sort.py:525-586: sorted output compared with numpy's sort of the same keys)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _digits(keys, lo_bit, bits=8):
    return (keys >> np.uint32(lo_bit)) & np.uint32((1 << bits) - 1)


@pytest.mark.parametrize('n', [0, 1, 31, 8192, 8193, 100000, (1 << 20) + 12345])
@pytest.mark.parametrize('lo_bit', [0, 8, 13, 24])
def test_single_pass_is_a_stable_partition(native, built, n, lo_bit):
    N = native
    from cuburn_b200.code.sort import Sorter
    rs = np.random.RandomState(n % 1000 + lo_bit)
    keys = rs.randint(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    d_src, d_dst = N.to_device(keys if n else np.zeros(1, np.uint32)), N.DeviceBuffer(4 * max(n, 1))
    srt = Sorter(max(n, 1))
    srt.sort(d_dst, d_src, n, lo_bit)
    got = N.from_device(d_dst, (max(n, 1),), np.uint32)[:n]
    order = np.argsort(_digits(keys, lo_bit), kind='stable')
    assert np.array_equal(got, keys[order])
    assert srt.nvalid() == n
    starts = srt.digit_starts()
    want = np.concatenate([[0], np.cumsum(np.bincount(_digits(keys, lo_bit), minlength=256))])
    assert np.array_equal(starts, want.astype(np.uint32))


def test_ignore_max_drops_sentinels(native, built):
    N = native
    from cuburn_b200.code.sort import Sorter
    rs = np.random.RandomState(3)
    n = 50000
    keys = rs.randint(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    keys[rs.rand(n) < 0.3] = 0xffffffff
    d_src, d_dst = N.to_device(keys), N.DeviceBuffer(4 * n)
    N.fill32(d_dst, n, 0)
    srt = Sorter(n)
    srt.sort(d_dst, d_src, n, 4, ignore_max=True)
    kept = keys[keys != 0xffffffff]
    assert srt.nvalid() == len(kept)
    got = N.from_device(d_dst, (n,), np.uint32)[:len(kept)]
    assert np.array_equal(got, kept[np.argsort(_digits(kept, 4), kind='stable')])


@pytest.mark.parametrize('dist', ['uniform', 'flame', 'constant'])
def test_multisort_is_a_full_sort(native, built, dist):
    """Four stable 8-bit passes, least significant digit first = an ascending sort (what
    the reference's multi-pass sort could not guarantee, sort.py:462-466)."""
    N = native
    from cuburn_b200.code.sort import Sorter
    rs = np.random.RandomState(5)
    n = 300000
    if dist == 'uniform':
        keys = rs.randint(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    elif dist == 'flame':           # (bin << 8 | colour) records of a concentrated histogram
        bins = (2155008 * rs.rand(n) ** 4).astype(np.uint32)
        keys = (bins << np.uint32(8)) | rs.randint(0, 256, n).astype(np.uint32)
    else:
        keys = np.full(n, 0xdeadbeef, np.uint32)
    d_src, d_a, d_b = N.to_device(keys), N.DeviceBuffer(4 * n), N.DeviceBuffer(4 * n)
    srt = Sorter(n)
    d_out = srt.multisort(d_a, d_b, d_src, n)
    assert d_out is d_a or d_out is d_b
    assert np.array_equal(N.from_device(d_out, (n,), np.uint32), np.sort(keys))
    assert np.array_equal(N.from_device(d_src, (n,), np.uint32), keys)        # untouched


def test_sorter_argument_errors(native, built):
    N = native
    from cuburn_b200.code.sort import Sorter
    srt = Sorter(1000)
    buf = N.DeviceBuffer(4000)
    with pytest.raises(ValueError):
        srt.sort(buf, buf, 2000)
    with pytest.raises(ValueError):
        srt.sort(buf, buf, 100)                      # dst == src
    with pytest.raises(ValueError):
        Sorter(1 << 20, offsets=N.DeviceBuffer(64))
