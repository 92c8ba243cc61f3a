"""
Radix sorting of 32-bit keys on the device: the reference's ``Sorter``
(cuburn/code/sort.py:385-520) on the C ABI (``cb_sort_pass``, csrc/cb_sort.cu).

The reference built this as the primitive for sorted accumulation (helpers/sortbench.cu)
and never wired it into the renderer; its single pass is not stable, so its multi-pass sort
is flagged as broken in its own docstring.  The pass here is stable, ``multisort`` is a
correct least-significant-digit sort, and ``tools/deferred_bench.py`` measures what a
tile-binned deferred accumulation built on it would cost next to the direct L2 reductions
the renderer uses (profiles/r02_deferred_accumulation.md).
"""
import ctypes

import numpy as np

from .. import _native as N


class Sorter(object):
    group_size = 8192
    radix_bits = 8

    def __init__(self, max_size, offsets=None):
        """
        A sorter for up to ``max_size`` keys.  Unlike the reference, ``max_size`` need not be
        a multiple of the group size.  ``offsets``: a device buffer to use as scratch
        (at least ``scratch_bytes(max_size)``), to share it between sorters.
        """
        self.max_size = int(max_size)
        self.radix_size = 1 << self.radix_bits
        self._words = self._scratch_words(self.max_size, self.radix_bits)
        if offsets is None:
            offsets = N.DeviceBuffer(4 * self._words)
        elif offsets.nbytes < 4 * self._words:
            raise ValueError('scratch buffer too small: %d < %d' % (offsets.nbytes, 4 * self._words))
        self.doffsets = offsets
        self._last = None               # (n, bits) of the last pass

    @staticmethod
    def _scratch_words(n, bits):
        words = ctypes.c_uint64()
        N.check(N.lib().cb_sort_scratch_words(n, bits, ctypes.byref(words)))
        return int(words.value)

    @classmethod
    def scratch_bytes(cls, max_size):
        return 4 * cls._scratch_words(int(max_size), cls.radix_bits)

    def sort(self, dst, src, size, lo_bit=0, ignore_max=False, stream=None, bits=None):
        """
        Group the ``size`` keys of ``src`` by the ``radix_bits`` bits from ``lo_bit`` up
        (0 = least significant) into ``dst``; equal digits keep their order.  With
        ``ignore_max`` keys equal to 0xffffffff are dropped (``nvalid()`` tells how many
        are left).  ``dst`` and ``src`` are device buffers (or pointers) and must differ.
        """
        size = int(size)
        if size > self.max_size:
            raise ValueError('size %d exceeds max_size %d' % (size, self.max_size))
        bits = self.radix_bits if bits is None else int(bits)
        N.check(N.lib().cb_sort_pass(int(dst), int(src), size, int(lo_bit), bits,
                                     int(bool(ignore_max)), self.doffsets.ptr,
                                     stream.handle if stream is not None else None))
        self._last = (size, bits)

    def _layout(self):
        if self._last is None:
            raise RuntimeError('no pass has run')
        n, bits = self._last
        groups = max(1, -(-n // self.group_size))
        return n, bits, groups, self._scratch_words(n, bits)

    def nvalid(self):
        """Keys the last pass kept (synchronises)."""
        n, bits, groups, words = self._layout()
        out = N.from_device(N.DeviceSlice(self.doffsets, 4 * (words - 8), 4), (1,), np.uint32)
        return int(out[0])

    def digit_starts(self):
        """Index in ``dst`` of the first key of every digit of the last pass, plus the end
        (uint32 [2^bits + 1]; synchronises)."""
        n, bits, groups, words = self._layout()
        table = N.from_device(N.DeviceSlice(self.doffsets, 0, 4 * (groups << bits)),
                              (1 << bits, groups), np.uint32)
        return np.concatenate([table[:, 0], [self.nvalid()]]).astype(np.uint32)

    def multisort(self, scratch_a, scratch_b, src, size, lo_bit=0, rounds=4, ignore_max=False,
                  stream=None):
        """
        Sort by ``rounds`` digits of ``radix_bits`` bits starting at ``lo_bit``, least
        significant first (``rounds=4, lo_bit=0``: a full ascending sort).  The result ends
        up in the buffer that is returned (``scratch_a`` or ``scratch_b``); ``src`` is not
        modified.
        """
        cur, out, other = src, scratch_a, scratch_b
        n = int(size)
        for r in range(rounds):
            self.sort(out, cur, n, lo_bit + r * self.radix_bits,
                      ignore_max=ignore_max and r == 0, stream=stream)
            if ignore_max and r == 0:
                n = self.nvalid()
            cur, out, other = out, (other if cur is src else cur), out
        return cur
