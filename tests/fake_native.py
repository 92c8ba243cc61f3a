"""
A recording stand-in for libcuburn_b200.so, for CPU tests of the host-side orchestration.

``install(monkeypatch)`` puts a ``RecordingLib`` where ``cuburn_b200._native.lib()`` looks.
Everything that needs no device goes to the real library (cb_calc_dim, cb_band_rows,
cb_convert_size, cb_sort_scratch_words, cb_last_error, NVRTC builds); every entry point that
would touch a device is recorded as ``(name, args)`` and answers success without computing
anything: allocations hand out addresses from a counter, pinned host memory is ordinary
memory, events are always complete and measure 1 ms per kernel launch issued between them.
What is tested with it is WHICH calls the render manager issues, in which order and with
which arguments -- never a pixel.
"""
import ctypes
import gc

PASSTHROUGH = ('cb_last_error', 'cb_version', 'cb_calc_dim', 'cb_band_rows', 'cb_convert_size',
               'cb_sort_scratch_words', 'cb_module_build', 'cb_module_destroy',
               'cb_module_get_cubin')

# entry points that launch kernels of ours (what cb_launch_count counts on the device);
# value: launches per call (callable on the argument tuple where it varies)
KERNEL_CALLS = {
    'cb_fill32': 1, 'cb_mwc_test': 1, 'cb_interp_rows': 1, 'cb_interp_params': 1,
    'cb_interp_palette': 1, 'cb_palette_pack': 1, 'cb_iterate': 1, 'cb_module_launch': 1,
    'cb_flush_packed': 1, 'cb_hist_unswizzle': 1, 'cb_hist_finish': 1, 'cb_hot_scan': 2,
    'cb_sort_pass': 3, 'cb_yuv_to_rgb': 1, 'cb_den_blur': 1, 'cb_den_blur_1c': 1,
    'cb_full_blur': 1, 'cb_bilateral': 2, 'cb_bilateral_direction': 3, 'cb_logscale': 1,
    'cb_apply_gamma': 1, 'cb_haloclip': 1, 'cb_apply_gamma_full_hi': 1, 'cb_smearclip': 1,
    'cb_plainclip': 1, 'cb_colorclip': 1, 'cb_logencode': 1, 'cb_convert': 1,
    'cb_convert_rows': 1,
}


def _out(arg):
    """The ctypes object behind a byref() argument."""
    return arg._obj


class RecordingLib(object):
    SM_COUNT = 148
    L2_BYTES = 126 * 1024 * 1024

    def __init__(self, real):
        self._real = real
        self.calls = []                 # (name, args)
        self.iterations = []            # field dicts of every cb_iterate
        self._next_dev = 0x7f0000000000
        self._next_handle = 1
        self._host = {}                 # address -> buffer kept alive
        self.device_sizes = {}          # address -> bytes
        self.d2h_hook = None            # f(dst_addr, src_addr, nbytes): fill host memory
        self.ctas_per_sm = 8
        self._launched = 0              # kernel launches so far (the stand-in's clock)
        self._event_at = {}             # event handle -> _launched when it was recorded

    # -- helpers for the tests -------------------------------------------------------------
    def names(self, start=0):
        return [c[0] for c in self.calls[start:]]

    def kernel_launches(self, start=0):
        return sum(KERNEL_CALLS.get(n, 0) for n in self.names(start))

    def mark(self):
        return len(self.calls)

    def args_of(self, name, start=0):
        return [c[1] for c in self.calls[start:] if c[0] == name]

    # -- the entry points --------------------------------------------------------------------
    def __getattr__(self, name):
        if name.startswith('_') or not name.startswith('cb_'):
            raise AttributeError(name)
        if name in PASSTHROUGH:
            return getattr(self._real, name)
        handler = getattr(self, '_do_' + name[3:], None)

        weight = KERNEL_CALLS.get(name, 0)

        def call(*args):
            self.calls.append((name, args))
            self._launched += weight
            return handler(*args) if handler else 0
        return call

    def _handle(self, out):
        _out(out).value = self._next_handle
        self._next_handle += 1
        return 0

    def _do_malloc(self, nbytes, out):
        size = (int(nbytes) + 511) // 512 * 512
        _out(out).value = self._next_dev
        self.device_sizes[self._next_dev] = int(nbytes)
        self._next_dev += size + 512
        return 0

    def _do_free(self, ptr):
        self.device_sizes.pop(int(ptr), None)
        return 0

    def _do_host_alloc(self, nbytes, out):
        # page-aligned like cudaHostAlloc; _host: address -> (length, backing buffer)
        n = max(int(nbytes), 1)
        buf = ctypes.create_string_buffer(n + 4096)
        addr = (ctypes.addressof(buf) + 4095) // 4096 * 4096
        self._host[addr] = (n, buf)
        _out(out).value = addr
        return 0

    def _do_host_free(self, ptr):
        self._host.pop(getattr(ptr, 'value', ptr), None)
        return 0

    def _do_stream_create(self, out):
        return self._handle(out)

    def _do_event_create(self, out):
        return self._handle(out)

    def _do_comm_create(self, ident, rank, world, out):
        return self._handle(out)

    def _do_event_record(self, evt, stream):
        self._event_at[getattr(evt, 'value', evt)] = self._launched
        return 0

    def _do_event_elapsed_ms(self, a, b, out):
        at = self._event_at
        _out(out).value = float(at[getattr(b, 'value', b)] - at[getattr(a, 'value', a)])
        return 0

    def _do_device_count(self, out):
        _out(out).value = 1
        return 0

    def _do_device_info(self, device, name, namelen, maj, mnr, sms, mem, l2):
        name.value = b'recording stand-in (no device)'
        _out(maj).value, _out(mnr).value = 10, 0
        _out(sms).value = self.SM_COUNT
        _out(mem).value = 180 * 1024 ** 3
        _out(l2).value = self.L2_BYTES
        return 0

    def _do_launch_count(self, out):
        _out(out).value = self.kernel_launches()
        return 0

    def _do_module_kernel_info(self, mod, kernel, threads, regs, smem, ctas):
        _out(regs).value, _out(smem).value, _out(ctas).value = 32, 16 * 1024, self.ctas_per_sm
        return 0

    def _do_module_kernel_local_bytes(self, mod, kernel, out):
        _out(out).value = 0
        return 0

    def _do_memcpy_d2h(self, dst, src, nbytes, stream):
        if self.d2h_hook is not None:
            self.d2h_hook(getattr(dst, 'value', dst), int(src), int(nbytes))
        return 0

    def _do_iterate(self, mod, args, grid, stream):
        a = _out(args)
        rec = dict((f[0], getattr(a, f[0])) for f in a._fields_ if f[0] != 'dim')
        rec.update(grid=int(grid), stream=getattr(stream, 'value', stream), module=mod,
                   dim=tuple(a.dim))
        self.iterations.append(rec)
        return 0


def install(monkeypatch):
    """Route cuburn_b200._native through a RecordingLib for the current test."""
    from cuburn_b200 import _native as N
    fake = RecordingLib(N.lib())
    monkeypatch.setattr(N, '_lib', fake)
    monkeypatch.setattr(N, '_initialised', None)
    return fake


def uninstall():
    """Run pending destructors while the stand-in is still in place (their handles are not
    real): call at the end of a test, before monkeypatch restores the library."""
    gc.collect()
