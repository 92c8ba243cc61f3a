"""
ctypes binding of the C-ABI library (include/cuburn_b200.h).

This is the only place the package talks to the device.  There is no fallback:
if ``csrc/libcuburn_b200.so`` is missing or a call fails, an exception is
raised (``NativeError``, or ``MemoryError`` for allocation failures, matching
the reference's use of pycuda exceptions in render.py:140-147).
"""
import ctypes
import math
import os
from ctypes import (c_int, c_int32, c_uint32, c_uint64, c_size_t, c_float,
                    c_char_p, c_void_p, POINTER, byref)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CUBURN_B200_LIB') or os.path.join(HERE, 'csrc', 'libcuburn_b200.so')

CB_OK = 0
CB_ERR_CUDA, CB_ERR_NVRTC, CB_ERR_INVALID, CB_ERR_NOMEM = -1, -2, -3, -4
CB_ERR_NOT_READY = 1

FMT_RGBA_U8, FMT_RGBA_U16, FMT_YUV444P, FMT_YUV444P10, FMT_YUV420P10, \
    FMT_YUV444P12 = range(6)


class NativeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__('cuburn_b200 native call failed (%d): %s' % (status, message))
        self.status = status


class CompileError(NativeError):
    """NVRTC rejected a generated module; the message carries the log."""


class Dims(ctypes.Structure):
    """cb_dims / render.Dimensions: w h aw ah astride."""
    _fields_ = [('w', c_int32), ('h', c_int32), ('aw', c_int32),
                ('ah', c_int32), ('astride', c_int32)]

    def __iter__(self):
        return iter((self.w, self.h, self.aw, self.ah, self.astride))

    def __len__(self):
        return 5

    def __repr__(self):
        return 'Dimensions(w=%d, h=%d, aw=%d, ah=%d, astride=%d)' % tuple(self)

    @property
    def nbins(self):
        return self.ah * self.astride


class IterArgs(ctypes.Structure):
    _fields_ = [('hist', c_uint64), ('seeds', c_uint64), ('points', c_uint64),
                ('params', c_uint64), ('palette', c_uint64), ('dim', Dims),
                ('param_stride', c_int32), ('nts', c_int32),
                ('pal_rows', c_int32), ('fuse_rounds', c_int32), ('swizzle_bins', c_int32),
                ('first_sample', c_uint64), ('nsamples', c_uint64),
                ('total_samples', c_uint64), ('cells', c_uint64),
                ('palette_packed', c_uint64), ('hot_tags', c_uint64),
                ('first_round', c_int32), ('spill', c_uint64), ('spill_bins', c_int32),
                ('spill_count', c_float), ('tickets', c_uint64), ('dynamic', c_int32)]


_SIGNATURES = {
    'cb_last_error': (c_char_p, []),
    'cb_version': (c_char_p, []),
    'cb_device_count': (c_int, [POINTER(c_int)]),
    'cb_device_info': (c_int, [c_int, c_char_p, c_size_t, POINTER(c_int), POINTER(c_int),
                               POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t)]),
    'cb_init': (c_int, [c_int]),
    'cb_device_sync': (c_int, []),
    'cb_launch_count': (c_int, [POINTER(c_uint64)]),
    'cb_calc_dim': (c_int, [c_int, c_int, POINTER(Dims)]),
    'cb_malloc': (c_int, [c_size_t, POINTER(c_uint64)]),
    'cb_free': (c_int, [c_uint64]),
    'cb_host_alloc': (c_int, [c_size_t, POINTER(c_void_p)]),
    'cb_host_free': (c_int, [c_void_p]),
    'cb_host_register': (c_int, [c_void_p, c_size_t]),
    'cb_host_unregister': (c_int, [c_void_p]),
    'cb_stream_create': (c_int, [POINTER(c_void_p)]),
    'cb_stream_destroy': (c_int, [c_void_p]),
    'cb_stream_sync': (c_int, [c_void_p]),
    'cb_stream_wait_event': (c_int, [c_void_p, c_void_p]),
    'cb_event_create': (c_int, [POINTER(c_void_p)]),
    'cb_event_destroy': (c_int, [c_void_p]),
    'cb_event_record': (c_int, [c_void_p, c_void_p]),
    'cb_event_query': (c_int, [c_void_p]),
    'cb_event_sync': (c_int, [c_void_p]),
    'cb_event_elapsed_ms': (c_int, [c_void_p, c_void_p, POINTER(c_float)]),
    'cb_memcpy_h2d': (c_int, [c_uint64, c_void_p, c_size_t, c_void_p]),
    'cb_memcpy_d2h': (c_int, [c_void_p, c_uint64, c_size_t, c_void_p]),
    'cb_memcpy_d2d': (c_int, [c_uint64, c_uint64, c_size_t, c_void_p]),
    'cb_fill32': (c_int, [c_uint64, c_size_t, c_uint32, c_void_p]),
    'cb_mwc_test': (c_int, [c_uint64, c_int, c_int, c_uint64, c_void_p]),
    'cb_interp_rows': (c_int, [c_uint64, c_uint64, c_uint64, c_uint64, c_int, c_float,
                               c_float, c_int, c_void_p]),
    'cb_interp_params': (c_int, [c_uint64, c_int, c_uint64, c_int, c_uint64, c_int,
                                 POINTER(Dims), c_int, c_void_p]),
    'cb_interp_palette': (c_int, [c_uint64, c_uint64, c_uint64, c_uint64, c_float,
                                  c_float, c_int, c_void_p]),
    'cb_module_build': (c_int, [c_char_p, c_char_p, POINTER(c_char_p), POINTER(c_char_p),
                                c_int, POINTER(c_char_p), c_int, POINTER(c_void_p)]),
    'cb_module_destroy': (c_int, [c_void_p]),
    'cb_module_get_cubin': (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    'cb_module_kernel_info': (c_int, [c_void_p, c_char_p, c_int, POINTER(c_int),
                                      POINTER(c_int), POINTER(c_int)]),
    'cb_module_kernel_local_bytes': (c_int, [c_void_p, c_char_p, POINTER(c_int)]),
    'cb_module_launch': (c_int, [c_void_p, c_char_p, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, POINTER(c_void_p), c_void_p]),
    'cb_module_set_global': (c_int, [c_void_p, c_char_p, c_uint64, c_size_t, c_void_p]),
    'cb_iterate': (c_int, [c_void_p, POINTER(IterArgs), c_int, c_void_p]),
    'cb_palette_pack': (c_int, [c_uint64, c_uint64, c_int, c_void_p]),
    'cb_flush_packed': (c_int, [c_uint64, c_uint64, POINTER(Dims), c_void_p]),
    'cb_hist_unswizzle': (c_int, [c_uint64, c_uint64, c_int, POINTER(Dims), c_void_p]),
    'cb_hist_finish': (c_int, [c_uint64, c_uint64, c_uint64, c_int, c_float,
                               POINTER(Dims), c_void_p]),
    'cb_hot_scan': (c_int, [c_uint64, c_uint64, c_uint64, c_uint64, c_uint64, c_int, c_float, c_float,
                            POINTER(Dims), c_void_p]),
    'cb_sort_scratch_words': (c_int, [c_uint64, c_int, POINTER(c_uint64)]),
    'cb_sort_pass': (c_int, [c_uint64, c_uint64, c_uint64, c_int, c_int, c_int, c_uint64, c_void_p]),
    'cb_yuv_to_rgb': (c_int, [c_uint64, c_uint64, POINTER(Dims), c_void_p]),
    'cb_den_blur': (c_int, [c_uint64, c_uint64, c_int, c_int, POINTER(c_float),
                            POINTER(Dims), c_void_p]),
    'cb_den_blur_1c': (c_int, [c_uint64, c_uint64, c_int, c_int, POINTER(c_float),
                               POINTER(Dims), c_void_p]),
    'cb_full_blur': (c_int, [c_uint64, c_uint64, c_int, c_int, POINTER(c_float),
                             POINTER(Dims), c_void_p]),
    'cb_bilateral': (c_int, [c_uint64, c_uint64, c_uint64, c_int, c_int, c_float, c_float,
                             c_float, c_float, c_float, POINTER(Dims), c_void_p]),
    'cb_bilateral_direction': (c_int, [c_uint64, c_uint64, c_uint64, c_int, c_int,
                                       POINTER(c_float), c_float, c_float, c_float, c_float,
                                       c_float, POINTER(Dims), c_void_p]),
    'cb_logscale': (c_int, [c_uint64, c_uint64, c_float, c_float, POINTER(Dims), c_void_p]),
    'cb_apply_gamma': (c_int, [c_uint64, c_uint64, c_float, POINTER(Dims), c_void_p]),
    'cb_haloclip': (c_int, [c_uint64, c_uint64, c_float, POINTER(Dims), c_void_p]),
    'cb_apply_gamma_full_hi': (c_int, [c_uint64, c_uint64, c_float, POINTER(Dims), c_void_p]),
    'cb_smearclip': (c_int, [c_uint64, c_uint64, c_float, c_float, c_float, POINTER(Dims),
                             c_void_p]),
    'cb_plainclip': (c_int, [c_uint64, c_float, c_float, c_float, c_float, POINTER(Dims),
                             c_void_p]),
    'cb_colorclip': (c_int, [c_uint64, c_float, c_float, c_float, c_float, c_float,
                             POINTER(Dims), c_void_p]),
    'cb_logencode': (c_int, [c_uint64, c_uint64, c_float, POINTER(Dims), c_void_p]),
    'cb_convert': (c_int, [c_int, c_uint64, c_uint64, c_int, POINTER(Dims), c_uint64,
                           c_int, c_void_p]),
    'cb_convert_rows': (c_int, [c_int, c_uint64, c_uint64, c_int, POINTER(Dims), c_uint64,
                                c_int, c_int, c_int, c_void_p]),
    'cb_convert_size': (c_int, [c_int, POINTER(Dims), POINTER(c_size_t)]),
    'cb_comm_version': (c_int, [POINTER(c_int)]),
    'cb_comm_unique_id': (c_int, [POINTER(ctypes.c_uint8)]),
    'cb_comm_create': (c_int, [POINTER(ctypes.c_uint8), c_int, c_int, POINTER(c_void_p)]),
    'cb_comm_destroy': (c_int, [c_void_p]),
    'cb_hist_reduce': (c_int, [c_void_p, c_uint64, POINTER(Dims), c_int, c_void_p]),
    'cb_band_rows': (c_int, [c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    'cb_band_gather': (c_int, [c_void_p, c_uint64, POINTER(Dims), c_int, c_void_p]),
}

EXPORTS = tuple(sorted(_SIGNATURES))

_lib = None


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                '%s is missing: build it with `python -m cuburn_b200.build` '
                '(there is no CPU fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status == CB_OK:
        return
    msg = lib().cb_last_error().decode('utf-8', 'replace')
    if status == CB_ERR_NOMEM:
        raise MemoryError(msg)
    if status == CB_ERR_NVRTC:
        raise CompileError(status, msg)
    if status == CB_ERR_INVALID:
        raise ValueError(msg)
    raise NativeError(status, msg)


_initialised = None


def init(device=0):
    global _initialised
    check(lib().cb_init(int(device)))
    _initialised = int(device)


def ensure_init(device=0):
    if _initialised is None:
        init(device)


# ---- thin object wrappers ----------------------------------------------------
def _sptr(stream):
    return stream.handle if stream is not None else None


class Stream(object):
    def __init__(self):
        h = c_void_p()
        check(lib().cb_stream_create(byref(h)))
        self.handle = h

    def synchronize(self):
        check(lib().cb_stream_sync(self.handle))

    def wait_for_event(self, evt):
        check(lib().cb_stream_wait_event(self.handle, evt.handle))

    def __del__(self):
        try:
            if self.handle:
                lib().cb_stream_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Event(object):
    def __init__(self):
        h = c_void_p()
        check(lib().cb_event_create(byref(h)))
        self.handle = h

    def record(self, stream=None):
        check(lib().cb_event_record(self.handle, _sptr(stream)))
        return self

    def query(self):
        r = lib().cb_event_query(self.handle)
        if r == CB_ERR_NOT_READY:
            return False
        check(r)
        return True

    def synchronize(self):
        check(lib().cb_event_sync(self.handle))

    def time_since(self, prior):
        ms = c_float()
        check(lib().cb_event_elapsed_ms(prior.handle, self.handle, byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self.handle:
                lib().cb_event_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class DeviceBuffer(object):
    """An owned device allocation; ``int(buf)`` / ``buf.ptr`` is the address."""
    def __init__(self, nbytes):
        p = c_uint64()
        check(lib().cb_malloc(int(nbytes), byref(p)))
        self.ptr, self.nbytes = p.value, int(nbytes)

    def __int__(self):
        return self.ptr

    def free(self):
        if self.ptr:
            check(lib().cb_free(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            if self.ptr:
                lib().cb_free(self.ptr)
                self.ptr = 0
        except Exception:
            pass

    def view(self, shape, typestr):
        """Object exposing __cuda_array_interface__ (for torch.as_tensor)."""
        return _CudaArrayView(self, shape, typestr)


def preload_nccl():
    """
    Make sure the NCCL this process ends up with is the one PyTorch ships, when PyTorch
    is installed: libtorch_cuda needs symbols of its own NCCL version, and the dynamic
    loader keeps whichever ``libnccl.so.2`` arrives first.  Call before the first
    ``cb_comm_*`` (``multigpu.NativeComm`` does).  Returns the path loaded, or None.
    """
    import importlib.util
    try:
        spec = importlib.util.find_spec('nvidia.nccl')
    except (ImportError, ValueError):
        spec = None
    for base in (spec.submodule_search_locations if spec else []):
        path = os.path.join(base, 'lib', 'libnccl.so.2')
        if os.path.exists(path):
            ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
            return path
    return None


class DeviceSlice(object):
    """A window into someone else's allocation (never freed here)."""
    def __init__(self, buf, offset, nbytes):
        assert 0 <= offset and offset + nbytes <= buf.nbytes
        self.ptr, self.nbytes, self._owner = int(buf) + int(offset), int(nbytes), buf

    def __int__(self):
        return self.ptr

    def view(self, shape, typestr):
        return _CudaArrayView(self, shape, typestr)


class _CudaArrayView(object):
    def __init__(self, buf, shape, typestr):
        self._buf = buf
        self.__cuda_array_interface__ = {
            'shape': tuple(shape), 'typestr': typestr,
            'data': (buf.ptr, False), 'version': 2, 'strides': None}


def to_device(arr, stream=None):
    arr = np.ascontiguousarray(arr)
    buf = DeviceBuffer(arr.nbytes)
    memcpy_htod(buf, arr, stream)
    if stream is None:
        check(lib().cb_stream_sync(None))
    return buf


def memcpy_htod(dst, arr, stream=None, nbytes=None):
    arr = np.ascontiguousarray(arr)
    n = arr.nbytes if nbytes is None else nbytes
    check(lib().cb_memcpy_h2d(int(dst), arr.ctypes.data_as(c_void_p), n, _sptr(stream)))
    if stream is None:
        check(lib().cb_stream_sync(None))


def memcpy_dtoh(arr, src, stream=None, nbytes=None):
    assert arr.flags['C_CONTIGUOUS']
    n = arr.nbytes if nbytes is None else nbytes
    check(lib().cb_memcpy_d2h(arr.ctypes.data_as(c_void_p), int(src), n, _sptr(stream)))
    if stream is None:
        check(lib().cb_stream_sync(None))
    return arr


def from_device(src, shape, dtype):
    out = np.empty(shape, dtype)
    return memcpy_dtoh(out, src)


def fill32(dst, nwords, value=0, stream=None):
    """Stream-ordered 32-bit fill; ``value`` may be an int or a float32."""
    if isinstance(value, (float, np.floating)):
        value = int(np.float32(value).view(np.uint32))
    check(lib().cb_fill32(int(dst), int(nwords), int(value) & 0xffffffff, _sptr(stream)))


class PinnedPool(object):
    """
    Page-locked host arrays, recycled by size once the array (and every view
    of it) has been garbage collected -- stands in for
    pycuda.tools.PageLockedMemoryPool (render.py:93).
    """
    def __init__(self):
        self._free = {}
        self._all = []

    def allocate(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        if not isinstance(shape, (tuple, list)):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        used = math.prod(shape) * dtype.itemsize
        nbytes = max(used, 1)
        bucket = self._free.setdefault(nbytes, [])
        if bucket:
            raw = bucket.pop()
        else:
            p = c_void_p()
            check(lib().cb_host_alloc(nbytes, byref(p)))
            raw = (ctypes.c_char * nbytes).from_address(p.value)
            self._all.append((raw, p.value))
        root = np.frombuffer(raw, dtype=np.uint8)
        weakref.finalize(root, bucket.append, raw)
        return root[:used].view(dtype).reshape(shape)

    def free_all(self):
        for raw, ptr in self._all:
            lib().cb_host_free(c_void_p(ptr))
        self._all, self._free = [], {}


class Module(object):
    """An NVRTC-compiled module (cb_module_build)."""
    def __init__(self, source, name, headers=(), header_names=(), options=()):
        n = len(headers)
        hs = (c_char_p * max(n, 1))(*[h.encode() for h in headers])
        hn = (c_char_p * max(n, 1))(*[h.encode() for h in header_names])
        opts = (c_char_p * max(len(options), 1))(*[o.encode() for o in options])
        h = c_void_p()
        check(lib().cb_module_build(source.encode(), name.encode(), hs, hn, n, opts,
                                    len(options), byref(h)))
        self.handle = h
        self.name = name

    @property
    def cubin(self):
        p, n = c_void_p(), c_size_t()
        check(lib().cb_module_get_cubin(self.handle, byref(p), byref(n)))
        return ctypes.string_at(p.value, n.value)

    def kernel_info(self, kernel, block_threads=256):
        regs, smem, ctas = c_int(), c_int(), c_int()
        check(lib().cb_module_kernel_info(self.handle, kernel.encode(), block_threads,
                                          byref(regs), byref(smem), byref(ctas)))
        return dict(num_regs=regs.value, static_smem=smem.value, ctas_per_sm=ctas.value)

    def local_bytes(self, kernel):
        """Local memory per thread (spills) of a kernel."""
        n = c_int()
        check(lib().cb_module_kernel_local_bytes(self.handle, kernel.encode(), byref(n)))
        return n.value

    def set_global(self, symbol, src, nbytes, stream=None):
        check(lib().cb_module_set_global(self.handle, symbol.encode(), int(src), int(nbytes),
                                         _sptr(stream)))

    def launch(self, kernel, grid, block, args, stream=None, dyn_smem=0):
        """args: list of ctypes values."""
        grid = tuple(grid) + (1,) * (3 - len(grid))
        block = tuple(block) + (1,) * (3 - len(block))
        arr = (c_void_p * len(args))(*[ctypes.cast(byref(a), c_void_p) for a in args])
        check(lib().cb_module_launch(self.handle, kernel.encode(), grid[0], grid[1], grid[2],
                                     block[0], block[1], block[2], dyn_smem, arr,
                                     _sptr(stream)))

    def __del__(self):
        try:
            if self.handle:
                lib().cb_module_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def calc_dim(width, height):
    d = Dims()
    check(lib().cb_calc_dim(int(width), int(height), byref(d)))
    return d


def device_info(device=0):
    name = ctypes.create_string_buffer(256)
    maj, mnr, sms = c_int(), c_int(), c_int()
    mem, l2 = c_size_t(), c_size_t()
    check(lib().cb_device_info(device, name, 256, byref(maj), byref(mnr), byref(sms),
                               byref(mem), byref(l2)))
    return dict(name=name.value.decode(), cc=(maj.value, mnr.value), sm_count=sms.value,
                total_mem=mem.value, l2_bytes=l2.value)


def launch_count():
    """Kernels launched by the library so far in this process."""
    n = c_uint64()
    check(lib().cb_launch_count(byref(n)))
    return n.value


def device_count():
    n = c_int()
    check(lib().cb_device_count(byref(n)))
    return n.value
