"""
Output modules: device-side pixel-format conversion, the device->host copy, and
host-side encoding of finished frames.

Mirrors the reference interface (cuburn/output.py:28-66, 411-434):
``Output.convert(fb, gprof, dim, stream)`` writes the converted frame to
``fb.d_back``; ``.copy(fb, dim, pool, stream)`` schedules the D2H copy into
pinned memory and returns the array; ``.encode(host_frame | None)`` returns
``({suffix: file-like}, [(key, log)])``.  JPEG / PNG go through Pillow; 16-bit
TIFF is written by a small built-in baseline-TIFF writer; the planar YUV
formats are available as raw planes (``type: raw``) for an external encoder --
the x264 / vpx / ffmpeg subprocess plumbing of the reference is out of scope.
"""
import io
import struct

import numpy as np

from . import _native as N


def _h(stream):
    return stream.handle if stream is not None else None


def launchC(fmt, dim, fb, stream):
    """Convert fb.d_front -> fb.d_back (output.py:21-26)."""
    N.check(N.lib().cb_convert(fmt, fb.d_back.ptr, fb.d_front.ptr, fb.gutter,
                               N.byref(dim), fb.d_seeds.ptr, fb.nstreams, _h(stream)))


class Output(object):
    fmt = None
    dtype = 'u1'

    def shape(self, dim):
        raise NotImplementedError()

    def convert(self, fb, gnm, dim, stream=None):
        launchC(self.fmt, dim, fb, stream)

    def copy(self, fb, dim, pool, stream=None):
        h_out = pool.allocate(self.shape(dim), self.dtype)
        N.memcpy_dtoh(h_out, fb.d_back, stream)
        return h_out

    def encode(self, host_frame):
        raise NotImplementedError()


class PILOutput(Output):
    fmt = N.FMT_RGBA_U8
    dtype = 'u1'

    def __init__(self, codec='jpeg', quality=100, alpha=False):
        from PIL import Image  # noqa: F401  (fail early if Pillow is missing)
        self.type, self.quality, self.alpha = codec, quality, alpha

    def shape(self, dim):
        return (dim.h, dim.w, 4)

    def _convert_buf(self, buf):
        from PIL import Image
        out = io.BytesIO()
        img = Image.fromarray(np.ascontiguousarray(buf))
        img.save(out, self.type, quality=self.quality)
        out.seek(0)
        return out

    def encode(self, buf):
        if buf is None:
            return {}, []
        if self.type == 'jpeg':
            out = self._convert_buf(buf[:, :, :3])
            if self.alpha:
                alpha = self._convert_buf(buf[:, :, 3])
                return {'_color.jpg': out, '_alpha.jpg': alpha}, []
            return {'.jpg': out}, []
        return {'.' + self.type: self._convert_buf(buf if self.alpha else buf[:, :, :3])}, []


def _tiff_bytes(arr):
    """Minimal baseline TIFF (little-endian, uncompressed, 16-bit RGB/RGBA)."""
    h, w, ch = arr.shape
    data = np.ascontiguousarray(arr.astype('<u2')).tobytes()
    tags = []

    def tag(code, typ, count, value):
        tags.append((code, typ, count, value))
    nent = 11 + (1 if ch == 4 else 0)
    ifd_off = 8
    bps_off = ifd_off + 2 + nent * 12 + 4
    data_off = bps_off + 2 * ch
    tag(256, 4, 1, w)
    tag(257, 4, 1, h)
    tag(258, 3, ch, bps_off)
    tag(259, 3, 1, 1)
    tag(262, 3, 1, 2)
    tag(273, 4, 1, data_off)
    tag(277, 3, 1, ch)
    tag(278, 4, 1, h)
    tag(279, 4, 1, len(data))
    tag(284, 3, 1, 1)
    tag(339, 3, 1, 1)
    if ch == 4:
        tag(338, 3, 1, 2)
    tags.sort()
    out = io.BytesIO()
    out.write(b'II' + struct.pack('<HI', 42, ifd_off))
    out.write(struct.pack('<H', len(tags)))
    for code, typ, count, value in tags:
        if typ == 3 and count == 1:
            out.write(struct.pack('<HHIHH', code, typ, count, value, 0))
        else:
            out.write(struct.pack('<HHII', code, typ, count, value))
    out.write(struct.pack('<I', 0))
    out.write(struct.pack('<%dH' % ch, *([16] * ch)))
    out.write(data)
    out.seek(0)
    return out


class TiffOutput(Output):
    fmt = N.FMT_RGBA_U16
    dtype = 'u2'

    def __init__(self, alpha=False):
        self.alpha = alpha

    def shape(self, dim):
        return (dim.h, dim.w, 4)

    def encode(self, buf):
        if buf is None:
            return {}, []
        if not self.alpha:
            buf = buf[:, :, :3]
        return {'.tiff': _tiff_bytes(buf)}, []


class RawPlanarOutput(Output):
    """Planar YUV frames as raw bytes, for an external encoder."""
    _FORMATS = {
        'yuv444p': (N.FMT_YUV444P, 'u1'), 'yuv444p10': (N.FMT_YUV444P10, 'u2'),
        'yuv420p10': (N.FMT_YUV420P10, 'u2'), 'yuv444p12': (N.FMT_YUV444P12, 'u2'),
        'rgba': (N.FMT_RGBA_U8, 'u1'), 'rgba16': (N.FMT_RGBA_U16, 'u2'),
    }

    def __init__(self, pix_fmt='yuv444p', **unused):
        if pix_fmt not in self._FORMATS:
            raise ValueError('Invalid pixel format "%s".' % pix_fmt)
        self.pix_fmt = pix_fmt
        self.fmt, self.dtype = self._FORMATS[pix_fmt]

    def shape(self, dim):
        if self.pix_fmt in ('rgba', 'rgba16'):
            return (dim.h, dim.w, 4)
        if self.pix_fmt == 'yuv420p10':
            return (dim.h * dim.w * 3 // 2,)
        return (3, dim.h, dim.w)

    def encode(self, buf):
        if buf is None:
            return {}, []
        return {'.' + self.pix_fmt: io.BytesIO(np.ascontiguousarray(buf).tobytes())}, []


_EXT = dict(jpeg='.jpg', png='.png', tiff='.tiff', x264='.h264', prores='.mov',
            vp8='.webm', vp9='.webm', raw='.raw')


def get_suffix_for_profile(gprof):
    opts = dict(gprof.output._val)
    kind = opts.get('type', 'jpeg')
    if kind == 'raw':
        return '.' + opts.get('pix_fmt', 'yuv444p')
    ext = _EXT[kind]
    if opts.get('alpha'):
        ext = '_color' + ext
    return ext


def get_output_for_profile(gprof):
    opts = dict(gprof.output._val)
    handler = opts.pop('type', 'jpeg')
    if handler in ('jpeg', 'png'):
        return PILOutput(codec=handler, **opts)
    if handler == 'tiff':
        return TiffOutput(**opts)
    if handler == 'raw':
        return RawPlanarOutput(**opts)
    if handler in ('x264', 'vp8', 'vp9', 'prores'):
        raise NotImplementedError(
            'output type "%s" pipes frames to an external encoder binary, which '
            'this build does not drive; use type "raw" with the matching pix_fmt '
            'and feed the planes to the encoder yourself' % handler)
    raise ValueError('Invalid output type "%s".' % handler)
