"""
Output modules: device-side pixel-format conversion, the device->host copy, and
host-side encoding of finished frames.

Mirrors the reference interface (cuburn/output.py:28-66, 411-434):
``Output.convert(fb, gprof, dim, stream)`` writes the converted frame to
``fb.d_back``; ``.copy(fb, dim, pool, stream)`` schedules the D2H copy into
pinned memory and returns the array; ``.encode(host_frame | None)`` returns
``({suffix: file-like}, [(key, log)])``.  JPEG / PNG go through Pillow; 16-bit
TIFF is written by a small built-in baseline-TIFF writer; the planar YUV
formats are available as raw planes (``type: raw``); the video types stream
frames into an external encoder process (``x264``, ``vpxenc``, ``ffmpeg``) with
the reference's command lines (output.py:139-409) -- the binaries are not part of
this package, and a missing one is reported when the first frame is encoded.
"""
import io
import shlex
import struct
import subprocess
import tempfile

import numpy as np

from . import _native as N


def _h(stream):
    return stream.handle if stream is not None else None


def launchC(fmt, dim, fb, stream, rows=None):
    """Convert fb.d_front -> fb.d_back (output.py:21-26); ``rows = (row0, row1)``
    restricts it to those output rows (multi-GPU stills: one band per GPU)."""
    row0, row1 = rows if rows is not None else (0, dim.h)
    N.check(N.lib().cb_convert_rows(fmt, fb.d_back.ptr, fb.d_front.ptr, fb.gutter,
                                    N.byref(dim), fb.d_seeds.ptr, fb.nstreams,
                                    int(row0), int(row1), _h(stream)))


class Output(object):
    fmt = None
    dtype = 'u1'

    def shape(self, dim):
        raise NotImplementedError()

    def convert(self, fb, gnm, dim, stream=None, rows=None):
        launchC(self.fmt, dim, fb, stream, rows)

    def copy(self, fb, dim, pool, stream=None, rows=None, out=None):
        """Schedule the D2H copy of the converted frame (or of its ``rows``) into pinned
        memory: a fresh array from ``pool``, or ``out`` -- e.g. a frame in shared memory
        that every GPU of a banded still copies its own rows into."""
        h_out = out if out is not None else pool.allocate(self.shape(dim), self.dtype)
        if rows is None:
            N.memcpy_dtoh(h_out, fb.d_back, stream)
            return h_out
        row0, row1 = rows
        if row1 > row0:
            band = h_out[row0:row1]                 # (h, w, 4) formats: rows are contiguous
            off = row0 * band.strides[0]
            N.memcpy_dtoh(band, N.DeviceSlice(fb.d_back, off, band.nbytes), stream)
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


class EncoderPipe(object):
    """
    One external encoder process: raw frames go to its stdin, its stdout is
    collected in an unnamed temporary file (or it writes a named file itself), its
    stderr -- the log -- goes to another temporary file (a pipe that nobody drains
    fills up after a few hundred frames of ``--log-level debug`` and blocks the
    encoder, and with it the render loop).  ``finish()`` closes stdin, waits, and returns
    ``(file positioned at 0, log)``; a non-zero exit status is an ``IOError``, as
    in the reference (output.py:172-173, 255-256).
    """
    def __init__(self, argv, name, named_suffix=None):
        self.name = name
        self.named = None
        if named_suffix is not None:
            # the encoder wants a seekable path (mov muxer): it writes the file itself
            self.named = tempfile.NamedTemporaryFile(suffix=named_suffix)
            argv = [a.replace('{fn}', self.named.name) for a in argv]
            self.outf = None
        else:
            self.outf = tempfile.TemporaryFile()
        self.errf = tempfile.TemporaryFile()
        try:
            self.proc = subprocess.Popen([str(a) for a in argv], stdin=subprocess.PIPE,
                                         stderr=self.errf,
                                         stdout=self.outf if self.outf is not None
                                         else subprocess.DEVNULL)
        except OSError as e:
            raise IOError('cannot start %s encoder "%s": %s' % (name, argv[0], e))

    def _read_log(self):
        self.errf.seek(0)
        return self.errf.read().decode(errors='replace')

    def write(self, buf):
        try:
            self.proc.stdin.write(memoryview(np.ascontiguousarray(buf)).cast('B'))
        except (IOError, OSError) as e:
            self.proc.wait()
            raise IOError('%s stopped reading frames: %s\n%s'
                          % (self.name, e, self._read_log()))

    def finish(self):
        self.proc.communicate()
        log = self._read_log()
        self.errf.close()
        if self.proc.returncode:
            raise IOError('%s exited with an error\n%s' % (self.name, log))
        if self.named is not None:
            outf = open(self.named.name, 'rb')      # a new handle; the name goes away
            self.named.close()
        else:
            outf = self.outf
            outf.seek(0)
        return outf, log


class X264Output(Output):
    """16-bit RGB frames into ``x264`` (output.py:196-311); with ``alpha`` a second
    encoder receives the alpha plane as the luma of a 4:2:0 stream with flat chroma."""
    fmt = N.FMT_RGBA_U16
    dtype = 'u2'
    profiles = {'normal': '--profile high444 --level 4.2', '': ''}
    base = ('--no-progress --input-depth 16 --sync-lookahead 0 '
            '--rc-lookahead 5 --muxer raw -o - - --log-level debug')

    def __init__(self, profile='normal', csp='i444', crf=15, command='x264', x264opts='',
                 alpha=False):
        self.args = shlex.split(' '.join([command, self.base, self.profiles[profile],
                                          '--crf', str(crf), x264opts]))
        self.alpha, self.csp = alpha, csp
        self.framesize = None
        self.color = self.matte = None

    def shape(self, dim):
        return (dim.h, dim.w, 4)

    def _spawn(self, framesize, matte):
        extras = ['--input-csp', 'yv12' if matte else 'rgb', '--demuxer', 'raw',
                  '--input-res', '%dx%d' % (framesize[1], framesize[0])]
        if matte:
            extras += ['--output-csp', 'i420', '--chroma-qp-offset', '24']
        else:
            extras += ['--output-csp', self.csp]
        return EncoderPipe(self.args + extras, 'x264')

    def _flush(self):
        if self.color is None:
            return {}, []
        outf, log = self.color.finish()
        self.color = None
        if self.matte is not None:
            aoutf, alog = self.matte.finish()
            self.matte = None
            return ({'_color.h264': outf, '_alpha.h264': aoutf},
                    [('x264_color', log), ('x264_alpha', alog)])
        return {'.h264': outf}, [('x264_color', log)]

    def encode(self, buf):
        out = ({}, [])
        if buf is None or self.framesize != buf.shape[:2]:
            out = self._flush()          # a new frame size starts a new stream
        if buf is None:
            return out
        if self.color is None:
            self.framesize = buf.shape[:2]
            self.color = self._spawn(self.framesize, False)
            if self.alpha:
                self.matte = self._spawn(self.framesize, True)
        self.color.write(buf[:, :, :3])
        if self.matte is not None:
            self.matte.write(buf[:, :, 3])
            self.matte.write(np.full(buf.shape[0] * buf.shape[1] // 2, 32767, 'u2'))
        return out


class VPxOutput(Output):
    """Planar YUV frames into ``vpxenc`` (output.py:313-409)."""
    base = ('--end-usage=3 -p 1 -q --cpu-used=-8 --lag-in-frames=5 '
            '--min-q=2 --disable-kf --arnr-maxframes=3 -o - -')
    _PIX = {  # pix_fmt: (device format, dtype, extra encoder flags)
        'yuv420p': (N.FMT_YUV444P, 'u1', []),       # subsampled on the host, see encode
        'yuv444p': (N.FMT_YUV444P, 'u1', ['--profile=1', '--i444']),
        'yuv420p10': (N.FMT_YUV420P10, 'u2', ['-b', '10', '--input-bit-depth=10', '--profile=2']),
        'yuv444p10': (N.FMT_YUV444P10, 'u2', ['-b', '10', '--input-bit-depth=10', '--profile=3',
                                              '--i444']),
        'yuv444p12': (N.FMT_YUV444P12, 'u2', ['-b', '12', '--input-bit-depth=12', '--profile=3',
                                              '--i444']),
    }

    def __init__(self, codec='vp9', fps=24, crf=15, pix_fmt='yuv420p', command='vpxenc'):
        if pix_fmt not in self._PIX:
            raise ValueError('Invalid pix_fmt: ' + pix_fmt)
        if pix_fmt != 'yuv420p' and codec != 'vp9':
            raise ValueError('%s needs codec vp9' % pix_fmt)
        self.codec, self.pix_fmt = codec, pix_fmt
        self.fmt, self.dtype, extra = self._PIX[pix_fmt]
        self.args = shlex.split(command) + self.base.split() + extra
        self.args += ['--codec=' + codec, '--cq-level=' + str(crf), '--fps=%d/1' % fps]
        if codec == 'vp9':
            self.args += ['-t', '4']
        self.dim = None
        self.pipe = None

    def shape(self, dim):
        if self.pix_fmt == 'yuv420p10':
            return (dim.h * dim.w * 6 // 4,)
        return (3, dim.h, dim.w)

    def convert(self, fb, gnm, dim, stream=None, rows=None):
        self.dim = dim
        launchC(self.fmt, dim, fb, stream, rows)

    def _spawn(self, w, h):
        extras = ['-w', w, '-h', h]
        columns = int(max(0, min(3, np.log2(w) - 8.9)))
        if columns:
            extras.append('--tile-columns=%d' % columns)
        return EncoderPipe(self.args + extras, 'vpxenc')

    def encode(self, buf):
        if buf is None:
            if self.pipe is None:
                return {}, []
            outf, log = self.pipe.finish()
            self.pipe = None
            return {'.webm': outf}, [('webm', log)]
        if self.pipe is None:
            if self.dim is not None:
                width, height = self.dim.w, self.dim.h
            elif buf.ndim == 3:
                height, width = buf.shape[1:]
            else:
                raise ValueError('frame size unknown: convert() has not run')
            self.pipe = self._spawn(width, height)
        if self.pix_fmt == 'yuv420p':
            # 4:4:4 planes from the device, chroma decimated here (output.py:394-398)
            self.pipe.write(buf[0])
            self.pipe.write(buf[1, ::2, ::2])
            self.pipe.write(buf[2, ::2, ::2])
        else:
            self.pipe.write(buf)
        return {}, []


class ProResOutput(Output):
    """12-bit 4:4:4 planes into ``ffmpeg -c:v prores`` (output.py:139-194); the mov
    muxer needs a seekable file, so ffmpeg writes a named temporary file."""
    fmt = N.FMT_YUV444P12
    dtype = 'u2'
    cmd = ('-loglevel panic -f rawvideo -pix_fmt yuv444p12le -s {w}x{h} -r {fps} -i - '
           '-c:v prores -f mov -y {fn}')

    def __init__(self, fps=24, command='ffmpeg'):
        self.fps, self.command = fps, command
        self.dim = None
        self.pipe = None

    def shape(self, dim):
        return (3, dim.h, dim.w)

    def convert(self, fb, gnm, dim, stream=None, rows=None):
        self.dim = dim
        launchC(self.fmt, dim, fb, stream, rows)

    def encode(self, buf):
        if buf is None:
            if self.pipe is None:
                return {}, []
            outf, _ = self.pipe.finish()
            self.pipe = None
            return {'.mov': outf}, []
        if self.pipe is None:
            h, w = (self.dim.h, self.dim.w) if self.dim is not None else buf.shape[1:]
            argv = shlex.split(self.command) + \
                [a.format(w=w, h=h, fps=self.fps, fn='{fn}') for a in self.cmd.split()]
            self.pipe = EncoderPipe(argv, 'ffmpeg', named_suffix='.mov')
        self.pipe.write(buf)
        return {}, []


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
    if handler == 'x264':
        return X264Output(**opts)
    if handler in ('vp8', 'vp9'):
        return VPxOutput(codec=handler, fps=gprof.fps, **opts)
    if handler == 'prores':
        return ProResOutput(fps=gprof.fps, **opts)
    raise ValueError('Invalid output type "%s".' % handler)
