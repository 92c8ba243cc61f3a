"""
Host plumbing of the video outputs (x264 / vpxenc / ffmpeg pipes,
cuburn/output.py:139-409).  The encoder binaries are not installed, so a stand-in
script plays the encoder: it copies stdin to stdout (or to the file named last on
its command line) and writes its argument vector to stderr, which is what
``Output.encode`` hands back as the log.  Checked: the exact bytes each output
streams per frame, the command lines, flush semantics, and error reporting.
"""
import json
import os
import stat
import sys
from types import SimpleNamespace as NS

import numpy as np
import pytest

from cuburn_b200 import output

FAKE = '''#!%s
import sys
args = sys.argv[1:]
sys.stderr.write(__import__('json').dumps(args))
data = sys.stdin.buffer.read()
if '--fail' in args:
    sys.exit(3)
if args and args[-1].endswith('.mov'):
    open(args[-1], 'wb').write(data)
else:
    sys.stdout.buffer.write(data)
''' % sys.executable


@pytest.fixture
def fake(tmp_path):
    path = tmp_path / 'fakeenc'
    path.write_text(FAKE)
    path.chmod(path.stat().st_mode | stat.S_IXUSR)
    return str(path)


def rgba16(h, w, seed):
    return np.random.RandomState(seed).randint(0, 65536, (h, w, 4)).astype('u2')


def test_x264_streams_rgb_and_flushes(fake):
    out = output.X264Output(command=fake, crf=17, x264opts='--tune grain')
    a, b = rgba16(6, 8, 1), rgba16(6, 8, 2)
    assert out.encode(a) == ({}, [])
    assert out.encode(b) == ({}, [])
    media, logs = out.encode(None)
    assert list(media) == ['.h264'] and [k for k, _ in logs] == ['x264_color']
    assert media['.h264'].read() == a[:, :, :3].tobytes() + b[:, :, :3].tobytes()
    argv = json.loads(logs[0][1])
    assert argv[:3] == ['--no-progress', '--input-depth', '16']
    for flag, val in (('--crf', '17'), ('--tune', 'grain'), ('--input-csp', 'rgb'),
                      ('--input-res', '8x6'), ('--output-csp', 'i444'), ('--profile', 'high444')):
        assert argv[argv.index(flag) + 1] == val
    assert out.encode(None) == ({}, [])          # nothing pending


def test_x264_alpha_uses_a_second_encoder(fake):
    out = output.X264Output(command=fake, alpha=True)
    a = rgba16(4, 6, 3)
    out.encode(a)
    media, logs = out.encode(None)
    assert sorted(media) == ['_alpha.h264', '_color.h264']
    assert media['_color.h264'].read() == a[:, :, :3].tobytes()
    matte = np.frombuffer(media['_alpha.h264'].read(), 'u2')
    assert np.array_equal(matte[:24], a[:, :, 3].ravel())          # alpha as luma
    assert matte.size == 36 and (matte[24:] == 32767).all()      # flat 4:2:0 chroma
    alog = json.loads(dict(logs)['x264_alpha'])
    assert alog[alog.index('--input-csp') + 1] == 'yv12'
    assert alog[alog.index('--chroma-qp-offset') + 1] == '24'


def test_x264_new_frame_size_starts_a_new_stream(fake):
    out = output.X264Output(command=fake)
    a, b = rgba16(4, 6, 4), rgba16(8, 6, 5)
    out.encode(a)
    media, _ = out.encode(b)                      # size changed: first stream comes back
    assert media['.h264'].read() == a[:, :, :3].tobytes()
    media, _ = out.encode(None)
    assert media['.h264'].read() == b[:, :, :3].tobytes()


def test_encoder_failure_is_an_ioerror(fake, tmp_path):
    out = output.X264Output(command=fake, x264opts='--fail')
    out.encode(rgba16(4, 6, 6))
    with pytest.raises(IOError, match='x264 exited with an error'):
        out.encode(None)
    missing = output.X264Output(command=str(tmp_path / 'no-such-x264'))
    with pytest.raises(IOError, match='cannot start x264'):
        missing.encode(rgba16(4, 6, 7))


def test_vpx_420_decimates_chroma_on_the_host(fake):
    out = output.VPxOutput(codec='vp8', fps=30, crf=12, command=fake)
    assert out.fmt == output.N.FMT_YUV444P and out.shape(NS(w=640, h=360)) == (3, 360, 640)
    out.dim = NS(w=8, h=4)
    buf = np.random.RandomState(8).randint(0, 256, (3, 4, 8)).astype('u1')
    out.encode(buf)
    media, logs = out.encode(None)
    want = buf[0].tobytes() + buf[1, ::2, ::2].tobytes() + buf[2, ::2, ::2].tobytes()
    assert media['.webm'].read() == want
    argv = json.loads(logs[0][1])
    assert '--codec=vp8' in argv and '--cq-level=12' in argv and '--fps=30/1' in argv
    assert argv[argv.index('-w') + 1] == '8' and argv[argv.index('-h') + 1] == '4'
    assert '-t' not in argv                       # threads flag is vp9 only


@pytest.mark.parametrize('pix_fmt,flags,dtype', [
    ('yuv444p', ['--profile=1', '--i444'], 'u1'),
    ('yuv420p10', ['--input-bit-depth=10', '--profile=2'], 'u2'),
    ('yuv444p10', ['--input-bit-depth=10', '--profile=3', '--i444'], 'u2'),
    ('yuv444p12', ['--input-bit-depth=12', '--profile=3', '--i444'], 'u2')])
def test_vp9_high_bit_depth_formats(fake, pix_fmt, flags, dtype):
    out = output.VPxOutput(codec='vp9', pix_fmt=pix_fmt, command=fake)
    dim = NS(w=1920, h=1080)
    out.dim = dim
    assert out.dtype == dtype
    shape = out.shape(dim)
    assert shape == ((1920 * 1080 * 3 // 2,) if pix_fmt == 'yuv420p10' else (3, 1080, 1920))
    buf = (np.arange(np.prod(shape)) % 251).astype(dtype).reshape(shape)
    out.encode(buf)
    media, logs = out.encode(None)
    assert media['.webm'].read() == buf.tobytes()
    argv = json.loads(logs[0][1])
    assert all(f in argv for f in flags) and '--tile-columns=2' in argv and '-t' in argv
    with pytest.raises(ValueError):
        output.VPxOutput(codec='vp8', pix_fmt=pix_fmt)


def test_prores_goes_through_a_named_file(fake):
    out = output.ProResOutput(fps=25, command=fake)
    out.dim = NS(w=6, h=4)
    buf = np.random.RandomState(9).randint(0, 4096, (3, 4, 6)).astype('u2')
    out.encode(buf)
    out.encode(buf)
    media, logs = out.encode(None)
    assert list(media) == ['.mov'] and logs == []
    assert media['.mov'].read() == buf.tobytes() * 2
    assert not os.path.exists(media['.mov'].name)            # only the handle remains


def test_profile_selects_the_encoder_outputs():
    from cuburn_b200 import profile, samples
    gnm = samples.g3()
    for kind, cls, suffix in (('x264', output.X264Output, '.h264'), ('vp9', output.VPxOutput, '.webm'),
                              ('vp8', output.VPxOutput, '.webm'), ('prores', output.ProResOutput, '.mov')):
        gprof = profile.wrap(dict(output=dict(type=kind), fps=30), gnm)
        out = output.get_output_for_profile(gprof)
        assert isinstance(out, cls)
        assert output.get_suffix_for_profile(gprof) == suffix
        if kind != 'x264':
            assert out.fps == 30 if kind == 'prores' else '--fps=30/1' in out.args
