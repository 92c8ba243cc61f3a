"""
Pixel formats: the reference's five assertions (cuburn/code/tests/test_output.py:23-124)
re-expressed against the device kernels, and bit-exact parity with the oracle for
all six formats.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FMTS = ['rgba_u8', 'rgba_u16', 'yuv444p', 'yuv444p10', 'yuv420p10', 'yuv444p12']


def _convert(N, fmt, ins, w, h, seeds, nstreams):
    from cuburn_b200 import _native
    code = FMTS.index(fmt)
    dim = N.calc_dim(w, h)
    src = N.to_device(np.ascontiguousarray(ins, np.float32))
    d_seeds = N.to_device(seeds)
    nbytes = N.c_size_t()
    N.check(N.lib().cb_convert_size(code, N.byref(dim), N.byref(nbytes)))
    dst = N.DeviceBuffer(nbytes.value)
    N.check(N.lib().cb_convert(code, dst.ptr, src.ptr, 12, N.byref(dim), d_seeds.ptr,
                               nstreams, None))
    N.check(N.lib().cb_device_sync())
    dt = np.uint8 if fmt in ('rgba_u8', 'yuv444p') else np.uint16
    out = N.from_device(dst, (nbytes.value // np.dtype(dt).itemsize,), dt)
    return out, N.from_device(d_seeds, seeds.shape, np.uint32)


@pytest.fixture(scope='module')
def fb(native, built):
    from cuburn_b200 import mwc
    dim = native.calc_dim(640, 360)
    return native, dim, mwc.make_seeds(262144, host_seed=13)


def test_clamping_below_0(fb):
    N, dim, seeds = fb
    ins = np.full((dim.ah, dim.astride, 4), -1, np.float32)
    outs = _convert(N, 'yuv444p', ins, 640, 360, seeds, 262144)[0].reshape(3, 360, 640)
    assert np.all(outs[0] == 0)
    assert np.all((outs[1] >= 127) & (outs[1] <= 128))
    assert np.all((outs[2] >= 127) & (outs[2] <= 128))


def test_clamping_above_1(fb):
    N, dim, seeds = fb
    ins = np.full((dim.ah, dim.astride, 4), 5, np.float32)
    outs = _convert(N, 'yuv444p', ins, 640, 360, seeds, 262144)[0].reshape(3, 360, 640)
    assert np.all(outs[0] == 255)
    assert np.all((outs[1] >= 127) & (outs[1] <= 128))
    assert np.all((outs[2] >= 127) & (outs[2] <= 128))


def test_yuv444p10_zero_passthru(fb):
    N, dim, seeds = fb
    ins = np.zeros((dim.ah, dim.astride, 4), np.float32)
    outs = _convert(N, 'yuv444p10', ins, 640, 360, seeds, 262144)[0].reshape(3, 360, 640)
    assert np.all(outs[0] == 0)
    assert np.all((510 < outs[1]) & (outs[1] < 513))
    assert np.all((510 < outs[2]) & (outs[2] < 513))


def test_yuv444p10_chroma_address_preservation(fb):
    N, dim, seeds = fb
    ins = np.zeros((dim.ah, dim.astride, 4), np.float32)
    ins[12, 12, :] = [0, 1, 0, 1]
    ins[13, 13, :] = [0, 1, 0, 1]
    outs = _convert(N, 'yuv444p10', ins, 640, 360, seeds, 262144)[0].reshape(3, 360, 640)
    assert outs[0, 0, 0] > 0 and outs[0, 1, 1] > 0
    assert outs[1, 0, 0] < 500 and outs[1, 1, 1] < 500


def test_yuv420p10_chroma_address_preservation(fb):
    N, dim, seeds = fb
    ins = np.zeros((dim.ah, dim.astride, 4), np.float32)
    ins[12, 12, :] = [0, 1, 0, 1]
    ins[14, 14, :] = [0, 1, 0, 1]
    ins[15, 15, :] = [1, 0, 0, 1]
    w, h = 640, 360
    flat = _convert(N, 'yuv420p10', ins, w, h, seeds, 262144)[0]
    luma = flat[:w * h].reshape(h, w)
    out_cr = flat[w * h: w * h + w * h // 4].reshape(h // 2, w // 2)
    assert luma[0, 0] > 0 and luma[1, 0] == 0 and luma[0, 1] == 0 and luma[1, 1] == 0
    assert luma[2, 2] > 0 and luma[3, 3] > 0
    assert 172 <= out_cr[0, 0] <= 174
    assert 511 <= out_cr[0, 1] <= 512 and 511 <= out_cr[1, 0] <= 512


@pytest.mark.parametrize('fmt', FMTS)
@pytest.mark.parametrize('w,h,nstreams', [(640, 360, 262144), (132, 70, 1000)])
def test_bit_exact_vs_oracle(fb, fmt, w, h, nstreams):
    """Random field incl. negatives, zeros and > 1; ragged sizes; few streams (many rounds)."""
    N, _, seeds = fb
    from oracle import output_ref as O
    dim = N.calc_dim(w, h)
    rs = np.random.RandomState(3)
    ins = rs.uniform(-0.2, 1.3, (dim.ah, dim.astride, 4)).astype(np.float32)
    ins[rs.rand(dim.ah, dim.astride) < 0.2] = 0
    ins[..., 3] = np.abs(ins[..., 3])
    got, gseeds = _convert(N, fmt, ins, w, h, seeds, nstreams)
    want, oseeds = O.convert(fmt, ins, w, h, seeds, nstreams)
    assert np.array_equal(got, want.reshape(-1))
    assert np.array_equal(gseeds, oseeds)


def test_output_module_shapes(native, built):
    from cuburn_b200 import output, profile
    dim = native.calc_dim(640, 360)
    for kind, opts, shape, dt in (('jpeg', {}, (360, 640, 4), 'u1'), ('tiff', {}, (360, 640, 4), 'u2'),
                                  ('raw', {'pix_fmt': 'yuv420p10'}, (640 * 360 * 3 // 2,), 'u2'),
                                  ('raw', {'pix_fmt': 'yuv444p12'}, (3, 360, 640), 'u2')):
        gprof = profile.wrap(dict(output=dict(type=kind, **opts)), {'type': 'animation'})
        out = output.get_output_for_profile(gprof)
        assert out.shape(dim) == shape and out.dtype == dt
    with pytest.raises(ValueError):
        output.get_output_for_profile(profile.wrap(dict(output=dict(type='gif')), {'type': 'animation'}))


@pytest.mark.parametrize('fmt', ['rgba_u8', 'rgba_u16'])
def test_convert_by_row_bands_equals_whole_frame(native, built, fmt):
    """cb_convert_rows: a frame converted in three bands is byte-identical to the frame
    converted at once, and the RNG streams end in the same state after every band."""
    N = native
    from cuburn_b200 import mwc
    w, h = 96, 40
    dim = N.calc_dim(w, h)
    rs = np.random.RandomState(8)
    src = rs.rand(dim.ah, dim.astride, 4).astype(np.float32)
    src[rs.rand(dim.ah, dim.astride) < 0.2] = 0
    code = {'rgba_u8': N.FMT_RGBA_U8, 'rgba_u16': N.FMT_RGBA_U16}[fmt]
    px = 4 if fmt == 'rgba_u8' else 8
    seeds = mwc.make_seeds(1024, host_seed=3)
    d_src = N.to_device(src)

    def run(bands):
        d_seeds, d_dst = N.to_device(seeds), N.DeviceBuffer(w * h * px)
        N.fill32(d_dst, w * h * px // 4, 0)
        states = []
        for r0, r1 in bands:
            N.check(N.lib().cb_convert_rows(code, d_dst.ptr, d_src.ptr, 12, N.byref(dim),
                                            d_seeds.ptr, 1024, r0, r1, None))
            N.check(N.lib().cb_device_sync())
            states.append(N.from_device(d_seeds, (1024, 3), np.uint32))
        return N.from_device(d_dst, (h, w * px), np.uint8), states
    whole, s_whole = run([(0, h)])
    parts = []
    for band in ((0, 13), (13, 14), (14, h)):
        got, st = run([band])
        assert np.array_equal(st[0], s_whole[0])
        assert np.array_equal(got[band[0]:band[1]], whole[band[0]:band[1]])
        assert not got[:band[0]].any() and not got[band[1]:].any()
        parts.append(got[band[0]:band[1]])
    assert np.array_equal(np.concatenate(parts), whole)
    # planar formats convert whole frames only
    with pytest.raises(ValueError):
        N.check(N.lib().cb_convert_rows(N.FMT_YUV444P, 0, d_src.ptr, 12, N.byref(dim), 0, 1024,
                                        0, 8, None))
