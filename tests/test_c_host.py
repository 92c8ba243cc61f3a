"""
The drop-in boundary used from C: examples/c_host.c drives the filter chain and the
RGBA8 output through the C ABI alone.  Without a GPU it must build against the
header and the library and bow out cleanly (exit status 77); with one it renders a
frame.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = str(tmp_path / 'c_host')
    csrc = os.path.join(ROOT, 'cuburn_b200', 'csrc')
    cmd = ['gcc', '-O2', '-Wall', '-Werror', '-I' + os.path.join(ROOT, 'include'),
           os.path.join(ROOT, 'examples', 'c_host.c'), '-L' + csrc, '-lcuburn_b200', '-lm',
           '-Wl,-rpath,' + csrc, '-o', exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_c_host_builds_and_reports_a_missing_gpu(built, tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip('a GPU is present: covered by the gpu test')
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 77 and 'no usable GPU' in r.stderr


@pytest.mark.gpu
def test_c_host_renders_a_frame(native, built, tmp_path):
    exe = build(tmp_path)
    out = tmp_path / 'frame.ppm'
    r = subprocess.run([exe, str(out)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    assert '320x180 RGBA8' in r.stdout
    data = out.read_bytes()
    assert data.startswith(b'P6\n320 180\n255\n')
    img = np.frombuffer(data[len(b'P6\n320 180\n255\n'):], 'u1').reshape(180, 320, 3)
    assert img.max() > 100 and (img.sum(axis=2) > 30).mean() > 0.05
    # three blobs of different colour: the brightest pixels are not grey
    bright = img[img.sum(axis=2) > 200].astype(int)
    assert (bright.max(axis=1) - bright.min(axis=1)).mean() > 10
