import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = '/root/reference'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'production_schedule: run with the shipped default of '
                            'RenderManager.schedule (dynamic unit claiming) instead of the '
                            'reproducible static schedule the other tests compare runs under')


@pytest.fixture(autouse=True)
def _static_schedule(request, monkeypatch):
    """Most GPU tests compare two runs of the chaos game sample for sample (float4 vs packed
    cells, hot bins vs plain, one launch vs chunks, reseed and repeat): they run under
    schedule = 'static', where a frame is a pure function of its seeds.  Tests marked
    production_schedule (statistical parity with the oracle at the benchmark sizes,
    conservation, frame PSNR) keep the shipped default."""
    if request.node.get_closest_marker('production_schedule') is None:
        from cuburn_b200 import render
        monkeypatch.setattr(render.RenderManager, 'schedule', 'static')


@pytest.fixture(scope='session')
def native():
    """The C-ABI library initialised on cuda:0 (GPU tests only)."""
    from cuburn_b200 import _native as N
    try:
        ndev = N.device_count()
    except Exception as e:          # no driver on this host
        pytest.skip('no CUDA driver: %s' % e)
    if ndev < 1:
        pytest.skip('no CUDA device')
    N.init(0)
    return N


@pytest.fixture(scope='session')
def built():
    """Make sure the native library, multiplier table and C oracle exist."""
    from cuburn_b200.build import build_all
    from oracle.build import build
    build_all()
    build()
    return True


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, 'cuburn'))
