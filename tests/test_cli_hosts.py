"""
The two host programs around the render path -- main.py (main.py:28-94) and the job
farm's worker (distribute.py:73-127) -- run on the CPU against a stand-in render
backend: the frame pipeline's ordering, what gets written where, the worker's wire
messages.  (The same programs run on the device in tests/test_render_gpu.py.)
"""
import importlib
import io
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class FakeEvent(object):
    def __init__(self, log, k, ms, polls):
        self.log, self.k, self.ms, self.polls = log, k, ms, polls

    def synchronize(self):
        self.log.append(('sync', self.k))

    def query(self):
        self.log.append(('query', self.k))
        self.polls -= 1
        return self.polls <= 0

    def time(self):
        return self.ms


class FakeManager(object):
    """queue_frame hands back frame k filled with the value k + 1."""
    frame_ms = 7.0
    polls = 3

    def __init__(self, *a, **kw):
        self.log = FakeManager.last_log = []
        self.k = 0

    def queue_frame(self, rdr, gnm, gprof, tc, copy=True):
        k, self.k = self.k, self.k + 1
        self.log.append(('queue', k, tc))
        frame = np.full((int(gprof.height), int(gprof.width), 4), k + 1, np.uint8)
        return FakeEvent(self.log, k, self.frame_ms, self.polls), frame


class FakeRenderer(object):
    def __init__(self, gnm, gprof, keep=False, arch=None):
        from cuburn_b200 import output
        self.out = output.get_output_for_profile(gprof)
        self.keep = keep


@pytest.fixture
def backend(monkeypatch):
    from cuburn_b200 import _native, render
    inits = []
    monkeypatch.setattr(_native, 'init', lambda dev=0: inits.append(dev))
    monkeypatch.setattr(render, 'RenderManager', FakeManager)
    monkeypatch.setattr(render, 'Renderer', FakeRenderer)
    FakeManager.frame_ms, FakeManager.polls = 7.0, 3
    return inits


def test_frame_pipeline_queues_one_frame_ahead():
    from cuburn_b200 import render
    mgr = FakeManager()
    gprof = types.SimpleNamespace(width=8, height=4)
    got = []
    for n, evt, frame in render.frame_pipeline(mgr, None, None, gprof, [0.1, 0.2, 0.3]):
        mgr.log.append(('yield', n))
        got.append((n, int(frame[0, 0, 0]), evt.k))
    assert got == [(1, 1, 0), (2, 2, 1), (3, 3, 2)]
    # frame k + 1 is on the device before the host waits for frame k
    assert mgr.log == [('queue', 0, 0.1), ('queue', 1, 0.2), ('sync', 0), ('yield', 1),
                       ('queue', 2, 0.3), ('sync', 1), ('yield', 2), ('sync', 2), ('yield', 3)]
    assert list(render.frame_pipeline(FakeManager(), None, None, gprof, [])) == []
    # a caller-supplied wait replaces synchronize
    waited = []
    mgr = FakeManager()
    for _ in render.frame_pipeline(mgr, None, None, gprof, [0.5, 0.6], wait=waited.append):
        pass
    assert [e.k for e in waited] == [0, 1] and not [x for x in mgr.log if x[0] == 'sync']


def load_main():
    sys.modules.pop('main', None)
    return importlib.import_module('main')


def test_main_renders_a_still_to_jpeg_and_raw_preview(backend, tmp_path, capsys):
    from PIL import Image
    main = load_main()
    raw = tmp_path / 'preview.raw'
    main.run(['sample:G3', '-P', 'preview', '--still', '-o', str(tmp_path), '--raw', str(raw),
              '--device', '0'])
    out = tmp_path / 'G3_00002.jpg'                 # --still renders frame 2 (Q4)
    assert out.exists(), os.listdir(tmp_path)
    img = np.array(Image.open(out))
    assert img.shape == (360, 640, 3) and np.all(img == 1)
    assert raw.stat().st_size == 640 * 360 * 4 and not (tmp_path / 'preview.raw.tmp').exists()
    err = capsys.readouterr().err
    assert '0: ' in err and 'G3_00002 (  1/  1), 7ms' in err
    assert backend == [0]
    assert FakeManager.last_log[0][:2] == ('queue', 0) and ('sync', 0) in FakeManager.last_log
    # --resume: the finished frame is not rendered again (no device is even initialised)
    main.run(['sample:G3', '-P', 'preview', '--still', '-o', str(tmp_path), '--resume'])
    assert backend == [0] and 'G3_00002' not in capsys.readouterr().err


def test_main_renders_an_animation_in_order_and_polls_slow_frames(backend, tmp_path, capsys,
                                                                  monkeypatch):
    from PIL import Image
    main = load_main()
    monkeypatch.setattr(main, 'POLL_INTERVAL_S', 0.0)
    FakeManager.frame_ms = 2500.0                   # slow frames: poll instead of blocking
    main.run(['sample:G6F', '-P', 'preview', '--duration', '1', '--fps', '3', '--skip', '0',
              '-o', str(tmp_path)])
    names = sorted(p.name for p in tmp_path.glob('*.jpg'))
    assert names == ['G6F_%05d.jpg' % i for i in (1, 2, 3)]
    for i, n in enumerate(names):
        assert np.all(np.array(Image.open(tmp_path / n)) == i + 1)      # frames in order
    log = FakeManager.last_log
    # every job is one frame here (an image per frame): queue, wait, next job
    assert [x[:2] for x in log if x[0] == 'queue'] == [('queue', 0), ('queue', 1), ('queue', 2)]
    # first wait blocks (nothing is known about frame times yet), later ones poll
    assert ('sync', 0) in log and ('sync', 1) not in log
    assert [x for x in log if x == ('query', 1)] == [('query', 1)] * 3
    err = capsys.readouterr().err
    assert 'G6F_00001 (  1/  1), 2500ms' in err and not err.startswith('0: ')


def test_main_print_and_missing_flame(backend, capsys):
    main = load_main()
    main.run(['sample:G3', '--print'])
    doc = json.loads(capsys.readouterr().out)
    assert doc['type'] == 'animation' and sorted(doc['xforms']) == ['0', '1', '2']
    assert backend == []
    with pytest.raises(SystemExit):
        main.run([])


def test_worker_speaks_the_wire_protocol(backend, capsys):
    """distribute.work: READY, then per frame OUTPUT_FILE + suffix + length-prefixed file,
    CLOSING_ENCODER, DONE (distribute.py:73-127)."""
    sys.modules.pop('distribute', None)
    D = importlib.import_module('distribute')
    from cuburn_b200 import samples
    from cuburn_b200 import profile as P
    gnm = samples.g6f(animated=True)
    prof = dict(P.BUILTIN['preview'], duration=1.0, fps=3, skip=0)
    gprof = P.wrap(prof, gnm)
    times = [t for _, ts in P.enumerate_times(gprof) for t in ts]
    job = dict(profile=prof, genome=gnm, times=times, name='whirl')
    stdin, stdout = io.BytesIO(), io.BytesIO()
    D.write_str(stdin, json.dumps(job))
    stdin.seek(0)
    D.work(types.SimpleNamespace(device=1), stdin, stdout)
    stdout.seek(0)
    assert D.read_str(stdout) == D.READY
    sizes = []
    for _ in times:
        assert D.read_str(stdout) == D.OUTPUT_FILE
        assert D.read_str(stdout) == '.jpg'
        blob = io.BytesIO()
        D.copy_filelike(stdout, blob)
        assert blob.getvalue()[:2] == b'\xff\xd8'                       # a JPEG
        sizes.append(len(blob.getvalue()))
    assert len(sizes) == 3
    assert D.read_str(stdout) == D.CLOSING_ENCODER
    assert D.read_str(stdout) == D.DONE and stdout.read() == b''
    assert backend == [1]
    err = capsys.readouterr().err
    assert '/1: whirl (  1/  3), 7ms' in err and 'whirl (  3/  3)' in err
    # a dispatcher with nothing left sends DONE instead of a job
    stdin, stdout = io.BytesIO(), io.BytesIO()
    D.write_str(stdin, D.DONE)
    stdin.seek(0)
    D.work(types.SimpleNamespace(device=0), stdin, stdout)
    stdout.seek(0)
    assert D.read_str(stdout) == D.READY and stdout.read() == b''
