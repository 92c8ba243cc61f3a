"""
Job farm (distribute.py): the reference's framing (distribute.py:34-71) and dispatch
semantics (:131-248) -- retries, worker back-off, atomic file appearance -- checked
on the CPU with a stand-in worker that speaks the protocol without rendering.
"""
import io
import json
import os
import stat
import struct
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distribute as D          # noqa: E402


def test_framing_matches_the_reference_wire_format():
    buf = io.BytesIO()
    D.write_str(buf, 'worker ready')
    assert buf.getvalue() == b'\x00\x00\x00\x0cworker ready'           # '>u4' length + bytes
    payload = io.BytesIO(b'x' * (D.CHUNK + 17))
    D.write_filelike(buf, payload)
    assert buf.getvalue()[16:24] == struct.pack('>Q', D.CHUNK + 17)     # '>u8' length
    buf.seek(0)
    assert D.read_str(buf) == 'worker ready'
    out = io.BytesIO()
    D.copy_filelike(buf, out)
    assert out.getvalue() == payload.getvalue()
    with pytest.raises(EOFError):
        D.read_str(io.BytesIO(b'\x00\x00\x00\x09short'))
    with pytest.raises(ValueError):
        D.read_str(io.BytesIO(b'\x01\x00\x00\x00'))
    with pytest.raises(EOFError):
        D.copy_filelike(io.BytesIO(struct.pack('>Q', 10) + b'abc'), io.BytesIO())


STUB = '''#!%s
import json, os, sys
sys.path.insert(0, %r)
import io, distribute as D
out, inp = sys.stdout.buffer, sys.stdin.buffer
assert sys.argv[1] == 'work' and sys.argv[2] == '--device'
D.write_str(out, D.READY)
text = D.read_str(inp)
if text == D.DONE:
    sys.exit(0)
job = json.loads(text)
flag = os.environ.get('STUB_FAIL_FLAG')
if os.environ.get('STUB_ALWAYS_FAIL') or (flag and not os.path.exists(flag)):
    if flag:
        open(flag, 'w').close()
    D.write_str(out, D.OUTPUT_FILE)
    D.write_str(out, '.jpg')
    out.write(b'\\x00\\x00\\x00\\x00\\x00\\x00\\x10\\x00partial')      # dies mid-file
    out.flush()
    sys.exit(3)
for i, t in enumerate(job['times']):
    D.write_str(out, D.OUTPUT_FILE)
    D.write_str(out, '.jpg' if len(job['times']) == 1 else '_%%d.jpg' %% i)
    body = json.dumps(dict(name=job['name'], t=t, device=sys.argv[3],
                           spp=job['profile'].get('spp'))).encode()
    D.write_filelike(out, io.BytesIO(body))
D.write_str(out, D.CLOSING_ENCODER)
D.write_str(out, D.DONE)
''' % (sys.executable, ROOT)


@pytest.fixture
def stub(tmp_path, monkeypatch):
    path = tmp_path / 'stubworker'
    path.write_text(STUB)
    path.chmod(path.stat().st_mode | stat.S_IXUSR)
    monkeypatch.setenv('CUBURN_WORKER_COMMAND', str(path))
    return path


def jobs(tmp_path, n):
    return [D.Job({'type': 'animation'}, str(tmp_path / ('flame_%05d' % i)), [i / 100.0], 0)
            for i in range(1, n + 1)]


def test_dispatch_spreads_jobs_over_workers(stub, tmp_path):
    d = D.Dispatcher({'spp': 77}, ['localhost/0', 'localhost/1', 'remote/0'], log=io.StringIO())
    assert d.run(iter(jobs(tmp_path, 12)))
    seen = set()
    for i in range(1, 13):
        body = json.loads((tmp_path / ('flame_%05d.jpg' % i)).read_text())
        assert body['t'] == i / 100.0 and body['spp'] == 77
        seen.add(body['device'])
    assert not list(tmp_path.glob('*.tmp'))
    assert seen <= {'0', '1'} and d.failures == {'localhost/0': 0, 'localhost/1': 0, 'remote/0': 0}


def test_failed_job_is_retried_and_leaves_no_partial_file(stub, tmp_path, monkeypatch):
    monkeypatch.setenv('STUB_FAIL_FLAG', str(tmp_path / 'failed-once'))
    log = io.StringIO()
    d = D.Dispatcher({}, ['localhost/0'], log=log)
    assert d.run(iter(jobs(tmp_path, 3)))
    assert (tmp_path / 'failed-once').exists() and 'incomplete read' in log.getvalue()
    assert sorted(p.name for p in tmp_path.glob('flame_*.jpg')) == [
        'flame_00001.jpg', 'flame_00002.jpg', 'flame_00003.jpg']
    assert all(json.loads(p.read_text())['name'].endswith(p.stem) for p in tmp_path.glob('*.jpg'))


def test_hopeless_worker_is_dropped_and_the_run_reports_failure(stub, tmp_path, monkeypatch):
    monkeypatch.setenv('STUB_ALWAYS_FAIL', '1')
    d = D.Dispatcher({}, ['localhost/0'], log=io.StringIO())
    assert d.run(iter(jobs(tmp_path, 1))) is False
    assert d.failures['localhost/0'] == D.Dispatcher.MAX_WORKER_FAILURES
    assert d.failed == [str(tmp_path / 'flame_00001')]
    assert not list(tmp_path.glob('flame_*.jpg'))


def test_multi_frame_job_returns_one_file_per_message(stub, tmp_path):
    job = D.Job({}, str(tmp_path / 'shard_00001'), [0.1, 0.2, 0.3], 0)
    assert D.Dispatcher({}, ['localhost/3'], log=io.StringIO()).run(iter([job]))
    assert sorted(p.name for p in tmp_path.glob('shard_*')) == [
        'shard_00001_0.jpg', 'shard_00001_1.jpg', 'shard_00001_2.jpg']


def test_worker_command_forms(monkeypatch):
    monkeypatch.delenv('CUBURN_WORKER_COMMAND', raising=False)
    assert D.worker_command('localhost/2')[-4:] == [os.path.join(ROOT, 'distribute.py'), 'work',
                                                    '--device', '2']
    assert D.worker_command('render7/1') == ['ssh', 'render7', '.cuburn_dist/distribute.py',
                                             'work', '--device', '1']
