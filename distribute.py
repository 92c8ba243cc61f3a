#!/usr/bin/env python3
"""
Job farm: render flames on several GPUs / hosts, one worker process per job.

Command line, wire format and job semantics of the reference's ``distribute.py``
(dispatch: distribute.py:131-248, work: distribute.py:73-127, framing: :34-71), so a
dispatcher here can drive reference workers and vice versa:

  dispatch ID [ID ...] [--worker HOST/DEVICE ...] [-d GENOMEDB] [profile options]
  work --device N

* a worker is started per job (``HOST/DEVICE``: ``localhost`` spawns this file, any
  other host runs ``.cuburn_dist/distribute.py work`` over ssh), announces itself,
  reads one job (JSON: profile, genome, times, name), renders it with the software
  pipelining of ``main.py`` and streams every encoded file back;
* strings travel as a 4-byte big-endian length + bytes, files as an 8-byte length +
  bytes in 1 MiB pieces;
* a failed job is retried up to three times, a worker address is dropped after four
  consecutive failures, finished frames are skipped (``enumerate_jobs(resume=True)``)
  and files appear under their final name only when complete.

Concurrency is plain threads and a bounded queue (the reference uses gevent).
"""
import argparse
import json
import os
import queue
import struct
import subprocess
import sys
import threading
import time
import traceback
from collections import namedtuple

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

READY = 'worker ready'
CLOSING_ENCODER = 'closing encoder'
OUTPUT_FILE = 'here is a file for you'
DONE = 'we done here'
CHUNK = 1024 * 1024


# ---- framing ---------------------------------------------------------------------
def write_str(out, val):
    data = val.encode() if isinstance(val, str) else bytes(val)
    out.write(struct.pack('>I', len(data)))
    out.write(data)
    out.flush()


def read_exact(infp, n, what):
    data = infp.read(n)
    if len(data) != n:
        raise EOFError('incomplete read of %s: expected %d bytes, got %d' % (what, n, len(data)))
    return data


def read_str(infp):
    head = read_exact(infp, 4, 'string size')
    if head[0] != 0:
        raise ValueError('no string should be that big')
    return read_exact(infp, struct.unpack('>I', head)[0], 'string').decode()


def write_filelike(out, filelike):
    filelike.seek(0, 2)
    out.write(struct.pack('>Q', filelike.tell()))
    filelike.seek(0)
    while True:
        buf = filelike.read(CHUNK)
        if not buf:
            break
        out.write(buf)
    out.flush()


def copy_filelike(infp, dst):
    head = read_exact(infp, 8, 'file size')
    if head[0] != 0:
        raise ValueError('no file should be that big')
    left = struct.unpack('>Q', head)[0]
    while left:
        chunk = read_exact(infp, min(CHUNK, left), 'file chunk')
        dst.write(chunk)
        left -= len(chunk)


# ---- worker ----------------------------------------------------------------------
def work(args, stdin=None, stdout=None):
    """Render one job read from stdin; frames and logs go back over stdout."""
    stdin = stdin or sys.stdin.buffer
    stdout = stdout or sys.stdout.buffer
    import socket
    addr = '%s/%s' % (socket.gethostname().split('.')[0], args.device)
    write_str(stdout, READY)
    job_text = read_str(stdin)
    if job_text == DONE:
        return
    job = json.loads(job_text)

    from cuburn_b200 import _native, profile, render
    _native.init(args.device or 0)
    gnm, times, name = job['genome'], job['times'], job['name']
    gprof = profile.wrap(job['profile'], gnm)
    rmgr = render.RenderManager()
    rdr = render.Renderer(gnm, gprof)

    def save(buf):
        media, _logs = rdr.out.encode(buf)
        for suffix, filelike in media.items():
            write_str(stdout, OUTPUT_FILE)
            write_str(stdout, suffix)
            write_filelike(stdout, filelike)
            if getattr(filelike, 'close', None):
                filelike.close()

    for idx, evt, buf in render.frame_pipeline(rmgr, rdr, gnm, gprof, times):
        print('%30s: %s (%3d/%3d), %dms' % (addr, name, idx, len(times), evt.time()),
              file=sys.stderr, flush=True)
        save(buf)
    write_str(stdout, CLOSING_ENCODER)
    save(None)
    write_str(stdout, DONE)


# ---- dispatcher -------------------------------------------------------------------
Job = namedtuple('Job', 'genome name times retry_count')


def worker_command(addr):
    """argv that starts a worker for ``HOST/DEVICE`` (distribute.py:161-183)."""
    host, device = addr.split('/')
    override = os.environ.get('CUBURN_WORKER_COMMAND')
    if override:
        return override.split() + ['work', '--device', str(device)]
    if host == 'localhost':
        return [sys.executable, os.path.abspath(__file__), 'work', '--device', str(device)]
    return ['ssh', host, '.cuburn_dist/distribute.py', 'work', '--device', str(device)]


class Dispatcher(object):
    MAX_RETRIES = 3             # per job
    MAX_WORKER_FAILURES = 4     # consecutive, per address

    def __init__(self, prof, workers, log=sys.stderr):
        self.prof, self.workers, self.log = prof, list(workers), log
        self.jobs = queue.Queue()       # unbounded: a retry must never block a worker thread
        self.backlog = 5                # ... the job source is throttled instead
        self.failures = dict.fromkeys(self.workers, 0)
        self.lock = threading.Lock()
        self.outstanding = 0            # jobs queued or running, retries included
        self.idle = threading.Condition(self.lock)
        self.failed = []
        self.threads = []

    def connect(self, addr):
        delay = 5
        while True:
            try:
                proc = subprocess.Popen(worker_command(addr), stdin=subprocess.PIPE,
                                        stdout=subprocess.PIPE)
                if read_str(proc.stdout) != READY:
                    raise IOError('worker %s did not announce itself' % addr)
                return proc
            except Exception:
                if addr.split('/')[0] == 'localhost' or os.environ.get('CUBURN_WORKER_COMMAND'):
                    raise               # a local worker that cannot start will not get better
                traceback.print_exc(file=self.log)
                time.sleep(delay)
                delay = min(600, delay * 2)

    def submit(self, job, throttle=False):
        with self.idle:
            while throttle and self.jobs.qsize() >= self.backlog and self._alive():
                self.idle.wait(0.2)
            self.outstanding += 1
        self.jobs.put(job)

    def _alive(self):
        return any(t.is_alive() for t in self.threads)

    def run_job(self, addr, job):
        """One job on one fresh worker; True when every file arrived."""
        proc = self.connect(addr)
        try:
            desc = dict(profile=self.prof, genome=job.genome, times=list(job.times), name=job.name)
            write_str(proc.stdin, json.dumps(desc))
            proc.stdin.close()
            while True:
                msg = read_str(proc.stdout)
                if msg == OUTPUT_FILE:
                    filename = job.name + read_str(proc.stdout)
                    with open(filename + '.tmp', 'wb') as fp:
                        copy_filelike(proc.stdout, fp)
                    os.rename(filename + '.tmp', filename)
                elif msg == DONE:
                    return True
                elif msg != CLOSING_ENCODER:
                    raise IOError('no known event ' + msg)
        finally:
            if proc.stdin and not proc.stdin.closed:
                proc.stdin.close()
            proc.stdout.close()
            proc.wait()

    def serve(self, addr):
        """Worker loop for one address: take jobs until told to stop or failing."""
        while self.failures[addr] < self.MAX_WORKER_FAILURES:
            job = self.jobs.get()
            if job is None:
                return
            try:
                self.run_job(addr, job)
                self.failures[addr] = 0
            except Exception:
                print(traceback.format_exc(), file=self.log)
                self.failures[addr] += 1
                if job.retry_count < self.MAX_RETRIES:
                    self.submit(job._replace(retry_count=job.retry_count + 1))
                else:
                    self.failed.append(job.name)
            finally:
                with self.idle:
                    self.outstanding -= 1
                    self.idle.notify_all()

    def run(self, job_source):
        self.threads = [threading.Thread(target=self.serve, args=(addr,), daemon=True)
                        for addr in self.workers]
        for t in self.threads:
            t.start()
        for job in job_source:
            if not self._alive():
                break
            self.submit(job, throttle=True)
        with self.idle:
            while self.outstanding and self._alive():
                self.idle.wait(0.5)
        for _ in self.threads:
            self.jobs.put(None)
        for t in self.threads:
            t.join(timeout=5)
        return not self.failed and not self.outstanding


def read_workers(args):
    workers = args.worker
    if not workers:
        try:
            with open(os.path.expanduser('~/.cuburn-workers')) as fp:
                workers = fp.read().split()
        except IOError:
            workers = []
    return workers


def dispatch(args):
    from cuburn_b200 import profile
    from cuburn_b200.genome import db
    pname, prof = profile.get_from_args(args)
    workers = read_workers(args)
    if not workers:
        sys.exit('No workers defined. Pass --worker or set up ~/.cuburn-workers with one '
                 'worker per line.')
    gdb = db.connect(args.genomedb)

    def jobs():
        for oid in args.flames:
            ids = [oid]
            if oid.startswith('@'):
                with open(oid[1:]) as fp:
                    ids = fp.read().split()
            for gid in ids:
                gnm, basename = gdb.get_anim(gid)
                gprof = profile.wrap(prof, gnm)
                for name, times in profile.enumerate_jobs(gprof, basename, args, resume=True):
                    yield Job(gnm, name, [float(t) for t in times], 0)

    ok = Dispatcher(prof, workers).run(jobs())
    sys.exit(0 if ok else 1)


def make_parser():
    from cuburn_b200 import profile
    parser = argparse.ArgumentParser(description='Render fractal flames on multiple GPUs.')
    commands = parser.add_subparsers(dest='command', required=True)
    d = commands.add_parser('dispatch', help='Dispatch tasks to workers.')
    d.add_argument('flames', metavar='ID', type=str, nargs='+',
                   help='Flames to render (prefix playlist with @)')
    d.add_argument('--worker', metavar='ADDRESS', nargs='*',
                   help='Worker address (in the form "host/device_id")')
    d.add_argument('-d', '--genomedb', metavar='PATH', type=str, default='.',
                   help="Path to genome database (file or directory, default '.')")
    profile.add_args(d)
    d.set_defaults(func=dispatch)
    w = commands.add_parser('work', help='Perform a task (controlled by a dispatcher).')
    w.add_argument('--device', metavar='NUM', type=int, help='GPU device number to use, 0-indexed.')
    w.set_defaults(func=work)
    return parser


if __name__ == '__main__':
    ns = make_parser().parse_args()
    ns.func(ns)
