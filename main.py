#!/usr/bin/env python
"""
Render fractal flames on a B200 -- the reference's command line
(main.py:110-137 + cuburn/profile.py:17-74) on top of cuburn_b200.

    python main.py FLAME.json -P 1080p --still -o out/
    python main.py sample:G6F -P 1080p --still          (built-in sample genomes)

Accepts what the reference accepts: flam3 XML (.flam3 / .flame), cuburn JSON nodes,
edges and animations, by file name or by ID inside a genome DB (-d).
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cuburn_b200 import profile  # noqa: E402

# A frame that took longer than this is waited for by polling (the host thread sleeps
# between polls instead of blocking inside the driver), as the reference does
# (main.py:67-71).
POLL_ABOVE_MS = 2000
POLL_INTERVAL_S = 0.2


def load_anim(args):
    """``(animation dict, basename)`` through the genome DB (main.py:29-30)."""
    if args.flame.startswith('sample:'):
        from cuburn_b200 import samples
        name = args.flame.split(':', 1)[1]
        return samples.GENOMES[name](), name
    from cuburn_b200.genome import db
    return db.connect(args.genomedb).get_anim(args.flame, args.half)


def write_media(rdr, basename, frame):
    """Encode one finished frame (``None``: flush the encoder) and write every file the
    output module hands back as ``basename + suffix``; encoder logs go to stderr."""
    media, logs = rdr.out.encode(frame)
    for suffix, filelike in media.items():
        with open(basename + suffix, 'wb') as fp:
            fp.write(filelike.read())
        close = getattr(filelike, 'close', None)
        if close:
            close()
    for key, text in logs:
        print('\n=== %s ===\n%s' % (key, text), file=sys.stderr)


def write_preview(path, frame):
    """--raw: the newest frame's raw buffer, replaced atomically (a viewer polls it)."""
    try:
        frame.tofile(path + '.tmp')
        os.rename(path + '.tmp', path)
    except Exception:
        import traceback
        print('Failed to write %s: %s' % (path, traceback.format_exc()), file=sys.stderr)


class Waiter(object):
    """How the host waits for a frame: block on its event while frames are quick, poll
    once the previous frame took more than POLL_ABOVE_MS."""
    def __init__(self):
        self.last_ms = 0

    def __call__(self, evt):
        if self.last_ms > POLL_ABOVE_MS:
            while not evt.query():
                time.sleep(POLL_INTERVAL_S)
        else:
            evt.synchronize()
        self.last_ms = evt.time()


def main(args, prof):
    gnm, basename = load_anim(args)
    if getattr(args, 'print'):
        from cuburn_b200.genome.util import json_encode
        sys.stdout.write(json_encode(gnm))
        return
    gprof = profile.wrap(prof, gnm)
    jobs = profile.enumerate_jobs(gprof, basename, args)
    if not jobs:
        return

    from cuburn_b200 import _native, render
    _native.init(args.device or 0)
    rmgr = render.RenderManager()
    rdr = render.Renderer(gnm, gprof, keep=args.keep)
    tag = '%d: ' % args.device if args.device is not None and args.device >= 0 else ''
    wait = Waiter()

    for name, times in jobs:
        for idx, evt, frame in render.frame_pipeline(rmgr, rdr, gnm, gprof, times, wait):
            write_media(rdr, name, frame)
            if args.rawfn:
                write_preview(args.rawfn, frame)
            print('%s%s (%3d/%3d), %dms' % (tag, name, idx, len(times), wait.last_ms),
                  file=sys.stderr)
        write_media(rdr, name, None)


def list_devices():
    from cuburn_b200 import _native
    for i in range(_native.device_count()):
        d = _native.device_info(i)
        print('Device %d (%s): compute %d.%d, %d SMs, mem %d, L2 %d' % (
            i, d['name'], d['cc'][0], d['cc'][1], d['sm_count'], d['total_mem'], d['l2_bytes']))


def build_parser():
    """The reference's options (main.py:110-130) plus the profile options."""
    parser = argparse.ArgumentParser(description='Render fractal flames.')
    opt = parser.add_argument
    opt('flame', metavar='ID', type=str, nargs='?',
        help='Filename of the genome to render (or sample:NAME)')
    opt('-d', '--genomedb', metavar='PATH', type=str, default='.',
        help="Path to genome database (file or directory, default '.')")
    opt('--raw', metavar='PATH', type=str, dest='rawfn',
        help='Target file for raw buffer, to enable previews.')
    opt('--half', action='store_true',
        help='Use half-loops when converting nodes to animations')
    opt('--print', action='store_true', help='Print the animation and exit.')
    opt('--list-devices', action='store_true', help='List devices and exit.')
    opt('--device', metavar='NUM', type=int, help='GPU device number to use.')
    opt('--keep', action='store_true', help='Keep the generated source and cubin in $TMPDIR')
    profile.add_args(parser)
    return parser


def run(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.list_devices:
        return list_devices()
    if not args.flame:
        parser.error('a flame is required')
    pname, prof = profile.get_from_args(args)
    main(args, prof)


if __name__ == '__main__':
    run()
