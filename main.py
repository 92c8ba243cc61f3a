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
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cuburn_b200 import profile  # noqa: E402


def load_anim(args):
    """``(animation dict, basename)`` through the genome DB (main.py:29-30)."""
    if args.flame.startswith('sample:'):
        from cuburn_b200 import samples
        name = args.flame.split(':', 1)[1]
        return samples.GENOMES[name](), name
    from cuburn_b200.genome import db
    gdb = db.connect(args.genomedb)
    return gdb.get_anim(args.flame, args.half)


def main(args, prof):
    gnm, basename = load_anim(args)
    if getattr(args, 'print'):
        from cuburn_b200.genome.util import json_encode
        sys.stdout.write(json_encode(gnm))
        return
    gprof = profile.wrap(prof, gnm)
    frames = profile.enumerate_jobs(gprof, basename, args)
    if not frames:
        return

    from cuburn_b200 import _native, render
    _native.init(args.device or 0)
    rmgr = render.RenderManager()
    rdr = render.Renderer(gnm, gprof, keep=args.keep)
    last_ms = 0

    for name, times in frames:
        def save(buf):
            out, log = rdr.out.encode(buf)
            for suffix, file_like in out.items():
                with open(name + suffix, 'wb') as fp:
                    fp.write(file_like.read())
                if getattr(file_like, 'close', None):
                    file_like.close()
            for key, val in log:
                print('\n=== %s ===\n%s' % (key, val), file=sys.stderr)

        evt = buf = next_evt = next_buf = None
        for idx, t in enumerate(list(times) + [None]):
            evt, buf = next_evt, next_buf
            if t is not None:
                next_evt, next_buf = rmgr.queue_frame(rdr, gnm, gprof, t)
            if not evt:
                continue
            if last_ms > 2000:
                while not evt.query():
                    time.sleep(0.2)
            else:
                evt.synchronize()
            last_ms = evt.time()
            save(buf)
            if args.rawfn:
                try:
                    buf.tofile(args.rawfn + '.tmp')
                    os.rename(args.rawfn + '.tmp', args.rawfn)
                except Exception:
                    import traceback
                    print('Failed to write %s: %s' % (args.rawfn, traceback.format_exc()),
                          file=sys.stderr)
            print('%s%s (%3d/%3d), %dms' % (
                ('%d: ' % args.device) if args.device is not None and args.device >= 0 else '',
                name, idx, len(times), last_ms), file=sys.stderr)
        save(None)


def list_devices():
    from cuburn_b200 import _native
    for i in range(_native.device_count()):
        d = _native.device_info(i)
        print('Device %d (%s): compute %d.%d, %d SMs, mem %d, L2 %d' % (
            i, d['name'], d['cc'][0], d['cc'][1], d['sm_count'], d['total_mem'], d['l2_bytes']))


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Render fractal flames.')
    parser.add_argument('flame', metavar='ID', type=str, nargs='?',
                        help='Filename of the genome to render (or sample:NAME)')
    parser.add_argument('-d', '--genomedb', metavar='PATH', type=str, default='.',
                        help="Path to genome database (file or directory, default '.')")
    parser.add_argument('--raw', metavar='PATH', type=str, dest='rawfn',
                        help='Target file for raw buffer, to enable previews.')
    parser.add_argument('--half', action='store_true',
                        help='Use half-loops when converting nodes to animations')
    parser.add_argument('--print', action='store_true',
                        help='Print the animation and exit.')
    parser.add_argument('--list-devices', action='store_true', help='List devices and exit.')
    parser.add_argument('--device', metavar='NUM', type=int, help='GPU device number to use.')
    parser.add_argument('--keep', action='store_true',
                        help='Keep the generated source and cubin in $TMPDIR')
    profile.add_args(parser)
    args = parser.parse_args()
    if args.list_devices:
        list_devices()
    else:
        if not args.flame:
            parser.error('a flame is required')
        pname, prof = profile.get_from_args(args)
        main(args, prof)
