"""
Render profiles: built-ins, CLI flags, genome-adjusted wrapping and the
enumeration of frame centre times / output jobs.

Drop-in for the reference module (cuburn/profile.py): same function names,
flags, defaults and return shapes.
"""
import os
import json
import argparse

import numpy as np

from .genome.specs import toplevels
from .genome.use import RefWrapper, SplineWrapper

BUILTIN = {
    '1080p': dict(width=1920, height=1080),
    '720p': dict(width=1280, height=720),
    '540p': dict(width=960, height=540),
    'preview': dict(width=640, height=360, spp=1200, skip=1),
}

_OVERRIDABLE = ('duration', 'fps', 'frame_width', 'start', 'end', 'skip',
                'shard', 'spp', 'width', 'height')


# (group title, [(flags, argparse keywords)]) -- the reference's options
# (profile.py:17-74), kept as data
_OPTION_GROUPS = (
    ('Profile options', [
        (('-P', '--builtin-profile'), dict(
            choices=list(BUILTIN.keys()), default='720p',
            help='Set parameters below from a builtin profile. (default: 720p)')),
        (('-p', '--profile'), dict(type=argparse.FileType(), metavar='PROFILE',
                                   help='Set profile from a JSON file.')),
    ]),
    ('Temporal options', [
        (('--duration',), dict(type=float, metavar='TIME',
                               help='Override base duration in seconds')),
        (('--fps',), dict(type=float, dest='fps', help='Override frames per second')),
        (('--start',), dict(metavar='FRAME_NO', type=int,
                            help='First frame to render (1-indexed, inclusive)')),
        (('--end',), dict(metavar='FRAME_NO', type=int,
                          help='Last frame to render (1-indexed, exclusive, negative from end)')),
        (('--skip',), dict(dest='skip', metavar='N', type=int,
                           help='Skip N frames between each rendered frame')),
        (('--shard',), dict(
            dest='shard', metavar='SECS', type=float,
            help="Write SECS of output into each file, instead of one frame per file. "
                 "If set, causes 'start', 'end', and 'skip' to be ignored.")),
        (('--frame_width',), dict(metavar='SCALE', type=float,
                                  help='Adjustment factor for temporal frame width.')),
        (('--still',), dict(action='store_true',
                            help='Override start, end, and temporal frame width to render '
                                 'one frame without motion blur.')),
    ]),
    ('Spatial options', [
        (('--spp',), dict(type=int, metavar='SPP', help='Set base samples per pixel')),
        (('--width',), dict(type=int, metavar='PX')),
        (('--height',), dict(type=int, metavar='PX')),
    ]),
    ('Output options', [
        (('--codec',), dict(choices=['jpeg', 'png', 'tiff', 'x264', 'vp8', 'vp9', 'prores',
                                     'raw'])),
        (('-n',), dict(metavar='NAME', type=str, dest='name',
                       help='Prefix to use when saving files (default is basename of input)')),
        (('--suffix',), dict(metavar='NAME', type=str, dest='suffix', default='',
                             help="Suffix to use when saving files (default '')")),
        (('-o',), dict(metavar='DIR', type=str, dest='dir', default='.',
                       help='Output directory')),
        (('--resume',), dict(action='store_true', dest='resume',
                             help="Don't overwrite output files that are newer than the input")),
        (('--subdir',), dict(action='store_true',
                             help='Use basename as subdirectory of out dir, instead of prefix')),
    ]),
)


def add_args(parser=None):
    """Add the profile option groups to ``parser`` (a new one if None)."""
    parser = argparse.ArgumentParser() if parser is None else parser
    for title, options in _OPTION_GROUPS:
        group = parser.add_argument_group(title)
        for flags, kw in options:
            group.add_argument(*flags, **kw)
    return parser


def get_from_args(args):
    """argparse result -> ``(name, profile dict)`` (profile.py:76-95)."""
    if args.profile:
        name = os.path.basename(args.profile.name).rsplit('.', 1)[0]
        base = json.load(args.profile)
    else:
        name = args.builtin_profile
        base = dict(BUILTIN[args.builtin_profile])

    if args.still:
        base.update(frame_width=0, start=1, end=2)
    for arg in _OVERRIDABLE:
        if getattr(args, arg, None) is not None:
            base[arg] = getattr(args, arg)
    if args.codec is not None:
        base.setdefault('output', {})['type'] = args.codec
    return name, base


def wrap(prof, gnm):
    """Profile view whose RefScalars are multiplied into the genome's splines."""
    scale = gnm.get('time', {}).get('duration', 1)
    return RefWrapper(prof, toplevels['profile'],
                      other=SplineWrapper(gnm, scale=scale))


def enumerate_times(gprof):
    """
    ``[(frame_no, [centre times])]``; numbering is assigned before start / end /
    skip are applied, so frame numbers may be non-contiguous (profile.py:107-127).
    """
    nframes = int(round(gprof.fps * gprof.duration))
    edges = np.linspace(0, 1, nframes + 1)
    centres = edges[:-1] + 0.5 * (edges[1] - edges[0])
    if gprof.shard:
        per = max(1, int(round(gprof.fps * gprof.shard)))
        return [(i, centres[t:t + per])
                for i, t in enumerate(range(0, len(centres), per), 1)]
    frames = list(enumerate([[t] for t in centres], 1))
    if gprof.end is not None:
        frames = frames[:gprof.end]
    if gprof.start is not None:
        frames = frames[gprof.start:]
    return frames[::gprof.skip + 1]


def enumerate_jobs(gprof, basename, args, resume=None):
    """``[(output path without extension, [centre times])]`` (profile.py:129-159)."""
    from . import output      # deferred: output imports the native library
    if args.name is not None:
        basename = args.name
    prefix = os.path.join(args.dir, basename)
    if args.subdir:
        if not os.path.isdir(prefix):
            os.mkdir(prefix)
        lead = prefix + '/'
    else:
        lead = prefix + '_'

    jobs = [('%s%05d%s' % (lead, i, args.suffix), t)
            for i, t in enumerate_times(gprof)]
    resume = args.resume if resume is None else resume
    if resume:
        ext = output.get_suffix_for_profile(gprof)
        jobs = [(n, t) for n, t in jobs if not os.path.isfile(n + ext)]
    return jobs
