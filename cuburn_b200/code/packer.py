"""
Genome packing: which knot rows go to the device, which parameter slots the
iterate kernel reads, and the precalc program that connects them.

Plays the role of the reference ``GenomePacker`` (cuburn/code/interp.py:125-282)
with a different mechanism.  The reference discovers the required splines as a
side effect of rendering code templates and emits a C struct plus a generated
interpolation kernel.  Here the genome *structure* (which xforms, variations,
post-affines and final xform exist) is walked once and turned into data:

  rows     unique spline paths -> rows of the ``times``/``knots`` [nrows][32]
           upload (pack(), same content as interp.py:207-232)
  slots    named floats of one temporal sample's parameter block
  program  int32 [nops][12] precalc ops interpreted by the static kernel
           ``cb_interp_params`` (csrc/cb_interp.cu)

Slot and row names are dotted genome paths, so a CPU restatement can be compared
by name without knowing the layout.
"""
import numpy as np

from ..genome import specs
from ..genome.use import SplineEval
from ..genome.util import resolve_spec
from ..genome.variations import var_param_order
from . import varlib
from .varlib import (OP_DIRECT, OP_DIRECT_MAG, OP_AFFINE, OP_CAMERA, OP_DENSITY, OP_XAOS)

SEARCH_ROUNDS = 5
MAX_KNOTS = 1 << SEARCH_ROUNDS
PROG_WIDTH = 12
AFFINE_COEFS = ('xx', 'xy', 'xo', 'yx', 'yy', 'yo')
_AFFINE_INPUTS = (('angle',), ('spread',), ('magnitude', 'x'), ('magnitude', 'y'),
                  ('offset', 'x'), ('offset', 'y'))


class GenomePacker(object):
    def __init__(self, gnm, spec=None):
        self.spec = spec or specs.anim
        self.row_paths = []          # tuple paths
        self.row_mag = []            # 1 for magnitude-domain rows
        self._row_index = {}
        self.slot_names = []         # dotted names
        self._slot_index = {}
        self.program = []            # lists of PROG_WIDTH ints

        # xform order = string sort of the keys; the last one takes the
        # remainder of the probability mass (iter.py:232,263-272)
        self.xform_ids = sorted(str(k) for k in gnm.get('xforms', {}).keys())
        if not self.xform_ids:
            raise ValueError('genome has no xforms')
        self.has_final = 'final_xform' in gnm
        # (xform path, variation names in sorted order, has post affine)
        self.xforms = []
        # xform path -> slot of its opacity, for xforms that carry one (specs.py:17)
        self.opacity = {}
        # xaos: the choice of the next xform depends on the previous one
        # (precalc_chaos, iter.py:32-54); used when any xform has a 'chaos' map
        self.xaos = any('chaos' in gnm['xforms'][xid] for xid in self.xform_ids)

        # weights first so the density op sees one contiguous block of rows
        wrows = [self._row(('xforms', xid, 'weight')) for xid in self.xform_ids]
        assert wrows == list(range(len(wrows)))

        # Every block of slots that is read together (one xform, the choice densities, the
        # camera) starts on a 16-byte boundary: with the parameters in shared memory
        # (motion blur) the kernel fetches them as aligned float4s, and a block that
        # straddles a boundary costs an extra shared-memory wavefront per warp and round.
        for xid in self.xform_ids:
            self._align()
            self._add_xform(('xforms', xid), gnm['xforms'][xid])
        if self.has_final:
            self._align()
            self._add_xform(('final_xform',), gnm['final_xform'])
        self._align()

        if len(self.xform_ids) > 1 and not self.xaos:
            first = None
            for xid in self.xform_ids[:-1]:
                s = self._slot(('xforms', xid, 'density'))
                first = s if first is None else first
            self._op(OP_DENSITY, first, [wrows[0]], aux0=len(self.xform_ids))
        if len(self.xform_ids) > 1 and self.xaos:
            # per previous xform p: cumulative normalised weight[n] * chaos[p][n]
            for pid in self.xform_ids:
                crows = [self._row(('xforms', pid, 'chaos', nid)) for nid in self.xform_ids]
                assert crows == list(range(crows[0], crows[0] + len(crows)))
                first = self._slot_block(('xforms', pid, 'chaos_den'), self.xform_ids[:-1])
                self._op(OP_XAOS, first, [wrows[0], crows[0]], aux0=len(self.xform_ids))

        self._align()
        cam_rows = [self._row(('camera', 'rotation')),
                    self._row(('camera', 'center', 'x')),
                    self._row(('camera', 'center', 'y')),
                    self._row(('camera', 'scale'))]
        cam_first = self._slot_block(('camera',), AFFINE_COEFS)
        self._op(OP_CAMERA, cam_first, cam_rows)

        self.nrows = len(self.row_paths)
        self.nslots = len(self.slot_names)
        # parameter block stride: padded to 16 B so blocks stay vector-aligned
        self.param_stride = (self.nslots + 3) // 4 * 4

    # ---- structure walk ------------------------------------------------------
    def _add_xform(self, xpath, xf):
        variations = sorted(xf.get('variations', {}).keys())
        for v in variations:
            if v not in var_param_order:
                raise KeyError('unknown variation %r' % v)
        has_post = 'post_affine' in xf
        self.xforms.append((xpath, variations, has_post))

        self._affine(xpath + ('pre_affine',))
        if has_post:
            self._affine(xpath + ('post_affine',))
        self._direct(xpath + ('color',))
        self._direct(xpath + ('color_speed',))
        if 'opacity' in xf:
            self._direct(xpath + ('opacity',))
            self.opacity[xpath] = self.slot(xpath + ('opacity',))
        for v in variations:
            vpath = xpath + ('variations', v)
            self._direct(vpath + ('weight',))
            for kind, name in varlib.var_args(v):
                if kind == 'p':
                    self._direct(vpath + (name,))
            if v in varlib.PRECALC:
                op, inputs, outputs = varlib.PRECALC[v]
                rows = []
                for inp in inputs:
                    if inp.startswith('^'):
                        rows.append(self._row(xpath + tuple(inp[1:].split('.'))))
                    else:
                        rows.append(self._row(vpath + (inp,)))
                first = self._slot_block(vpath, outputs)
                self._op(op, first, rows)

    def _affine(self, apath):
        rows = [self._row(apath + sub) for sub in _AFFINE_INPUTS]
        first = self._slot_block(apath, AFFINE_COEFS)
        self._op(OP_AFFINE, first, rows)

    def _direct(self, path):
        row = self._row(path)
        slot = self._slot(path)
        self._op(OP_DIRECT_MAG if self.row_mag[row] else OP_DIRECT, slot, [row])

    # ---- tables ----------------------------------------------------------------
    def _row(self, path):
        if path not in self._row_index:
            sp = resolve_spec(self.spec, path)
            self._row_index[path] = len(self.row_paths)
            self.row_paths.append(path)
            self.row_mag.append(1 if sp.interp == 'mag' else 0)
        return self._row_index[path]

    def _slot(self, path):
        name = '.'.join(path)
        if name in self._slot_index:
            raise AssertionError('slot %s allocated twice' % name)
        self._slot_index[name] = len(self.slot_names)
        self.slot_names.append(name)
        return self._slot_index[name]

    def _align(self, n=4):
        """Pad with unnamed slots up to a multiple of ``n``."""
        while len(self.slot_names) % n:
            self.slot_names.append('_pad.%d' % len(self.slot_names))

    def named_slots(self):
        """[(slot index, dotted name)] of the slots that carry a parameter."""
        return [(i, n) for i, n in enumerate(self.slot_names) if not n.startswith('_pad.')]

    def _slot_block(self, path, names):
        first = None
        for n in names:
            s = self._slot(path + (n,))
            first = s if first is None else first
        return first

    def _op(self, op, out, rows, aux0=0, aux1=0):
        assert len(rows) <= 8
        word = [op, out] + list(rows) + [0] * (8 - len(rows)) + [aux0, aux1]
        self.program.append(word)

    def slot(self, *path):
        """Index of a parameter slot by path, e.g. slot('camera', 'xx')."""
        if len(path) == 1 and isinstance(path[0], (tuple, list)):
            path = tuple(path[0])
        return self._slot_index['.'.join(path)]

    def __len__(self):
        return self.nslots

    # ---- data ------------------------------------------------------------------
    def program_array(self):
        """The precalc program as int32 [nops][PROG_WIDTH] (built once: the program is
        complete when the constructor returns)."""
        cached = getattr(self, '_program_array', None)
        if cached is None or len(cached) != len(self.program):
            cached = np.asarray(self.program, dtype=np.int32).reshape(-1, PROG_WIDTH)
            cached.setflags(write=False)
            self._program_array = cached
        return cached

    def _lookup(self, gnm, path):
        """The genome's value at ``path``, or the schema default where a key is missing."""
        attr = gnm
        for name in path:
            if isinstance(attr, dict) and name not in attr and name.isdigit() \
                    and int(name) in attr:
                name = int(name)            # chaos maps straight from the converter
            if not isinstance(attr, dict) or name not in attr:
                return resolve_spec(self.spec, path).default
            attr = attr[name]
        return attr

    def pack(self, gnm, times=None, knots=None):
        """
        Knot times and values for every row: two float32 [nrows][32] arrays
        (times padded with 1e9).  Missing genome keys take the schema default
        (interp.py:207-232).  ``times`` / ``knots`` may be preallocated (pinned)
        arrays to fill in place.
        """
        if times is None:
            times = np.empty((self.nrows, MAX_KNOTS), np.float32)
        if knots is None:
            knots = np.empty((self.nrows, MAX_KNOTS), np.float32)
        times.fill(1e9)
        knots.fill(0)
        scale = gnm.get('time', {}).get('duration', 1)
        const_rows, const_vals = [], []
        for idx, path in enumerate(self.row_paths):
            attr = gnm
            try:
                for name in path:           # the common case: every key is there
                    attr = attr[name]
            except (KeyError, TypeError, IndexError):
                attr = self._lookup(gnm, path)
            if type(attr) in (int, float):
                # a constant normalises to four equal knots at t = -2, 0, 1, 3; rows
                # like this are most of a genome and are written in one go below
                const_rows.append(idx)
                const_vals.append(attr)
                continue
            kn = SplineEval.normalize(attr, scale)
            n = kn.shape[1]
            if n > MAX_KNOTS:
                raise ValueError('%s has %d knots; at most %d are supported'
                                 % ('.'.join(path), n, MAX_KNOTS))
            times[idx, :n] = kn[0]
            knots[idx, :n] = kn[1]
        if const_rows:
            times[const_rows, :4] = (-2.0, 0.0, 1.0, 3.0)
            knots[const_rows, :4] = np.asarray(const_vals, np.float64)[:, None]
        return times, knots
