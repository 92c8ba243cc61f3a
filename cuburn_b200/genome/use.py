"""
Schema-driven accessors over plain-dict genomes and profiles.

API-compatible with the reference (cuburn/genome/use.py): ``Wrapper``,
``RefWrapper``, ``SplineWrapper`` and ``SplineEval`` keep their names,
constructor arguments and container behaviour (sorted keys, defaults for
missing leaves, KeyError for names outside the schema).
"""
from bisect import bisect_left

import numpy as np

from .spectypes import Enum, Spline, Scalar, RefScalar, Map, List
from .specs import toplevels


class Wrapper(object):
    """
    A lazy view of ``val`` shaped by ``spec``.  Attribute / item access
    descends one level and re-wraps according to the spec type found there.
    Extra keyword arguments are handed to every wrapper created on the way
    down (use.py:6-24).
    """
    def __init__(self, val, spec=None, path=(), **params):
        if spec is None:
            assert val.get('type') in toplevels, 'Unrecognized dict type'
            spec = toplevels[val['type']]
        self._val, self.spec, self.path, self._params = val, spec, path, params

    # -- dispatch: spec node type -> wrap_<kind> --------------------------------------
    # (the hook names are the extension interface: subclasses override wrap_spline,
    # wrap_refscalar, ... as in the reference, use.py:26-60)
    _KINDS = ((Enum, 'wrap_enum'), (Spline, 'wrap_spline'), (Scalar, 'wrap_scalar'),
              (RefScalar, 'wrap_refscalar'), (dict, 'wrap_dict'), (Map, 'wrap_Map'),
              (List, 'wrap_List'))
    _hook_of_type = {}          # exact spec type -> hook name, filled on first sight

    def wrap(self, name, spec, val):
        hook = self._hook_of_type.get(type(spec))
        if hook is None:
            hook = 'wrap_default'
            for kind, method in self._KINDS:
                if isinstance(spec, kind):
                    hook = method
                    break
            self._hook_of_type[type(spec)] = hook
        return getattr(self, hook)(self.path + (name,), spec, val)

    def wrap_default(self, path, spec, val):
        return val

    def wrap_spline(self, path, spec, val):
        return val

    def wrap_enum(self, path, spec, val):
        return val or spec.default

    def wrap_scalar(self, path, spec, val):
        return spec.default if val is None else val

    wrap_refscalar = wrap_scalar

    def wrap_dict(self, path, spec, val):
        return type(self)(val or {}, spec, path, **self._params)

    wrap_Map = wrap_dict

    def wrap_List(self, path, spec, val):
        items = spec.default if val is None else val
        return [self.wrap(path, spec.type, v) for v in items]

    def get_spec(self, name):
        if isinstance(self.spec, Map):
            return self.spec.type
        return self.spec[name]

    @classmethod
    def visit(cls, obj):
        """Deep-copy a wrapped tree back into plain containers."""
        if isinstance(obj, (Wrapper, dict)):
            return dict((k, cls.visit(obj[k])) for k in obj)
        if isinstance(obj, list):
            return [cls.visit(o) for o in obj]
        return obj

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return self.wrap(name, self.get_spec(name), self._val.get(name))

    # -- container protocol (only keys present on the underlying dict) --------
    def keys(self):
        return sorted(self._val.keys())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __contains__(self, name):
        self.get_spec(name)     # KeyError when the schema has no such field
        return name in self._val

    def __iter__(self):
        return iter(sorted(self._val))

    def __getitem__(self, name):
        return getattr(self, str(name))


class RefWrapper(Wrapper):
    """
    Profile view: a RefScalar leaf evaluates to ``profile value x genome
    spline`` (use.py:100-110).  Needs ``other=<SplineWrapper of the genome>``.
    """
    def wrap_refscalar(self, path, spec, val):
        spev = self._params['other']
        for part in spec.ref.split('.'):
            spev = spev[part]
        spev *= val if val is not None else spec.default
        return spev


class SplineWrapper(Wrapper):
    """Genome view whose spline leaves are callable ``SplineEval`` objects."""
    def wrap_spline(self, path, spec, val):
        return SplineEval(val if val is not None else spec.default,
                          self._params['scale'], spec.interp)


class SplineEval(object):
    """
    Host-side (float64) evaluation of one animated parameter.

    ``normalize`` turns any of the three JSON spellings of a spline into a
    ``(2, nknots)`` array of sorted times and positions, adding the guard knots
    at t=-2 / t=3 that make the end tangents equal the stored velocities
    (use.py:129-158).  ``__call__`` is the plain Catmull-Rom form and, like
    the reference, ignores the 'mag' domain (use.py:174-185).
    """
    def __init__(self, knots, scale, interp='linear'):
        self.knots, self.interp = self.normalize(knots, scale), interp

    @staticmethod
    def normalize(knots, scale):
        if isinstance(knots, (int, float, np.number)):
            # a constant: knots at 0 and 1, guard knots at -2 and 3, zero velocity
            k = float(knots)
            return np.array([[-2.0, 0.0, 1.0, 3.0], [k, k, k, k]])
        if len(knots) % 2 != 0:
            raise ValueError('List with odd number of elements given')
        if len(knots) == 2:
            v0 = v1 = 0.0
            pts = [(0.0, knots[0]), (1.0, knots[1])]
        else:
            p0, v0, p1, v1 = knots[:4]
            pts = [(0.0, p0), (1.0, p1)]
            pts += list(zip(knots[4::2], knots[5::2]))
        v0 *= scale
        v1 *= scale
        pts.sort()

        # guard knots 2 time units outside [0, 1], placed so that the end tangents of the
        # Catmull-Rom segments equal the stored velocities
        lead = 2.0
        if pts[0][0] >= 0:
            t1, p1_ = pts[1]
            pts.insert(0, (-lead, p1_ - (t1 + lead) * v0))
        if pts[-1][0] <= 1:
            t2, p2_ = pts[-2]
            pts.append((1 + lead, p2_ + (1 + lead - t2) * v1))
        return np.array(pts, dtype=np.float64).T.copy()

    def find_knots(self, itime):
        t_all = self.knots[0]
        idx = int(np.searchsorted(t_all, itime)) - 2
        idx = max(0, min(idx, len(t_all) - 4))
        times = t_all[idx:idx + 4]
        vals = self.knots[1][idx:idx + 4]
        t = itime - times[1]
        times = times - times[1]
        scale = 1.0 / times[2]
        return times * scale, vals, t * scale, scale

    # Hermite basis as polynomials in t (highest power first) and their derivatives;
    # evaluated with Horner's rule in plain floats (this runs ~50 times per frame on the
    # host: filter parameters, spp, frame width)
    _BASIS = [[list(np.poly1d(c).deriv(k).coeffs) if k else list(c) for c in
               ([1., -2, 1, 0], [2., -3, 0, 1], [1., -1, 0, 0], [-2., 3, 0, 0])]
              for k in range(4)]

    def __call__(self, itime, deriv=0):
        # find_knots in plain floats (the same IEEE operations in the same order: results
        # are bit-identical to the array form, at a fifth of the cost for 4-knot segments)
        itime = float(itime)
        ts, vs = self.knots[0].tolist(), self.knots[1].tolist()
        idx = max(0, min(bisect_left(ts, itime) - 2, len(ts) - 4))
        t0, t1, t2, t3 = ts[idx:idx + 4]
        v0, v1, v2, v3 = vs[idx:idx + 4]
        scale = 1.0 / (t2 - t1)
        t = (itime - t1) * scale
        m1 = (v2 - v0) / (1.0 - (t0 - t1) * scale)
        m2 = (v3 - v1) / ((t3 - t1) * scale)
        mult = scale ** deriv if deriv else 1.0
        total = 0.0
        for coef, poly in zip((m1, v1, m2, v2), self._BASIS[deriv]):
            acc = 0.0
            for c in poly:
                acc = acc * t + (c * mult if deriv else c)
            total += coef * acc
        return total

    def __imul__(self, other):
        self.knots[1] *= other
        return self
