"""
Schema node types for genome / profile documents.

Mirrors the type vocabulary of the reference (cuburn/genome/spectypes.py:279-326)
so that ``specs`` reads the same and third-party tools that walk the schema
keep working.  A *spec* is a nested dict whose leaves are instances of the
tuples below.
"""
from collections import namedtuple

Map = namedtuple('Map', 'type doc')
List = namedtuple('List', 'type default doc')
Spline = namedtuple('Spline', 'default min max interp period doc var')
Scalar = namedtuple('Scalar', 'default doc')
RefScalar = namedtuple('RefScalar', 'default ref doc')
String = namedtuple('String', 'doc')
Enum = namedtuple('Enum', 'choices default doc')
Palette = namedtuple('Palette', '')


def map_(type, d=None):
    return Map(type, d)


def list_(type, default=(), d=None):
    return List(type, default, d)


def scalar(default, d=None):
    return Scalar(default, d)


def refscalar(default, ref, d=None):
    return RefScalar(default, ref, d)


def spline(default=0, min=None, max=None, interp='linear', period=None, d=None):
    """A plain (linear-domain) animated parameter."""
    return Spline(default, min, max, interp, period, d, False)


def scalespline(default=1, min=0, max=None, d=None):
    """An animated scale factor: interpolated in the magnitude domain."""
    return Spline(default, min, None, 'mag', None, d, False)


def enum(choices, default=None, d=None):
    if isinstance(choices, str):
        choices = choices.split()
    return Enum(list(choices), default, d)


class XYPair(dict):
    """Two splines of the same type under the keys ``x`` and ``y``."""
    def __init__(self, type):
        super().__init__(x=type, y=type)
        self.type = type


def export_spec(spec):
    """JSON-serialisable view of a spec tree."""
    if isinstance(spec, dict):
        return {k: export_spec(v) for k, v in spec.items()}
    if isinstance(spec, str):
        return spec
    out = spec._asdict()
    out['type'] = type(spec).__name__
    return out
