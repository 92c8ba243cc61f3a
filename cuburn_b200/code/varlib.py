"""
Call signatures of the device variation library (csrc/device/variations.cuh).

For every variation: whether it draws random numbers, and where each extra
argument of ``var_<name>(tx, ty, w, ox, oy[, rng], args...)`` comes from:

  ('p',  name)  the variation's own animated parameter `name`
  ('pre', c)    coefficient `c` of the owning xform's pre-affine (xx xy xo yx yy yo)
  ('pc', name)  a value precalculated once per temporal sample (see PRECALC)

PRECALC[name] = (op, [input parameter paths relative to the xform or variation],
[outputs]).  The ops are evaluated on the device by the interpolation kernel
(csrc/kernels_interp.cu) right after the splines, as in the reference's precalc
hunks (cuburn/code/variations.py:135-140,267-273,292-294,307-309,630-634).
"""
from ..genome.variations import var_param_order

# variations that consume random numbers ("RNG draws")
RNG_USERS = frozenset("""julia noise julian juliascope blur gaussian_blur
    radial_blur pie arch square rays blade super_shape flower conic parabola
    boarders cpow pre_blur""".split())

# precalc op codes shared with the device interpreter
OP_DIRECT, OP_DIRECT_MAG, OP_AFFINE, OP_CAMERA, OP_DENSITY = 0, 1, 2, 3, 4
OP_WAVES, OP_PERSPECTIVE, OP_JULIAN_CN, OP_CURVE, OP_XAOS = 5, 6, 7, 8, 9

# name -> (op, inputs, outputs); inputs starting with '^' are relative to the
# xform (not the variation).
PRECALC = {
    'waves': (OP_WAVES, ['^pre_affine.offset.x', '^pre_affine.offset.y'],
              ['dx2', 'dy2']),
    'perspective': (OP_PERSPECTIVE, ['angle', 'dist'], ['mdist', 'sin', 'cos']),
    'julian': (OP_JULIAN_CN, ['dist', 'power'], ['cn']),
    'juliascope': (OP_JULIAN_CN, ['dist', 'power'], ['cn']),
    'curve': (OP_CURVE, ['xlength', 'ylength'], ['x2', 'y2']),
}

_SPECIAL_ARGS = {
    'waves': [('pre', 'xy'), ('pre', 'yy'), ('pc', 'dx2'), ('pc', 'dy2')],
    'popcorn': [('pre', 'xo'), ('pre', 'yo')],
    'rings': [('pre', 'xo')],
    'fan': [('pre', 'xo'), ('pre', 'yo')],
    'perspective': [('pc', 'mdist'), ('pc', 'sin'), ('pc', 'cos')],
    'julian': [('p', 'power'), ('pc', 'cn')],
    'juliascope': [('p', 'power'), ('pc', 'cn')],
    'curve': [('p', 'xamp'), ('p', 'yamp'), ('pc', 'x2'), ('pc', 'y2')],
}


def var_args(name):
    """Argument sources for ``var_<name>`` after the common prefix."""
    if name in _SPECIAL_ARGS:
        return list(_SPECIAL_ARGS[name])
    return [('p', p) for p in var_param_order[name]]


def uses_rng(name):
    return name in RNG_USERS
