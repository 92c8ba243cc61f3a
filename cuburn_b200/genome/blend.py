"""
Nodes / edges -> animations.

Drop-in for the reference blender (cuburn/genome/blend.py): ``node_to_anim``,
``edge_to_anim``, ``resolve``, ``apply_temporal_offset``, ``blend``,
``merge_edits``, ``tospline``, ``merge_nodes``, ``blend_xform``,
``padding_xform``, ``sort_xforms``.  A node fixes position and velocity of every
spline at one instant; blending two nodes yields the ``[p0, v0, p1, v1, t, p, ...]``
animation splines the renderer packs, with periodic parameters (angles) extended
by whole turns so that the interpolated motion matches the end velocities
(blend.py:167-191), missing xforms padded by identity-like xforms
(blend.py:223-258) and xforms paired by the configured sort (blend.py:266-314).
"""
from itertools import zip_longest

from . import spectypes, specs, variations
from .use import Wrapper
from .util import get, resolve_spec, flatten, unflatten


def node_to_anim(gdb, node, half):
    """A node looped onto itself (one full period, or half of it centred on t=0)."""
    node = resolve(gdb, node)
    osrc, odst = (-0.25, 0.25) if half else (0, 1)
    src = apply_temporal_offset(node, osrc)
    dst = apply_temporal_offset(node, odst)
    edge = dict(blend=dict(duration=odst - osrc, xform_sort='natural'))
    return blend(src, dst, edge)


def edge_to_anim(gdb, edge):
    edge = resolve(gdb, edge)
    src, osrc = _split_ref_id(edge['link']['src'])
    dst, odst = _split_ref_id(edge['link']['dst'])
    src = apply_temporal_offset(resolve(gdb, gdb.get(src)), osrc)
    dst = apply_temporal_offset(resolve(gdb, gdb.get(dst)), odst)
    return blend(src, dst, edge)


def resolve(gdb, item):
    """Merge an item with its chain of ``base`` documents (later overrides earlier;
    on edges, spline and list values concatenate instead)."""
    is_edge = item['type'] == 'edge'
    spec = specs.toplevels[item['type']]

    def chain(i):
        if i.get('base') is not None:
            return chain(gdb.get(i['base'])) + [i]
        return [i]
    items = [flatten(i) for i in chain(item)]
    out = {}
    for k in set(key for i in items for key in i):
        sp = resolve_spec(spec, k.split('.'))
        vs = [i[k] for i in items if k in i]
        if is_edge and isinstance(sp, (spectypes.Spline, spectypes.List)):
            merged = []
            for v in vs:
                merged += v
            out[k] = merged
        else:
            out[k] = vs[-1]
    return unflatten(out)


def _split_ref_id(s):
    parts = s.split('@')
    if len(parts) == 1:
        return parts[0], 0
    return parts[0], float(parts[1])


def apply_temporal_offset(node, offset=0):
    """Advance every periodic ``[position, velocity]`` spline by offset x velocity."""
    class _Offset(Wrapper):
        def wrap_spline(self, path, spec, val):
            if spec.period is not None and isinstance(val, list) and val[1]:
                position, velocity = val
                return [position + offset * velocity, velocity]
            return val
    wr = _Offset(node)
    return wr.visit(wr)


def blend(src, dst, edit=None):
    """
    Blend two pre-merged nodes (and an optional pre-merged edge ``edit``) into an
    animation document (blend.py:80-133).
    """
    edit = edit or {}
    opts = {}
    for d in (src, dst, edit):
        opts.update(d.get('blend', {}))
    opts = Wrapper(opts, specs.blend)

    blended = merge_nodes(specs.node, src, dst, edit, opts.duration)
    pairs = sort_xforms(src.get('xforms', {}), dst.get('xforms', {}), opts.xform_sort,
                        explicit=opts.xform_map)
    blended['xforms'] = {}
    for sxf_key, dxf_key in pairs:
        bxf_key = (sxf_key or 'pad') + '_' + (dxf_key or 'pad')
        xf_edits = merge_edits(specs.xform,
                               get(edit, {}, 'xforms', 'src', sxf_key),
                               get(edit, {}, 'xforms', 'dst', dxf_key))
        # 'dup' pairs an xform with a copy of its partner whose weight fades in/out
        if sxf_key == 'dup':
            xf_edits.setdefault('weight', []).extend([0, 0])
        if dxf_key == 'dup':
            xf_edits.setdefault('weight', []).extend([1, 0])
        blended['xforms'][bxf_key] = blend_xform(
            src.get('xforms', {}).get(sxf_key), dst.get('xforms', {}).get(dxf_key),
            xf_edits, opts.duration)

    if 'final_xform' in src or 'final_xform' in dst:
        blended['final_xform'] = blend_xform(src.get('final_xform'), dst.get('final_xform'),
                                             edit.get('final_xform'), opts.duration, True)
    blended['type'] = 'animation'
    blended.setdefault('time', {})['duration'] = opts.duration
    return blended


def merge_edits(sv, av, bv):
    """Merge two edit trees according to the spec ``sv``."""
    if isinstance(sv, (dict, spectypes.Map)):
        av, bv = av or {}, bv or {}

        def sub(k):
            return sv.type if isinstance(sv, spectypes.Map) else sv[k]
        return dict((k, merge_edits(sub(k), av.get(k), bv.get(k)))
                    for k in set(av) | set(bv))
    if isinstance(sv, (spectypes.List, spectypes.Spline)):
        return (av or []) + (bv or [])
    return bv if bv is not None else av


def split_node_val(spl, val):
    """A node spline value -> (position, velocity)."""
    if val is None:
        return spl.default, 0
    if isinstance(val, (int, float)):
        return val, 0
    return val


def tospline(spl, src, dst, edit, duration):
    """Two node values and the edge's knots -> one animation spline value."""
    sp, sv = split_node_val(spl, src)
    dp, dv = split_node_val(spl, dst)
    # variation parameters copy the other side instead of falling back to the
    # default, which could make a variation explode mid-blend
    if spl.var:
        if src is None:
            sp = dp
        if dst is None:
            dp = sp

    knots = dict(zip(edit[::2], edit[1::2])) if edit else {}
    e0, e1 = knots.pop(0, None), knots.pop(1, None)
    rest = []
    for k, v in knots.items():
        if v is not None:
            rest += [k, v]

    if spl.period:
        # periodic extension: pick the number of whole turns that best matches
        # the mean end velocity (blend.py:167-184)
        def sign(x):
            return 1. if x >= 0 else -1.
        movement = duration * (sv + dv) / (2.0 * spl.period)
        angdiff = (float(dp - sp) / spl.period) % (sign(movement))
        dp = sp + (round(movement - angdiff) + angdiff) * spl.period
        if e0 is not None:
            sp += round(float(e0 - sp) / spl.period) * spl.period
        if e1 is not None:
            dp += round(float(e1 - dp) / spl.period) * spl.period
    if rest or sv or dv or e0 or e1:
        return [sp, sv, dp, dv] + rest
    if sp != dp:
        return [sp, dp]
    return sp


def merge_nodes(sp, src, dst, edit, duration):
    if isinstance(sp, dict):
        src, dst, edit = src or {}, dst or {}, edit or {}
        return dict((k, merge_nodes(sp[k], src.get(k), dst.get(k), edit.get(k), duration))
                    for k in set(src) | set(dst) | set(edit) if k in sp)
    if isinstance(sp, spectypes.Map):
        # maps (xform tables) are blended pairwise by the caller
        return None
    if isinstance(sp, spectypes.Spline):
        return tospline(sp, src, dst, edit, duration)
    if isinstance(sp, spectypes.List):
        if isinstance(sp.type, spectypes.Palette):
            if src is not None:
                src = [[0] + src]
            if dst is not None:
                dst = [[1] + dst]
        return (src or []) + (dst or []) + (edit or [])
    return edit if edit is not None else dst if dst is not None else src


def blend_xform(sxf, dxf, edits, duration, isfinal=False):
    if sxf is None:
        sxf = padding_xform(dxf, isfinal)
    if dxf is None:
        dxf = padding_xform(sxf, isfinal)
    return merge_nodes(specs.xform, sxf, dxf, edits, duration)


# an xform using one of these is padded with the inverted identity
hole_variations = ('spherical ngon julian juliascope polar '
                   'wedge_sph wedge_julia bipolar').split()
# identity functions at their default parameter values
ident_variations = 'rectangles fan2 blob perspective super_shape'.split()


def padding_xform(xf, isfinal):
    """The do-nothing partner an unmatched xform is blended with (blend.py:223-258)."""
    vs = {}
    out = {'variations': vs, 'pre_affine': {'angle': 45}}
    if isfinal:
        out.update(weight=0, color_speed=0)
    if get(xf, 45, 'pre_affine', 'spread') > 90:
        out['pre_affine'] = {'angle': 135, 'spread': 135}
    if get(xf, 45, 'post_affine', 'spread') > 90:
        out['post_affine'] = {'angle': 135, 'spread': 135}
    for k in xf.get('variations', {}):
        if k in hole_variations:
            out['pre_affine']['angle'] += 180
            vs.clear()
            vs['linear'] = dict(weight=-1)
            return out
        if k in ident_variations:
            vs[k] = dict((pk, pv.default) for pk, pv in variations.var_params[k].items())
    if vs:
        n = float(len(vs))
        for k in vs:
            vs[k]['weight'] = 1 / n
    else:
        vs['linear'] = dict(weight=1)
    return out


def halfhearted_human_sort_key(key):
    try:
        return (0, int(key), '')
    except (TypeError, ValueError):
        return (1, 0, str(key))


def sort_xforms(sxfs, dxfs, sortmethod, explicit=()):
    """Yield (src key | None, dst key | None) pairs (blend.py:266-314)."""
    fwd, rev = {}, {}
    for sx, dx in explicit:
        if sx not in ('pad', 'dup') and sx in fwd:
            rev.pop(fwd.pop(sx, None), None)
        if dx not in ('pad', 'dup') and dx in rev:
            fwd.pop(rev.pop(dx, None), None)
        fwd[sx] = dx
        rev[dx] = sx
    for sd in sorted(fwd.items(), key=lambda kv: (str(kv[0]), str(kv[1]))):
        yield sd

    # remaining xforms are matched within classes: (pre flipped?, post flipped?)
    scl, dcl = {}, {}
    for cl, xfs, taken in ((scl, sxfs, fwd), (dcl, dxfs, rev)):
        for k, v in xfs.items():
            if k in taken:
                continue
            xcl = (get(v, 45, 'pre_affine', 'spread') > 90,
                   get(v, 45, 'post_affine', 'spread') > 90)
            cl.setdefault(xcl, []).append(k)

    def _num(v):
        return v[0] if isinstance(v, (list, tuple)) else v

    def order(keys, dct):
        if sortmethod in ('weight', 'weightflip'):
            return sorted(keys, key=lambda k: _num(dct[k].get('weight', 0)))
        if sortmethod == 'color':
            return sorted(keys, key=lambda k: _num(dct[k].get('color', 0)))
        return sorted(keys, key=halfhearted_human_sort_key)

    for cl in sorted(set(scl) | set(dcl)):
        ssort = order(scl.get(cl, []), sxfs)
        dsort = order(dcl.get(cl, []), dxfs)
        if sortmethod == 'weightflip':
            dsort = list(reversed(dsort))
        for sd in zip_longest(ssort, dsort):
            yield sd
