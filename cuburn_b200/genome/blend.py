"""
Nodes and edges -> animation documents.

Same entry points as the reference blender (cuburn/genome/blend.py): ``node_to_anim``,
``edge_to_anim``, ``resolve``, ``apply_temporal_offset``, ``blend``, ``merge_edits``,
``tospline``, ``merge_nodes``, ``blend_xform``, ``padding_xform``, ``sort_xforms``.
Written from the behaviour those functions have on documents -- pinned by sixteen
documents produced by executing the reference's blender (tests/golden/blend_golden.json)
and by tests/test_convert.py -- not from its text.

Model.  A *node* fixes, for every animated parameter, a position and a velocity at one
instant: a bare number ``p`` (velocity 0) or ``[p, v]``.  An *edge* adds knots
``[t, value, ...]`` in between (t = 0 and t = 1 override the ends).  The blend of two
nodes over ``duration`` is, per parameter, the animation spline

    [p0, v0, p1, v1, t, value, ...]      (collapsed to [p0, p1] or p0 when nothing moves)

which is what the renderer packs into knot rows.  Periodic parameters (angles) first
move the far end by whole periods so that the interpolated motion is the one the end
velocities describe.  Xforms are blended in pairs; one without a partner is blended
with a stand-in that draws nothing new (``padding_xform``).

The document walk is driven by the schema: ``_schema_walk`` visits the positions the
schema (genome/specs.py) defines and hands each leaf to a rule chosen by the leaf's
type, so documents may carry extra keys (ids, links, blend options) without those
leaking into the animation.
"""
from itertools import zip_longest

from . import spectypes, specs, variations
from .use import Wrapper
from .util import get, resolve_spec, flatten, unflatten

_ABSENT = object()


# ---- walking documents along the schema -------------------------------------------
def _schema_walk(schema, docs, leaf):
    """
    Combine several parallel documents (``None`` where one has nothing) into one.
    ``leaf(spec, values)`` produces the value at a schema leaf; dictionaries of the
    schema are descended into for every key at least one document carries and the
    schema knows; maps (tables keyed by the document, like ``xforms``) are left to the
    caller and come out as ``None``.
    """
    if isinstance(schema, spectypes.Map):
        return None
    if not isinstance(schema, dict):
        return leaf(schema, docs)
    docs = [d if d is not None else {} for d in docs]
    keys = set()
    for d in docs:
        keys.update(d)
    return {k: _schema_walk(schema[k], [d.get(k) for d in docs], leaf)
            for k in keys if k in schema}


# ---- one parameter -----------------------------------------------------------------
def split_node_val(spl, val):
    """Node value -> (position, velocity); missing values sit at the schema default."""
    if val is None:
        return spl.default, 0
    if isinstance(val, (int, float)):
        return val, 0
    position, velocity = val
    return position, velocity


def _edge_knots(edit):
    """Edge knots ``[t, value, ...]`` -> (value at 0 | None, value at 1 | None, the rest
    as a flat list); knots whose value is ``None`` are deletions and vanish."""
    at = {}
    for t, value in zip(edit[::2], edit[1::2]):
        at[t] = value
    start, end = at.pop(0, None), at.pop(1, None)
    inner = []
    for t, value in at.items():
        if value is not None:
            inner.extend((t, value))
    return start, end, inner


def _unwind(period, p0, v0, p1, v1, duration, start, end):
    """
    Far end of a periodic parameter, moved by whole periods.  The mean of the two end
    velocities says how many periods the parameter travels during the blend; the far
    end keeps its phase and takes the number of turns nearest to that.  Edge knots at
    the ends are then brought to the same branch.
    """
    turns = duration * (v0 + v1) / (2.0 * period)
    direction = 1.0 if turns >= 0 else -1.0
    phase = (float(p1 - p0) / period) % direction       # in [0, 1) or (-1, 0]
    p1 = p0 + (round(turns - phase) + phase) * period
    if start is not None:
        p0 += round(float(start - p0) / period) * period
    if end is not None:
        p1 += round(float(end - p1) / period) * period
    return p0, p1


def tospline(spl, src, dst, edit, duration):
    """Two node values (+ the edge's knots) -> one animation spline value."""
    p0, v0 = split_node_val(spl, src)
    p1, v1 = split_node_val(spl, dst)
    if spl.var:
        # a variation parameter missing on one side copies the other side: falling back
        # to the default could push the variation through a singularity mid-blend
        if src is None:
            p0 = p1
        if dst is None:
            p1 = p0
    start, end, inner = _edge_knots(edit) if edit else (None, None, [])
    if spl.period:
        p0, p1 = _unwind(spl.period, p0, v0, p1, v1, duration, start, end)
    if inner or v0 or v1 or start or end:
        return [p0, v0, p1, v1] + inner
    return p0 if p0 == p1 else [p0, p1]


# ---- whole documents ---------------------------------------------------------------
def merge_nodes(sp, src, dst, edit, duration):
    """Blend ``src`` and ``dst`` (with the edge's ``edit``) wherever schema ``sp`` reaches."""
    def leaf(spec, values):
        s, d, e = values
        if isinstance(spec, spectypes.Spline):
            return tospline(spec, s, d, e, duration)
        if isinstance(spec, spectypes.List):
            if isinstance(spec.type, spectypes.Palette):
                # palettes become keyframes at the two ends of the blend
                s = [[0] + s] if s is not None else s
                d = [[1] + d] if d is not None else d
            return (s or []) + (d or []) + (e or [])
        for v in (e, d, s):                 # scalars: the edge wins, then the far node
            if v is not None:
                return v
        return None
    return _schema_walk(sp, [src, dst, edit], leaf)


def merge_edits(sv, av, bv):
    """Overlay two edit trees: knot lists concatenate, other values take the later one."""
    def leaf(spec, values):
        a, b = values
        if isinstance(spec, (spectypes.List, spectypes.Spline)):
            return (a or []) + (b or [])
        return a if b is None else b

    def walk(schema, a, b):
        # unlike _schema_walk, edit trees keep keys the schema does not list and descend
        # into maps (their entries all have the map's type)
        if isinstance(schema, (dict, spectypes.Map)):
            a, b = a or {}, b or {}
            sub = (lambda k: schema.type) if isinstance(schema, spectypes.Map) else schema.__getitem__
            return {k: walk(sub(k), a.get(k), b.get(k)) for k in set(a) | set(b)}
        return leaf(schema, (a, b))
    return walk(sv, av, bv)


def resolve(gdb, item):
    """
    An item with its chain of ``base`` documents folded in, oldest first: a later
    document overrides an earlier one -- except on edges, where knot lists and lists
    accumulate along the chain.
    """
    lineage = [item]
    while lineage[0].get('base') is not None:
        lineage.insert(0, gdb.get(lineage[0]['base']))
    schema = specs.toplevels[item['type']]
    accumulate = item['type'] == 'edge'
    merged = {}
    for doc in lineage:
        for path, value in flatten(doc).items():
            kind = resolve_spec(schema, path.split('.'))
            if accumulate and path in merged and \
                    isinstance(kind, (spectypes.Spline, spectypes.List)):
                merged[path] = merged[path] + value
            elif accumulate and isinstance(kind, (spectypes.Spline, spectypes.List)):
                merged[path] = list(value)
            else:
                merged[path] = value
    return unflatten(merged)


def apply_temporal_offset(node, offset=0):
    """The node ``offset`` blend-durations later: periodic ``[position, velocity]``
    values advance along their velocity, everything else stays."""
    class Shifted(Wrapper):
        def wrap_spline(self, path, spec, val):
            moving = spec.period is not None and isinstance(val, list) and val[1]
            return [val[0] + offset * val[1], val[1]] if moving else val
    view = Shifted(node)
    return view.visit(view)


def _ref_and_offset(ref):
    """'id@0.25' -> ('id', 0.25); a bare id has offset 0."""
    ident, _, offset = ref.partition('@')
    return ident, (float(offset) if offset else 0)


def node_to_anim(gdb, node, half):
    """A node blended with itself one period later (``half``: half a period, centred)."""
    node = resolve(gdb, node)
    t0, t1 = (-0.25, 0.25) if half else (0, 1)
    loop = {'blend': {'duration': t1 - t0, 'xform_sort': 'natural'}}
    return blend(apply_temporal_offset(node, t0), apply_temporal_offset(node, t1), loop)


def edge_to_anim(gdb, edge):
    edge = resolve(gdb, edge)
    ends = []
    for side in ('src', 'dst'):
        ident, offset = _ref_and_offset(edge['link'][side])
        ends.append(apply_temporal_offset(resolve(gdb, gdb.get(ident)), offset))
    return blend(ends[0], ends[1], edge)


def blend(src, dst, edit=None):
    """Two resolved nodes (and the resolved edge between them) -> an animation."""
    edit = edit or {}
    options = {}
    for doc in (src, dst, edit):
        options.update(doc.get('blend', {}))
    options = Wrapper(options, specs.blend)
    duration = options.duration

    anim = merge_nodes(specs.node, src, dst, edit, duration)
    sxfs, dxfs = src.get('xforms', {}), dst.get('xforms', {})
    anim['xforms'] = {}
    paired = {}
    for skey, dkey in sort_xforms(sxfs, dxfs, options.xform_sort, explicit=options.xform_map):
        knots = merge_edits(specs.xform, get(edit, {}, 'xforms', 'src', skey),
                            get(edit, {}, 'xforms', 'dst', dkey))
        # 'dup': the partner is a copy of the xform on the other side whose weight
        # fades in from (or out to) nothing
        if skey == 'dup':
            knots.setdefault('weight', []).extend([0, 0])
        if dkey == 'dup':
            knots.setdefault('weight', []).extend([1, 0])
        name = '%s_%s' % (skey or 'pad', dkey or 'pad')
        anim['xforms'][name] = blend_xform(sxfs.get(skey), dxfs.get(dkey), knots, duration)
        paired[name] = (skey, dkey)
    _blend_chaos(anim['xforms'], paired, sxfs, dxfs, duration)
    if 'final_xform' in src or 'final_xform' in dst:
        anim['final_xform'] = blend_xform(src.get('final_xform'), dst.get('final_xform'),
                                          edit.get('final_xform'), duration, True)
    anim['type'] = 'animation'
    anim.setdefault('time', {})['duration'] = duration
    return anim


# ---- xforms --------------------------------------------------------------------------
def _blend_chaos(blended, pairs, sxfs, dxfs, duration):
    """
    Xaos tables (an addition: the reference's blender predates its kernel's xaos support
    and drops them).  ``chaos[n]`` of an xform multiplies the weight of xform ``n`` when
    the trajectory comes from this one; in the animation the xforms are the *pairs*, so
    the entry for pair (s, d) blends the source's entry for ``s`` into the destination's
    entry for ``d``.  A side that says nothing about a target counts as 1.
    """
    spec = specs.xform['chaos'].type
    for name, (skey, dkey) in pairs.items():
        sch = (sxfs.get(skey) or {}).get('chaos')
        dch = (dxfs.get(dkey) or {}).get('chaos')
        blended[name].pop('chaos', None)            # the schema walk leaves maps to us
        if sch is None and dch is None:
            continue
        table = {}
        for target, (sk2, dk2) in pairs.items():
            sv = None if sch is None else sch.get(sk2, sch.get(_as_int(sk2)))
            dv = None if dch is None else dch.get(dk2, dch.get(_as_int(dk2)))
            if sv is not None or dv is not None:
                table[target] = tospline(spec, sv, dv, None, duration)
        blended[name]['chaos'] = table


def _as_int(key):
    try:
        return int(key)
    except (TypeError, ValueError):
        return None


def blend_xform(sxf, dxf, edits, duration, isfinal=False):
    if sxf is None:
        sxf = padding_xform(dxf, isfinal)
    if dxf is None:
        dxf = padding_xform(sxf, isfinal)
    out = merge_nodes(specs.xform, sxf, dxf, edits, duration)
    if out.get('chaos', 0) is None:
        del out['chaos']                            # maps are blended by the caller
    return out


# Variations with a hole at the origin: shrinking their weight to zero opens the hole
# over the whole picture, so their stand-in is a point reflection instead.
hole_variations = frozenset(('spherical', 'ngon', 'julian', 'juliascope', 'polar',
                             'wedge_sph', 'wedge_julia', 'bipolar'))
# Variations that are the identity at their default parameters: the stand-in keeps them
# (at defaults) so that only their parameters, not their weights, move during the blend.
ident_variations = frozenset(('rectangles', 'fan2', 'blob', 'perspective', 'super_shape'))


def _is_flipped(xf, which):
    return get(xf, 45, which, 'spread') > 90


def padding_xform(xf, isfinal):
    """The stand-in an unpaired xform is blended with: an identity-like xform matching
    the orientation of ``xf``'s affines; a final xform's stand-in also has no colour pull."""
    pad = {'pre_affine': {'angle': 45}, 'variations': {}}
    if isfinal:
        pad['weight'] = 0
        pad['color_speed'] = 0
    for which in ('pre_affine', 'post_affine'):
        if _is_flipped(xf, which):
            pad[which] = {'angle': 135, 'spread': 135}
    used = list(xf.get('variations', {}))
    for name in used:
        if name in hole_variations:
            pad['pre_affine']['angle'] += 180
            pad['variations'] = {'linear': {'weight': -1}}
            return pad
        if name in ident_variations:
            pad['variations'][name] = {p: sp.default
                                       for p, sp in variations.var_params[name].items()}
    kept = pad['variations']
    if not kept:
        kept['linear'] = {'weight': 1}
    else:
        share = 1 / float(len(kept))
        for params in kept.values():
            params['weight'] = share
    return pad


def halfhearted_human_sort_key(key):
    """Numeric keys in numeric order, then the others alphabetically."""
    try:
        return (0, int(key), '')
    except (TypeError, ValueError):
        return (1, 0, str(key))


def _scalar(value):
    return value[0] if isinstance(value, (list, tuple)) else value


def sort_xforms(sxfs, dxfs, sortmethod, explicit=()):
    """
    Pair the xforms of two nodes: a list of ``(src key | None, dst key | None)``.
    Explicit pairs come first (a later pair releases an earlier claim on either xform;
    'pad' and 'dup' are not xforms and can repeat).  The rest pair up rank by rank
    inside classes of equal orientation (pre and post affine flipped or not), ranked by
    ``sortmethod``: 'weight', 'weightflip' (descending on the far side), 'color', or
    key order.
    """
    by_src, by_dst = {}, {}
    for skey, dkey in explicit:
        if skey not in ('pad', 'dup') and skey in by_src:
            by_dst.pop(by_src.pop(skey), None)
        if dkey not in ('pad', 'dup') and dkey in by_dst:
            by_src.pop(by_dst.pop(dkey), None)
        by_src[skey], by_dst[dkey] = dkey, skey
    pairs = sorted(by_src.items(), key=lambda p: (str(p[0]), str(p[1])))

    if sortmethod in ('weight', 'weightflip'):
        def rank(xfs):
            return lambda k: _scalar(xfs[k].get('weight', 0))
    elif sortmethod == 'color':
        def rank(xfs):
            return lambda k: _scalar(xfs[k].get('color', 0))
    else:
        def rank(xfs):
            return halfhearted_human_sort_key

    def classes(xfs, taken):
        out = {}
        for key, xf in xfs.items():
            if key not in taken:
                cls = (_is_flipped(xf, 'pre_affine'), _is_flipped(xf, 'post_affine'))
                out.setdefault(cls, []).append(key)
        return {cls: sorted(keys, key=rank(xfs)) for cls, keys in out.items()}
    sclasses, dclasses = classes(sxfs, by_src), classes(dxfs, by_dst)
    for cls in sorted(set(sclasses) | set(dclasses)):
        near, far = sclasses.get(cls, []), dclasses.get(cls, [])
        if sortmethod == 'weightflip':
            far = far[::-1]
        pairs.extend(zip_longest(near, far))
    return pairs
