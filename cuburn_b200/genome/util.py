"""
Small genome helpers: nested-dict utilities, structure hash, palette codec and
the compact JSON writer (reference cuburn/genome/util.py).
"""
import base64
import json
from hashlib import sha1

import numpy as np

from . import spectypes


def get(dct, default, *keys):
    if len(keys) == 1:
        keys = keys[0].split('.')
    for k in keys:
        if k in dct:
            dct = dct[k]
        else:
            return default
    return dct


def flatten(src):
    """{'a': {'b': 1}} -> {'a.b': 1}"""
    out = {}

    def go(dct, ctx):
        for k, v in dct.items():
            k = str(k)
            if isinstance(v, dict):
                go(v, ctx + (k,))
            else:
                out['.'.join(ctx + (k,))] = v
    go(src, ())
    return out


def unflatten(dct):
    out = {}
    for key, v in dct.items():
        parts = key.split('.')
        d = out
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return out


def hash(gnm):
    """
    Structure hash: two genomes with the same set of keys share a compiled
    iterate module (genome/util.py:55-65).  Key order is canonicalised here so
    that dict ordering cannot split the cache.
    """
    keys = sorted(flatten(gnm).keys())
    return sha1('\n'.join(keys).encode('utf-8')).hexdigest()


def resolve_spec(sp, path):
    for name in path:
        if isinstance(sp, spectypes.Map):
            sp = sp.type
        else:
            sp = sp[name]
    return sp


def palette_decode(datastrs):
    """['rgb8', b64, b64, ...] -> float32 (256, 4) RGBA in [0,1], alpha = 1."""
    if datastrs[0] != 'rgb8':
        raise NotImplementedError(datastrs[0])
    raw = base64.b64decode(''.join(datastrs[1:]))
    rgb = np.frombuffer(raw, np.uint8).reshape(256, 3)
    out = np.ones((256, 4), np.float32)
    out[:, :3] = rgb / 255.0
    return out


def palette_encode(data, format='rgb8'):
    if format != 'rgb8':
        raise NotImplementedError(format)
    q = np.clip(np.round(np.asarray(data)[:, :3] * 255.0), 0, 255)
    enc = base64.b64encode(q.astype(np.uint8).tobytes()).decode('ascii')
    return ['rgb8'] + [enc[i:i + 64] for i in range(0, len(enc), 64)]


def json_encode(obj):
    """Readable JSON for genomes: %.6g numbers, sorted keys, short lines."""
    text = _enc(obj, 0).lstrip()
    return '\n'.join(line.rstrip() for line in text.split('\n')) + '\n'


def _isnum(v):
    return isinstance(v, (float, int, np.number)) and not isinstance(v, bool)


def _enc(obj, indent):
    pad = ' ' * indent

    def fold(parts, opener, closer):
        flat = opener + ', '.join(parts) + closer
        if '\n' not in flat and len(flat) + indent < 70:
            return flat
        return '\n' + pad + opener + ' ' + ('\n' + pad + ', ').join(parts) + '\n' + pad + closer

    if isinstance(obj, dict):
        if not obj:
            return '{}'
        def order(kv):
            k = kv[0]
            return (0, int(k), '') if str(k).isdigit() else (1, 0, str(k))
        items = sorted(obj.items(), key=order)
        return fold(['%s: %s' % (json.dumps(str(k)), _enc(v, indent + 2))
                     for k, v in items], '{', '}')
    if isinstance(obj, (list, tuple)):
        parts = [_enc(v, indent + 2) for v in obj]
        if parts and len(parts) % 2 == 0 and _isnum(obj[1]):
            parts = [a + ', ' + b for a, b in zip(parts[::2], parts[1::2])]
        return fold(parts, '[', ']')
    if isinstance(obj, str):
        return json.dumps(obj)
    if _isnum(obj):
        return '%.6g' % obj
    raise TypeError("Don't know how to serialize %r" % (obj,))
