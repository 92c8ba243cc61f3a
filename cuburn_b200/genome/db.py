"""
Genome look-up by file name or ID, and conversion of whatever is found into an
animation.  Same surface as the reference module (cuburn/genome/db.py):
``connect(path)`` -> ``GenomeDB`` with ``get / stash / get_anim``; ``OneFileDB``
holds many documents in one JSON file, ``FilesystemDB`` is a directory of
``<id>.json`` files.
"""
import json
import os
import warnings

from . import convert

_XML_EXTENSIONS = ('flam3', 'flame')
_KNOWN_EXTENSIONS = _XML_EXTENSIONS + ('json',)


def _split_name(name):
    """'dir/spark.flam3' -> ('spark', 'flam3'); unknown extensions stay in the name."""
    base = os.path.basename(name)
    head, dot, ext = base.rpartition('.')
    if dot and ext in _KNOWN_EXTENSIONS:
        return head, ext
    return base, ''


def _load_xml(path):
    with open(path) as fp:
        flames = convert.XMLGenomeParser.parse(fp.read())
    if len(flames) != 1:
        warnings.warn('%d flames in file, only using one.' % len(flames))
    return convert.flam3_to_node(flames[0])


class GenomeDB(object):
    """Base class: documents by ID, with an in-memory overlay (``stash``)."""
    def __init__(self):
        self.stashed = {}

    def _get(self, id):
        raise NotImplementedError()

    def stash(self, id, gnm):
        self.stashed[id] = gnm

    def get(self, id):
        try:
            return self.stashed[id]
        except KeyError:
            return self._get(id)

    def load(self, name):
        """The raw document behind ``name``: a file on disk wins over a DB id."""
        _, ext = _split_name(name)
        if os.path.isfile(name):
            if ext in _XML_EXTENSIONS:
                return _load_xml(name)
            if ext == 'json':
                with open(name) as fp:
                    return json.load(fp)
        return self.get(name)

    def to_anim(self, gnm, half=False):
        kind = gnm['type']
        if kind == 'node':
            gnm = convert.node_to_anim(self, gnm, half=half)
        elif kind == 'edge':
            gnm = convert.edge_to_anim(self, gnm)
        assert gnm['type'] == 'animation', 'Unrecognized genome type.'
        return gnm

    def get_anim(self, name, half=False):
        """``(animation dict, basename suitable for output files)`` (db.py:22-54)."""
        return self.to_anim(self.load(name), half), _split_name(name)[0]


class FilesystemDB(GenomeDB):
    def __init__(self, path):
        super().__init__()
        self.path = path

    def _get(self, id):
        fname = id if id.endswith('.json') else id + '.json'
        with open(os.path.join(self.path, fname)) as fp:
            return json.load(fp)


class OneFileDB(GenomeDB):
    def __init__(self, dct):
        super().__init__()
        if dct.get('type') != 'onefiledb':
            raise AssertionError("Doesn't look like a OneFileDB.")
        self.dct = dct

    @classmethod
    def read(cls, path):
        with open(path) as fp:
            return cls(json.load(fp))

    def _get(self, id):
        return self.dct[id]


def connect(path):
    """A file is a OneFileDB (or, failing that, its directory); a directory a FilesystemDB."""
    if not os.path.isfile(path):
        return FilesystemDB(path)
    try:
        return OneFileDB.read(path)
    except (ValueError, AssertionError):
        return FilesystemDB(os.path.dirname(path) or '.')
