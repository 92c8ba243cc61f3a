"""
Genome look-up by file name or ID (reference cuburn/genome/db.py): a directory of
``<id>.json`` files, or one JSON file holding many documents; ``get_anim`` turns
whatever it finds (flam3 XML, node, edge, animation) into an animation.
"""
import json
import os
import warnings

from . import convert


class GenomeDB(object):
    def __init__(self):
        self.stashed = {}

    def _get(self, id):
        raise NotImplementedError()

    def get(self, id):
        if id in self.stashed:
            return self.stashed[id]
        return self._get(id)

    def stash(self, id, gnm):
        self.stashed[id] = gnm

    def get_anim(self, name, half=False):
        """``(animation dict, basename suitable for output files)``."""
        basename = os.path.basename(name)
        head, dot, ext = basename.rpartition('.')
        if dot and ext in ('json', 'flam3', 'flame'):
            basename = head
        else:
            ext = ext if dot else ''

        if os.path.isfile(name) and ext in ('flam3', 'flame'):
            with open(name) as fp:
                flames = convert.XMLGenomeParser.parse(fp.read())
            if len(flames) != 1:
                warnings.warn('%d flames in file, only using one.' % len(flames))
            gnm = convert.flam3_to_node(flames[0])
        elif os.path.isfile(name) and ext == 'json':
            with open(name) as fp:
                gnm = json.load(fp)
        else:
            gnm = self.get(name)

        if gnm['type'] == 'node':
            gnm = convert.node_to_anim(self, gnm, half=half)
        elif gnm['type'] == 'edge':
            gnm = convert.edge_to_anim(self, gnm)
        assert gnm['type'] == 'animation', 'Unrecognized genome type.'
        return gnm, basename


class OneFileDB(GenomeDB):
    def __init__(self, dct):
        super().__init__()
        assert dct.get('type') == 'onefiledb', "Doesn't look like a OneFileDB."
        self.dct = dct

    @classmethod
    def read(cls, path):
        with open(path) as fp:
            return cls(json.load(fp))

    def _get(self, id):
        return self.dct[id]


class FilesystemDB(GenomeDB):
    def __init__(self, path):
        super().__init__()
        self.path = path

    def _get(self, id):
        if not id.endswith('.json'):
            id += '.json'
        with open(os.path.join(self.path, id)) as fp:
            return json.load(fp)


def connect(path):
    if os.path.isfile(path):
        try:
            return OneFileDB.read(path)
        except (ValueError, AssertionError):
            pass
        return FilesystemDB(os.path.dirname(path) or '.')
    return FilesystemDB(path)
