"""Genome schema, accessors and codecs (host side, CPU only)."""
