"""Turns the public headers into cffi cdef text, the way the reference's build script does
(reference: libgdf/python/libgdf_cffi/libgdf_build.py:4-8 - "cdef" = concatenation of the headers)."""
import os
import re

from . import INCLUDE_DIR


def _strip(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)      # block comments
    text = re.sub(r"//[^\n]*", "", text)                   # line comments
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    return text


def header_cdef(*relative_paths):
    parts = []
    for rel in relative_paths:
        with open(os.path.join(INCLUDE_DIR, rel)) as fh:
            parts.append(_strip(fh.read()))
    return "\n".join(parts)
