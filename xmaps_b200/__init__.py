"""Importable alias of the ``x-maps_b200/`` package directory.

The product lives in ``x-maps_b200/`` (the directory name mandated for this repository); a
hyphen cannot appear in a Python import statement, so ``import xmaps_b200`` resolves to the
same files by pointing this package's search path at that directory.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "x-maps_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py"), "r") as _fh:
    exec(compile(_fh.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _os, _fh
