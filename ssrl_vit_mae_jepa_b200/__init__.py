"""Importable alias of the ``ssrl-vit-mae-jepa_b200/`` package directory (hyphens are not valid in a Python
module name).  All code lives there; this file only redirects the package search path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ssrl-vit-mae-jepa_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
