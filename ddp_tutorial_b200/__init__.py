"""Importable alias of the ``ddp-tutorial_b200/`` package directory.

The product lives in ``ddp-tutorial_b200/`` (the name the build contract fixes); a
hyphen is not importable, so this shim points the package search path there and
re-exports its public surface: ``import ddp_tutorial_b200 as b2``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ddp-tutorial_b200")
__path__.insert(0, _REAL)

with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
del _f
