"""Importable alias of the package directory ``hyperspectral_super-resolution_b200/``.

The directory name required by the project layout contains a hyphen and therefore cannot be
written in an ``import`` statement; ``import hsr_b200`` resolves every submodule
(``hsr_b200.kernels``, ``hsr_b200.s2_emit.srf`` ...) inside that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "hyperspectral_super-resolution_b200")
__path__[:] = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
