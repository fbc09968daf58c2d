"""Importable alias for the package directory ``vit-vs-raw-iq_b200/`` (a hyphen cannot appear in a
Python module name).  ``import vit_vs_raw_iq_b200`` runs that directory's ``__init__.py`` with this
package's ``__path__`` pointing there, so ``vit_vs_raw_iq_b200.modules`` etc. resolve inside it.
All code lives in ``vit-vs-raw-iq_b200/``; this file holds none."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vit-vs-raw-iq_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init, _real
