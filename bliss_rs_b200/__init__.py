"""Importable alias of the `bliss-rs_b200/` package directory (a hyphen cannot be imported).

`import bliss_rs_b200` exposes everything `bliss-rs_b200/__init__.py` defines.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bliss-rs_b200"))
_init = _os.path.join(__path__[0], "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _os, _f, _init
