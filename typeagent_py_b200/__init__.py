"""Import shim: ``import typeagent_py_b200`` -> the package in ``typeagent-py_b200/``.

The package directory keeps the project's name (with its hyphen), which Python cannot
import directly; this shim re-points the package path there and runs its ``__init__``.
"""

import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "typeagent-py_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f, _os, _real
