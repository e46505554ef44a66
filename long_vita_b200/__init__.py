"""Import shim: the package sources live in ``long-vita_b200/`` (the directory name the project
layout prescribes, which is not a valid Python identifier).  ``import long_vita_b200`` resolves
here and executes the real package ``__init__`` with ``__path__`` pointing at that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "long-vita_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
