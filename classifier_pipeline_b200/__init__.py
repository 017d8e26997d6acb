"""Import alias for the hyphenated package directory ``classifier-pipeline_b200/``.

Python identifiers cannot contain ``-``; this shim makes the package importable as
``classifier_pipeline_b200`` by pointing ``__path__`` at the real directory and
executing its ``__init__.py`` in this module's namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "classifier-pipeline_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
