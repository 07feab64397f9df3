"""Importable alias for the package directory ``vit-inductive-bias-distillation_b200``
(a hyphenated directory name cannot appear in an ``import`` statement).

``import basd_b200`` executes that directory's ``__init__.py`` under this name, so
``basd_b200.losses``, ``basd_b200.synthetic`` ... resolve to the files that live there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "vit-inductive-bias-distillation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
