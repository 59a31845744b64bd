"""Importable alias of the product package.

The product lives in the directory ``multi-robot-informative-path-planning_b200/`` (the name the
build contract fixes); hyphens make that directory un-importable by name, so this shim exposes it
as ``mripp_b200``: ``import mripp_b200.gp``, ``from mripp_b200.src.rrt.rrt import MR_IPP`` ...
"""
import os as _os

_IMPL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "multi-robot-informative-path-planning_b200")
__path__ = [_IMPL]
with open(_os.path.join(_IMPL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_IMPL, "__init__.py"), "exec"))
del _f
