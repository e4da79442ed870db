"""Overlay package: ``src`` here only provides ``src.models.simclr``; every other ``src.*`` module resolves to
the reference checkout's own ``src`` package when that is on sys.path (its training scripts put their repo
root there, tools/backbone_train.py:11-13)."""
import os as _os
import sys as _sys

_mine = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_mine]
for _p in list(_sys.path):
    _cand = _os.path.abspath(_os.path.join(_p or ".", "src"))
    if _cand != _mine and _cand not in __path__ and _os.path.isfile(_os.path.join(_cand, "__init__.py")):
        __path__.append(_cand)
