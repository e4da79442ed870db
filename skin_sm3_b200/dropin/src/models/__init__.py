"""``src.models`` overlay: simclr.py comes from here, the rest (projector, evaluator, resnet, ...) from the
reference's src/models when present, and a torchvision-backed ``resnet`` registry as the last resort."""
import os as _os
import sys as _sys

_mine = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_mine]
for _p in list(_sys.path):
    _cand = _os.path.abspath(_os.path.join(_p or ".", "src", "models"))
    if _cand != _mine and _cand not in __path__ and _os.path.isfile(_os.path.join(_cand, "__init__.py")):
        __path__.append(_cand)
__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_mine)), "_fallback", "src", "models"))
