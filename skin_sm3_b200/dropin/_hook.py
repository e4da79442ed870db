"""Import hook that serves ``src.models.simclr`` from this package, whatever sys.path order the caller has.

No heavy imports here: sitecustomize loads this file by path at interpreter start-up (also in every
``mp.spawn`` child, which is how the reference's scripts create their per-GPU processes,
tools/backbone_train.py:626-631)."""
import importlib.abc
import importlib.util
import os
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
TARGET = "src.models.simclr"
SHADOW = os.path.join(_DIR, "src", "models", "simclr.py")


class _ShadowFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == TARGET:
            return importlib.util.spec_from_file_location(fullname, SHADOW)
        return None


def install_hook() -> None:
    repo_root = os.path.dirname(os.path.dirname(_DIR))
    if repo_root not in sys.path:
        sys.path.append(repo_root)            # makes `skin_sm3_b200` importable from the shadow module
    if not any(isinstance(f, _ShadowFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _ShadowFinder())
    if _DIR not in sys.path:
        sys.path.append(_DIR)                 # fallback `src` / `src.models` packages when no reference checkout
    stale = sys.modules.get(TARGET)
    if stale is not None and os.path.abspath(getattr(stale, "__file__", "") or "") != SHADOW:
        del sys.modules[TARGET]
