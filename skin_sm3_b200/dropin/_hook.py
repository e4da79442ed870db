"""Import hook that lets the reference's scripts run UNCHANGED on the fused kernels, whatever sys.path order the caller
has.  No heavy imports here: sitecustomize loads this file by path at interpreter start-up (also in every ``mp.spawn``
child, which is how the reference's scripts create their per-GPU processes, tools/backbone_train.py:626-631).

What it installs (each switchable by environment, read at install time):
  * ``src.models.simclr`` is served from this package                                   (SM3_DROPIN=0 disables);
  * after ``src.utils.data.datasets`` has been imported, the synthetic Derm7pt-shaped dataset ``SM3SyntheticPairs`` is
    registered in its namespace, where ``init_dataset`` looks datasets up (src/utils/misc.py:433)   (SM3_SHIMS=0 disables);
  * a minimal ``torchmetrics`` is appended to the END of sys.path, so an installed torchmetrics always wins; the scripts
    import four of its names at module level (tools/backbone_train.py:32-37)                (SM3_SHIMS=0 disables);
  * after ``src.utils.misc`` has been imported, ``init_distributed_mode`` (called first thing by every script's
    ``main``, e.g. tools/mlc_train.py:303) is wrapped: on return it replaces the running script's module-level
    ``cluster_memory`` (tools/mlc_train.py:116-189) by ``skin_sm3_b200.cluster_memory``     (SM3_DROPIN_KMEANS=0 disables)
    and gives the script's own ``Model`` class (tools/mlc_train.py:58-89) the fused normalise + prototype-heads tail
    (``skin_sm3_b200.mlc_model_forward``)                                                   (SM3_DROPIN_HEADS=0 disables).
The first time the shadow module is served a one-line notice goes to stderr (SM3_DROPIN_QUIET=1 silences it), so a
run whose hook silently did not install (e.g. another sitecustomize earlier on the path) is easy to spot.
"""
import importlib.abc
import importlib.util
import os
import sys

_DIR = os.path.dirname(os.path.abspath(__file__))
TARGET = "src.models.simclr"
SHADOW = os.path.join(_DIR, "src", "models", "simclr.py")
SHIMS = os.path.join(_DIR, "_shims")


def _note(msg: str) -> None:
    if os.environ.get("SM3_DROPIN_QUIET", "0") != "1":
        print(f"[skin_sm3_b200] {msg}", file=sys.stderr, flush=True)


class _ShadowFinder(importlib.abc.MetaPathFinder):
    _announced = False

    def find_spec(self, fullname, path=None, target=None):
        if fullname == TARGET:
            if not _ShadowFinder._announced:
                _ShadowFinder._announced = True
                _note(f"drop-in active: {TARGET} is served from {SHADOW} (pid {os.getpid()})")
            return importlib.util.spec_from_file_location(fullname, SHADOW)
        return None


# ---------------------------------------------------------------------------------------------------------------
# post-import patches: the module is loaded by whichever finder owns it, then the callback runs on the module object
# ---------------------------------------------------------------------------------------------------------------
def _register_synthetic_dataset(mod) -> None:
    spec = importlib.util.spec_from_file_location("_sm3_dropin_synthetic", os.path.join(_DIR, "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    mod.__dict__.setdefault("SM3SyntheticPairs", syn.SM3SyntheticPairs)


def _wrap_init_distributed(mod) -> None:
    orig = getattr(mod, "init_distributed_mode", None)
    if orig is None or getattr(orig, "_sm3_wrapped", False):
        return

    def init_distributed_mode(args):
        out = orig(args)
        main = sys.modules.get("__main__")
        cm = getattr(main, "cluster_memory", None)
        if (os.environ.get("SM3_DROPIN_KMEANS", "1") == "1" and callable(cm)
                and getattr(cm, "__module__", "") in ("__main__", "__mp_main__")):
            from skin_sm3_b200.functional import cluster_memory
            main.cluster_memory = cluster_memory
            _note("drop-in active: cluster_memory of the running script replaced by skin_sm3_b200.cluster_memory")
        # the multi-label script defines its `Model` itself (tools/mlc_train.py:58-89): give it the fused prototype tail
        model_cls = getattr(main, "Model", None)
        if (os.environ.get("SM3_DROPIN_HEADS", "1") == "1" and isinstance(model_cls, type)
                and getattr(model_cls, "__module__", "") in ("__main__", "__mp_main__")
                and {"mlc_sa", "prototypes", "projectors", "extractor"} <= set(model_cls.__init__.__code__.co_names)):
            from skin_sm3_b200.functional import mlc_model_forward
            model_cls.forward = mlc_model_forward
            _note("drop-in active: Model.forward of the running script uses the fused prototype heads")
        return out

    init_distributed_mode._sm3_wrapped = True
    init_distributed_mode.__doc__ = orig.__doc__
    mod.init_distributed_mode = init_distributed_mode


class _PatchLoader(importlib.abc.Loader):
    def __init__(self, inner, callback):
        self._inner, self._callback = inner, callback

    def create_module(self, spec):
        return self._inner.create_module(spec)

    def exec_module(self, module):
        self._inner.exec_module(module)
        self._callback(module)


class _PostImportFinder(importlib.abc.MetaPathFinder):
    def __init__(self, patches):
        self.patches = patches
        self._busy = False

    def find_spec(self, fullname, path=None, target=None):
        cb = self.patches.get(fullname)
        if cb is None or self._busy:
            return None
        self._busy = True
        try:
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                spec = finder.find_spec(fullname, path, target)
                if spec is not None and spec.loader is not None:
                    spec.loader = _PatchLoader(spec.loader, cb)
                    return spec
        finally:
            self._busy = False
        return None


def install_hook() -> None:
    repo_root = os.path.dirname(os.path.dirname(_DIR))
    if repo_root not in sys.path:
        sys.path.append(repo_root)            # makes `skin_sm3_b200` importable from the shadow module
    if os.environ.get("SM3_DROPIN", "1") == "1":
        if not any(isinstance(f, _ShadowFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _ShadowFinder())
        if _DIR not in sys.path:
            sys.path.append(_DIR)             # fallback `src` / `src.models` packages when no reference checkout
        stale = sys.modules.get(TARGET)
        if stale is not None and os.path.abspath(getattr(stale, "__file__", "") or "") != SHADOW:
            del sys.modules[TARGET]
    patches = {}
    if os.environ.get("SM3_SHIMS", "1") == "1":
        patches["src.utils.data.datasets"] = _register_synthetic_dataset
        if SHIMS not in sys.path:
            sys.path.append(SHIMS)            # END of the path: an installed torchmetrics is found first
    if os.environ.get("SM3_DROPIN", "1") == "1" and (os.environ.get("SM3_DROPIN_KMEANS", "1") == "1" or
                                                   os.environ.get("SM3_DROPIN_HEADS", "1") == "1"):
        patches["src.utils.misc"] = _wrap_init_distributed
    if patches and not any(isinstance(f, _PostImportFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _PostImportFinder(patches))
    for name, cb in patches.items():          # modules that were imported before the hook: patch them now
        if name in sys.modules:
            cb(sys.modules[name])
