"""Put this directory on PYTHONPATH to run the reference's training scripts UNCHANGED on the fused kernels:

    PYTHONPATH=/path/to/skin_sm3_b200/dropin/_site python tools/backbone_train.py -a resnet50 ...

Every interpreter (including mp.spawn workers) then resolves ``src.models.simclr`` to the drop-in module, finds the
synthetic dataset / torchmetrics stand-in when the real ones are absent, and swaps the script's ``cluster_memory``
(see ../_hook.py for the switches: SM3_DROPIN, SM3_SHIMS, SM3_DROPIN_KMEANS, SM3_DROPIN_QUIET).

Python imports only the FIRST ``sitecustomize`` on sys.path, so after installing the hook this file chains to the next
one (a distro / conda / cluster-launcher hook that this directory would otherwise shadow)."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))

_hook_path = os.path.join(os.path.dirname(_HERE), "_hook.py")
_spec = importlib.util.spec_from_file_location("_sm3_dropin_hook", _hook_path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
_mod.install_hook()

# chain to the sitecustomize this one shadows, if any
for _p in sys.path:
    _d = os.path.abspath(_p or ".")
    if _d == _HERE:
        continue
    _cand = os.path.join(_d, "sitecustomize.py")
    if os.path.isfile(_cand):
        _s = importlib.util.spec_from_file_location("_sm3_chained_sitecustomize", _cand)
        _m = importlib.util.module_from_spec(_s)
        try:
            _s.loader.exec_module(_m)
        except Exception as _e:  # a broken foreign hook must not take the interpreter down (site.py prints and goes on)
            print(f"[skin_sm3_b200] chained sitecustomize {_cand} failed: {_e!r}", file=sys.stderr)
        break
