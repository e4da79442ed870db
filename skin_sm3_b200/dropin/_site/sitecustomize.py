"""Put this directory on PYTHONPATH to run the reference's training scripts UNCHANGED on the fused kernels:

    PYTHONPATH=/path/to/skin_sm3_b200/dropin/_site python tools/backbone_train.py -a resnet50 ...

Every interpreter (including mp.spawn workers) then resolves ``src.models.simclr`` to the drop-in module.
Set SM3_DROPIN=0 to disable without touching PYTHONPATH."""
import importlib.util
import os

if os.environ.get("SM3_DROPIN", "1") == "1":
    _hook_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_hook.py")
    _spec = importlib.util.spec_from_file_location("_sm3_dropin_hook", _hook_path)
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    _mod.install_hook()
