"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/sm3_b200.h declares, the ctypes table covers them, and the product refuses CPU tensors
(no fallback).  No compute calls: there is no GPU in the build container."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sm3_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sm3_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    import skin_sm3_b200 as sm3
    if not os.path.exists(sm3.LIB_PATH):
        sm3.build()
    lib = ctypes.CDLL(sm3.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in sm3_b200.h but not exported"
    from skin_sm3_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"
    lib.sm3_version.restype = ctypes.c_int
    assert lib.sm3_version() == 1


def test_argument_validation_without_gpu():
    """Pure host-side validation paths return error codes + messages, never crash."""
    from skin_sm3_b200 import _lib
    l = _lib.lib()
    rc = l.sm3_l2norm_fwd(None, 4, None, 0, 8, 0, None, 0, None, 1e-12, None)
    assert rc == -1 and b"null" in l.sm3_last_error()
    rc = l.sm3_multihead_ce(None, 0, None, 4, 8, None, None, 1.0, 0, -100, None, None, 1.0, None, 0, None)
    assert rc == -1
    assert l.sm3_infonce_host_scratch_bytes(0, 128, 2, 0) == 0
    assert l.sm3_multihead_ce_workspace_bytes(512, 8) > 0
    # pipelined host entry: size query and handle validation are host-only
    assert l.sm3_host_pipe_scratch_bytes(4096, 128, 2, 0, 2) > l.sm3_infonce_host_scratch_bytes(4096, 128, 2, 0)
    assert l.sm3_host_pipe_scratch_bytes(4096, 128, 2, 0, 9) == 0
    h = ctypes.c_void_p()
    assert l.sm3_host_pipe_create(ctypes.byref(h), 4096, 128, 2, 0, 2, None, 0) == -1 and not h
    assert l.sm3_host_pipe_wait(None, 0) == -1 and b"null handle" in l.sm3_last_error()
    assert l.sm3_host_pipe_destroy(None) == 0


def test_product_refuses_cpu_tensors():
    import skin_sm3_b200 as sm3
    p = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        sm3.cal_logits(p, p, 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        sm3.l2_normalize(p)
    with pytest.raises(RuntimeError, match="CUDA"):
        sm3.multihead_ce(torch.randn(4, 24), torch.zeros(4, 8, dtype=torch.long))


def test_product_never_imports_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import skin_sm3_b200, skin_sm3_b200.dropin; "
            "bad=[m for m in sys.modules if m.startswith('oracle')]; assert not bad, bad" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "skin_sm3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")
def test_dropin_module_matches_reference_structure():
    """Same class names, ctor signatures, state_dict keys and shapes as the reference module (CPU, no compute)."""
    import importlib.util
    import inspect
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from skin_sm3_b200 import dropin
        dropin.install()
        import src.models.simclr as mine
        assert mine.__file__.startswith(os.path.join(ROOT, "skin_sm3_b200"))
        spec = importlib.util.spec_from_file_location("ref_simclr", "/root/reference/src/models/simclr.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        for name in ["SimCLR", "SimCLRSkin", "SimCLRSkinV2", "SimCLRSkinV21", "SimCLRSkinV22", "SimCLRSkinV23",
                     "SimCLRSkinV3", "SimCLRSkinV32"]:
            a, b = getattr(mine, name), getattr(ref, name)
            assert str(inspect.signature(a.__init__)) == str(inspect.signature(b.__init__)), name
        for name in ["SimCLRSkinV3", "SimCLRSkinV32", "SimCLRSkinV2", "SimCLRSkinV23"]:
            ma = getattr(mine, name)("resnet18", proj_dim=16, temperature=0.1)
            mb = getattr(ref, name)("resnet18", proj_dim=16, temperature=0.1)
            sa, sb = ma.state_dict(), mb.state_dict()
            assert list(sa.keys()) == list(sb.keys()), name
            assert all(sa[k].shape == sb[k].shape for k in sa), name
            ma.load_state_dict(sb)
        m = mine.SimCLRSkinV32("resnet18", proj_dim=16)
        assert (m.derm_feat_dim, m.clinic_feat_dim, m.temperature) == (512, 512, 0.5)
        m.derm_backbone.projector = m.clinic_backbone.projector = m.cross_proj = None   # mlc_train.py:344-346
        assert len(m.extract(torch.randn(2, 3, 32, 32), torch.randn(2, 3, 32, 32))) == 2
    finally:
        sys.path.remove("/root/reference")


def test_dropin_rejects_out_of_range_settings_at_construction():
    """D > 256 or 1/T >= 83 are outside the fused kernels' range: the drop-in says so when the model is built (CPU, no
    compute), pointing at SM3_DROPIN=0, instead of failing at the first forward."""
    import sys
    import pytest
    from oracle import vendor_ref
    root = vendor_ref.ref_root()
    if root is None:
        pytest.skip("no reference checkout / vendored copy (src.models.resnet comes from it)")
    sys.path.insert(0, root)
    try:
        from skin_sm3_b200 import dropin
        dropin.install()
        import src.models.simclr as mine
        assert mine.__file__.startswith(os.path.join(ROOT, "skin_sm3_b200"))
        with pytest.raises(ValueError, match="proj_dim"):
            mine.SimCLR("resnet18", None, proj_dim=512)
        with pytest.raises(ValueError, match="temperature"):
            mine.SimCLRSkinV32("resnet18", None, proj_dim=128, temperature=0.01)
        mine.SimCLR("resnet18", None, proj_dim=128, temperature=0.1)
    finally:
        sys.path.remove(root)
