"""Recipe that makes the UNMODIFIED reference travel to the GPU box  --  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (Dylan-H-Wang/skin-sm3) is pure Python with no setup.py / pyproject.toml, so the base contract's
``pip install --target`` cannot install it.  This script is its equivalent: it copies the reference's own Python
sources, byte for byte, from the read-only checkout into ``oracle/_ref/skin_sm3/`` -- a git-ignored build artefact
(like the compiled .so files; it is listed in .gitignore, not in .gpurunignore, so it ships with the gpurun snapshot
but never enters the history).  Nothing is modified, nothing is copied anywhere else in the repository.

What uses it (never the product):
  * ``bench.py --impl reference`` and ``cpu_baseline`` run the real ``SimCLRSkinV3._cal_logits`` + ``nn.CrossEntropyLoss``
    (src/models/simclr.py:290-322, tools/backbone_train.py:531) on the box's host cores  -> ``kind: "reference"``;
  * ``tests/test_scripts_unchanged.py`` launches the real ``tools/backbone_train.py`` / ``tools/mlc_train.py``
    with ``PYTHONPATH=skin_sm3_b200/dropin/_site`` to prove they run unchanged on the fused kernels.

Run:  python oracle/vendor_ref.py [/root/reference]      (``__graft_entry__.build()`` runs it when the checkout exists)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "skin_sm3")
WHAT = ("src", "tools", "inference.py", "resnet.py", "run.sh", "LICENSE", "README.md")


def ref_root():
    """Directory holding the unmodified reference sources: the vendored copy if present, else the checkout, else None."""
    for cand in (DEST, os.environ.get("SM3_REFERENCE_ROOT", ""), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "src", "models", "simclr.py")):
            return cand
    return None


def vendor(src_root: str = "/root/reference") -> str:
    if not os.path.isfile(os.path.join(src_root, "src", "models", "simclr.py")):
        raise FileNotFoundError(f"{src_root} is not a skin-sm3 checkout")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    manifest = []
    for item in WHAT:
        s = os.path.join(src_root, item)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DEST, item), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copy2(s, os.path.join(DEST, item))
    for dirpath, _, files in sorted(os.walk(DEST)):
        for f in sorted(files):
            p = os.path.join(dirpath, f)
            with open(p, "rb") as fh:
                manifest.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {os.path.relpath(p, DEST)}")
    with open(os.path.join(DEST, "MANIFEST.sha256"), "w") as fh:
        fh.write("# sha256 of the files copied unmodified from the reference checkout by oracle/vendor_ref.py\n")
        fh.write("\n".join(manifest) + "\n")
    return DEST


if __name__ == "__main__":
    print(vendor(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
