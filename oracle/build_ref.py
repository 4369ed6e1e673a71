"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe for `oracle/_ref/`: the UNMODIFIED reference's own implementation of
the render path, taken from /root/reference so that it can travel to the GPU box (which has no /root/reference).

    python -m oracle.build_ref

Copies, byte for byte, the reference files the path lives in (renderer.py, models/tensorBase.py,
models/tensoRF.py and what they import: models/ref.py, ref_utils.py, image.py, utils.py, ray_utils.py,
dataLoader/ray_utils.py) into oracle/_ref/, which is git-ignored (never part of the history) but not gpurun-ignored.
`oracle/ref_import.py` imports it with the same stub shim as the in-container reference.  Used by
`bench.py --impl reference` (cpu_baseline.kind = "reference") and nothing else; the product never imports it.
"""
import hashlib
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["renderer.py", "utils.py", "ray_utils.py", "models/__init__.py", "models/tensorBase.py", "models/tensoRF.py",
         "models/ref.py", "models/ref_utils.py", "models/image.py", "models/sh.py", "dataLoader/ray_utils.py"]


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[build_ref] {SRC} not present (GPU box): keeping the prebuilt oracle/_ref as is")
        return os.path.isdir(DST)
    manifest = []
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue                                   # optional modules (e.g. models/__init__.py)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"[build_ref] {len(manifest)} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
