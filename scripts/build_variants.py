#!/usr/bin/env python
"""Build tuning variants of libtvm_b200.so (same sources, different -D knobs) into iffnerf_b200/variants/.
    python scripts/build_variants.py TAG=DEF1,DEF2 TAG2=DEF ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iffnerf_b200 import build

out_dir = os.path.join(ROOT, "iffnerf_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
for spec in sys.argv[1:]:
    tag, defs = spec.split("=", 1)
    out = os.path.join(out_dir, f"libtvm_{tag}.so")
    build.build(defines=[d for d in defs.split(",") if d], out=out)
    print(out)
