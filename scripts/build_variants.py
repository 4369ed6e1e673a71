#!/usr/bin/env python
"""Build tuning variants of libtvm_b200.so (same sources, different -D knobs) into iffnerf_b200/variants/."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iffnerf_b200 import build

out_dir = os.path.join(ROOT, "iffnerf_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
combos = [32, 64, 128]
for rays in combos:
    tag = f"refbwd_r{rays}"
    out = os.path.join(out_dir, f"libtvm_{tag}.so")
    build.build(defines=[f"TVM_REF_BWD_RAYS={rays}"], out=out)
    print(out)
