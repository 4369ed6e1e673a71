#!/usr/bin/env python
"""Build tuning variants of libtvm_b200.so (same sources, different -D knobs) into iffnerf_b200/variants/."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iffnerf_b200 import build

out_dir = os.path.join(ROOT, "iffnerf_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
combos = [5, 6]
for mb in combos:
    tag = f"bwd_pose_b{mb}"
    out = os.path.join(out_dir, f"libtvm_{tag}.so")
    build.build(defines=[f"TVM_BWD_MIN_BLOCKS_POSE={mb}"], out=out)
    print(out)
