#!/usr/bin/env python
"""Build tuning variants of libtvm_b200.so (same sources, different -D knobs) into iffnerf_b200/variants/."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iffnerf_b200 import build

out_dir = os.path.join(ROOT, "iffnerf_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
combos = [(4, 32, 3), (4, 32, 4), (8, 32, 2), (8, 64, 2)]
for warps, rpc, mb in combos:
    tag = f"w{warps}_r{rpc}_b{mb}"
    out = os.path.join(out_dir, f"libtvm_{tag}.so")
    build.build(defines=[f"TVM_MARCH_WARPS={warps}", f"TVM_MARCH_RAYS_PER_CTA={rpc}", f"TVM_MARCH_MIN_BLOCKS={mb}"], out=out)
    print(out)
