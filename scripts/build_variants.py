#!/usr/bin/env python
"""Build tuning variants of libtvm_b200.so (same sources, different -D knobs) into iffnerf_b200/variants/."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iffnerf_b200 import build

out_dir = os.path.join(ROOT, "iffnerf_b200", "variants")
os.makedirs(out_dir, exist_ok=True)
combos = [(1, 12, 1), (1, 12, 0), (2, 6, 2), (4, 3, 0)]
for warps, mb, rpc in combos:
    tag = f"bwd_w{warps}_b{mb}_r{rpc}"
    out = os.path.join(out_dir, f"libtvm_{tag}.so")
    defs = [f"TVM_BWD_WARPS={warps}", f"TVM_BWD_MIN_BLOCKS={mb}"] + ([f"TVM_BWD_RPC_FIXED={rpc}"] if rpc else [])
    build.build(defines=defs, out=out)
    print(out)
