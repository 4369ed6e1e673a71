#!/usr/bin/env python
"""L1-resident gather ceiling: the quad-granule LDG.128 micro-benchmark (tvm_gather_microbench) over a working set that
fits the SM's L1 (every SM reads the same few KB), vs the L2-resident 69 MB set.  Prints GB/s per (bytes, granule)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
buf = torch.randn(70 << 18, device=dev)            # 70 MiB of floats
sink = torch.zeros(4, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for nbytes in (16 << 10, 64 << 10, 128 << 10, 1 << 20, 69 << 20):
    for gran in (64, 192):
        moved = C.c_ulonglong(0)
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(_lib.load_bench().tvm_gather_microbench(_lib.ptr(buf), nbytes // gran * gran, gran, 256, _lib.ptr(sink), C.byref(moved), st), "mb")
            e1.record(); torch.cuda.synchronize()
            if rep: best = max(best, moved.value / (e0.elapsed_time(e1) / 1e3) / 1e9)
        print(json.dumps({"working_set_bytes": nbytes, "granule": gran, "gbs": round(best, 1)}), flush=True)
