"""CUDA-graph capture of launch-bound steps.

A 4096-ray training step is ~36 launches for ~1.0 ms of kernels and an iNeRF refinement step (1024 rays,
inerf/estimate_pose_inerf.py:103-186) is ~25 launches for ~0.2 ms: at these sizes the Python + launch overhead of
the host is the step time.  Every entry point of libtvm_b200.so is a plain sequence of stream-ordered launches
(no allocation, no synchronisation, no host read-back), so a whole step — ray generation, forward, loss, backward,
optimiser — can be captured once and replayed:

    static = dict(pixels=..., target=..., bg=...)          # tensors the caller refreshes in place every step
    def step():
        opt.zero_grad(set_to_none=True)
        rays = pixel_rays(K, cam_transf(), static["pixels"])
        rgb = model(rays, bg_color=static["bg"], is_train=False)[0]
        loss = torch.mean((rgb - static["target"]) ** 2)
        loss.backward(); opt.step()                         # optimiser built with capturable=True
        return loss
    graphed = CapturedStep(step, models=[model])
    for it in range(n_iters):
        static["pixels"].copy_(...); static["target"].copy_(...)
        loss = graphed()                                    # one cudaGraphLaunch
"""
from __future__ import annotations

import torch


class CapturedStep:
    """Captures `fn()` (no arguments; it reads and writes tensors that stay alive, i.e. "static" buffers) into a CUDA
    graph after `warmup` eager runs on a side stream, and replays it on every call.

    `models`: TensorVMSplit modules used inside `fn`.  Their packed parameter shadows are refreshed INSIDE the graph
    (packing is captured unconditionally), so after a replay the Python-side cache keys are invalidated: the next eager
    call re-packs from the current parameters."""

    def __init__(self, fn, models=(), warmup: int = 3):
        if not torch.cuda.is_available():
            raise RuntimeError("CapturedStep needs a CUDA device")
        self.models = list(models)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._invalidate()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()
        self._invalidate()

    def _invalidate(self):
        for m in self.models:
            m.invalidate_packed()

    def __call__(self):
        self.graph.replay()
        self._invalidate()
        return self.outputs
