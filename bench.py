#!/usr/bin/env python
"""bench.py — TensoRF-VM render throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full-image render per GPU (800x800 = 640 000 rays, lego-shaped 300^3 VM field, 16x3/48x3 components,
S=1036 samples/ray, synthetic sphere occupancy) — BASELINE.json configs[1].
  value         rays/s through OctreeRender_trilinear_fast with the rays already resident in HBM (march + shade
                kernels, device-timed, L2 flushed between steps)
  e2e           the same through OctreeRender_trilinear_fast with HOST (pinned) rays: H2D of the rays and D2H of
                rgb+depth inside the timed region
  roofline      the appearance-gather kernel: bytes it FETCHES (counted by a counting build of the same kernel) / its
                device time vs the L1-resident quad-gather ceiling measured in the same run; `stages` carries the
                sigma-march kernel and the HBM-denominated figures
  cpu_baseline  the reference's CPU renderer (oracle/_ref when present, else the oracle port) on a bounded ray sample
  gpu_eager_baseline  the reference's op sequence (oracle port, chunk 4096) as eager ATen kernels on the same GPU
N>1 (torchrun): the job is N views = N x 640 000 rays as ONE ray set, dealt to the ranks in cyclic 4096-ray tiles
(iffnerf_b200.sharding: replicated factors, rays are the sharded unit, no data-path collective); weak scaling, value =
all rays / max-over-ranks time.  `parity`: the sharded render re-assembled == a single-GPU render, and the data-parallel
gradients == the single-process gradients of the global batch.  `strong` (config 4, one 1080p image over N GPUs) and
`dp_train` (config 3 under data parallelism) are reported beside it.
--impl reference: times the reference's CPU renderer on the host cores (rank 0 only).
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "TensoRF VM render rays/s"
UNIT = "rays/s"
H = W = 800
GRID = [300, 300, 300]
TILE = 4096


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=16384, help="rays in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational legs (other configs, eager baseline)")
    return ap.parse_args()


def workload_config(extra=None):
    cfg = {"workload": "lego-shaped TensoRF VM 300^3 (density 16x3, app 48x3, app_dim 27, MLP_Fea 150-128-128-3), "
                       "800x800 full-image render, S=1036, forward only",
           "rays_per_step": H * W, "n_samples": 1036, "grid": GRID, "density_shift": 0.0,
           "occupancy": "sphere r=1.0 on 200^3", "white_bg": True, "cache": "L2 flushed between timed steps"}
    if extra:
        cfg.update(extra)
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.idx = gpu_index
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference legs
def reference_cpu_renderer():
    """(render(rays) -> None, kind): the reference's own CPU implementation of the path.  kind = "reference" when the
    unmodified reference files are at hand (oracle/_ref built by `python -m oracle.build_ref`, or /root/reference),
    else "port" (the oracle restatement, asserted bit-identical to the reference by oracle/make_golden.py)."""
    import contextlib
    import io
    import torch
    from oracle import fixtures as fx, ref_import
    fld = fx.make_field(GRID, density_shift=0.0)
    if ref_import.reference_root() is not None:
        R = ref_import.import_reference()
        torch.manual_seed(fx.SEED)
        with contextlib.redirect_stdout(io.StringIO()):
            m = R.TensorVMSplit(fld.aabb.clone(), list(fld.grid), "cpu", density_n_comp=[16] * 3,
                                appearance_n_comp=[48] * 3, app_dim=27, near_far=list(fld.near_far),
                                shadingMode="MLP_Fea", alphaMask_thres=1e-4, density_shift=0.0, distance_scale=25,
                                pos_pe=6, view_pe=2, fea_pe=2, featureC=128, step_ratio=0.5, fea2denseAct="softplus")
        m.alphaMask = R.AlphaGridMask("cpu", fld.occupancy.aabb.clone(), fld.occupancy.volume.clone())

        def render(rays):
            with torch.no_grad():
                R.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True, ndc_ray=False,
                                              device="cpu")
        return render, "reference"
    from oracle import tensorf_oracle as orc

    def render(rays):
        with torch.no_grad():
            orc.render_rays(fld, rays, chunk=4096, white_bg=True)
    return render, "port"


def cpu_baseline_rate(n_rays):
    """Bounded sample of the SAME 800x800 ray set on all host threads: (rays/s, threads, kind, description)."""
    import torch
    from oracle import fixtures as fx
    torch.set_num_threads(os.cpu_count())
    render, kind = reference_cpu_renderer()
    rays = fx.config2_rays(H, W)
    stride = max(1, rays.shape[0] // n_rays)
    sample = rays[::stride][:n_rays].contiguous()
    render(sample[:4096])                                                       # warm-up
    t0 = time.perf_counter()
    render(sample)
    dt = time.perf_counter() - t0
    return sample.shape[0] / dt, torch.get_num_threads(), kind, \
        f"{sample.shape[0]} rays (every {stride}th ray of the 800x800 image), chunk 4096, 1 warm-up + 1 timed pass"


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU renderer, all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    import torch
    from oracle import fixtures as fx
    torch.set_num_threads(os.cpu_count())
    render, kind = reference_cpu_renderer()
    rays = fx.config2_rays(H, W)
    n = 8192
    stride = rays.shape[0] // n
    times = []
    for s in range(args.warmup + args.steps):
        sample = rays[(s % stride)::stride][:n].contiguous()
        t0 = time.perf_counter()
        render(sample)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    sample_desc = f"{n} rays per step (strided sample of the 800x800 image), chunk 4096"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config({"rays_per_step": n, "note": "the reference's CPU renderer "
                                       "(OctreeRender_trilinear_fast, chunk 4096) on the host cores"}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                             "sample": sample_desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- timing helpers
def device_timer(torch, dev, flush, barrier):
    def timed(fn, steps, warmup):
        """Sum of the device times (ms) of `steps` calls, L2 flushed before each, after `warmup` untimed calls."""
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)
    return timed


def plain_timer(torch, dev):
    def timeit(fn, steps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps
    return timeit


# ---------------------------------------------------------------------------------------------- informational legs
def gpu_eager_baseline(dev, rays_dev):
    """The reference's own op sequence (TensorBase.forward, tensorBase.py:775-917, restated by the oracle port) as eager
    ATen kernels on this GPU, chunk 4096 like renderer.py:12-25, timed with CUDA events the way the reference's
    profile_performance.py:143-155 does.  This is the same-box GPU number the kernels should be compared with."""
    import torch
    from oracle import fixtures as fx, tensorf_oracle as orc
    fld = fx.make_field(GRID, density_shift=0.0)
    for name in ("density_plane", "density_line", "app_plane", "app_line", "mlp_w", "mlp_b"):
        setattr(fld, name, [t.to(dev) for t in getattr(fld, name)])
    fld.basis = fld.basis.to(dev)
    fld.aabb = fld.aabb.to(dev)
    fld.occupancy.aabb = fld.occupancy.aabb.to(dev)
    fld.occupancy.volume = fld.occupancy.volume.to(dev)
    orig_geo = orc.step_geometry

    def geo_on_dev(aabb, grid, step_ratio):            # host scalars (CPU fp32, as the reference) moved to the device
        with torch.device("cpu"):
            geo = orig_geo(aabb, grid, step_ratio)
        return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in geo.items()}
    orc.step_geometry = geo_on_dev
    try:
        with torch.no_grad(), torch.device(dev):
            def full():
                return orc.render_rays(fld, rays_dev, chunk=4096, white_bg=True)
            out = full()                                                        # warm-up (allocator)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                out = full()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / 2
    finally:
        orc.step_geometry = orig_geo
    n = rays_dev.shape[0]
    return {"value": n / (ms / 1e3), "unit": UNIT, "ms_per_image": ms, "chunk": 4096,
            "kind": "oracle port of the reference's op sequence, eager ATen CUDA kernels, fp32",
            "note": "1 warm-up + 2 timed full images, CUDA events"}, out["rgb_map"]


def extra_configs(model, dev, syn):
    """Informational timings of BASELINE configs 3 and 5 on the same field (their parity is in tests/)."""
    import torch
    import iffnerf_b200 as I
    out = {}
    timeit = plain_timer(torch, dev)
    allrays = syn.config2_rays()
    g = torch.Generator().manual_seed(0)
    rays = allrays[torch.randint(0, allrays.shape[0], (4096,), generator=g)].to(dev)
    target = torch.rand(4096, 3, device=dev)
    ones = torch.ones(3, device=dev)

    def train_step():
        model.zero_grad(set_to_none=True)
        rgb, _, _, alpha, _, _ = model(rays, bg_color=ones, is_train=True, N_samples=1039)
        (torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))).backward()
    model.train()
    ms = timeit(train_step)
    out["config3_train_step"] = {"rays": 4096, "n_samples": 1039, "ms_fwd_bwd": ms, "rays_per_s": 4096 / (ms / 1e3)}
    model.eval()
    model.zero_grad(set_to_none=True)
    for p in model.parameters():
        p.requires_grad_(False)
    prays = allrays[torch.randint(0, allrays.shape[0], (64 * 1024,), generator=g)].to(dev)
    ptarget = torch.rand(64 * 1024, 3, device=dev)
    bg = torch.rand(3, device=dev)

    def pose_step():
        r = prays.clone().requires_grad_(True)
        rgb = model(r, bg_color=bg, is_train=False)[0]
        torch.mean((rgb - ptarget) ** 2).backward()
    ms = timeit(pose_step, steps=5, warm=2)
    model.eval_sample_outputs = False      # the pose loop reads rgb / opacity only: early termination in both directions
    ms_et = timeit(pose_step, steps=5, warm=2)
    out["config5_pose_step_64x1024"] = {"rays": 64 * 1024, "ms_fwd_bwd_to_rays": ms, "rays_per_s": 65536 / (ms / 1e3),
                                        "ms_without_sample_outputs": ms_et,
                                        "rays_per_s_without_sample_outputs": 65536 / (ms_et / 1e3)}
    gq = torch.Generator().manual_seed(11)
    pq = torch.randn(540000, 3, generator=gq)
    pq = pq / pq.norm(dim=-1, keepdim=True) * (0.55 + 0.5 * torch.rand(540000, 1, generator=gq))
    dq = torch.randn(540000, 3, generator=gq)
    rays6 = torch.cat([pq, dq / dq.norm(dim=-1, keepdim=True)], -1).to(dev)

    def raybank():
        with torch.no_grad():
            return model(rays6, N_samples=20, sample_func=model.sample_point_color, white_bg=True)
    ms = timeit(raybank, steps=10, warm=3)
    out["iffnerf_raybank_540k_x20"] = {"rays": 540000, "n_samples": 20, "ms": ms, "rays_per_s": 540000 / (ms / 1e3)}
    focal = 400.0 / math.tan(0.5 * syn.FOV_X)
    Kc = torch.tensor([[[focal, 0.0, 400.0], [0.0, focal, 400.0], [0.0, 0.0, 1.0]]])
    base = torch.cat([syn.orbit_pose(), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0).to(dev)
    delta = torch.zeros(3, 4, device=dev, requires_grad=True)
    opt = torch.optim.Adam([delta], lr=1e-3, capturable=True)
    pix = torch.stack([torch.randint(200, 600, (1024,), generator=g), torch.randint(200, 600, (1024,), generator=g)],
                      -1).to(device=dev, dtype=torch.int32)
    tgt = torch.rand(1024, 3, device=dev)

    def inerf_step():
        opt.zero_grad(set_to_none=True)
        pose = base + torch.cat([delta, torch.zeros(1, 4, device=dev)], 0)
        r = I.pixel_rays(Kc, pose, pix)
        rgb = model(r, bg_color=bg, is_train=False)[0]
        loss = torch.mean((rgb - tgt) ** 2)
        loss.backward()
        opt.step()
        return loss
    ms_eager = timeit(inerf_step, steps=20, warm=3)
    graphed = I.graphs.CapturedStep(inerf_step, models=[model], warmup=1)
    ms_graph = timeit(graphed, steps=20, warm=3)
    out["config5_inerf_step_1024"] = {"rays": 1024, "sample_outputs": False, "ms_eager": ms_eager,
                                      "ms_cuda_graph": ms_graph, "rays_per_s_cuda_graph": 1024 / (ms_graph / 1e3)}
    model.eval_sample_outputs = True
    for p in model.parameters():
        p.requires_grad_(True)
    model.train()
    model.zero_grad(set_to_none=True)
    try:
        topt = torch.optim.Adam(model.get_optparam_groups(0.02, 1e-3), betas=(0.9, 0.99), capturable=True, fused=True)
        adam_kind = "torch fused"
    except (RuntimeError, TypeError, ValueError):
        topt = torch.optim.Adam(model.get_optparam_groups(0.02, 1e-3), betas=(0.9, 0.99), capturable=True)
        adam_kind = "torch foreach"
    jit = torch.rand(4096, device=dev)

    def full_train_step():
        topt.zero_grad(set_to_none=True)
        rgb, _, _, alpha, _, _ = model(rays, bg_color=ones, is_train=True, N_samples=1039, jitter=jit)
        loss = torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))
        loss.backward()
        topt.step()
        return loss
    ms_eager = timeit(full_train_step, steps=10, warm=3)
    graphed_t = I.graphs.CapturedStep(full_train_step, models=[model], warmup=1)
    ms_graph = timeit(graphed_t, steps=10, warm=3)
    out["config3_train_step_with_adam"] = {"rays": 4096, "adam": adam_kind, "ms_eager": ms_eager,
                                           "ms_cuda_graph": ms_graph, "rays_per_s_cuda_graph": 4096 / (ms_graph / 1e3)}
    model.eval()
    model.zero_grad(set_to_none=True)
    del graphed, graphed_t
    rm = syn.build_model(GRID, dev, shading="Ref", occ_res=None)
    rm.alphaMask = model.alphaMask
    rm.train()

    def ref_train_step():
        rm.zero_grad(set_to_none=True)
        rgb, _, _, alpha, _, _ = rm(rays, bg_color=ones, is_train=True, N_samples=1039)
        (torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))).backward()
    out["ref_head_train_step"] = {"rays": 4096, "n_samples": 1039, "ms_fwd_bwd": timeit(ref_train_step)}
    del rm
    return out


def dp_train(model, dev, syn, world, rank, dist):
    """BASELINE configs[2] under ray-sharded data parallelism: 4096 rays PER RANK (global batch 4096 N), the packed
    factor-gradient bucket and the small MLP bucket all-reduced over NCCL.  Also the parity of that path: the
    all-reduced (averaged) gradients of every rank == the gradients a single process computes for the global batch."""
    import torch
    from iffnerf_b200 import sharding
    allrays = syn.config2_rays()
    n_local = 4096
    rays_all, target_all, jit_all = [], [], []
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        rays_all.append(allrays[torch.randint(0, allrays.shape[0], (n_local,), generator=g)])
        target_all.append(torch.rand(n_local, 3, generator=g))
        jit_all.append(torch.rand(n_local, generator=g))
    rays, target, jit = rays_all[rank].to(dev), target_all[rank].to(dev), jit_all[rank].to(dev)
    ones = torch.ones(3, device=dev)
    model.train()

    def loss_of(rgb, alpha, tgt):
        return torch.mean((rgb - tgt) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))

    state = {"sync": None}

    def step():
        sync = state["sync"]
        model.zero_grad(set_to_none=True)
        rgb, _, _, alpha, _, _ = model(rays, bg_color=ones, is_train=True, N_samples=1039, jitter=jit)
        loss_of(rgb, alpha, target).backward()
        if sync is not None:
            sync.finish()

    def time_steps():
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    ms_nccl = None
    if world > 1:
        state["sync"] = sharding.GradSync(model, average=True, transport="nccl").install()
        ms_nccl = time_steps()
        state["sync"].remove()
        state["sync"] = sharding.GradSync(model, average=True).install()     # "auto": NVLink peer memory when available
    sync = state["sync"]
    ms = time_steps()
    out = {"rays_global": n_local * world, "rays_per_rank": n_local, "ms_fwd_bwd_allreduce": ms,
           "rays_per_s": n_local * world / (ms / 1e3), "scaling": "weak"}
    if world > 1:
        out["allreduce_bytes_per_step"] = sync.bytes // max(sync.calls, 1)
        out["allreduce_transport"] = sync.transport_used
        out["ms_fwd_bwd_allreduce_nccl"] = ms_nccl
        # parity: gradients after the all-reduce vs the single-process gradients of the global batch (every rank computes
        # them rank-batch by rank-batch; both losses are means, so the global loss is the mean of the per-rank losses)
        names = [n for n, _ in model.named_parameters()]
        dp_grads = [p.grad.detach().clone() for _, p in model.named_parameters()]
        sync.remove()
        model.zero_grad(set_to_none=True)
        for r in range(world):
            rgb, _, _, alpha, _, _ = model(rays_all[r].to(dev), bg_color=ones, is_train=True, N_samples=1039,
                                           jitter=jit_all[r].to(dev))
            (loss_of(rgb, alpha, target_all[r].to(dev)) / world).backward()
        worst, worst_name = 0.0, ""
        for nme, p, gdp in zip(names, model.parameters(), dp_grads):
            err = float((p.grad - gdp).abs().max() / p.grad.abs().max().clamp_min(1e-30))
            if err > worst:
                worst, worst_name = err, nme
        w = torch.tensor([worst], dtype=torch.float64, device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        out["parity"] = {"max_rel_grad_err_vs_single_process": w.item(), "worst_param_rank0": worst_name,
                         "tolerance": 1e-4, "ok": bool(w.item() <= 1e-4)}
    model.eval()
    model.zero_grad(set_to_none=True)
    return out


def strong_config4(dev, syn, world, dist):
    """BASELINE configs[3]: ONE 1920x1080 image of the truck-shaped (non-cubic ~300^3-voxel) field, ray-sharded over the
    ranks in cyclic 4096-ray tiles, every rank ends up with the full image (16 B/ray).  Strong scaling: total work
    fixed, time = max over ranks.  Parity: the assembled image == rank 0's own single-GPU render (torch.equal)."""
    import torch
    import iffnerf_b200 as I
    from iffnerf_b200 import sharding
    m, rays = syn.config4(dev)
    m.eval()
    rays = rays.to(dev)

    def step(placement="auto"):
        if world > 1:
            return sharding.render_sharded(rays, m, I.OctreeRender_trilinear_fast, white_bg=True, device=dev,
                                           placement=placement)
        rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev)
        return rgb, depth

    def time_it(placement):
        for _ in range(2):
            out = step(placement)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = step(placement)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out
    ms, out = time_it("auto")
    res = {"rays": int(rays.shape[0]), "grid": m.gridSize.tolist(), "ms_per_image": ms,
           "rays_per_s": rays.shape[0] / (ms / 1e3), "scaling": "strong", "gathered": world > 1}
    if world > 1:
        res["placement"] = ("nccl all_gather (symmetric memory unavailable: %s)" % sharding._peer_fallback_reason
                            if sharding._peer_fallback_reason else
                            "shading epilogue stores into every rank's image over NVLink peer memory + 1 barrier")
        rgb1, _, depth1, _, _ = I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev)
        ok = torch.equal(out[0], rgb1) and torch.equal(out[1], depth1)
        ms_g, out_g = time_it("gather")
        ok_g = torch.equal(out_g[0], rgb1) and torch.equal(out_g[1], depth1)
        res["ms_per_image_nccl_gather"] = ms_g
        same = torch.tensor([int(ok), int(ok_g)], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        res["parity"] = {"sharded_image_equals_single_gpu_render": bool(same[0].item()),
                         "nccl_gather_path_equals_single_gpu_render": bool(same[1].item()),
                         "compared": "rgb and depth, every rank"}
    del m
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------- the product arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import iffnerf_b200 as I
    from iffnerf_b200 import _lib, build, sharding, synthetic as syn

    build.build()
    lib = _lib.load()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly ONE JSON line: whatever NCCL prints at the level the caller chose (NCCL_DEBUG is left
        # alone) goes to stderr while the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    model = syn.config2_model(dev)
    model.eval()
    S = model.nSamples
    # the job: `world` views of the orbit as ONE ray set, dealt to the ranks in cyclic 4096-ray tiles
    views = torch.cat([syn.config2_rays(H, W, theta_deg=35.0 + 45.0 * v) for v in range(world)])
    n_global = views.shape[0]
    mine = sharding.shard_index(n_global, world, rank, TILE)
    rays_host = views[mine].contiguous().pin_memory()
    rays_dev = rays_host.to(dev)
    n = rays_dev.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)            # > 126 MB L2
    rgb_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    depth_host = torch.empty((n,), dtype=torch.float32).pin_memory()

    def stream():
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
    timed = device_timer(torch, dev, flush, barrier)

    def step_device():
        # the reference-facing call (renderer.py:12-25) on device-resident rays
        rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays_dev, model, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device=dev)
        return {"rgb_map": rgb, "depth_map": depth}

    d, keep = model.field_desc()
    need = C.c_size_t(0)
    lib.tvm_workspace_bytes(C.byref(d), n, _lib.F_SPLIT_APP, C.byref(need))
    ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
    bg = model._bg(None, True, dev)
    fetch = torch.zeros((n,), dtype=torch.int32, device=dev)

    def march(flags, counts=None):
        _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays_dev), n, rays_dev.shape[1], S, None, _lib.ptr(bg),
                                      _lib.F_EARLY_TERM | _lib.F_NO_SHADE | _lib.F_SPLIT_APP | flags, None, None, None,
                                      None, None, None, None, None, _lib.ptr(counts), _lib.ptr(ws), ws.numel(),
                                      stream()), "march")

    def step_march():
        march(0)

    def step_gather():
        march(_lib.F_GATHER_ONLY)

    def step_e2e():
        I.OctreeRender_trilinear_fast(rays_host, model, chunk=4096, N_samples=-1, white_bg=True, ndc_ray=False,
                                      device=dev, out_host=(rgb_host, depth_host))
        torch.cuda.current_stream(dev).synchronize()

    with ClockSampler(local_rank) as clk:
        launches0 = lib.tvm_launch_count()
        ms_dev = timed(step_device, args.steps, args.warmup)
        launches = (lib.tvm_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    ms_march = timed(step_march, args.steps, 1)
    ms_gather = timed(step_gather, args.steps, 1)             # the lists of the last march are still in `ws`
    ms_e2e = timed(step_e2e, args.steps, 1)

    # work counters of one step (from the march workspace) and the gather kernel's fetch count (counting build)
    march(0)
    march(_lib.F_GATHER_ONLY | _lib.F_COUNT_FETCH, fetch)
    wsv = model.workspace_views(d, ws, n)
    cnt = torch.stack([wsv["occ_count"].sum(), wsv["sigma_count"].sum(), wsv["app_count"].sum(), fetch.sum(),
                       (wsv["app_count"] > 0).sum()]).to(torch.float64)
    t = torch.tensor([ms_dev, ms_march, ms_gather, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_dev, ms_march, ms_gather, ms_e2e = t.tolist()
    v0, v, a_samples, fetched16, lit = (int(x) for x in cnt.tolist())

    # parity of the sharded job (world > 1): the tiles of view 0 gathered from all ranks == rank 0's own render of view 0
    parity = None
    if world > 1:
        view0 = views[:H * W].to(dev)
        full = sharding.render_sharded(view0, model, I.OctreeRender_trilinear_fast, white_bg=True, device=dev)
        rgb1, _, depth1, _, _ = I.OctreeRender_trilinear_fast(view0, model, white_bg=True, device=dev)
        same = torch.tensor([int(torch.equal(full[0], rgb1) and torch.equal(full[1], depth1))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        parity = {"sharded_view_equals_single_gpu_render": bool(same.item()), "rays": H * W,
                  "compared": "rgb and depth of view 0, cyclic 4096-ray tiles over %d ranks, every rank" % world}

    # measured gather ceilings for the kernels' access shape — quads of lanes, LDG.128, random 64-B
    # pieces: over the resident factor set (L2 -> SM ceiling) and over an L1-resident 64 KB set (L1 data-pipe ceiling)
    gather_peak = {}
    if rank == 0:
        bl = _lib.load_bench()
        pf = model.packed_factors()
        sink = torch.zeros(4, device=dev)
        for name, nbytes, gran in (("l2_resident_64B_gbs", pf.numel() * 4, 64), ("l2_resident_192B_gbs", pf.numel() * 4, 192),
                                   ("l1_resident_64B_gbs", 64 << 10, 64)):
            moved = C.c_ulonglong(0)
            best = 0.0
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(bl.tvm_gather_microbench(_lib.ptr(pf), nbytes // gran * gran, gran, 128, _lib.ptr(sink),
                                                    C.byref(moved), stream()), "tvm_gather_microbench")
                e1.record()
                torch.cuda.synchronize(dev)
                if rep > 0:
                    best = max(best, moved.value / (e0.elapsed_time(e1) / 1e3) / 1e9)
            gather_peak[name] = best

    # the other shading kernels and the informational configs (single GPU only: they are not part of the scaled job)
    modes = {}
    eager = None
    others = {}
    if world == 1 and not args.no_extras:
        shade_default = model._shade_mode()
        model.mlp_precision = "fp32"
        ref_rgb = step_device()["rgb_map"].clone()
        ms_simt = timed(step_device, args.steps, 1)
        model.mlp_precision = "auto"
        def_err = float((step_device()["rgb_map"] - ref_rgb).abs().max())
        model.mlp_precision = "bf16"
        ms_tc = timed(step_device, args.steps, 1)
        tc_err = float((step_device()["rgb_map"] - ref_rgb).abs().max())
        model.mlp_precision = "auto"
        model.split_app = False
        ms_fused = timed(step_device, args.steps, 1)
        model.split_app = True
        K = args.steps
        modes = {"shading_kernel": shade_default, "max_abs_rgb_default_vs_fp32_simt": def_err,
                 "fp32_simt_mlp_mode": {"value": n * K / (ms_simt / 1e3), "unit": UNIT, "ms_per_step": ms_simt / K},
                 "bf16_mlp_mode": {"value": n * K / (ms_tc / 1e3), "unit": UNIT, "ms_per_step": ms_tc / K,
                                   "max_abs_rgb_vs_fp32": tc_err},
                 "fused_march_mode": {"value": n * K / (ms_fused / 1e3), "unit": UNIT, "ms_per_step": ms_fused / K,
                                      "note": "split_app=False: the one-kernel march of round 1"}}
        try:
            eager, eager_rgb = gpu_eager_baseline(dev, rays_dev)
            eager["max_abs_rgb_vs_ours"] = float((eager_rgb - step_device()["rgb_map"]).abs().max())
            del eager_rgb
        except Exception as exc:        # informational leg: never lose the bench line over it
            eager = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        others = extra_configs(model, dev, syn)
    torch.cuda.empty_cache()
    dp = dp_train(model, dev, syn, world, rank, dist)
    strong = strong_config4(dev, syn, world, dist)

    if rank == 0:
        K = args.steps
        hbm_peak, hbm_src = peaks()
        t_gather = ms_gather / K / 1e3
        t_sigma = max(ms_march - ms_gather, 1e-6) / K / 1e3
        # per launch and per GPU (the kernels of one rank); the counters above are summed over ranks
        per = 1.0 / world
        fetched_bytes = 16.0 * fetched16 * per
        alg_app = 3456.0 * a_samples * per
        alg_sigma = (44.0 * n_global + 32.0 * v0 + 1152.0 * v) * per
        list_bytes = 8.0 * a_samples * per
        feat_bytes = 576.0 * lit * per
        factor_touch = 4.0 * model.packed_factors().numel() * 0.8
        l1_peak = gather_peak.get("l1_resident_64B_gbs") or None
        clk_mhz = clk.result.get("sm_mhz") or 1965.0
        roofline = {
            "bound": "l1-gather", "kernel": "app_gather_kernel",
            "achieved": fetched_bytes / t_gather / 1e9, "peak": l1_peak, "unit": "GB/s",
            "frac": (fetched_bytes / t_gather / 1e9 / l1_peak) if l1_peak else None,
            "peak_source": "tvm_gather_microbench over an L1-resident 64 KB set, measured in this run (LDG.128 by quads, "
                           "8 texels per warp request)",
            "traffic": list_bytes + feat_bytes + factor_touch,
            "traffic_source": "derived: appearance lists read (8 B/sample) + ray_feat rows written (576 B/lit ray) + "
                              "first touch of the appearance factors (80 % of the 69.6 MB set); ncu dram__bytes of the "
                              "same launch is in profiles/",
            "ms_per_launch": ms_gather / K,
            "fetched_bytes_per_launch": fetched_bytes,
            "algorithmic_bytes_per_launch": alg_app,
            "algorithmic_gbs": alg_app / t_gather / 1e9,
            "fetch_reduction": fetched_bytes / alg_app if alg_app else None,
            "units_per_launch": {"app_samples": a_samples * per, "lit_rays": lit * per,
                                 "texel_fetches_16B": fetched16 * per},
            "frac_of_hbm_algorithmic": alg_app / t_gather / 1e9 / hbm_peak,
            "dram_frac": (list_bytes + feat_bytes + factor_touch) / t_gather / 1e9 / hbm_peak,
            "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
            "l1_data_pipe_nominal_gbs": 148 * 128 * clk_mhz * 1e6 / 1e9,
            "measured_gather_ceilings": gather_peak,
            "note": "SURVEY 8d counts 3456 B per appearance sample; the kernel keeps the texels of the current cell in "
                    "registers and fetches `fetch_reduction` of that.  `achieved` = bytes the kernel actually requests "
                    "from L1 (counted by the COUNT build of the same kernel) / its device time; the factor set is "
                    "L2-resident, so HBM is not the bound (`dram_frac`).",
            "stages": {
                "sigma_march_kernel": {
                    "ms_per_launch": (ms_march - ms_gather) / K, "bound": "instruction issue (ncu: issue-active 75 %)",
                    "algorithmic_bytes_per_launch": alg_sigma, "algorithmic_gbs": alg_sigma / t_sigma / 1e9,
                    "frac_of_l1_gather": (alg_sigma / t_sigma / 1e9 / l1_peak) if l1_peak else None,
                    "frac_of_hbm_algorithmic": alg_sigma / t_sigma / 1e9 / hbm_peak,
                    "units_per_launch": {"rays": n_global * per, "occupancy_tests": v0 * per, "sigma_samples": v * per},
                    "note": "includes the ~8 us overflow pass; time = march - gather-only"},
                "march_total": {"ms_per_launch": ms_march / K,
                                "algorithmic_bytes_per_launch": alg_sigma + alg_app,
                                "algorithmic_gbs": (alg_sigma + alg_app) / (ms_march / K / 1e3) / 1e9,
                                "frac_of_l1_gather_algorithmic": ((alg_sigma + alg_app) / (ms_march / K / 1e3) / 1e9
                                                                  / l1_peak) if l1_peak else None,
                                "note": "SURVEY 8d bytes of the whole march / its device time against the measured "
                                        "L1 gather ceiling (round 1: 0.74); the gather kernel's own `frac` above "
                                        "counts only the bytes it still fetches",
                                "frac_of_hbm_algorithmic": (alg_sigma + alg_app) / (ms_march / K / 1e3) / 1e9 / hbm_peak}}}
        line = {"metric": METRIC, "value": n_global * K / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": K, "warmup": args.warmup, "ms_per_step": ms_dev / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config({"parallelism": f"{world} view(s) = {n_global} rays as one ray set, cyclic "
                                                          f"{TILE}-ray tiles over {world} rank(s), replicated factors, "
                                                          "no data-path collective",
                                           "rays_per_step": n_global, "rays_per_gpu_per_step": n_global // world,
                                           "early_term_eps": model.early_term_eps,
                                           "march": "split: sigma-march + appearance gather (+ overflow pass)",
                                           "shading_kernel": model._shade_mode(),
                                           "mlp_arithmetic": "tcgen05 MMA on bf16x3 split operands, fp32 accumulate"}),
                "samples_per_s_nominal": n_global * S * K / (ms_dev / 1e3),
                "sigma_samples_per_s": v * K / (ms_dev / 1e3), "app_samples_per_s": a_samples * K / (ms_dev / 1e3),
                "e2e": {"value": n_global * K / (ms_e2e / 1e3), "unit": UNIT,
                        "h2d_bytes_per_step": int(rays_host.numel() * 4) * world, "d2h_bytes_per_step": int(n * 16) * world,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": int(launches),
                "gpu_launches_note": "tvm_launch_count() delta over the timed region of `value`, rank 0 "
                                     "(per step: sigma-march, appearance gather, overflow pass, shade)",
                "roofline": roofline, "clocks": clk.result,
                "dp_train": dp, "strong": strong}
        if parity is not None:
            line["parity"] = parity
        line.update(modes)
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        if others:
            line["other_configs"] = others
        if not args.no_cpu_baseline and world == 1:
            rate, cores, kind, desc = cpu_baseline_rate(args.cpu_rays)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun when called plainly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
