#!/usr/bin/env python
"""Benchmark of the ray-tracing hot path (driver contract: one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference] [--no-extras]

A *step* is one pass of the hot path over one batch of synthetic input: one ``render`` of the
workload's source grid (a ``render_response_matrix`` for the ``ct3_matrix_*`` workloads, loss +
gradient for ``ct5_cfg5_*``).  The default workload is BASELINE.json configs[1]: HESS CT5, 4096
off-axis point sources on a 64x64 grid of field angles, MCIntegrator(115) -> 876*115 = 100 740
(~1e5) rays per source, hex camera (sensor 0).  ``metric`` = traced rays/s, a ray being one
(source, facet, sample) triple.

``value`` (headline, unchanged since round 1): WEAK scaling -- every rank renders its own 4096-source
grid (a rank-specific sub-pixel shift of the field angles) and the partial images are summed with one
NCCL all-reduce inside the step.

The same JSON line also carries
* ``workloads``: device ms and end-to-end ms of the other BASELINE configurations on ONE GPU (rank 0's):
  the CT3 response matrix at M = 64 and M = 1000 (config 4), the Cassegrain at 1e9 rays (config 3), the CT5
  render on the square lid, and config 5 (loss + gradient w.r.t. the 876 x 3 facet rotations);
* ``strong`` (only under torchrun, N > 1): FIXED total work split N ways -- config 2's 4096 sources (+ all-reduce
  of the image), config 4's matrix rows (no collective; + all-gather as a second figure), config 3's 10 000
  directions (+ all-reduce) -- with the wall time per step as the max over ranks.

``--impl reference``: the reference is pure JAX and JAX is not installable in this image, so the
reference arm times the oracle's C restatement of the reference algorithm (``oracle/cport``,
OpenMP over all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# Brute-force algorithmic flops per ray (FMA = 2): fixed part + per-primitive tests.
F_FIXED = {("hex", "point"): 104, ("square", "point"): 76, ("hex", "parallel"): 92, ("square", "parallel"): 64}
F_CYL, F_BOX = 89, 31

WORKLOADS = {
    # name: scene, M, sources, source_type, sensor_idx, mode
    "ct5_point_4096x115_hex": dict(scene="CT5", M=115, grid=("point", 64, 1.5), sensor=0, mode="render"),
    "ct5_point_4096x115_square": dict(scene="CT5", M=115, grid=("point", 64, 1.5), sensor=2, mode="render"),
    "ct5_point_4096x4096_hex": dict(scene="CT5", M=4096, grid=("point", 64, 1.5), sensor=0, mode="render"),
    # BASELINE configs[1] read literally: "4096 off-axis point sources x 1e5 samples each" per facet = 3.59e11 rays; the samples
    # are streamed from the PRNG key in L2-sized windows (core/streaming.py), no 2.8 GB table
    "ct5_point_4096x100000_hex": dict(scene="CT5", M=100000, grid=("point", 64, 1.5), sensor=0, mode="render"),
    "ct3_matrix_64x64_M64": dict(scene="CT3", M=64, grid=("parallel", 64, 5.5), sensor=0, mode="matrix", roughness=24, seed=42),
    "ct3_matrix_64x64_M1000": dict(scene="CT3", M=1000, grid=("parallel", 64, 5.5), sensor=0, mode="matrix", roughness=24, seed=42),
    # examples/ResponseMatrix.ipynb cell 11 at full size: 512x512 directions x 380 facets x 64 samples = 6.4e9 rays,
    # output 262144 x 960 f32 = 1.0 GB (the notebook reports 30.3 s wall on unstated hardware)
    "ct3_matrix_512x512_M64": dict(scene="CT3", M=64, grid=("parallel", 512, 5.5), sensor=0, mode="matrix", roughness=24, seed=42),
    # BASELINE config 3: Cassegrain (examples/Cassegrain.ipynb cell 3) + synthetic obstructions, 1e9 rays
    "cassegrain_1e9": dict(scene="cassegrain", M=16667, grid=("stars", 10000, 3.0), sensor=0, mode="render"),
    # BASELINE config 5: CT5 + DifferentiableHexagonalSensor(0.5, 1); loss 1/2 |img(theta) - img(theta*)|^2 and its
    # gradient w.r.t. the 876 x 3 facet rotations (theta* = apply_misalignment_to_group(0, 15, 10, key 4242))
    "ct5_cfg5_loss_grad_4096x115_softhex": dict(scene="CT5", M=115, grid=("point", 64, 1.5), sensor=0, mode="grad"),
}
DEFAULT_WORKLOAD = "ct5_point_4096x115_hex"
EXTRA_WORKLOADS = ("ct3_matrix_64x64_M64", "ct3_matrix_64x64_M1000", "cassegrain_1e9", "ct5_point_4096x115_square",
                   "ct5_cfg5_loss_grad_4096x115_softhex", "ct5_point_4096x100000_hex")
STRONG_WORKLOADS = ("ct5_point_4096x115_hex", "ct3_matrix_64x64_M64", "ct3_matrix_64x64_M1000", "cassegrain_1e9",
                    "ct5_point_4096x100000_hex")


def make_sources(w, rank=0):
    from iactrace_b200.workloads import point_grid, parallel_grid, star_field
    kind, n_side, ang = w["grid"]
    if kind == "stars":  # Cassegrain.ipynb cell 8: uniform directions in a 3 deg box, z = -1, normalised
        return star_field(n_side, ang, seed=42 + rank)[0], "parallel"
    if kind == "point":
        src = point_grid(n_side, ang)
        if rank:  # rank-specific sub-pixel shift of the field angles (weak scaling: distinct work per rank)
            src[:, 0] += np.float32(1e10 * np.tan(np.deg2rad(0.003 * rank)))
        return src, "point"
    src = parallel_grid(n_side, ang)
    if rank:
        src[:, 0] += np.float32(5e-5 * rank)
        src /= np.linalg.norm(src, axis=1, keepdims=True)
    return src.astype(np.float32), "parallel"


def load_scene_config(name):
    if name == "cassegrain":
        from iactrace_b200.workloads import cassegrain_config
        return cassegrain_config(True)
    from iactrace_b200.io import load_packed_config
    return load_packed_config(name)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(w, seconds_target=12.0, threads=None, variant="exact"):
    """Time the oracle (CPU restatement of the reference algorithm) on a bounded sample of the workload:
    the C/OpenMP port on all host threads where it applies (single-stage telescopes), else the NumPy form.
    ``variant``: "exact" = the bit-exact build (-O2 -march=x86-64-v2 -ffp-contract=off, one float op per reference
    op) the parity tests use; "native" = the same source at -O3 -march=native -fno-math-errno with FMA contraction,
    built on this host -- the honest speed of the port on these cores."""
    threads = threads or host_threads()      # explicit: torchrun exports OMP_NUM_THREADS=1
    from oracle import cport, prng, scene as oscene, trace as otrace
    cfg = load_scene_config(w["scene"])
    src, stype = make_sources(w)
    val = np.ones(len(src), np.float32)
    if any(m.get("stage", 0) for m in cfg["mirrors"]):
        sc = oscene.build_scene(cfg, min(w["M"], 64), prng.key(w.get("seed", 0)))
        F = sum(len(g["positions"]) for g in sc["groups"] if g["stage"] == 0)
        M = sc["groups"][0]["points"].shape[1]
        n = max(2, min(len(src), int(seconds_target * 2e4 / (F * M))))
        sel = np.arange(n)
        render = lambda idx: otrace.render(sc, src[idx], val[idx], stype, w["sensor"], np.float32)
        t0 = time.perf_counter(); render(sel); dt = time.perf_counter() - t0
        rays = n * F * M
        return rays / dt, 1, f"NumPy oracle: {n} of {len(src)} sources x {F} facets x {M} samples ({rays:.3g} rays, {dt:.1f} s)", (render, sel, rays)
    sc = oscene.build_scene(cfg, min(w["M"], 115), prng.key(w.get("seed", 0)))
    if w.get("roughness"):
        sc = oscene.apply_roughness(sc, w["roughness"])
    prep = cport.prepare(sc, w["sensor"])
    F, M = prep["tp"].shape[:2]
    render = lambda idx: cport.render(prep, src[idx], val[idx], stype, threads=threads, variant=variant)
    # calibrate on 2 sources, then size the sample for ~seconds_target
    t0 = time.perf_counter(); _, nt = render(np.arange(2)); dt = time.perf_counter() - t0
    n = int(max(2, min(len(src), seconds_target / max(dt / 2, 1e-6))))
    sel = np.linspace(0, len(src) - 1, n).astype(int)
    t0 = time.perf_counter(); render(sel); dt = time.perf_counter() - t0
    rays = n * F * M
    return rays / dt, nt, f"C/OpenMP oracle ({variant} build): {n} of {len(src)} sources x {F} facets x {M} samples ({rays:.3g} rays, {dt:.1f} s)", (render, sel, rays)


def run_reference(args, w, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (the oracle) on the host cores."""
    if rank != 0:
        return
    rate, nt, sample, (render, sel, rays_per_step) = cpu_reference_rate(w, seconds_target=2.0)
    for _ in range(args.warmup):
        render(sel[:2])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        render(sel)
    dt = time.perf_counter() - t0
    value = rays_per_step * args.steps / dt
    native = None
    try:
        r2, nt2, sample2, _ = cpu_reference_rate(w, seconds_target=4.0, variant="native")
        native = {"value": r2, "unit": "rays/s", "cores": nt2, "kind": "port", "sample": sample2}
    except Exception as e:  # pragma: no cover
        native = {"unavailable": str(e)[:200]}
    line = {"impl": "reference", "metric": "traced_rays_per_second", "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "reference = JAX (not installable here); timed: the oracle's CPU restatement of the reference algorithm (bit-exact build), brute-force obstruction tests as in the reference; each step is a bounded sample of the workload's sources, so only the rays/s rates compare with the GPU arm, not ms_per_step"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": nt, "kind": "port", "sample": sample},
            "cpu_baseline_native_build": native,
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
class Workload:
    """One benchmark workload on this rank's GPU: the telescope, host + device inputs and the step functions."""

    def __init__(self, name, dev, rank=0, shard=None):
        import torch
        import iactrace_b200 as I
        from iactrace_b200.io import build_telescope
        self.name, self.w, self.dev = name, WORKLOADS[name], dev
        w = self.w
        self.tel = build_telescope(load_scene_config(w["scene"]), I.MCIntegrator(w["M"]), I.random.key(w.get("seed", 0)))
        if w.get("roughness"):
            self.tel = self.tel.apply_roughness(w["roughness"])
        src_np, self.stype = make_sources(w, rank)
        val_np = np.ones(len(src_np), np.float32)
        if w["grid"][0] == "stars":
            from iactrace_b200.workloads import star_field
            val_np = star_field(len(src_np))[1]
        if shard is not None:                                        # strong scaling: this rank's slice of the FIXED job
            from iactrace_b200.parallel import shard_bounds
            a, b = shard_bounds(len(src_np), *shard)
            src_np, val_np = src_np[a:b], val_np[a:b]
        self.n_sources = len(src_np)
        self.src_host = torch.from_numpy(np.ascontiguousarray(src_np)).pin_memory()
        self.val_host = torch.from_numpy(np.ascontiguousarray(val_np)).pin_memory()
        self.src_dev, self.val_dev = self.src_host.to(dev), self.val_host.to(dev)
        self.F = sum(len(g) for g in self.tel.mirror_groups if g.optical_stage == 0)
        self.rays_per_step = self.n_sources * self.F * w["M"]
        self.mode = w["mode"]
        self.out_host = None
        if self.mode == "grad":
            from iactrace_b200.sensors import DifferentiableHexagonalSensor
            from iactrace_b200.core import render
            hard = self.tel.sensors[0]
            self.tel = self.tel.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers,
                                                                             0.5, 1, grid=hard.grid_constants()), 0)
            with torch.no_grad():
                self.target = render(self.tel.apply_misalignment_to_group(0, 15, 10, I.random.key(4242)), self.src_dev,
                                     self.val_dev, self.stype, 0)
            self.rot = self.tel.mirror_groups[0].rotations
            self.rot.requires_grad_(True)
            self.grad_host = torch.empty((self.F, 3), dtype=torch.float32, pin_memory=True)

    def step(self, src, val, world=1, collective=True):
        import torch.distributed as dist
        from iactrace_b200.core import render, render_response_matrix
        if self.mode == "matrix":
            return render_response_matrix(self.tel, src, val, self.stype, self.w["sensor"])
        if self.mode == "grad":
            self.rot.grad = None
            loss = 0.5 * ((render(self.tel, src, val, self.stype, 0) - self.target) ** 2).sum()
            loss.backward()
            return loss.detach()
        img = render(self.tel, src, val, self.stype, self.w["sensor"])
        if world > 1 and collective:
            dist.all_reduce(img)
        return img

    def step_device(self, world=1, collective=True):
        return self.step(self.src_dev, self.val_dev, world, collective)

    def step_e2e(self, world=1):
        """Through the public API with HOST buffers: pinned sources/values -> device, result (image / matrix / loss +
        gradient) -> pinned host memory, synchronised before the next step starts."""
        import torch
        s = self.src_host.to(self.dev, non_blocking=True)
        v = self.val_host.to(self.dev, non_blocking=True)
        out = self.step(s, v, world)
        if self.out_host is None:        # pinned result buffer, allocated once (a pageable .cpu() copy of the 15.7 MB /
            self.out_host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)   # 1 GB matrices runs at a tenth of PCIe)
        self.out_host.copy_(out, non_blocking=True)
        if self.mode == "grad":
            self.grad_host.copy_(self.rot.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.out_host

    def h2d_bytes(self):
        return (self.src_host.numel() + self.val_host.numel()) * 4

    def d2h_bytes(self):
        n = self.out_host.numel() * 4 if self.out_host is not None else 0
        return n + (self.grad_host.numel() * 4 if self.mode == "grad" else 0)


def time_device(wl, steps, warmup, flush_buf, barrier, world=1, collective=True):
    """Per-step CUDA-event times (ms) of ``steps`` device-resident steps; the L2 is flushed (untimed) before each."""
    import torch
    for _ in range(warmup):
        wl.step_device(world, collective)
    barrier()
    evs = []
    for _ in range(steps):
        flush_buf.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wl.step_device(world, collective)
        e1.record()
        evs.append((e0, e1))
    barrier()
    return [a.elapsed_time(b) for a, b in evs]


def time_e2e(wl, steps, barrier, world=1):
    for _ in range(2):
        wl.step_e2e(world)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        wl.step_e2e(world)
    barrier()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `workloads` and `strong` blocks")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist
    from iactrace_b200 import _native as N
    from iactrace_b200.core.render import build_scene

    N.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = Workload(args.workload, dev, rank)
    tel, src_np, stype = wl.tel, wl.src_host.numpy(), wl.stype
    rays_per_step, F = wl.rays_per_step, wl.F
    matrix = wl.mode == "matrix"
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    warmup = max(args.warmup, 3)

    # ---- timed region: exactly K steps; L2 flushed (untimed) between steps; device time, max over ranks
    for _ in range(warmup):
        wl.step_device(world)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = N.lib().iact_launch_count()
    t_wall0 = time.perf_counter()
    step_ms = time_device(wl, args.steps, 0, flush_buf, barrier, world)
    t_wall = time.perf_counter() - t_wall0
    launches = N.lib().iact_launch_count() - launches0
    clocks = sampler.stop()
    total_ms = max_over_ranks(sum(step_ms))
    value = rays_per_step * world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers (H2D + D2H inside the timed region)
    e2e_s = max_over_ranks(time_e2e(wl, args.steps, barrier, world))
    e2e_value = rays_per_step * world * args.steps / e2e_s
    h2d, d2h = wl.h2d_bytes(), wl.d2h_bytes()

    # ---- fixed-size jobs split over the ranks (strong scaling)
    strong = None
    if world > 1 and not args.no_extras:
        strong = {}
        for name in STRONG_WORKLOADS:
            s_wl = Workload(name, dev, 0, shard=(rank, world))
            k = 20 if s_wl.rays_per_step * world < 2e9 else (10 if s_wl.rays_per_step * world < 1e10 else 2)
            ms = max_over_ranks(float(np.median(time_device(s_wl, k, 3 if k > 2 else 1, flush_buf, barrier, world))))
            total_rays = s_wl.F * WORKLOADS[name]["M"] * len(make_sources(WORKLOADS[name])[0])
            entry = {"n_gpus": world, "ms_per_step": ms, "rays_per_step_total": total_rays, "rays_per_s": total_rays / (ms * 1e-3),
                     "steps": k, "split": "sources (matrix rows)" if s_wl.mode == "matrix" else "sources",
                     "collective": "none (rows are rank-owned)" if s_wl.mode == "matrix" else "NCCL all-reduce of the image inside the step"}
            if s_wl.mode != "matrix":
                ms_nc = max_over_ranks(float(np.median(time_device(s_wl, k, 1, flush_buf, barrier, world, collective=False))))
                entry["ms_per_step_without_allreduce"] = ms_nc
            else:
                from iactrace_b200.parallel import response_matrix_sharded
                full_src, _ = make_sources(WORKLOADS[name])
                fs = torch.from_numpy(full_src).to(dev)
                fv = torch.ones(len(full_src), device=dev)
                ev = []
                for i in range(3 + 5):
                    flush_buf.fill_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    response_matrix_sharded(s_wl.tel, fs, fv, s_wl.stype, s_wl.w["sensor"], gather=True)
                    e1.record()
                    ev.append((e0, e1))
                barrier()
                entry["ms_per_step_with_allgather"] = max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in ev[3:]])))
            strong[name] = entry
            del s_wl
        torch.cuda.empty_cache()

    if rank == 0:
        # ---- the other BASELINE configurations on this GPU: device ms (CUDA events) and end-to-end ms (host buffers)
        workloads = None
        if not args.no_extras and world == 1:           # one-GPU figures; under torchrun the other ranks would idle
            workloads = {}
            for name in EXTRA_WORKLOADS:
                x = Workload(name, dev, 0)
                k = 10 if x.rays_per_step < 1e10 else 2
                ms = float(np.median(time_device(x, k, 3 if k > 2 else 1, flush_buf, torch.cuda.synchronize)))
                e2e_ms = 1e3 * time_e2e(x, k, torch.cuda.synchronize) / k
                workloads[name] = {"device_ms": ms, "e2e_ms": e2e_ms, "rays_per_step": x.rays_per_step,
                                   "rays_per_s_device": x.rays_per_step / (ms * 1e-3), "steps": k,
                                   "h2d_bytes_per_step": x.h2d_bytes(), "d2h_bytes_per_step": x.d2h_bytes(),
                                   "step": {"matrix": "render_response_matrix", "grad": "loss + backward (VJP kernel)"}.get(x.mode, "render")}
                del x
            torch.cuda.empty_cache()

        # ---- roofline of the dominant kernel (trace_kernel): FP32 CUDA-core bound, not HBM / tensor
        d = C.c_double()
        N.check(N.lib().iact_probe_fp32(8192, C.byref(d), None))
        fp32_peak = d.value / 1e12
        N.check(N.lib().iact_probe_smem_atomics(4096, 32, C.byref(d), None))
        atom_peak = d.value
        # candidate-list lengths actually tested per ray after exact culling
        keep = []
        sc, _ = build_scene(tel, w["sensor"], keep)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        N.check(N.lib().iact_cull_stats(sc, N.ptr(wl.src_dev), len(src_np), 0 if stype == "point" else 1,
                                        stats.data_ptr(), None))
        torch.cuda.synchronize()
        n_cyl_kept, n_oth_kept, n_rays_stat, n_lvl1 = [int(x) for x in stats.tolist()]
        n_pairs = max(n_rays_stat, 1)                  # per-ray averages
        n_lvl1 = n_lvl1 * w["M"]
        sensor_kind = "hex" if hasattr(tel.sensors[w["sensor"]], "hex_size") else "square"
        n_cyl = sc.n_cyl
        n_oth = sc.n_box + sc.n_sph + sc.n_obox + sc.n_tri
        f_fixed = F_FIXED[(sensor_kind, stype)]
        n_sec = sum(len(g) for g in tel.mirror_groups if g.optical_stage > 0)
        # each extra optical stage: 530 flop per (ray, mirror) + a brute-force shadow test of that leg (App. C)
        f_stage = n_sec * 530 + (F_CYL * n_cyl + F_BOX * n_oth if n_sec else 0)
        f_brute = f_fixed + F_CYL * n_cyl + F_BOX * n_oth + f_stage
        f_culled = f_fixed + F_CYL * n_cyl_kept / max(n_pairs, 1) + F_BOX * n_oth_kept / max(n_pairs, 1) + f_stage
        kern_s = total_ms * 1e-3 / args.steps
        achieved = rays_per_step * f_culled / kern_s / 1e12
        out_elems = wl.out_host.numel() if wl.out_host is not None else 0
        table_bytes = F * w["M"] * 32 + F * 16            # packed world table (two float4 per sample) + facet bounds
        alg_bytes = out_elems * 4 + h2d + table_bytes     # image / matrix + sources + the sample table, each once
        ncu = _ncu_figures(args.workload)
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                    "traffic": ncu.get("dram_bytes"),
                    "peak_source": "measured live: iact_probe_fp32 (dependent-free FFMA chains, all SMs); MEASURED_PEAKS.json has no FP32 CUDA-core figure",
                    "flops_per_ray": {"after_exact_culling": f_culled, "brute_force_reference": f_brute,
                                      "mean_cylinders_tested": n_cyl_kept / max(n_pairs, 1),
                                      "mean_other_tested": n_oth_kept / max(n_pairs, 1),
                                      "mean_level1_list": n_lvl1 / max(n_pairs, 1)},
                    "brute_force_equivalent_tflops": rays_per_step * f_brute / kern_s / 1e12,
                    # what the SMs actually executed, from the committed ncu capture of this command (not live):
                    "executed_fp32_frac": ncu.get("executed_fp32_frac"), "issue_slot_util": ncu.get("issue_slot_util"),
                    "thread_inst_per_ray": ncu.get("thread_inst_per_ray"), "ncu_source": ncu.get("source"),
                    "atomic": {"unit": "shared-memory atomic instructions/s (one f32 atomicAdd per lane, 32 distinct addresses)",
                               "peak": atom_peak, "peak_source": "measured live: iact_probe_smem_atomics",
                               "achieved": (ncu["shared_atom_inst"] * 32 / kern_s) if ncu.get("shared_atom_inst") else None,
                               "frac": (ncu["shared_atom_inst"] * 32 / kern_s / atom_peak) if ncu.get("shared_atom_inst") else None,
                               "note": "the register pixel cache leaves one shared atomic per ~200 rays: the atomic side of the roofline is idle"},
                    "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / kern_s / 1e9,
                            "peak_gbs": _hbm_peak(), "note": "output + sources + sample table, each once; per-ray HBM bytes ~ 0: not the bound"}}
        cpu_baseline = cpu_native = None
        if not args.no_cpu_baseline and world == 1:
            rate, nt, sample, _ = cpu_reference_rate(w, seconds_target=10.0)
            cpu_baseline = {"value": rate, "unit": "rays/s", "cores": nt, "kind": "port", "sample": sample,
                            "build": "bit-exact: gcc -O2 -march=x86-64-v2 -ffp-contract=off"}
            try:
                r2, nt2, sample2, _ = cpu_reference_rate(w, seconds_target=6.0, variant="native")
                cpu_native = {"value": r2, "unit": "rays/s", "cores": nt2, "kind": "port", "sample": sample2,
                              "build": "gcc -O3 -march=native -fno-math-errno (FMA contraction on), built on this host"}
            except Exception as e:  # pragma: no cover
                cpu_native = {"unavailable": str(e)[:200]}
        line = {"metric": "traced_rays_per_second", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scene": w["scene"], "n_sources_per_gpu": len(src_np), "n_facets": F,
                           "n_samples_per_facet": w["M"], "rays_per_step_per_gpu": rays_per_step, "sensor_idx": w["sensor"],
                           "mode": w["mode"], "source_type": stype, "seed": w.get("seed", 0),
                           "l2": "flushed between steps (256 MiB write, outside the per-step CUDA events)",
                           "world_table": "transform_to_world output cached on the Telescope across steps (the reference re-runs it inside every render, mirrors.py:64-79; O(F*M), < 1 % of a step)",
                           "parallelism": f"sources sharded x{world}" + (", NCCL all-reduce of the image per step" if world > 1 and not matrix else "")},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_baseline_native_build": cpu_native,
                "workloads": workloads, "strong": strong,
                "wall_s_timed_region": t_wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _ncu_figures(workload):
    """Figures of the dominant kernel from the committed ncu capture of this command (profiles/ncu_figures.json,
    written by tools/summarize_ncu.py): DRAM bytes per launch, executed FP32 fraction, issue-slot utilisation."""
    try:
        return json.loads((ROOT / "profiles" / "ncu_figures.json").read_text())[workload]
    except Exception:
        return {}


def _hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return json.loads(p.read_text())["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    main()
