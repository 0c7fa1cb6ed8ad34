#!/usr/bin/env python
"""Benchmark of the ray-tracing hot path (driver contract: one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic input: one ``render`` of the
workload's source grid (a ``render_response_matrix`` for the ``ct3_matrix_*`` workloads).  The
default workload is BASELINE.json configs[1]: HESS CT5, 4096 off-axis point sources on a 64x64 grid
of field angles, MCIntegrator(115) -> 876*115 = 100 740 (~1e5) rays per source, hex camera
(sensor 0).  ``metric`` = traced rays/s, a ray being one (source, facet, sample) triple.

Multi-GPU (torchrun, one rank per GPU): weak scaling -- every rank renders its own 4096-source
grid (a rank-specific sub-pixel shift of the field angles) and the partial images are summed
with one NCCL all-reduce inside the step.

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
    "ct3_matrix_64x64_M64": dict(scene="CT3", M=64, grid=("parallel", 64, 5.5), sensor=0, mode="matrix", roughness=24, seed=42),
    "ct3_matrix_64x64_M1000": dict(scene="CT3", M=1000, grid=("parallel", 64, 5.5), sensor=0, mode="matrix", roughness=24, seed=42),
    # examples/ResponseMatrix.ipynb cell 11 at full size: 512x512 directions x 380 facets x 64 samples = 6.4e9 rays,
    # output 262144 x 960 f32 = 1.0 GB (the notebook reports 30.3 s wall on unstated hardware)
    "ct3_matrix_512x512_M64": dict(scene="CT3", M=64, grid=("parallel", 512, 5.5), sensor=0, mode="matrix", roughness=24, seed=42),
    # BASELINE config 3: Cassegrain (examples/Cassegrain.ipynb cell 3) + synthetic obstructions, 1e9 rays
    "cassegrain_1e9": dict(scene="cassegrain", M=16667, grid=("stars", 10000, 3.0), sensor=0, mode="render"),
}
DEFAULT_WORKLOAD = "ct5_point_4096x115_hex"


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


def cpu_reference_rate(w, seconds_target=12.0, threads=None):
    """Time the oracle (CPU restatement of the reference algorithm) on a bounded sample of the workload:
    the C/OpenMP port on all host threads where it applies (single-stage telescopes), else the NumPy form."""
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
    render = lambda idx: cport.render(prep, src[idx], val[idx], stype, threads=threads)
    # calibrate on 2 sources, then size the sample for ~seconds_target
    t0 = time.perf_counter(); _, nt = render(np.arange(2)); dt = time.perf_counter() - t0
    n = int(max(2, min(len(src), seconds_target / max(dt / 2, 1e-6))))
    sel = np.linspace(0, len(src) - 1, n).astype(int)
    t0 = time.perf_counter(); render(sel); dt = time.perf_counter() - t0
    rays = n * F * M
    return rays / dt, nt, f"C/OpenMP oracle: {n} of {len(src)} sources x {F} facets x {M} samples ({rays:.3g} rays, {dt:.1f} s)", (render, sel, rays)


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
    line = {"impl": "reference", "metric": "traced_rays_per_second", "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "reference = JAX (not installable here); timed: the oracle's CPU restatement of the reference algorithm, brute-force obstruction tests as in the reference"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": nt, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    import iactrace_b200 as I
    from iactrace_b200 import _native as N
    from iactrace_b200.core import render as render_fn, render_response_matrix
    from iactrace_b200.core.render import build_scene
    from iactrace_b200.io import build_telescope, load_packed_config

    N.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    tel = build_telescope(load_scene_config(w["scene"]), I.MCIntegrator(w["M"]), I.random.key(w.get("seed", 0)))
    if w.get("roughness"):
        tel = tel.apply_roughness(w["roughness"])
    src_np, stype = make_sources(w, rank)
    val_np = np.ones(len(src_np), np.float32)
    if w["grid"][0] == "stars":
        from iactrace_b200.workloads import star_field
        val_np = star_field(len(src_np))[1]
    src_host = torch.from_numpy(src_np).pin_memory()
    val_host = torch.from_numpy(val_np).pin_memory()
    src_dev, val_dev = src_host.to(dev), val_host.to(dev)
    F = sum(len(g) for g in tel.mirror_groups if g.optical_stage == 0)
    rays_per_step = len(src_np) * F * w["M"]
    matrix = w["mode"] == "matrix"

    def step_device():
        if matrix:
            return render_response_matrix(tel, src_dev, val_dev, stype, w["sensor"])
        img = render_fn(tel, src_dev, val_dev, stype, w["sensor"])
        if world > 1:
            dist.all_reduce(img)
        return img

    out_host = []                                   # pinned result buffer, allocated once (a pageable .cpu() copy of the
                                                    # 15.7 MB / 1 GB response matrices runs at a tenth of the PCIe rate)

    def step_e2e():
        s = src_host.to(dev, non_blocking=True)
        v = val_host.to(dev, non_blocking=True)
        if matrix:
            out = render_response_matrix(tel, s, v, stype, w["sensor"])
        else:
            out = render_fn(tel, s, v, stype, w["sensor"])
            if world > 1:
                dist.all_reduce(out)
        if not out_host:
            out_host.append(torch.empty(out.shape, dtype=out.dtype, pin_memory=True))
        out_host[0].copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the step's result is on the host before the next step starts
        return out_host[0]

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---- timed region: exactly K steps; L2 flushed (untimed) between steps; device time, max over ranks
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = N.lib().iact_launch_count()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_buf.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = N.lib().iact_launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = rays_per_step * world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers (H2D + D2H inside the timed region)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = rays_per_step * world * args.steps / float(e2e_s.item())
    h2d = src_host.numel() * 4 + val_host.numel() * 4
    d2h = out.numel() * 4

    if rank == 0:
        # ---- roofline of the dominant kernel (trace_kernel): FP32 CUDA-core bound, not HBM / tensor
        d = C.c_double()
        N.check(N.lib().iact_probe_fp32(8192, C.byref(d), None))
        fp32_peak = d.value / 1e12
        # candidate-list lengths actually tested per ray after exact culling
        keep = []
        sc, _ = build_scene(tel, w["sensor"], keep)
        stats = torch.zeros(4, dtype=torch.int64, device=dev)
        N.check(N.lib().iact_cull_stats(sc, N.ptr(src_dev), len(src_np), 0 if stype == "point" else 1,
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
        out_bytes = (out.numel() * 4) + h2d  # algorithmic HBM bytes per launch: image + sources (tables are L2-resident)
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                    "traffic": _traffic(args.workload),
                    "peak_source": "measured live: iact_probe_fp32 (dependent-free FFMA chains, all SMs); MEASURED_PEAKS.json has no FP32 CUDA-core figure",
                    "flops_per_ray": {"after_exact_culling": f_culled, "brute_force_reference": f_brute,
                                      "mean_cylinders_tested": n_cyl_kept / max(n_pairs, 1),
                                      "mean_other_tested": n_oth_kept / max(n_pairs, 1),
                                      "mean_level1_list": n_lvl1 / max(n_pairs, 1)},
                    "brute_force_equivalent_tflops": rays_per_step * f_brute / kern_s / 1e12,
                    "hbm": {"algorithmic_bytes_per_launch": out_bytes, "achieved_gbs": out_bytes / kern_s / 1e9,
                            "peak_gbs": _hbm_peak(), "note": "per-ray HBM bytes ~ 0: not the bound"}}
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            rate, nt, sample, _ = cpu_reference_rate(w, seconds_target=12.0)
            cpu_baseline = {"value": rate, "unit": "rays/s", "cores": nt, "kind": "port", "sample": sample}
        line = {"metric": "traced_rays_per_second", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "scene": w["scene"], "n_sources_per_gpu": len(src_np), "n_facets": F,
                           "n_samples_per_facet": w["M"], "rays_per_step_per_gpu": rays_per_step, "sensor_idx": w["sensor"],
                           "mode": w["mode"], "source_type": stype, "seed": w.get("seed", 0),
                           "l2": "flushed between steps (256 MiB write, outside the per-step CUDA events)",
                           "parallelism": f"sources sharded x{world}" + (", NCCL all-reduce of the image per step" if world > 1 and not matrix else "")},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "wall_s_timed_region": t_wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (null if not captured)."""
    try:
        return json.loads((ROOT / "profiles" / "traffic.json").read_text())[workload]["bytes"]
    except Exception:
        return None


def _hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return json.loads(p.read_text())["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    main()
