"""Ad-hoc GPU exploration: timings of the main workloads + roofline probes (not a bench contract)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import iactrace_b200 as I
from iactrace_b200 import _native as N
from iactrace_b200 import config as Rm
from iactrace_b200.core import render, render_response_matrix
from iactrace_b200.io import build_telescope, load_packed_config
from _bridge import point_grid, parallel_grid


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    print(torch.cuda.get_device_name(0))
    d = C.c_double()
    N.check(N.lib().iact_probe_fp32(4096, C.byref(d), None)); print(f"fp32 fma probe: {d.value/1e12:.2f} TFLOP/s")
    for nd in (1, 2, 4, 32):
        N.check(N.lib().iact_probe_smem_atomics(4096, nd, C.byref(d), None)); print(f"smem atomics ({nd} distinct/warp): {d.value/1e9:.1f} G/s")

    t0 = time.time()
    ct5 = load_packed_config("CT5")
    for M in (115, 1024):
        tel = build_telescope(ct5, I.MCIntegrator(M), I.random.key(0)); torch.cuda.synchronize()
        print(f"CT5 load+sample M={M}: {time.time()-t0:.2f}s")
        src = torch.from_numpy(point_grid(64, 1.5)).cuda(); val = torch.ones(len(src), device="cuda")
        rays = len(src) * 876 * M
        for sensor in (0, 2):
            for cull in (True, False):
                if not cull and M > 115:
                    continue
                Rm.cull_obstructions = cull
                med, mn = timeit(lambda: render(tel, src, val, "point", sensor), n=3 if not cull else 5, warm=1)
                print(f"CT5 render S=4096 M={M} sensor={sensor} cull={cull}: {med:.2f} ms  -> {rays/med/1e6:.1f} Grays/s")
        Rm.cull_obstructions = True
    # host-side cost of one render call (ctypes + scene packing + launches), and back-to-back device time
    tel = build_telescope(ct5, I.MCIntegrator(115), I.random.key(0))
    src = torch.from_numpy(point_grid(64, 1.5)).cuda(); val = torch.ones(len(src), device="cuda")
    render(tel, src, val, "point", 0); torch.cuda.synchronize()
    host = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        t = time.perf_counter(); render(tel, src, val, "point", 0); host.append(time.perf_counter() - t)
    e1.record(); torch.cuda.synchronize()
    print(f"20 back-to-back renders: {e0.elapsed_time(e1)/20:.2f} ms/call device; host call median {np.median(host)*1e6:.0f} us max {np.max(host)*1e6:.0f} us")
    host = []
    for _ in range(10):
        torch.cuda.synchronize()
        t = time.perf_counter(); render(tel, src, val, "point", 0); host.append(time.perf_counter() - t)
    torch.cuda.synchronize()
    print(f"render after sync: host call median {np.median(host)*1e6:.0f} us max {np.max(host)*1e6:.0f} us")
    ct3 = load_packed_config("CT3")
    for M in (64, 1000):
        tel = build_telescope(ct3, I.MCIntegrator(M), I.random.key(42)).apply_roughness(24)
        src = torch.from_numpy(parallel_grid(64, 5.5)).cuda(); val = torch.ones(len(src), device="cuda")
        rays = len(src) * 380 * M
        med, mn = timeit(lambda: render_response_matrix(tel, src, val, "parallel", 0))
        print(f"CT3 response matrix 64x64 M={M}: {med:.2f} ms -> {rays/med/1e6:.1f} Grays/s")
        med, mn = timeit(lambda: render(tel, src, val, "parallel", 0))
        print(f"CT3 render 64x64 M={M}: {med:.2f} ms -> {rays/med/1e6:.1f} Grays/s")
    # BASELINE config 5: CT5 + soft hex sensor, S=64, M=115: forward + VJP w.r.t. facet rotations
    from iactrace_b200.sensors import DifferentiableHexagonalSensor
    from iactrace_b200._util import replace
    tel = build_telescope(ct5, I.MCIntegrator(115), I.random.key(0))
    hard = tel.sensors[0]
    tel = tel.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, 0.5, 1, grid=hard.grid_constants()), 0)
    for S_side in (8, 64):
        src = torch.from_numpy(point_grid(S_side, 1.5)).cuda(); val = torch.ones(len(src), device="cuda")
        target = render(tel.apply_misalignment_to_group(0, 15, 10, I.random.key(4242)), src, val, "point", 0)
        g = tel.mirror_groups[0]
        rot = g.rotations.detach().clone().requires_grad_(True)
        def fwd_bwd():
            rot.grad = None
            t = replace(tel, mirror_groups=[replace(g, rotations=rot)])
            loss = 0.5 * ((render(t, src, val, "point", 0) - target) ** 2).sum()
            loss.backward()
        med, mn = timeit(fwd_bwd)
        rays = len(src) * 876 * 115
        print(f"CT5 config5 soft-hex S={len(src)}: forward+VJP {med:.2f} ms -> {rays/med/1e6:.1f} Grays/s (fwd+bwd), |grad| max {float(rot.grad.abs().max()):.3g}")
    tel = build_telescope(ct3, I.MCIntegrator(1000), I.random.key(0))
    s1 = torch.tensor([[0., 0., 1e10]], device="cuda"); v1 = torch.ones(1, device="cuda")
    for sensor in (0, 1):
        med, mn = timeit(lambda: render(tel, s1, v1, "point", sensor))
        print(f"CT3 config1 sensor={sensor}: {med*1e3:.1f} us")


if __name__ == "__main__":
    main()
